/*
 * umhs.h -- C ABI of libumhs_sm100.so: the B200 (sm_100a) kernels behind the UMHS
 * nerfstudio plugin's hot path.
 *
 * The reference (Factral/unsupervised-hyperspectral-nerf) is pure Python: it owns no
 * FFI.  Each entry point below replaces the chain of ATen / nerfacc / nerfstudio calls
 * the reference makes at the cited file:line; the Python host code in
 * unsupervised-hyperspectral-nerf_b200/umhs_b200/ binds them with ctypes and keeps the
 * reference's UMHSField / SpectralRenderer / UMHSModel interfaces.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless said
 *    otherwise; the CALLER allocates every buffer (torch does, in the plugin);
 *  - kernels never allocate, never synchronise the device, and are enqueued on
 *    `stream` (a cudaStream_t passed as void*); no global mutable state: safe to call
 *    from any host thread on any stream;
 *  - return 0 on success, a negative code on error; umhs_last_error() returns a
 *    thread-local message.  There is no CPU fallback.
 *  - packed layout: N samples sorted by ray; packed_info[r] = (start, count), int64;
 *    ray_indices int64 (nerfacc 0.5.2 dtypes).
 */
#ifndef UMHS_H_
#define UMHS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UMHS_OK 0
#define UMHS_ERR_INVALID (-1)
#define UMHS_ERR_CUDA (-2)
#define UMHS_ERR_UNSUPPORTED (-3)

#define UMHS_MAX_LEVELS 16
#define UMHS_MAX_CLASSES 16
#define UMHS_MAX_BANDS 512
#define UMHS_MAX_RANKS 8 /* GPUs of one NVSwitch box */
#define UMHS_GEO_FEAT 15 /* geo_feat_dim of NerfactoField */
#define UMHS_PE_DIM 12   /* NeRFEncoding(3, 2 freqs) */
#define UMHS_SH_DIM 16   /* SHEncoding(levels=4) */
#define UMHS_HIDDEN 64   /* hidden_dim / hidden_dim_color */
#define UMHS_SPEC_HIDDEN 16

const char *umhs_last_error(void);
int umhs_version(void);

/* ---------------------------------------------------------------------------------
 * Sample points.  A "point set" is either explicit world positions [n,3], or the
 * packed samples of a ray batch: position = o[ray] + d[ray] * (t0 + t1) / 2
 * (nerfstudio Frustums.get_positions, used at ref umhsnerf/umhs_field.py:183,303).
 * ------------------------------------------------------------------------------- */
typedef struct {
    const float *positions;     /* [n,3] world positions, or NULL to use the ray form */
    const float *directions;    /* [n,3] per-point directions (explicit form), may be NULL */
    const float *rays_o;        /* [R,3] */
    const float *rays_d;        /* [R,3] */
    const int64_t *ray_indices; /* [n]   */
    const float *t_starts;      /* [n]   */
    const float *t_ends;        /* [n]   */
    int32_t contraction;        /* 1: SceneContraction(inf) then (x+2)/4; 0: aabb-normalise */
    float aabb[6];              /* min xyz, max xyz (used when contraction == 0) */
    const int32_t *enc_index;   /* optional [n]: row of point i in the `enc` array handed to the MLP entry points
                                   (lets the kept samples reuse the encodings of the sampler's candidate pass) */
    const int64_t *n_dev;       /* optional DEVICE count: kernels process min(n, *n_dev) points, so a step can run
                                   on capacity-sized buffers without a host sync on the sample count */
} umhs_points;

/* ---------------------------------------------------------------------------------
 * (1) Multiresolution hash grid  -- replaces HashEncoding.pytorch_fwd reached from
 *     ref umhsnerf/umhs_field.py:320 (self.mlp_base(positions_flat)) and its autograd
 *     backward (index_put accumulate).
 *     table [L*T, 2] f32, level l at rows [l*T,(l+1)*T); scalings [L] f32 (state-dict
 *     buffer encoder.scalings); enc / d_enc [n, 2L] f32 row-major, level-major columns.
 * ------------------------------------------------------------------------------- */
typedef struct {
    const float *table;
    const float *scalings;
    int32_t num_levels;
    int32_t log2_table_size;
} umhs_hash_grid;

int umhs_hash_encode_fwd(const umhs_points *pts, int64_t n, const umhs_hash_grid *grid, float *enc, void *stream);
/* The same with the encodings / their gradients in the bf16 tile layout (16 levels x 2 features; element (n, c) at byte
 * (n >> 7) * 8192 + (c >> 3) * 2048 + ((n & 127) >> 3) * 128 + (n & 7) * 16 + (c & 7) * 2): the format the tensor-core
 * MLP kernels consume without conversion (umhs_mlp_params.enc_tiles = 1), half the bytes of the fp32 rows.  Buffers
 * hold ceil(n / 128) * 8192 bytes, 16-byte aligned. */
int umhs_hash_encode_fwd_tiles(const umhs_points *pts, int64_t n, const umhs_hash_grid *grid, void *enc_tiles,
                               void *stream);
int umhs_hash_encode_bwd_tiles(const umhs_points *pts, int64_t n, const umhs_hash_grid *grid, const void *d_enc_tiles,
                               float *d_table, void *stream);
/* ... levels [level_begin, level_end) only: the table gradient of the other levels is untouched.  The data-parallel
 * trainer scatters the fine half first and exchanges it over NVLink while the coarse half is being scattered. */
int umhs_hash_encode_bwd_tiles_levels(const umhs_points *pts, int64_t n, const umhs_hash_grid *grid,
                                      const void *d_enc_tiles, float *d_table, int32_t level_begin, int32_t level_end,
                                      void *stream);
int umhs_hash_encode_bwd(const umhs_points *pts, int64_t n, const umhs_hash_grid *grid, const float *d_enc,
                         float *d_table, void *stream);
/* Same two ops on already-normalised coordinates x01 [n,3] in [0,1] (the HashEncoding module boundary). */
int umhs_hash_encode_x01_fwd(const float *x01, int64_t n, const umhs_hash_grid *grid, float *enc, void *stream);
int umhs_hash_encode_x01_bwd(const float *x01, int64_t n, const umhs_hash_grid *grid, const float *d_enc,
                             float *d_table, void *stream);

/* ---------------------------------------------------------------------------------
 * (2) Fused field MLPs -- replaces, per sample, ref umhsnerf/umhs_field.py:
 *     :320-328  base MLP 32->64->16, trunc_exp, selector      -> sigma, emb[15]
 *     :183-184  NeRF position encoding of the WORLD position  -> pe[12]
 *     :211-212  mlp_head 27->64->64->K, sigmoid               -> scalar[K]
 *     :214-226  feature_mlp 27->64->64->K(+1), softmax(/T), sigmoid -> abund[K], s1
 *     :242-250  mlp_directional first layer [SH16,pe12]->16 ReLU    -> spec_hidden[16]
 *               (its 16->B sigmoid layer is evaluated inside the compositing kernel)
 *     All weights are nn.Linear layout [out,in] f32 + bias, state-dict order (A.9).
 * ------------------------------------------------------------------------------- */
typedef struct {
    const float *base_w0, *base_b0, *base_w1, *base_b1;                     /* 64x32, 64, 16x64, 16 */
    const float *head_w0, *head_b0, *head_w1, *head_b1, *head_w2, *head_b2; /* 64x27,64,64x64,64,Kx64,K */
    const float *feat_w0, *feat_b0, *feat_w1, *feat_b1, *feat_w2, *feat_b2; /* ..., (K+spec)x64, K+spec */
    const float *dir_w0, *dir_b0;                                           /* 16x28, 16 (specular only) */
    int32_t num_classes; /* K */
    int32_t specular;    /* pred_specular */
    float temperature;
    int32_t enc_tiles;   /* tensor-core entry points only: 1 = `enc` (and `d_enc`) are bf16 TILE buffers, 128 samples x
                          * 32 columns per 8 KB tile in the tensor core's K-major operand order (written by
                          * umhs_hash_encode_fwd_tiles, read by umhs_hash_encode_bwd_tiles); 0 = fp32 rows [n,32] */
} umhs_mlp_params;

typedef struct {
    float *base_w0, *base_b0, *base_w1, *base_b1;
    float *head_w0, *head_b0, *head_w1, *head_b1, *head_w2, *head_b2;
    float *feat_w0, *feat_b0, *feat_w1, *feat_b1, *feat_w2, *feat_b2;
    float *dir_w0, *dir_b0;
} umhs_mlp_grads;

typedef struct {
    float *sigma;       /* [n]            density = trunc_exp(h0) * selector          */
    float *embedding;   /* [n,15] or NULL base_mlp_out                                */
    float *scalar;      /* [n,K]  or NULL sigmoid(mlp_head)                           */
    float *abund;       /* [n,K]  or NULL softmax(logits / T)                         */
    float *s1;          /* [n]    or NULL (specular)                                  */
    float *spec_hidden; /* [n,16] or NULL (specular) ReLU(first layer mlp_directional) */
    void *act_tiles;    /* NULL, or (umhs_field_mlp_tc_fwd only) ceil(n/128) * UMHS_ACT_TILE_BYTES bytes, 16-byte aligned:
                         * the bf16 operand tiles of the hidden activations are kept for
                         * umhs_field_mlp_tc_bwd_saved, which then does not recompute the forward */
} umhs_field_out;
/* per 128-sample tile: H0 [128x64] | X = [pe, emb, 1] [128x32] | head hidden 1, 2 | feature hidden 1, 2 [128x64 each] */
#define UMHS_ACT_TILE_BYTES 90112

typedef struct {
    const float *d_sigma;       /* [n]    */
    const float *d_embedding;   /* [n,15] or NULL */
    const float *d_scalar;      /* [n,K]  */
    const float *d_abund;       /* [n,K]  */
    const float *d_s1;          /* [n]    or NULL */
    const float *d_spec_hidden; /* [n,16] or NULL */
} umhs_field_out_grads;

/* density only: sampler sigma_fn (ref umhs_model.py:229-237), occupancy update (:549-554), density_fn */
int umhs_density_fwd(const umhs_points *pts, int64_t n, const float *enc, const umhs_mlp_params *p, float *sigma,
                     void *stream);
int umhs_field_mlp_fwd(const umhs_points *pts, int64_t n, const float *enc, const umhs_mlp_params *p,
                       const umhs_field_out *out, void *stream);
/* Tensor-core variants (tcgen05 + TMEM, bf16 operands, fp32 accumulation): same interface and outputs as the
 * fp32 entry points above within bf16 tolerance.  These are the plugin's default; the fp32 ones are the
 * "precise" mode used by the tight parity tests. */
int umhs_density_tc_fwd(const umhs_points *pts, int64_t n, const float *enc, const umhs_mlp_params *p, float *sigma,
                        void *stream);
int umhs_field_mlp_tc_fwd(const umhs_points *pts, int64_t n, const float *enc, const umhs_mlp_params *p,
                          const umhs_field_out *out, void *stream);
/* Backward: two 128-sample tiles in flight per SM (two data groups + one MMA-issuing warp), weight-gradient
 * accumulators resident in TMEM for the whole persistent CTA. */
int umhs_field_mlp_tc_bwd(const umhs_points *pts, int64_t n, const float *enc, const umhs_mlp_params *p,
                          const umhs_field_out_grads *g, float *d_enc, const umhs_mlp_grads *dp, void *stream);
/* Backward from the activations the training forward kept (umhs_field_out.act_tiles) and from its outputs: 8 tensor-core
 * rounds per 128-sample tile instead of 16 (nothing is recomputed; the tiles come back by bulk copies).  Same results
 * as umhs_field_mlp_tc_bwd up to the order of the weight-gradient reductions.  Needs p->enc_tiles = 1. */
typedef struct {
    const void *act_tiles;    /* what umhs_field_mlp_tc_fwd wrote: ceil(n/128) * UMHS_ACT_TILE_BYTES */
    const float *sigma;       /* [n]    the forward's outputs */
    const float *scalar;      /* [n,K]  */
    const float *abund;       /* [n,K]  */
    const float *s1;          /* [n]    (specular) */
    const float *spec_hidden; /* [n,16] (specular) */
} umhs_field_saved;
int umhs_field_mlp_tc_bwd_saved(const umhs_points *pts, int64_t n, const float *enc, const umhs_mlp_params *p,
                                const umhs_field_out_grads *g, const umhs_field_saved *saved, void *d_enc_tiles,
                                const umhs_mlp_grads *dp, void *stream);
/* The same backward AND the hash-grid scatter of its d_enc (umhs_hash_encode_bwd_tiles) in one kernel: 11 more warps
 * per CTA read the rows of d_enc_tiles back as the data warps publish them and add them into d_table (accumulating,
 * [L * 2^log2_T, 2] f32) -- the autograd backward of `self.mlp_base(positions_flat)` (ref umhsnerf/umhs_field.py:320)
 * through both of its halves.  Needs p->enc_tiles = 1 and the 16 x 2 grid; d_enc_tiles is still written (scratch). */
int umhs_field_mlp_tc_bwd_scatter(const umhs_points *pts, int64_t n, const float *enc, const umhs_mlp_params *p,
                                  const umhs_field_out_grads *g, void *d_enc_tiles, const umhs_mlp_grads *dp,
                                  const umhs_hash_grid *grid, float *d_table, void *stream);
/* Developer aid: install (or with NULL remove) a device buffer of 4*capacity_slots int64 in which CTA 0 of the
 * tensor-core MLP backward records (tag, clock64) pairs: slot-major, [who = 0 data thread 0 | 1 issuer][tag, clock]. */
int umhs_debug_trace(int64_t *buffer, int32_t capacity_slots, void *stream);
/* One-CTA self-test of the three tcgen05 GEMM forms the MLP kernels are built from (fp32 in/out, bf16 operands):
 * mode 0: D[128xN] = A[128xK] B[NxK]^T; mode 1: D[128xN] = A[128xK] W[KxN]; mode 2: D[64xN] = P[128x64]^T Q[128xN]. */
int umhs_tc05_selftest(int32_t mode, const float *A, const float *B, float *D, int32_t N, int32_t K, void *stream);
/* Developer aid: cycles (device int64 [2]: issue only, issue + execution) of `reps` back-to-back GEMMs of one of the
 * forms above (mode 3: one K = 16 bias step) on zeroed operands, one CTA. */
int umhs_tc05_microbench(int32_t mode, int32_t N, int32_t K, int32_t reps, int64_t *cycles, void *stream);

/* parameter gradients are ACCUMULATED (+=) into *dp; d_enc [n,32] is overwritten (may be NULL).
 * Activations are recomputed from enc; nothing else is kept from the forward pass. */
int umhs_field_mlp_bwd(const umhs_points *pts, int64_t n, const float *enc, const umhs_mlp_params *p,
                       const umhs_field_out_grads *g, float *d_enc, const umhs_mlp_grads *dp, void *stream);

/* ---------------------------------------------------------------------------------
 * (3) Compositing x unmixing x spec->sRGB, per ray over packed samples -- replaces
 *     ref umhsnerf/umhs_model.py:245-313: pack weights (nerfacc render_weight_from_density),
 *     depth / accumulation renderers, SpectralRenderer.forward x4 (umhs_renderer.py:15-30),
 *     the linear mixing model of umhs_field.py:229-251 (spec = sum_k s_k a_k E[k,:],
 *     + s1*sigmoid(W h + b) when specular), ColourSystem.forward (utils/spec_to_rgb.py:112-127)
 *     and ClusterLookup.forward (utils/clusterprobe.py:17-38).
 * ------------------------------------------------------------------------------- */
typedef struct {
    int64_t n_rays, n_samples;
    int32_t num_classes, num_bands, specular, grad_scaling;
    const int64_t *packed_info; /* [R,2] */
    const float *t_starts, *t_ends, *sigma;
    const float *scalar, *abund; /* [N,K] */
    const float *s1;             /* [N]    (specular) */
    const float *spec_hidden;    /* [N,16] (specular) */
    const float *endmembers;     /* [K,B]  */
    const float *dir_w1, *dir_b1; /* [B,16], [B] (specular) */
    const float *rgb_matrix;     /* [B,3]  */
    const float *palette;        /* [>=K,3] class colours of ref umhs_model.py:146-162, or NULL */
    const float *specular_in;    /* [R,B] or NULL: the composited specular term, computed beforehand by
                                  * umhs_specular_tc_fwd (then `specular` = 0 here): spectral2 = the unmixed part,
                                  * spectral = spectral2 + specular_in, and rgb / seg_* follow spectral */
} umhs_composite_in;

typedef struct {
    float *trans;        /* [N]   transmittance per sample, or NULL */
    float *weights;      /* [N]   or NULL */
    float *spectral;     /* [R,B] */
    float *spectral2;    /* [R,B] (specular) unmixed part */
    float *specular;     /* [R,B] (specular) */
    float *abundances;   /* [R,K] */
    float *accumulation; /* [R]   */
    float *depth;        /* [R]   expected depth, unclipped */
    float *t_minmax;     /* [2]   caller-initialised to (+inf, -inf); min/max of (t0+t1)/2 */
    float *rgb;          /* [R,3] */
    float *seg_probs;    /* [R,K] */
    float *seg_raw;      /* [R]   */
    float *seg_pred;     /* [R,3] */
} umhs_composite_out;

typedef struct {
    const float *d_spectral;  /* [R,B] or NULL */
    const float *d_spectral2; /* [R,B] or NULL */
    const float *d_rgb;       /* [R,3] or NULL (gradient w.r.t. the gamma-corrected, clamped rgb) */
    const float *d_accumulation; /* [R] or NULL */
    const float *spectral;    /* [R,B] forward output (needed when d_rgb is given) */
    const float *trans;       /* [N] forward output `trans`, or NULL (the transmittance is then recomputed) */
    const float *d_weights;   /* [N] or NULL: extra dL/dw_n (umhs_specular_tc_bwd), added before the reverse scan */
} umhs_composite_grads_in;

typedef struct {
    float *d_sigma;       /* [N]    */
    float *d_scalar;      /* [N,K]  */
    float *d_abund;       /* [N,K]  */
    float *d_s1;          /* [N]    (specular) */
    float *d_spec_hidden; /* [N,16] (specular) */
    float *d_endmembers;  /* [K,B]  accumulated (+=) */
    float *d_dir_w1;      /* [B,16] accumulated (+=) (specular) */
    float *d_dir_b1;      /* [B]    accumulated (+=) (specular) */
} umhs_composite_grads_out;

int umhs_composite_fwd(const umhs_composite_in *in, const umhs_composite_out *out, void *stream);
int umhs_depth_clip(float *depth, int64_t n_rays, const float *t_minmax, void *stream);
/* Evaluation: the reference renders an image as separate 512-ray forward calls (ref umhsnerf/umhs_model.py:605-616,
 * eval_num_rays_per_chunk of ref umhs_config.py:50), so nerfstudio's DepthRenderer clips every call's depth to that
 * call's own [min, max] sample mid-point.  A fused whole-image pass reproduces it: chunk(r) = (ray_ids ? ray_ids[r] : r)
 * / chunk_rays; bounds [n_chunks,2] caller-initialised to (+inf, -inf); a chunk without samples clips to [1, 1]
 * (nerfstudio's fabricated sample).  depth is clipped in place. */
int umhs_depth_clip_chunks(float *depth, int64_t n_rays, const int64_t *packed_info, const float *t_starts,
                           const float *t_ends, const int64_t *ray_ids, int64_t chunk_rays, float *bounds,
                           void *stream);
int umhs_composite_bwd(const umhs_composite_in *in, const umhs_composite_grads_in *gin,
                       const umhs_composite_grads_out *gout, void *stream);

/* ---------------------------------------------------------------------------------
 * Specular head, last layer, on the tensor cores (csrc/specular_tc.cu) -- replaces, over packed samples,
 *     specular[r,b] = sum_{n in r} w_n s1_n sigmoid(W1[b,:] . h_n + b1[b])
 * i.e. mlp_directional layer 1 + gate of ref umhsnerf/umhs_field.py:240-251 composited as ref
 * umhsnerf/umhs_model.py:277-286 (SpectralRenderer.forward of the per-sample `specular`), and its backward.
 * spec_hidden [N,16] = relu(layer 0) and s1 [N] come from umhs_field_mlp_tc_fwd, weights [N] from
 * umhs_render_weights_fwd; ray_indices [N] sorted by ray; n_dev (optional device count) bounds n.
 * bf16 operands / fp32 accumulation.  fwd: specular [R,B] is overwritten.  bwd: d_spec [R,B] = dL/d(composited
 * spectrum) (umhs_spectral_grad); outputs d_spec_hidden [N,16], d_s1 [N], d_weights [N] (feed
 * umhs_composite_grads_in.d_weights); d_dir_w1 [B,16] / d_dir_b1 [B] accumulated (+=).  grad_scaling:
 * scale_gradients_by_distance_squared (ref umhs_model.py:241-242) from t_starts / t_ends.
 * ------------------------------------------------------------------------------- */
int umhs_specular_tc_fwd(int64_t n, const int64_t *n_dev, int64_t n_rays, int32_t num_bands, const float *spec_hidden,
                         const float *s1, const float *weights, const int64_t *ray_indices, const float *dir_w1,
                         const float *dir_b1, float *specular, void *stream);
int umhs_specular_tc_bwd(int64_t n, const int64_t *n_dev, int64_t n_rays, int32_t num_bands, int32_t grad_scaling,
                         const float *spec_hidden, const float *s1, const float *weights, const float *t_starts,
                         const float *t_ends, const int64_t *ray_indices, const float *dir_w1, const float *dir_b1,
                         const float *d_spec, float *d_spec_hidden, float *d_s1, float *d_weights, float *d_dir_w1,
                         float *d_dir_b1, void *stream);
/* d_spec [R,B] = d_spectral (or 0) + the spec -> sRGB chain of d_rgb (ColourSystem backward, ref
 * utils/spec_to_rgb.py:112-127): the per-ray band phase of the compositing backward as a kernel of its own. */
int umhs_spectral_grad(int64_t n_rays, int32_t num_bands, const float *d_spectral, const float *d_rgb,
                       const float *spectral, const float *rgb_matrix, float *d_spec, void *stream);

/* Un-fused pieces at the reference's module boundaries (UMHSField.get_outputs returns per-sample
 * spectra; SpectralRenderer.forward accumulates arbitrary [N,C] values). */
int umhs_unmix_fwd(int64_t n, int32_t K, int32_t B, int32_t specular, const float *scalar, const float *abund,
                   const float *s1, const float *spec_hidden, const float *endmembers, const float *dir_w1,
                   const float *dir_b1, float *spec /*[n,B] unmixed*/, float *specular_out /*[n,B] s1*spc or NULL*/,
                   void *stream);
int umhs_unmix_bwd(int64_t n, int32_t K, int32_t B, int32_t specular, const float *scalar, const float *abund,
                   const float *s1, const float *spec_hidden, const float *endmembers, const float *dir_w1,
                   const float *dir_b1, const float *d_spec /*[n,B]*/, const float *d_specular /*[n,B] or NULL*/,
                   float *d_scalar, float *d_abund, float *d_s1, float *d_spec_hidden, float *d_endmembers,
                   float *d_dir_w1, float *d_dir_b1, void *stream);
int umhs_render_weights_fwd(const int64_t *packed_info, int64_t n_rays, const float *t_starts, const float *t_ends,
                            const float *sigma, float *weights, float *trans, void *stream);
int umhs_render_weights_bwd(const int64_t *packed_info, int64_t n_rays, const float *t_starts, const float *t_ends,
                            const float *sigma, const float *trans, const float *d_weights, float *d_sigma,
                            void *stream);
int umhs_accumulate_fwd(const float *weights, const float *values /*[N,C] or NULL*/, const int64_t *ray_indices,
                        int64_t n, int32_t C, float *out /*[R,C], zero-filled by caller*/, void *stream);
int umhs_accumulate_bwd(const float *weights, const float *values, const int64_t *ray_indices, int64_t n, int32_t C,
                        const float *d_out, float *d_weights /*or NULL*/, float *d_values /*or NULL*/, void *stream);

/* ---------------------------------------------------------------------------------
 * (4) Ray-sample packing -- replaces VolumetricSampler / nerfacc OccGridEstimator.sampling
 *     reached from ref umhsnerf/umhs_model.py:229-237, and nerfacc.pack_info (:245).
 *     Integer outputs and t values are bit-exact against oracle/pack_oracle.c.
 * ------------------------------------------------------------------------------- */
typedef struct {
    const float *rays_o, *rays_d;    /* [R,3] */
    const float *near_planes, *far_planes; /* [R] */
    const uint8_t *binaries;         /* [levels,res,res,res] bool */
    const float *aabbs;              /* [levels,6] */
    int32_t levels, resolution;
    float step_size, cone_angle;
    int64_t n_rays;
    const int32_t *level_any;        /* [8] device flags from umhs_march_level_any, or NULL (always fetch occupancy) */
    const uint32_t *occ_bits;        /* optional bit-packed copy of binaries from umhs_occ_pack_bits
                                        ([levels, ceil(res^3 / 32)] words, bit b of word w = cell 32 w + b), or NULL:
                                        256 KB per 128^3 level instead of 2 MiB, so the march's lookups hit L1 */
    /* optional staging: umhs_march_count also records the first stage_stride interval starts of every ray in
     * stage[R, stage_stride]; umhs_march_write then copies them (stage_counts = the counts array of that call)
     * instead of marching again.  Identical outputs; NULL / 0 disables. */
    float *stage;
    const int64_t *stage_counts;
    int32_t stage_stride;
} umhs_march_args;

/* level_any[l] = 1 iff grid level l holds an occupied cell (lets the march skip occupancy fetches there) */
int umhs_march_level_any(const umhs_march_args *a, int32_t *level_any /*[8]*/, void *stream);

/* bits + level_any in one pass over binaries (the march reads the bits when a->occ_bits is set) */
int umhs_occ_pack_bits(const umhs_march_args *a, uint32_t *bits, int32_t *level_any /*[8]*/, void *stream);

int umhs_march_count(const umhs_march_args *a, int64_t *counts, void *stream);
/* packed_info[r] = (exclusive_scan(counts)[r], counts[r]); counts [R] int64 in, packed_info [R,2] int64 out.
 * total (device int64[2], may be NULL): total[0] = number of packed samples, total[1] = the unclamped sum.
 * capacity > 0 clamps the packing to `capacity` samples (trailing rays are truncated; total[1] > capacity tells the
 * host, after the fact, that the buffers were too small); capacity <= 0: unlimited. */
int umhs_pack_scan(const int64_t *counts, int64_t n_rays, int64_t *packed_info, int64_t *total, int64_t capacity,
                   void *stream);
/* writes exactly packed_info[r].count samples of ray r at packed_info[r].start */
int umhs_march_write(const umhs_march_args *a, const int64_t *packed_info, int64_t *ray_indices, float *t_starts,
                     float *t_ends, void *stream);
/* visibility: keep = (T >= early_stop_eps) & (alpha >= thre), thre = min(alpha_thre, *occ_mean) when the
 * device scalar occ_mean is given (nerfacc: min(alpha_thre, occs.mean())); also per-ray kept counts */
int umhs_visibility_mask(const int64_t *packed_info, int64_t n_rays, const float *t_starts, const float *t_ends,
                         const float *sigma, float early_stop_eps, float alpha_thre, const float *occ_mean,
                         uint8_t *keep, int64_t *keep_counts, void *stream);
/* src_index (optional, int32 [n_out]): index of each kept sample in the input arrays */
int umhs_compact_samples(const int64_t *packed_info_in, const int64_t *packed_info_out, int64_t n_rays,
                         const uint8_t *keep, const float *t_starts_in, const float *t_ends_in, int64_t *ray_indices,
                         float *t_starts, float *t_ends, int32_t *src_index, void *stream);
/* nerfacc.pack_info for arbitrary (sorted or not) ray_indices; counts is scratch [R] */
int umhs_pack_info(const int64_t *ray_indices, int64_t n, int64_t n_rays, int64_t *counts, int64_t *packed_info,
                   void *stream);

/* Occupancy grid update (ref umhs_model.py:549-554 -> nerfacc update_every_n_steps):
 * occs[idx] = max(occs[idx]*decay, occ); then binaries = occs > min(mean(occs), thre). */
/* cell -> jittered point inside it.  jitter [n,3] in [0,1), or NULL: drawn in the kernel from (seed, index) -- a
 * counter-based generator, identical on every data-parallel rank that shares the seed; seed_dev (optional device
 * uint64) is mixed in, so that a captured CUDA graph draws fresh numbers on every replay.  cell id < 0: slot skipped. */
int umhs_occ_cell_points(const int64_t *cell_ids /*or NULL: 0..n-1*/, const float *jitter /*[n,3] or NULL*/,
                         uint64_t seed, const uint64_t *seed_dev, int64_t n, int32_t resolution,
                         const float *aabb6 /*device, level aabb*/, float *positions, void *stream);
/* Steady-state cell choice of nerfacc's update (OccGridEstimator._sample_uniform_and_occupied_cells, one level):
 * cell_ids [2 n_uniform] = n_uniform uniform cells (floor(u * n_cells)), then the occupied cells: all of them in
 * increasing order when there are at most n_uniform (unused slots = -1), else n_uniform picks with replacement
 * (rank = floor(u * count)).  bits = the level's bit-packed binaries (umhs_occ_pack_bits); u_uniform / u_pick
 * [n_uniform] inject the draws (parity tests) or NULL (drawn from seed); prefix = scratch int32 [n_cells/32 + 2 + ceil(n_cells/32768)];
 * n_valid (device int64, optional) = n_uniform + min(count, n_uniform).  Replaces nonzero() + randint() + a full sort. */
int umhs_occ_select_cells(const uint32_t *bits, int64_t n_cells, int64_t n_uniform, const float *u_uniform,
                          const float *u_pick, uint64_t seed, const uint64_t *seed_dev, int32_t *prefix,
                          int64_t *cell_ids, int64_t *n_valid, void *stream);
int umhs_occ_ema_update(float *occs, const int64_t *cell_ids, int64_t level_offset, const float *sigma, float scale,
                        float decay, int64_t n, float *tmp /*[n] scratch*/, void *stream);
int umhs_occ_binarize(const float *occs, int64_t n_cells, float occ_thre, double *scratch /*[2]*/, uint8_t *binaries,
                      float *mean_out /*device scalar or NULL: mean(occs)*/, void *stream);

/* ---------------------------------------------------------------------------------
 * Training batch from device-resident images -- replaces UMHSDataManager.next_train (ref
 * umhsnerf/data/umhs_datamanager.py:95-108): nerfstudio PixelSampler (floor(rand * [N,H,W]); {key: value[c,y,x]};
 * indices[:,0] = image_idx[c]) + RayGenerator / Cameras.generate_rays (perspective, pixel centres, no distortion),
 * on H x W x B hyperspectral cubes (ref umhsnerf/data/utils/hs_dataloader.py:46-50).  SURVEY 8(f) row 2.
 * The uniform draws u are an input (the caller owns the random stream).  Any output pointer may be NULL.
 * ------------------------------------------------------------------------------- */
#define UMHS_DTYPE_F32 0
#define UMHS_DTYPE_F16 1
#define UMHS_DTYPE_BF16 2
#define UMHS_DTYPE_U8 3
typedef struct {
    const float *u;            /* [R,3] uniform draws in [0,1): image slot, row, column */
    int64_t n_rays;
    int32_t num_images, height, width;
    int32_t num_bands, hs_dtype;      /* UMHS_DTYPE_F32 / F16 / BF16 */
    int32_t image_channels, image_dtype; /* 3 or 4; UMHS_DTYPE_F32 or U8 (u8 -> /255) */
    const int64_t *image_idx;  /* [num_images] dataset (camera) index of each resident image, or NULL = identity */
    const void *hs;            /* [num_images,H,W,B] or NULL */
    const void *image;         /* [num_images,H,W,C] or NULL */
    const float *c2w;          /* [num_cameras,3,4] camera-to-world, or NULL (no rays) */
    const float *intrinsics;   /* [num_cameras,4] fx, fy, cx, cy (16-byte aligned) */
} umhs_batch_in;

typedef struct {
    int64_t *indices;       /* [R,3] (camera, row, column) */
    float *hs_image;        /* [R,B] */
    float *image;           /* [R,C] */
    float *origins;         /* [R,3] */
    float *directions;      /* [R,3] unit */
    float *pixel_area;      /* [R]   */
    float *directions_norm; /* [R]   */
} umhs_batch_out;

int umhs_sample_batch(const umhs_batch_in *in, const umhs_batch_out *out, void *stream);

/* ---------------------------------------------------------------------------------
 * Training losses + their gradients in one pass -- replaces UMHSModel.get_loss_dict (ref
 * umhsnerf/umhs_model.py:329-383: 5 * MSE(spectral, hs_image) + w * MSE(rgb blended over a background, image
 * blended over the same background)) and the autograd backward of those ~15 elementwise / reduction launches.
 * losses [2] = (spectral_loss, rgb_loss); d_* = d(sum of losses)/d(spectral | rgb | accumulation), the inputs
 * of umhs_composite_bwd.  background [R,3] (NULL = black); blend_pred = 1 (background_color "random": the
 * reference draws it with rand_like and adds background * (1 - accumulation) to the prediction) or 0 (named
 * colours: nerfstudio leaves the prediction alone); image [R,3] or [R,4] (RGBA: blended over the background as
 * nerfstudio's blend_background does, in both cases).
 * w_rgb = 0 (method "spectral"): the rgb term and its buffers are skipped.
 * ------------------------------------------------------------------------------- */
int umhs_train_loss(int64_t n_rays, int32_t num_bands, int32_t image_channels, const float *spectral,
                    const float *hs_image, const float *rgb, const float *accumulation, const float *image,
                    const float *background, int32_t blend_pred, float w_spectral, float w_rgb, float *losses,
                    float *d_spectral, float *d_rgb, float *d_accumulation, void *stream);

/* ---------------------------------------------------------------------------------
 * Optimiser: dense Adam over one flat fp32 parameter buffer (torch.optim.Adam
 * semantics, configured at ref umhsnerf/umhs_config.py:59-64), with the endmember
 * clamp of ref umhsnerf/umhs_model.py:568-570 folded in for [clamp_begin, clamp_end).
 * grad is read, multiplied by *inv_scale (device float, may be NULL) and then zeroed.
 * ------------------------------------------------------------------------------- */
int umhs_adam_step(float *param, float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float lr, float beta1,
                   float beta2, float eps, int32_t step, const float *inv_scale, const float *found_inf,
                   int64_t clamp_begin, int64_t clamp_end, int32_t zero_grad, void *stream);
/* Same step with the per-step scalars read from DEVICE memory: hyper = [lr, 1 - beta1^t, sqrt(1 - beta2^t)].
 * Lets a whole training step be replayed as one CUDA graph while lr / bias corrections change every step. */
int umhs_adam_step_dev(float *param, float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float beta1,
                       float beta2, float eps, const float *hyper, const float *inv_scale, const float *found_inf,
                       int64_t clamp_begin, int64_t clamp_end, int32_t zero_grad, void *stream);

/* ---------------------------------------------------------------------------------
 * Data-parallel step: gradient exchange + Adam in ONE kernel over NVLink peer memory.
 * The reference has no multi-GPU path (world_size forced to 1: ref umhsnerf/umhs_pipeline.py:86,108-113); this
 * replaces the all-reduce + optimizer.step() a DDP trainer would run.
 * peer_grads / peer_params: HOST arrays of `world` device pointers, entry w = rank w's flat gradient / parameter
 * buffer mapped into this process (symmetric memory), all of length n (multiple of 4).  Rank r reduces the shard
 * [r n / world, (r+1) n / world) of every gradient (x 1/world: DDP mean), applies Adam with its own exp_avg /
 * exp_avg_sq (only that shard is touched) and stores the new parameters into EVERY rank's buffer.
 * Bracket it with umhs_dp_barrier: channel A before (all gradients complete), channel B after (all parameters
 * landed, all gradients consumed -- only then may a rank zero its gradient buffer or start the next forward).
 * signal_pads: DEVICE array of `world` pointers to each rank's zero-initialised pad of >= 8 x 8 uint32;
 * epoch: this rank's device counter for the channel pair (zero-initialised, one per signal-pad set).
 * mc_grad / mc_param (optional, NULL = off): NVLS multicast addresses of the same gradient / parameter buffers.  The
 * reduce-scatter then is one multimem.ld_reduce per 16 bytes (summed inside the NVSwitch) and the all-gather one
 * multimem.st: 1.0 instead of 1.75 buffer sizes on each GPU's links per direction at 8 GPUs.
 * ------------------------------------------------------------------------------- */
int umhs_dp_barrier(uint32_t *const *signal_pads, int32_t rank, int32_t world, int32_t channel, uint32_t *epoch,
                    void *stream);
int umhs_dp_adam_step(const float *const *peer_grads, float *const *peer_params, float *exp_avg, float *exp_avg_sq,
                      int64_t n, int32_t rank, int32_t world, float beta1, float beta2, float eps, const float *hyper,
                      int64_t clamp_begin, int64_t clamp_end, const float *mc_grad, float *mc_param, void *stream);
/* The same step on the elements [range_begin, range_end) of the flat buffers only (multiples of 4; rank r then owns
 * the r-th of `world` equal parts OF THE RANGE): a trainer can exchange a slice whose gradients are already complete
 * while the backward still produces the rest.  Each slice needs its own pair of barrier channels. */
int umhs_dp_adam_step_range(const float *const *peer_grads, float *const *peer_params, float *exp_avg,
                            float *exp_avg_sq, int64_t n, int64_t range_begin, int64_t range_end, int32_t rank,
                            int32_t world, float beta1, float beta2, float eps, const float *hyper,
                            int64_t clamp_begin, int64_t clamp_end, const float *mc_grad, float *mc_param, void *stream);
/* out[i] += sum_s slabs[s*stride + i] */
int umhs_reduce_slabs(const float *slabs, int64_t n_slabs, int64_t stride, int64_t n, float *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* UMHS_H_ */
