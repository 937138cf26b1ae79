// specular_tc.cu -- the specular head's last layer on the 5th-generation tensor cores, composited per ray.
//
// Replaces, for the packed samples of a ray batch, the per-sample term of ref umhsnerf/umhs_field.py:240-251
//     specular_n[b] = s1_n * sigmoid(W1[b,:] . h_n + b1[b])          (mlp_directional layer 1, 16 -> B, + gate)
// composited along the rays as ref umhsnerf/umhs_model.py:277-286 does through SpectralRenderer.forward
//     specular[r, b] = sum_{n in r} w_n * specular_n[b]
// and its backward.  h = relu(layer 0) comes from the fused MLP kernel (field_mlp_tc.cu), w from the transmittance
// scan.  This [N x 16] . [16 x B] product is the largest GEMM of the model (4096 FLOP per sample at B = 128); it used to
// run as FP32 FMAs inside the warp-per-ray compositing kernel (composite.cu), 2-3x the cost of everything else there.
//
// One CTA = 128 data threads + one MMA-issuing warp, two CTAs per SM (one is in its epilogue while the other's MMAs
// run).  Per 128-sample tile and per chunk of <= 128 bands:
//   forward   z[128 x 128] = [h, 1][128 x 32] . [W1, b1]^T          tcgen05.mma, bf16 operands, fp32 in TMEM
//             thread = sample row: v = sigmoid(z) * w s1 -> shared memory V[sample][band]  (one MUFU per element)
//             thread = band: walks the 128 samples of V, one running sum per ray, red.global.add at ray boundaries
//   backward  z recomputed; thread = sample: g = dL/dspectral[ray(n), b] (rows shared by the lanes of a warp: L1),
//             q_n += g sigmoid, dz = g w s1 gs sigmoid' -> bf16 operand tile
//             d h [128 x 32] += dz . [W1, b1]        (dgrad form, K = bands)
//             d[W1, b1][128 bands x 32] += dz^T . [h, 1]   (wgrad form, M = 128, K = samples; TMEM-resident for the
//             whole kernel, flushed once per CTA) ; d s1 = gs w q ; d w += s1 q (input of the compositing backward)
// gs = clamp(((t0 + t1) / 2)^2, 0, 1) is scale_gradients_by_distance_squared (ref umhs_model.py:241-242).
#include "common.cuh"
#include "tc05.cuh"

namespace umhs {
using namespace tc05;

namespace spec {
constexpr int TM = 128;            // samples per tile == data threads
constexpr int CH_ACT = 2048;       // chunk stride of a 128-row operand tile
constexpr int kThreads = TM + 32;  // + the issuer warp
constexpr int kBarReady = 1, kBarData = 2;
constexpr int VSTRIDE = 129;       // fp32 words per V row: conflict-free by rows (writes) and by columns (reads)

// shared-memory plan (bytes)
struct Smem {
    int w, a, dz, v, rid, ctrl, total;
};
__host__ __device__ inline Smem plan(int nch, bool bwd) {
    Smem s;
    int o = 0;
    s.w = o; o += nch * 128 * 32 * 2;                 // [nch*128 bands x 32] K-major: cols 0..15 W1, col 16 b1
    s.a = o; o += TM * 32 * 2;                        // [128 samples x 32]: cols 0..15 h, col 16 = 1
    s.dz = o; o += bwd ? TM * 128 * 2 : 0;            // backward: [128 samples x 128 bands] dz of one band chunk
    s.v = o; o += bwd ? 0 : TM * VSTRIDE * 4;         // forward: V[sample][band] fp32
    s.rid = o; o += TM * 4;                           // ray of each sample of the tile (-1: no sample)
    s.ctrl = o; o += 64;                              // mbarrier `done`, TMEM address
    s.total = o;
    return s;
}

__device__ __forceinline__ float fast_sigmoid(float x) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
    return fmaf(0.5f, t, 0.5f);
}

// stage [W1 | b1] as the K-major weight tile (rows = bands, zero rows behind B)
__device__ __forceinline__ void stage_w(uint8_t *tile, int rows, const float *__restrict__ w1, const float *__restrict__ b1,
                                        int B, int tid, int nthr) {
    const int ch = rows * 16;
    for (int i = tid; i < rows * 4; i += nthr) {
        const int b = i >> 2, c0 = (i & 3) * 8;
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int c = c0 + q;
            v[q] = (b < B) ? (c < 16 ? __ldg(w1 + b * 16 + c) : (c == 16 ? __ldg(b1 + b) : 0.f)) : 0.f;
        }
        store8(tile, b, c0, ch, v);
    }
}

// this thread's row of the A tile: [h (16), 1, 0 x 15]
__device__ __forceinline__ void stage_a_row(uint8_t *tile, int row, const float (&h)[16]) {
    store8(tile, row, 0, CH_ACT, h);
    store8(tile, row, 8, CH_ACT, h + 8);
    const float one[8] = {1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    store8(tile, row, 16, CH_ACT, one);
    store8(tile, row, 24, CH_ACT, z);
}

struct Sample {
    bool valid;
    int64_t n;
    int rid;
    float w, s1;
};
__device__ __forceinline__ Sample load_sample(int64_t tile, int t, int64_t n_eff, const int64_t *__restrict__ ray_indices,
                                              const float *__restrict__ weights, const float *__restrict__ s1) {
    Sample s;
    s.n = tile * TM + t;
    s.valid = s.n < n_eff;
    s.rid = s.valid ? (int)__ldg(ray_indices + s.n) : -1;
    s.w = s.valid ? __ldg(weights + s.n) : 0.f;
    s.s1 = s.valid ? __ldg(s1 + s.n) : 0.f;
    return s;
}
__device__ __forceinline__ void load_h(const float *__restrict__ spec_hidden, const Sample &s, float (&h)[16]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s.valid) v = __ldg(reinterpret_cast<const float4 *>(spec_hidden + s.n * 16) + q);
        h[4 * q] = v.x; h[4 * q + 1] = v.y; h[4 * q + 2] = v.z; h[4 * q + 3] = v.w;
    }
}

// D[128 bands x N] += P[128 samples x 128 bands]^T Q[128 samples x N]: gemm_wgrad with M = 128
__device__ __forceinline__ void gemm_wgrad128(uint32_t d_tmem, uint32_t b16, uint32_t p_off, uint32_t q_off, int N,
                                              uint32_t accumulate) {
    const uint32_t idesc = make_idesc(128, N, 1, 1);
#pragma unroll
    for (int s = 0; s < 128; s += 16)
        mma_ss(d_tmem, desc_at(b16, p_off + (uint32_t)(s >> 3) * 128, 128, CH_ACT),
               desc_at(b16, q_off + (uint32_t)(s >> 3) * 128, 128, CH_ACT), idesc, (s > 0) ? 1u : accumulate);
}

struct Common {
    int64_t n, n_rays;
    const int64_t *n_dev;
    int B, nch;
    const float *spec_hidden, *s1, *weights;
    const int64_t *ray_indices;
    const float *w1, *b1;
};

__device__ __forceinline__ int64_t eff_n(const Common &c) {
    if (c.n_dev == nullptr) return c.n;
    const int64_t m = __ldg(c.n_dev);
    return m < c.n ? (m < 0 ? 0 : m) : c.n;
}

// ===================================================================================== forward
__global__ void __launch_bounds__(kThreads, 2)
specular_tc_fwd_kernel(Common c, float *__restrict__ specular) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const Smem L = plan(c.nch, false);
    const int tid = threadIdx.x, warp = tid >> 5;
    uint64_t *done = reinterpret_cast<uint64_t *>(smem + L.ctrl);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L.ctrl + 16);
    float *V = reinterpret_cast<float *>(smem + L.v);
    int *rid_s = reinterpret_cast<int *>(smem + L.rid);
    stage_w(smem + L.w, c.nch * 128, c.w1, c.b1, c.B, tid, kThreads);
    if (tid == 0) {
        mbar_init(done, 1);
        fence_mbar_init();
    }
    if (warp == 4) tmem_alloc<128>(tmem_slot);
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tbase = *tmem_slot;
    const int64_t n_eff = eff_n(c);
    const int64_t n_tiles = (n_eff + TM - 1) / TM;
    const uint32_t b16 = desc_base16(smem_u32(smem));

    if (warp == 4) {
        // ------------------------------------------------ issuer: one round per tile and band chunk
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int ch = 0; ch < c.nch; ++ch) {
                named_bar_sync(kBarReady, kThreads);
                fence_after_sync();
                if (elect_one()) {
                    const int nb = c.B - 128 * ch;
                    const int N = nb >= 128 ? 128 : ((nb + 15) & ~15);
                    gemm_kmajor(tbase, b16, (uint32_t)L.a, (uint32_t)L.w, c.nch * 128, 128 * ch, N, 32, 0);
                    commit(done);
                }
                __syncwarp();
            }
        }
    } else {
        const int t = tid;
        const uint32_t trow = tbase + ((uint32_t)(warp * 32) << 16);
        uint32_t parity = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const Sample s = load_sample(tile, t, n_eff, c.ray_indices, c.weights, c.s1);
            float h[16];
            load_h(c.spec_hidden, s, h);
            const float coef = s.w * s.s1;
            named_bar_sync(kBarData, TM);  // the previous tile's readers of V / rid_s are done
            rid_s[t] = s.rid;
            stage_a_row(smem + L.a, t, h);
            for (int ch = 0; ch < c.nch; ++ch) {
                fence_proxy_async();
                fence_before_sync();
                named_bar_arrive(kBarReady, kThreads);
                mbar_wait(done, parity);
                parity ^= 1u;
                fence_after_sync();
                const int nb = min(128, c.B - 128 * ch);
                // ---- thread = sample: sigmoid, weight, to shared memory
#pragma unroll 1
                for (int q = 0; q < 4; ++q) {
                    if (32 * q >= nb) break;
                    float z[32];
                    tmem_ld32(trow + 32u * q, z);
#pragma unroll
                    for (int i = 0; i < 32; ++i) V[t * VSTRIDE + 32 * q + i] = coef * fast_sigmoid(z[i]);
                }
                fence_before_sync();           // TMEM reads done before the next round's MMA overwrites the columns
                named_bar_sync(kBarData, TM);
                // ---- thread = band: one running sum per ray over the tile's samples
                if (t < nb) {
                    float *out = specular + 128 * ch + t;
                    float acc = 0.f;
                    int cur = rid_s[0];
#pragma unroll 4
                    for (int n = 0; n < TM; ++n) {
                        const int r = rid_s[n];
                        if (r != cur) {
                            if (cur >= 0) atomicAdd(out + (int64_t)cur * c.B, acc);
                            acc = 0.f;
                            cur = r;
                        }
                        acc += V[n * VSTRIDE + t];
                    }
                    if (cur >= 0) atomicAdd(out + (int64_t)cur * c.B, acc);
                }
                if (ch + 1 < c.nch) named_bar_sync(kBarData, TM);  // V is rewritten by the next chunk
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 4) tmem_dealloc<128>(tbase);
}

// ===================================================================================== backward
struct BwdArgs {
    const float *t_starts, *t_ends;  // for the gradient scaling
    int grad_scaling;
    const float *d_spec;             // [R, B] dL / d (composited spectrum)
    float *d_spec_hidden, *d_s1, *d_weights;  // [N,16], [N], [N]
    float *d_w1, *d_b1;              // [B,16], [B] accumulated (+=)
};
constexpr int T_Z = 0, T_DH = 128, T_DW = 160;  // TMEM columns: z chunk, d h, d[W1|b1] (32 per band chunk)

template <int NCOLS>
__global__ void __launch_bounds__(kThreads, NCOLS <= 256 ? 2 : 1)
specular_tc_bwd_kernel(Common c, BwdArgs g) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const Smem L = plan(c.nch, true);
    const int tid = threadIdx.x, warp = tid >> 5;
    uint64_t *done = reinterpret_cast<uint64_t *>(smem + L.ctrl);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L.ctrl + 16);
    stage_w(smem + L.w, c.nch * 128, c.w1, c.b1, c.B, tid, kThreads);
    if (tid == 0) {
        mbar_init(done, 1);
        fence_mbar_init();
    }
    if (warp == 4) tmem_alloc<NCOLS>(tmem_slot);
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tbase = *tmem_slot;
    const int64_t n_eff = eff_n(c);
    const int64_t n_tiles = (n_eff + TM - 1) / TM;
    const uint32_t b16 = desc_base16(smem_u32(smem));
    const bool any_tile = (int64_t)blockIdx.x < n_tiles;

    if (warp == 4) {
        // rounds per tile: 0: z(0) | 1..nch-1: grads(ch-1) + z(ch) | nch: grads(nch-1)
        bool first = true;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int r = 0; r <= c.nch; ++r) {
                named_bar_sync(kBarReady, kThreads);
                fence_after_sync();
                if (elect_one()) {
                    if (r > 0) {
                        const int ch = r - 1;
                        const int nb = c.B - 128 * ch;
                        const int Kp = nb >= 128 ? 128 : ((nb + 15) & ~15);
                        gemm_dgrad(tbase + T_DH, b16, (uint32_t)L.dz, (uint32_t)L.w, c.nch * 128, 128 * ch, 32, Kp,
                                   ch > 0 ? 1u : 0u);
                        gemm_wgrad128(tbase + T_DW + 32u * ch, b16, (uint32_t)L.dz, (uint32_t)L.a, 32, first ? 0u : 1u);
                    }
                    if (r < c.nch) {
                        const int nb = c.B - 128 * r;
                        const int N = nb >= 128 ? 128 : ((nb + 15) & ~15);
                        gemm_kmajor(tbase + T_Z, b16, (uint32_t)L.a, (uint32_t)L.w, c.nch * 128, 128 * r, N, 32, 0);
                    }
                    commit(done);
                }
                __syncwarp();
            }
            first = false;
        }
    } else {
        const int t = tid;
        const uint32_t trow = tbase + ((uint32_t)(warp * 32) << 16);
        uint32_t parity = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const Sample s = load_sample(tile, t, n_eff, c.ray_indices, c.weights, c.s1);
            float h[16];
            load_h(c.spec_hidden, s, h);
            float gs = 1.f;
            if (g.grad_scaling && s.valid) {
                const float tm = (__ldg(g.t_starts + s.n) + __ldg(g.t_ends + s.n)) / 2.0f;
                gs = fminf(fmaxf(tm * tm, 0.f), 1.f);
            }
            const float coef = s.w * s.s1 * gs;
            const float *grow = g.d_spec + (int64_t)(s.valid ? s.rid : 0) * c.B;
            stage_a_row(smem + L.a, t, h);  // (the previous tile's MMAs on this tile were covered by its last commit)
            float q = 0.f;
            for (int r = 0; r <= c.nch; ++r) {
                fence_proxy_async();
                fence_before_sync();
                named_bar_arrive(kBarReady, kThreads);
                mbar_wait(done, parity);
                parity ^= 1u;
                fence_after_sync();
                if (r == c.nch) break;
                const int nb = min(128, c.B - 128 * r);
                // ---- thread = sample: dz of this band chunk -> bf16 operand tile (zero behind the last band)
#pragma unroll 1
                for (int qq = 0; qq < 4; ++qq) {
                    float dz[32];
                    if (32 * qq < nb) {
                        float z[32];
                        tmem_ld32(trow + T_Z + 32u * qq, z);
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const int b = 128 * r + 32 * qq + i;
                            const float gb = (s.valid && b < c.B) ? __ldg(grow + b) : 0.f;
                            const float sg = fast_sigmoid(z[i]);
                            q = fmaf(gb, sg, q);
                            dz[i] = gb * coef * sg * (1.0f - sg);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) dz[i] = 0.f;
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) store8(smem + L.dz, t, 32 * qq + 8 * k, CH_ACT, dz + 8 * k);
                }
            }
            // ---- d h (cols 0..15 of the accumulator), d s1, d w
            float dh[16];
            tmem_ld16(trow + T_DH, dh);
            if (s.valid) {
                float4 *dst = reinterpret_cast<float4 *>(g.d_spec_hidden + s.n * 16);
#pragma unroll
                for (int k = 0; k < 4; ++k) dst[k] = make_float4(dh[4 * k], dh[4 * k + 1], dh[4 * k + 2], dh[4 * k + 3]);
                g.d_s1[s.n] = gs * s.w * q;
                g.d_weights[s.n] = s.s1 * q;
            }
        }
    }
    // ---- flush d[W1 | b1]: TMEM lane = band within its chunk
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (warp < 4 && any_tile) {
        const uint32_t trow = tbase + ((uint32_t)(warp * 32) << 16);
        for (int ch = 0; ch < c.nch; ++ch) {
            float a[16], bcol[16];
            tmem_ld16(trow + T_DW + 32u * ch, a);
            tmem_ld16(trow + T_DW + 32u * ch + 16u, bcol);
            const int b = 128 * ch + tid;
            if (b < c.B) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (a[i] != 0.f) atomicAdd(g.d_w1 + b * 16 + i, a[i]);
                if (bcol[0] != 0.f) atomicAdd(g.d_b1 + b, bcol[0]);
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 4) tmem_dealloc<NCOLS>(tbase);
}

// ===================================================================================== per-ray helpers
// dL / d (composited spectrum) = d_spectral + (spec -> sRGB chain of d_rgb): the band phase of the compositing
// backward as a kernel of its own (the tensor-core specular backward needs it before the per-sample sweep runs)
__device__ __forceinline__ float gamma_f(float x) {
    return x < 0.0031308f ? 12.92f * x : 1.055f * powf(fmaxf(x, 1e-6f), 1.0f / 2.4f) - 0.055f;
}
__device__ __forceinline__ float gamma_g(float x) {
    if (x < 0.0031308f) return 12.92f;
    return x > 1e-6f ? (1.055f / 2.4f) * powf(x, 1.0f / 2.4f - 1.0f) : 0.f;
}
__global__ void spectral_grad_kernel(int64_t R, int B, const float *__restrict__ d_spectral,
                                     const float *__restrict__ d_rgb, const float *__restrict__ spectral,
                                     const float *__restrict__ M, float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= R) return;
    float dl[3] = {0.f, 0.f, 0.f};
    if (d_rgb) {
        float lin[3] = {0.f, 0.f, 0.f};
        for (int b = lane; b < B; b += 32) {
            const float sp = spectral[r * B + b];
#pragma unroll
            for (int k = 0; k < 3; ++k) lin[k] += sp * __ldg(M + 3 * b + k);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            lin[k] = warp_sum(lin[k]);
            const float gx = gamma_f(lin[k]);
            const float pass = (gx >= 0.f && gx <= 1.f) ? 1.f : 0.f;
            dl[k] = __ldg(d_rgb + r * 3 + k) * gamma_g(lin[k]) * pass;
        }
    }
    for (int b = lane; b < B; b += 32) {
        float v = d_spectral ? d_spectral[r * B + b] : 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) v += dl[k] * __ldg(M + 3 * b + k);
        out[r * B + b] = v;
    }
}

}  // namespace spec
}  // namespace umhs

using namespace umhs;

static int spec_check(int64_t n, int64_t n_rays, int32_t B, const void *a, const void *b, const void *c, const void *d,
                      const void *e, const void *f) {
    if (n < 0 || n_rays < 0 || B < 1 || B > 512) return -1;
    return (a && b && c && d && e && f) ? 0 : -1;
}

extern "C" int umhs_specular_tc_fwd(int64_t n, const int64_t *n_dev, int64_t n_rays, int32_t num_bands,
                                    const float *spec_hidden, const float *s1, const float *weights,
                                    const int64_t *ray_indices, const float *dir_w1, const float *dir_b1,
                                    float *specular, void *stream) {
    UMHS_CHECK_ARG(spec_check(n, n_rays, num_bands, spec_hidden, s1, weights, ray_indices, dir_w1, dir_b1) == 0 && specular,
                   "bad argument");
    UMHS_CHECK_CUDA(cudaMemsetAsync(specular, 0, sizeof(float) * n_rays * num_bands, as_stream(stream)));
    if (n <= 0) return UMHS_OK;
    spec::Common c{n, n_rays, n_dev, num_bands, (num_bands + 127) / 128, spec_hidden, s1, weights, ray_indices, dir_w1,
                   dir_b1};
    const spec::Smem L = spec::plan(c.nch, false);
    UMHS_CHECK_CUDA(cudaFuncSetAttribute(spec::specular_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    const int64_t tiles = (n + spec::TM - 1) / spec::TM;
    const int grid = (int)(tiles < 2 * kNumSMs ? tiles : 2 * kNumSMs);
    spec::specular_tc_fwd_kernel<<<grid, spec::kThreads, L.total, as_stream(stream)>>>(c, specular);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

extern "C" int umhs_specular_tc_bwd(int64_t n, const int64_t *n_dev, int64_t n_rays, int32_t num_bands,
                                    int32_t grad_scaling, const float *spec_hidden, const float *s1,
                                    const float *weights, const float *t_starts, const float *t_ends,
                                    const int64_t *ray_indices, const float *dir_w1, const float *dir_b1,
                                    const float *d_spec, float *d_spec_hidden, float *d_s1, float *d_weights,
                                    float *d_dir_w1, float *d_dir_b1, void *stream) {
    UMHS_CHECK_ARG(spec_check(n, n_rays, num_bands, spec_hidden, s1, weights, ray_indices, dir_w1, dir_b1) == 0,
                   "bad argument");
    UMHS_CHECK_ARG(d_spec && d_spec_hidden && d_s1 && d_weights && d_dir_w1 && d_dir_b1, "null gradient pointer");
    UMHS_CHECK_ARG(!grad_scaling || (t_starts && t_ends), "gradient scaling needs t_starts / t_ends");
    if (n <= 0) return UMHS_OK;
    spec::Common c{n, n_rays, n_dev, num_bands, (num_bands + 127) / 128, spec_hidden, s1, weights, ray_indices, dir_w1,
                   dir_b1};
    spec::BwdArgs g{t_starts, t_ends, grad_scaling, d_spec, d_spec_hidden, d_s1, d_weights, d_dir_w1, d_dir_b1};
    const spec::Smem L = spec::plan(c.nch, true);
    const int64_t tiles = (n + spec::TM - 1) / spec::TM;
    if (spec::T_DW + 32 * c.nch <= 256) {
        auto k = spec::specular_tc_bwd_kernel<256>;
        UMHS_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
        const int grid = (int)(tiles < 2 * kNumSMs ? tiles : 2 * kNumSMs);
        k<<<grid, spec::kThreads, L.total, as_stream(stream)>>>(c, g);
    } else {
        auto k = spec::specular_tc_bwd_kernel<512>;  // all of TMEM: one CTA per SM
        const int bytes = L.total > 120 * 1024 ? L.total : 120 * 1024;  // (a second CTA could never allocate its TMEM)
        UMHS_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
        k<<<grid, spec::kThreads, bytes, as_stream(stream)>>>(c, g);
    }
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

extern "C" int umhs_spectral_grad(int64_t n_rays, int32_t num_bands, const float *d_spectral, const float *d_rgb,
                                  const float *spectral, const float *rgb_matrix, float *d_spec, void *stream) {
    UMHS_CHECK_ARG(num_bands >= 1 && d_spec && (d_spectral || d_rgb), "bad argument");
    UMHS_CHECK_ARG(!d_rgb || (spectral && rgb_matrix), "the rgb chain needs the forward spectrum and the colour matrix");
    if (n_rays <= 0) return UMHS_OK;
    spec::spectral_grad_kernel<<<div_up(n_rays * 32, 256), 256, 0, as_stream(stream)>>>(n_rays, num_bands, d_spectral,
                                                                                       d_rgb, spectral, rgb_matrix, d_spec);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}
