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
constexpr int kStageBytes = 8192 + 512 + 512 + 1024 + 1024;  // h, s1, w, ray ids (int64), t0, t1 of one tile
struct Smem {
    int w, a, dz, v, rid, seg, stage, ctrl, total;
};
__host__ __device__ inline Smem plan(int nch, bool bwd) {
    Smem s;
    int o = 0;
    s.w = o; o += nch * 128 * 32 * 2;                 // [nch*128 bands x 32] K-major: cols 0..15 W1, col 16 b1
    s.a = o; o += TM * 32 * 2;                        // [128 samples x 32]: cols 0..15 h, col 16 = 1
    s.dz = o; o += bwd ? TM * 128 * 2 : 0;            // backward: [128 samples x 128 bands] dz of one band chunk
    s.v = o; o += bwd ? 0 : TM * VSTRIDE * 4;         // forward: V[sample][band] fp32
    s.rid = o; o += TM * 4;                           // ray of each sample of the tile (-1: no sample)
    s.seg = o; o += bwd ? 0 : (TM + TM + 16) * 4;     // forward: segment heads + segment start list
    s.stage = o; o += 2 * kStageBytes;                // two tiles of raw inputs, filled by the copy engine (TMA)
    s.ctrl = o; o += 64;                              // mbarriers `done`, `full[2]`; TMEM address
    s.total = o;
    return s;
}

// Inputs of a tile arrive by 1-D bulk copies (cp.async.bulk, the TMA engine) two tiles ahead: spec_hidden rows
// (128 x 64 B, contiguous), s1, weights, ray indices, and for the backward t0 / t1 -- no thread waits on a global load
// at the top of a tile.  A tile that reaches past the end of the arrays is read with plain loads instead.
struct Stage {
    static constexpr int h = 0, s1 = 8192, w = 8192 + 512, rid = 8192 + 1024, t0 = 8192 + 2048, t1 = 8192 + 2560;
};
struct Loader {
    uint8_t *stage;       // 2 x kStageBytes
    uint64_t *full;       // [2]
    const float *spec_hidden, *s1, *weights, *t0, *t1;
    const int64_t *ray_indices;
    int64_t n_alloc;
    __device__ __forceinline__ bool bulk_ok(int64_t tile) const { return (tile + 1) * TM <= n_alloc; }
    // one thread: arm the barrier and start the copies of `tile` into buffer `buf`
    __device__ __forceinline__ void issue(int64_t tile, int buf) const {
        if (!bulk_ok(tile)) return;
        uint8_t *dst = stage + buf * kStageBytes;
        const int64_t n0 = tile * TM;
        const uint32_t bytes = 8192 + 512 + 512 + 1024 + (t0 ? 1024 : 0);
        fence_proxy_async();  // the buffer's last readers (generic proxy) are ordered before the engine's writes
        mbar_expect_tx(full + buf, bytes);
        bulk_load(dst + Stage::h, spec_hidden + n0 * 16, 8192, full + buf);
        bulk_load(dst + Stage::s1, s1 + n0, 512, full + buf);
        bulk_load(dst + Stage::w, weights + n0, 512, full + buf);
        bulk_load(dst + Stage::rid, ray_indices + n0, 1024, full + buf);
        if (t0) {
            bulk_load(dst + Stage::t0, t0 + n0, 512, full + buf);
            bulk_load(dst + Stage::t1, t1 + n0, 512, full + buf);
        }
    }
};

__device__ __forceinline__ float fast_tanh_half(float x) {  // tanh(x / 2): one MUFU
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
    return t;
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
// the same from a staged tile (bulk-copied): row t of the buffer
__device__ __forceinline__ Sample staged_sample(const uint8_t *buf, int64_t tile, int t, int64_t n_eff, float (&h)[16],
                                                float &tm) {
    Sample s;
    s.n = tile * TM + t;
    s.valid = s.n < n_eff;
    const float4 *hp = reinterpret_cast<const float4 *>(buf + Stage::h) + 4 * t;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 v = s.valid ? hp[q] : make_float4(0.f, 0.f, 0.f, 0.f);
        h[4 * q] = v.x; h[4 * q + 1] = v.y; h[4 * q + 2] = v.z; h[4 * q + 3] = v.w;
    }
    s.rid = s.valid ? (int)reinterpret_cast<const int64_t *>(buf + Stage::rid)[t] : -1;
    s.w = s.valid ? reinterpret_cast<const float *>(buf + Stage::w)[t] : 0.f;
    s.s1 = s.valid ? reinterpret_cast<const float *>(buf + Stage::s1)[t] : 0.f;
    tm = (reinterpret_cast<const float *>(buf + Stage::t0)[t] + reinterpret_cast<const float *>(buf + Stage::t1)[t]) / 2.0f;
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
    uint64_t *full = done + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L.ctrl + 32);
    float *V = reinterpret_cast<float *>(smem + L.v);
    int *rid_s = reinterpret_cast<int *>(smem + L.rid);
    int *head_s = reinterpret_cast<int *>(smem + L.seg);        // [128] 1 where a new ray starts (sample 0 aside)
    int *seg_s = head_s + TM;                                    // [129] segment starts, n_seg at seg_s[TM + 1]
    stage_w(smem + L.w, c.nch * 128, c.w1, c.b1, c.B, tid, kThreads);
    if (tid == 0) {
        mbar_init(done, 1);
        mbar_init(full, 1);
        mbar_init(full + 1, 1);
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
        const Loader ld{smem + L.stage, full, c.spec_hidden, c.s1, c.weights, nullptr, nullptr, c.ray_indices, c.n};
        if (t == 0) {  // two tiles ahead
            ld.issue(blockIdx.x, 0);
            if ((int64_t)blockIdx.x + gridDim.x < n_tiles) ld.issue((int64_t)blockIdx.x + gridDim.x, 1);
        }
        int64_t it = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int buf = (int)(it & 1);
            Sample s;
            float h[16], tm_unused;
            if (ld.bulk_ok(tile)) {
                mbar_wait(full + buf, (uint32_t)((it >> 1) & 1));
                s = staged_sample(smem + L.stage + buf * kStageBytes, tile, t, n_eff, h, tm_unused);
            } else {
                s = load_sample(tile, t, n_eff, c.ray_indices, c.weights, c.s1);
                load_h(c.spec_hidden, s, h);
            }
            const float half_coef = 0.5f * s.w * s.s1;  // w s1 sigmoid(z) = c tanh(z / 2) + c, c = w s1 / 2
            named_bar_sync(kBarData, TM);  // the previous tile's readers of V / rid_s are done; this tile's buffer is read
            if (t == 0 && tile + 2 * (int64_t)gridDim.x < n_tiles) ld.issue(tile + 2 * (int64_t)gridDim.x, buf);
            rid_s[t] = s.rid;
            stage_a_row(smem + L.a, t, h);
            // ray segments of the tile: sample t opens one when its ray differs from sample t-1's (rays are contiguous
            // runs of samples); seg_s[k] = first sample of segment k, seg_s[n_seg] = 128.  Built with one ballot per warp.
            {
                const int prev = __shfl_up_sync(kFull, s.rid, 1);
                const bool head_in_warp = (t & 31) != 0 && prev != s.rid;
                // (the first lane of a warp compares with the previous warp's last sample through shared memory below)
                head_s[t] = head_in_warp ? 1 : 0;
            }
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
                    for (int i = 0; i < 32; ++i) V[t * VSTRIDE + 32 * q + i] = fmaf(half_coef, fast_tanh_half(z[i]), half_coef);
                }
                fence_before_sync();           // TMEM reads done before the next round's MMA overwrites the columns
                named_bar_sync(kBarData, TM);
                if (ch == 0 && warp == 0) {    // segment list of the tile (one warp; the others start summing after it)
                    int n_seg = 0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int n = 32 * k + t;
                        bool head = head_s[n] != 0;
                        if (t == 0) head = (k == 0) || rid_s[n] != rid_s[n - 1];
                        const unsigned m = __ballot_sync(kFull, head);
                        if (head) seg_s[n_seg + __popc(m & ((1u << t) - 1u))] = n;
                        n_seg += __popc(m);
                    }
                    if (t == 0) {
                        seg_s[n_seg] = TM;
                        seg_s[TM + 1] = n_seg;
                    }
                }
                if (ch == 0) named_bar_sync(kBarData, TM);
                // ---- thread = band: one sum per ray segment over the tile's samples (4 independent partial sums)
                if (t < nb) {
                    float *out = specular + 128 * ch + t;
                    const int n_seg = seg_s[TM + 1];
                    for (int k = 0; k < n_seg; ++k) {
                        const int a = seg_s[k], e = seg_s[k + 1];
                        const int r = rid_s[a];
                        if (r < 0) continue;  // rows behind the last sample
                        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                        int n = a;
                        for (; n + 4 <= e; n += 4) {
                            a0 += V[(n + 0) * VSTRIDE + t];
                            a1 += V[(n + 1) * VSTRIDE + t];
                            a2 += V[(n + 2) * VSTRIDE + t];
                            a3 += V[(n + 3) * VSTRIDE + t];
                        }
                        for (; n < e; ++n) a0 += V[n * VSTRIDE + t];
                        atomicAdd(out + (int64_t)r * c.B, (a0 + a1) + (a2 + a3));
                    }
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
    uint64_t *full = done + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L.ctrl + 32);
    stage_w(smem + L.w, c.nch * 128, c.w1, c.b1, c.B, tid, kThreads);
    if (tid == 0) {
        mbar_init(done, 1);
        mbar_init(full, 1);
        mbar_init(full + 1, 1);
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
        const bool use_t = g.grad_scaling != 0;
        const Loader ld{smem + L.stage, full, c.spec_hidden, c.s1, c.weights, use_t ? g.t_starts : nullptr,
                        use_t ? g.t_ends : nullptr, c.ray_indices, c.n};
        if (t == 0) {
            ld.issue(blockIdx.x, 0);
            if ((int64_t)blockIdx.x + gridDim.x < n_tiles) ld.issue((int64_t)blockIdx.x + gridDim.x, 1);
        }
        int64_t it = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int buf = (int)(it & 1);
            Sample s;
            float h[16], tm = 0.f;
            if (ld.bulk_ok(tile)) {
                mbar_wait(full + buf, (uint32_t)((it >> 1) & 1));
                s = staged_sample(smem + L.stage + buf * kStageBytes, tile, t, n_eff, h, tm);
            } else {
                s = load_sample(tile, t, n_eff, c.ray_indices, c.weights, c.s1);
                load_h(c.spec_hidden, s, h);
                if (use_t && s.valid) tm = (__ldg(g.t_starts + s.n) + __ldg(g.t_ends + s.n)) / 2.0f;
            }
            const float gs = (use_t && s.valid) ? fminf(fmaxf(tm * tm, 0.f), 1.f) : 1.f;
            const float coef = s.w * s.s1 * gs;
            const float *grow = g.d_spec + (int64_t)(s.valid ? s.rid : 0) * c.B;
            const bool vec_g = (c.B & 3) == 0 && (reinterpret_cast<uintptr_t>(g.d_spec) & 15u) == 0;
            stage_a_row(smem + L.a, t, h);  // (the previous tile's MMAs on this tile were covered by its last commit)
            named_bar_sync(kBarData, TM);   // every thread has read its row of the staged tile
            if (t == 0 && tile + 2 * (int64_t)gridDim.x < n_tiles) ld.issue(tile + 2 * (int64_t)gridDim.x, buf);
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
                        const int b0 = 128 * r + 32 * qq;
                        float gv[32];
                        if (vec_g && b0 + 32 <= c.B) {  // whole 16-byte pieces of the ray's gradient row, all in flight
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                const float4 v = s.valid ? __ldg(reinterpret_cast<const float4 *>(grow + b0) + k)
                                                         : make_float4(0.f, 0.f, 0.f, 0.f);
                                gv[4 * k] = v.x; gv[4 * k + 1] = v.y; gv[4 * k + 2] = v.z; gv[4 * k + 3] = v.w;
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i) gv[i] = (s.valid && b0 + i < c.B) ? __ldg(grow + b0 + i) : 0.f;
                        }
                        float z[32];
                        tmem_ld32(trow + T_Z + 32u * qq, z);
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float gb = gv[i];
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
