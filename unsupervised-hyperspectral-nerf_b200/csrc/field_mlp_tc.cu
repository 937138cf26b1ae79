// field_mlp_tc.cu -- the fused field MLPs on the 5th-generation tensor cores (tcgen05 + TMEM), bf16 operands,
// fp32 accumulation.  Same arithmetic and interface as field_mlp.cu (the fp32 SIMT path); this is the default.
//
// One CTA = 128 threads = one tile of 128 samples; thread t owns sample row t (it is also TMEM lane t of every
// [128 x N] accumulator, so the epilogue -- bias, ReLU, softmax, sigmoid, trunc_exp -- is thread-private).
// All layer weights are staged once per CTA in shared memory as bf16 UMMA operand tiles; activations are
// written by their owning threads straight into operand tiles (tc05.cuh layout), one elected thread issues
// tcgen05.mma, completion comes back through an mbarrier (tcgen05.commit), accumulators are read with
// tcgen05.ld.  Layers that share an input run as one MMA (mlp_head/feature_mlp layer 0: N = 128) or are
// issued back to back under one commit.
//
// Forward rounds per tile:   enc->[base0] -> [base1 | dir0] -> [head0+feat0] -> [head1, feat1] -> [head2, feat2]
#include "common.cuh"
#include "tc05.cuh"

namespace umhs {
using namespace tc05;

constexpr int TM = 128;          // samples per tile == threads per CTA
constexpr int CH_ACT = 2048;     // chunk stride of a 128-row activation tile

// ---- shared-memory plan (bytes).  Weight tiles: [rows x K] bf16, CH = rows * 16.
struct TcSmem {
    // weights
    static constexpr int w_base0 = 0;                       // [64 x 32]
    static constexpr int w_base1 = w_base0 + 64 * 32 * 2;   // [16 x 64]
    static constexpr int w_hf0 = w_base1 + 16 * 64 * 2;     // [128 x 32]  rows 0-63 mlp_head.0, 64-127 feature_mlp.0
    static constexpr int w_h1 = w_hf0 + 128 * 32 * 2;       // [64 x 64]
    static constexpr int w_f1 = w_h1 + 64 * 64 * 2;         // [64 x 64]
    static constexpr int w_h2 = w_f1 + 64 * 64 * 2;         // [16 x 64]
    static constexpr int w_f2 = w_h2 + 16 * 64 * 2;         // [32 x 64]
    static constexpr int w_d0 = w_f2 + 32 * 64 * 2;         // [16 x 32]
    static constexpr int w_end = w_d0 + 16 * 32 * 2;
    // biases (fp32): base0 64, base1 16, head0 64, feat0 64, head1 64, feat1 64, head2 16, feat2 32, dir0 16
    static constexpr int b_base0 = w_end;
    static constexpr int b_base1 = b_base0 + 64 * 4;
    static constexpr int b_h0 = b_base1 + 16 * 4;
    static constexpr int b_f0 = b_h0 + 64 * 4;
    static constexpr int b_h1 = b_f0 + 64 * 4;
    static constexpr int b_f1 = b_h1 + 64 * 4;
    static constexpr int b_h2 = b_f1 + 64 * 4;
    static constexpr int b_f2 = b_h2 + 16 * 4;
    static constexpr int b_d0 = b_f2 + 32 * 4;
    static constexpr int b_end = b_d0 + 16 * 4;
    static constexpr int ctrl = (b_end + 127) / 128 * 128;  // mbarrier (8 B) + tmem address (4 B)
    static constexpr int act = ctrl + 128;
    // forward activation tiles
    static constexpr int f_x = act;                 // [128 x 32]  enc, then [pe, emb]
    static constexpr int f_x2 = f_x + TM * 32 * 2;  // [128 x 32]  [sh16, pe12] (specular)
    static constexpr int f_h = f_x2 + TM * 32 * 2;  // [128 x 64]
    static constexpr int f_f = f_h + TM * 64 * 2;   // [128 x 64]
    static constexpr int f_end = f_f + TM * 64 * 2;
};

// stage an nn.Linear weight [out x in] (fp32, global) as a bf16 K-major operand tile with `rows` rows at row0
__device__ __forceinline__ void stage_weight(uint8_t *tile, int rows, int row0, const float *w, int out, int in,
                                             int k_pad, int tid) {
    const int ch = rows * 16;
    for (int i = tid; i < out * (k_pad / 8); i += TM) {
        const int j = i / (k_pad / 8), c0 = (i - j * (k_pad / 8)) * 8;
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = (c0 + q < in) ? __ldg(w + j * in + c0 + q) : 0.f;
        store8(tile, row0 + j, c0, ch, v);
    }
}
__device__ __forceinline__ void zero_rows(uint8_t *tile, int rows, int row_from, int row_to, int k_pad, int tid) {
    const int ch = rows * 16;
    const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int i = tid; i < (row_to - row_from) * (k_pad / 8); i += TM) {
        const int j = i / (k_pad / 8), c0 = (i - j * (k_pad / 8)) * 8;
        store8(tile, row_from + j, c0, ch, z);
    }
}
__device__ __forceinline__ void stage_bias(float *dst, const float *b, int n, int n_pad, int tid) {
    for (int i = tid; i < n_pad; i += TM) dst[i] = (b != nullptr && i < n) ? __ldg(b + i) : 0.f;
}

__device__ __forceinline__ void stage_all_weights(uint8_t *smem, const umhs_mlp_params &mp, bool density_only, int tid) {
    const int K = mp.num_classes, spec = mp.specular;
    stage_weight(smem + TcSmem::w_base0, 64, 0, mp.base_w0, 64, 32, 32, tid);
    stage_weight(smem + TcSmem::w_base1, 16, 0, mp.base_w1, 16, 64, 64, tid);
    stage_bias(reinterpret_cast<float *>(smem + TcSmem::b_base0), mp.base_b0, 64, 64, tid);
    stage_bias(reinterpret_cast<float *>(smem + TcSmem::b_base1), mp.base_b1, 16, 16, tid);
    if (density_only) return;
    stage_weight(smem + TcSmem::w_hf0, 128, 0, mp.head_w0, 64, 27, 32, tid);
    stage_weight(smem + TcSmem::w_hf0, 128, 64, mp.feat_w0, 64, 27, 32, tid);
    stage_weight(smem + TcSmem::w_h1, 64, 0, mp.head_w1, 64, 64, 64, tid);
    stage_weight(smem + TcSmem::w_f1, 64, 0, mp.feat_w1, 64, 64, 64, tid);
    stage_weight(smem + TcSmem::w_h2, 16, 0, mp.head_w2, K, 64, 64, tid);
    zero_rows(smem + TcSmem::w_h2, 16, K, 16, 64, tid);
    stage_weight(smem + TcSmem::w_f2, 32, 0, mp.feat_w2, K + spec, 64, 64, tid);
    zero_rows(smem + TcSmem::w_f2, 32, K + spec, 32, 64, tid);
    stage_bias(reinterpret_cast<float *>(smem + TcSmem::b_h0), mp.head_b0, 64, 64, tid);
    stage_bias(reinterpret_cast<float *>(smem + TcSmem::b_f0), mp.feat_b0, 64, 64, tid);
    stage_bias(reinterpret_cast<float *>(smem + TcSmem::b_h1), mp.head_b1, 64, 64, tid);
    stage_bias(reinterpret_cast<float *>(smem + TcSmem::b_f1), mp.feat_b1, 64, 64, tid);
    stage_bias(reinterpret_cast<float *>(smem + TcSmem::b_h2), mp.head_b2, K, 16, tid);
    stage_bias(reinterpret_cast<float *>(smem + TcSmem::b_f2), mp.feat_b2, K + spec, 32, tid);
    if (spec) {
        stage_weight(smem + TcSmem::w_d0, 16, 0, mp.dir_w0, 16, 28, 32, tid);
        stage_bias(reinterpret_cast<float *>(smem + TcSmem::b_d0), mp.dir_b0, 16, 16, tid);
    }
}

__device__ __forceinline__ void pos_encoding_tc(const Point &p, float (&pe)[12]) {
    const float two_pi = 6.283185307179586f, half_pi = 1.5707963267948966f;
    const float s[3] = {two_pi * p.wx, two_pi * p.wy, two_pi * p.wz};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float a = s[d], b = s[d] * 2.0f;
        // the encoding is rounded to bf16 right after: the fast sine (abs. error ~1e-6 for |x| < 48) is exact enough
        if (fabsf(b) < 48.0f) {
            pe[2 * d] = __sinf(a);
            pe[2 * d + 1] = __sinf(b);
            pe[6 + 2 * d] = __sinf(a + half_pi);
            pe[6 + 2 * d + 1] = __sinf(b + half_pi);
        } else {
            pe[2 * d] = sinf(a);
            pe[2 * d + 1] = sinf(b);
            pe[6 + 2 * d] = sinf(a + half_pi);
            pe[6 + 2 * d + 1] = sinf(b + half_pi);
        }
    }
}

__device__ __forceinline__ void sh_deg3_tc(float x, float y, float z, float (&c)[16]) {
    const float xx = x * x, yy = y * y, zz = z * z;
    c[0] = 0.28209479177387814f;
    c[1] = 0.4886025119029199f * y;
    c[2] = 0.4886025119029199f * z;
    c[3] = 0.4886025119029199f * x;
    c[4] = 1.0925484305920792f * x * y;
    c[5] = 1.0925484305920792f * y * z;
    c[6] = 0.9461746957575601f * zz - 0.31539156525251999f;
    c[7] = 1.0925484305920792f * x * z;
    c[8] = 0.5462742152960396f * (xx - yy);
    c[9] = 0.5900435899266435f * y * (3.0f * xx - yy);
    c[10] = 2.890611442640554f * x * y * z;
    c[11] = 0.4570457994644658f * y * (5.0f * zz - 1.0f);
    c[12] = 0.3731763325901154f * z * (5.0f * zz - 3.0f);
    c[13] = 0.4570457994644658f * x * (5.0f * zz - 1.0f);
    c[14] = 1.445305721320277f * z * (xx - yy);
    c[15] = 0.5900435899266435f * x * (xx - 3.0f * yy);
}

// epilogue helper: 64 accumulator columns [col0, col0+64) of this thread's row -> + bias, ReLU -> bf16 operand tile
__device__ __forceinline__ void acc_to_tile_relu(uint32_t trow, int col0, int ncols, const float *bias, uint8_t *tile,
                                                 int row, int tile_col0) {
    (void)ncols;  // always 64
    float v[64];
    tmem_ld64(trow + (uint32_t)col0, v);
#pragma unroll
    for (int c = 0; c < 64; c += 8) {
        const float4 b0 = *reinterpret_cast<const float4 *>(bias + c), b1 = *reinterpret_cast<const float4 *>(bias + c + 4);
        float o[8] = {fmaxf(v[c] + b0.x, 0.f),     fmaxf(v[c + 1] + b0.y, 0.f), fmaxf(v[c + 2] + b0.z, 0.f),
                      fmaxf(v[c + 3] + b0.w, 0.f), fmaxf(v[c + 4] + b1.x, 0.f), fmaxf(v[c + 5] + b1.y, 0.f),
                      fmaxf(v[c + 6] + b1.z, 0.f), fmaxf(v[c + 7] + b1.w, 0.f)};
        store8(tile, row, tile_col0 + c, CH_ACT, o);
    }
}

// Optional cycle trace (developer aid, off unless umhs_debug_trace() installs a buffer): CTA 0 stamps clock64()
// at the stages of every round so that the latency chain of a tile can be read back.
__device__ long long *g_trace = nullptr;
__device__ int g_trace_cap = 0;
__device__ __forceinline__ void trace(int &slot, int tag) {
    if (g_trace != nullptr && blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == 96) && slot < g_trace_cap) {
        g_trace[2 * (slot * 2 + (threadIdx.x ? 1 : 0))] = tag;
        g_trace[2 * (slot * 2 + (threadIdx.x ? 1 : 0)) + 1] = clock64();
        ++slot;
    }
}

struct RoundSync {
    uint64_t *bar;
    uint32_t parity;
    int slot = 0;
    // all threads: operand tiles written & accumulators consumed -> one thread may issue MMAs
    __device__ __forceinline__ void before_issue() {
        trace(slot, 1);  // epilogue of the previous round finished
        fence_proxy_async();
        fence_before_sync();
        __syncthreads();
        trace(slot, 2);  // CTA barrier passed
    }
    __device__ __forceinline__ void wait_done() {
        trace(slot, 3);  // (thread 0: MMAs issued + commit) waiting for the tensor core
        mbar_wait(bar, parity);
        parity ^= 1u;
        fence_after_sync();
        trace(slot, 4);  // accumulators ready
    }
};

template <bool DENSITY_ONLY>
__global__ void __launch_bounds__(TM)
field_mlp_tc_fwd_kernel(umhs_points P, int64_t n_total, const float *__restrict__ enc, umhs_mlp_params mp,
                        umhs_field_out out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int K = mp.num_classes, spec = DENSITY_ONLY ? 0 : mp.specular;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + TcSmem::ctrl);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + TcSmem::ctrl + 16);
    constexpr int NCOLS = 128;

    stage_all_weights(smem, mp, DENSITY_ONLY, tid);
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<NCOLS>(tmem_slot);
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tbase = *tmem_slot;
    const uint32_t trow = tbase + ((uint32_t)(warp * 32) << 16);  // this warp's lane quarter
    RoundSync rs{bar, 0u};

    uint8_t *tX = smem + TcSmem::f_x, *tX2 = smem + TcSmem::f_x2, *tH = smem + TcSmem::f_h, *tF = smem + TcSmem::f_f;
    const uint32_t sX = smem_u32(tX), sX2 = smem_u32(tX2), sH = smem_u32(tH), sF = smem_u32(tF);
    const uint32_t sWb0 = smem_u32(smem + TcSmem::w_base0), sWb1 = smem_u32(smem + TcSmem::w_base1);
    const uint32_t sWhf0 = smem_u32(smem + TcSmem::w_hf0), sWh1 = smem_u32(smem + TcSmem::w_h1);
    const uint32_t sWf1 = smem_u32(smem + TcSmem::w_f1), sWh2 = smem_u32(smem + TcSmem::w_h2);
    const uint32_t sWf2 = smem_u32(smem + TcSmem::w_f2), sWd0 = smem_u32(smem + TcSmem::w_d0);
    const float *Bb0 = reinterpret_cast<const float *>(smem + TcSmem::b_base0);
    const float *Bb1 = reinterpret_cast<const float *>(smem + TcSmem::b_base1);
    const float *Bh0 = reinterpret_cast<const float *>(smem + TcSmem::b_h0);
    const float *Bf0 = reinterpret_cast<const float *>(smem + TcSmem::b_f0);
    const float *Bh1 = reinterpret_cast<const float *>(smem + TcSmem::b_h1);
    const float *Bf1 = reinterpret_cast<const float *>(smem + TcSmem::b_f1);
    const float *Bh2 = reinterpret_cast<const float *>(smem + TcSmem::b_h2);
    const float *Bf2 = reinterpret_cast<const float *>(smem + TcSmem::b_f2);
    const float *Bd0 = reinterpret_cast<const float *>(smem + TcSmem::b_d0);
    const float invT = DENSITY_ONLY ? 1.f : 1.0f / mp.temperature;

    n_total = effective_n(P, n_total);
    const int64_t n_tiles = (n_total + TM - 1) / TM;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t n0 = tile * TM;
        const bool valid = n0 + tid < n_total;
        const int64_t n = valid ? n0 + tid : n_total - 1;  // tail threads shadow the last sample, write nothing
        const Point pt = eval_point(P, n);
        float pe[12];
        // ---------------- round 0: enc -> tX ; (specular) [sh, pe] -> tX2 ; MMA base0
        {
            const float4 *e4 = reinterpret_cast<const float4 *>(enc + enc_row(P, n) * 32);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float4 a = __ldg(e4 + 2 * c), b = __ldg(e4 + 2 * c + 1);
                const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                store8(tX, tid, 8 * c, CH_ACT, v);
            }
            if (!DENSITY_ONLY) {
                pos_encoding_tc(pt, pe);
                if (spec) {
                    float x2[32];
                    float sh[16];
                    sh_deg3_tc((pt.dx + 1.0f) / 2.0f, (pt.dy + 1.0f) / 2.0f, (pt.dz + 1.0f) / 2.0f, sh);
#pragma unroll
                    for (int i = 0; i < 16; ++i) x2[i] = sh[i];
#pragma unroll
                    for (int i = 0; i < 12; ++i) x2[16 + i] = pe[i];
#pragma unroll
                    for (int i = 28; i < 32; ++i) x2[i] = 0.f;
#pragma unroll
                    for (int c = 0; c < 4; ++c) store8(tX2, tid, 8 * c, CH_ACT, x2 + 8 * c);
                }
            }
        }
        rs.before_issue();
        if (tid == 0) {
            fence_after_sync();
            gemm_kmajor(tbase + 0, sX, sWb0, 64, 0, 64, 32, 0);
            commit(bar);
        }
    if (warp == 0) __syncwarp();  // lanes 1-31 must not spin on the mbarrier while lane 0 is still issuing
        rs.wait_done();
        // ---------------- round 1: H0 = relu(. + b) -> tH ; MMA base1 (N=16) [+ dir0 (N=16) at cols 64..79]
        acc_to_tile_relu(trow, 0, 64, Bb0, tH, tid, 0);
        rs.before_issue();
        if (tid == 0) {
            fence_after_sync();
            gemm_kmajor(tbase + 0, sH, sWb1, 16, 0, 16, 64, 0);
            if (spec) gemm_kmajor(tbase + 64, sX2, sWd0, 16, 0, 16, 32, 0);
            commit(bar);
        }
        if (warp == 0) __syncwarp();  // lanes 1-31 must not spin on the mbarrier while lane 0 is still issuing
        rs.wait_done();
        // ---------------- round 2: sigma, emb ; X = [pe, emb] -> tX ; MMA head0|feat0 (N=128)
        float o16[16];
        tmem_ld16(trow + 0, o16);
#pragma unroll
        for (int i = 0; i < 16; ++i) o16[i] += Bb1[i];
        if (valid) {
            out.sigma[n] = expf(o16[0]) * pt.sel;  // trunc_exp forward == exp; density * selector
            if (out.embedding) {
#pragma unroll
                for (int i = 0; i < 15; ++i) out.embedding[n * 15 + i] = o16[i + 1];
            }
        }
        if (DENSITY_ONLY) {
            fence_before_sync();  // accumulator reads done before the next tile's MMA overwrites them
            continue;
        }
        if (spec) {
            float h[16];
            tmem_ld16(trow + 64, h);
            if (valid && out.spec_hidden) {
                float4 *dst = reinterpret_cast<float4 *>(out.spec_hidden + n * 16);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    dst[q] = make_float4(fmaxf(h[4 * q] + Bd0[4 * q], 0.f), fmaxf(h[4 * q + 1] + Bd0[4 * q + 1], 0.f),
                                         fmaxf(h[4 * q + 2] + Bd0[4 * q + 2], 0.f),
                                         fmaxf(h[4 * q + 3] + Bd0[4 * q + 3], 0.f));
            }
        }
        {
            float x[32];
#pragma unroll
            for (int i = 0; i < 12; ++i) x[i] = pe[i];
#pragma unroll
            for (int i = 0; i < 15; ++i) x[12 + i] = o16[1 + i];
#pragma unroll
            for (int i = 27; i < 32; ++i) x[i] = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) store8(tX, tid, 8 * c, CH_ACT, x + 8 * c);
        }
        rs.before_issue();
        if (tid == 0) {
            fence_after_sync();
            gemm_kmajor(tbase + 0, sX, sWhf0, 128, 0, 128, 32, 0);
            commit(bar);
        }
        if (warp == 0) __syncwarp();  // lanes 1-31 must not spin on the mbarrier while lane 0 is still issuing
        rs.wait_done();
        // ---------------- round 3: hidden 0 of both MLPs -> tH, tF ; MMA head1, feat1
        acc_to_tile_relu(trow, 0, 64, Bh0, tH, tid, 0);
        acc_to_tile_relu(trow, 64, 64, Bf0, tF, tid, 0);
        rs.before_issue();
        if (tid == 0) {
            fence_after_sync();
            gemm_kmajor(tbase + 0, sH, sWh1, 64, 0, 64, 64, 0);
            gemm_kmajor(tbase + 64, sF, sWf1, 64, 0, 64, 64, 0);
            commit(bar);
        }
        if (warp == 0) __syncwarp();  // lanes 1-31 must not spin on the mbarrier while lane 0 is still issuing
        rs.wait_done();
        // ---------------- round 4: hidden 1 -> tH, tF ; MMA head2 (N=16), feat2 (N=32)
        acc_to_tile_relu(trow, 0, 64, Bh1, tH, tid, 0);
        acc_to_tile_relu(trow, 64, 64, Bf1, tF, tid, 0);
        rs.before_issue();
        if (tid == 0) {
            fence_after_sync();
            gemm_kmajor(tbase + 0, sH, sWh2, 16, 0, 16, 64, 0);
            gemm_kmajor(tbase + 64, sF, sWf2, 32, 0, 32, 64, 0);
            commit(bar);
        }
        if (warp == 0) __syncwarp();  // lanes 1-31 must not spin on the mbarrier while lane 0 is still issuing
        rs.wait_done();
        // ---------------- round 5: outputs
        {
            float zh[16], zf[16], zf2[16];
            tmem_ld16(trow + 0, zh);
            tmem_ld16(trow + 64, zf);
            if (K + spec > 16) tmem_ld16(trow + 80, zf2);
            if (valid) {
                float mx = -3.0e38f;
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    zf[k] = (zf[k] + Bf2[k]) * invT;
                    if (k < K) mx = fmaxf(mx, zf[k]);
                }
                float den = 0.f, e[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    e[k] = (k < K) ? __expf(zf[k] - mx) : 0.f;
                    den += e[k];
                }
                const float rden = __frcp_rn(den);
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    if (k < K) {
                        out.scalar[n * K + k] = __frcp_rn(1.0f + __expf(-(zh[k] + Bh2[k])));
                        out.abund[n * K + k] = e[k] * rden;
                    }
                }
                if (spec && out.s1) {
                    float z = zf2[0] + Bf2[16];
#pragma unroll
                    for (int k = 0; k < 16; ++k)
                        if (k == K) z = zf[k] * mp.temperature;
                    out.s1[n] = __frcp_rn(1.0f + __expf(-z));
                }
            }
        }
        fence_before_sync();
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc<NCOLS>(tbase);
}

// ------------------------------------------------------------------ descriptor self-test (one CTA)
// mode 0: D[128xN] = A[128xK] * B[NxK]^T ; mode 1: D[128xN] = A[128xK] * W[KxN] ; mode 2: D[64xN] = P[128x64]^T Q[128xN]
__global__ void __launch_bounds__(TM)
tc05_selftest_kernel(int mode, const float *__restrict__ A, const float *__restrict__ B, float *__restrict__ D, int N,
                     int K) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *tA = smem, *tB = smem + 32768;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 65536);
    uint32_t *slot = reinterpret_cast<uint32_t *>(smem + 65536 + 16);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<128>(slot);
    const int acols = (mode == 2) ? 64 : K;
    for (int c = 0; c < acols; c += 8) {
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = A[tid * acols + c + q];
        store8(tA, tid, c, CH_ACT, v);
    }
    if (mode == 0) {  // B [N x K] -> tile with N rows
        for (int i = tid; i < N * (K / 8); i += TM) {
            const int j = i / (K / 8), c0 = (i % (K / 8)) * 8;
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = B[j * K + c0 + q];
            store8(tB, j, c0, N * 16, v);
        }
    } else if (mode == 1) {  // W [K(out) x N(in)] -> forward-style tile with K rows, N columns
        for (int i = tid; i < K * (N / 8); i += TM) {
            const int j = i / (N / 8), c0 = (i % (N / 8)) * 8;
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = B[j * N + c0 + q];
            store8(tB, j, c0, K * 16, v);
        }
    } else {  // Q [128 x N] activation tile
        for (int c = 0; c < N; c += 8) {
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = B[tid * N + c + q];
            store8(tB, tid, c, CH_ACT, v);
        }
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tbase = *slot;
    if (tid == 0) {
        if (mode == 0) gemm_kmajor(tbase, smem_u32(tA), smem_u32(tB), N, 0, N, K, 0);
        else if (mode == 1) gemm_dgrad(tbase, smem_u32(tA), smem_u32(tB), K, 0, N, K, 0);
        else gemm_wgrad(tbase, smem_u32(tA), 0, smem_u32(tB), 0, N, 0);
        commit(bar);
    }
    if (warp == 0) __syncwarp();  // lanes 1-31 must not spin on the mbarrier while lane 0 is still issuing
    mbar_wait(bar, 0);
    fence_after_sync();
    const uint32_t trow = tbase + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < N; c += 16) {
        float v[16];
        tmem_ld16(trow + (uint32_t)c, v);
        if (mode == 2) {
            if (lane < 16)
                for (int i = 0; i < 16; ++i) D[(warp * 16 + lane) * N + c + i] = v[i];
        } else {
            for (int i = 0; i < 16; ++i) D[tid * N + c + i] = v[i];
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<128>(tbase);
}

__global__ void set_trace_kernel(long long *buf, int cap) {
    g_trace = buf;
    g_trace_cap = cap;
}

static int tc_grid(int64_t n, int ctas_per_sm) {
    const int64_t want = (n + TM - 1) / TM;
    const int64_t cap = (int64_t)kNumSMs * ctas_per_sm;
    return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

}  // namespace umhs

using namespace umhs;

extern "C" int umhs_debug_trace(int64_t *buffer, int32_t capacity_slots, void *stream) {
    set_trace_kernel<<<1, 1, 0, as_stream(stream)>>>(reinterpret_cast<long long *>(buffer), buffer ? capacity_slots : 0);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

extern "C" int umhs_tc05_selftest(int32_t mode, const float *A, const float *B, float *D, int32_t N, int32_t K,
                                  void *stream) {
    UMHS_CHECK_ARG(A && B && D, "null argument");
    UMHS_CHECK_ARG(mode >= 0 && mode <= 2 && N >= 16 && N <= 128 && N % 16 == 0 && K >= 16 && K <= 128 && K % 16 == 0,
                   "bad shape");
    const int bytes = 65536 + 128;
    UMHS_CHECK_CUDA(cudaFuncSetAttribute(tc05_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    tc05_selftest_kernel<<<1, TM, bytes, as_stream(stream)>>>(mode, A, B, D, N, K);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

extern "C" int umhs_density_tc_fwd(const umhs_points *pts, int64_t n, const float *enc, const umhs_mlp_params *p,
                                   float *sigma, void *stream) {
    if (n <= 0) return UMHS_OK;
    UMHS_CHECK_ARG(pts && enc && sigma && p && p->base_w0 && p->base_b0 && p->base_w1 && p->base_b1, "null argument");
    umhs_field_out out = {};
    out.sigma = sigma;
    const int bytes = TcSmem::f_x2;  // weights + ctrl + tX ... tH is needed too
    const int need = TcSmem::f_f;    // up to and including tH
    (void)bytes;
    UMHS_CHECK_CUDA(cudaFuncSetAttribute(field_mlp_tc_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         need));
    field_mlp_tc_fwd_kernel<true><<<tc_grid(n, 3), TM, need, as_stream(stream)>>>(*pts, n, enc, *p, out);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

extern "C" int umhs_field_mlp_tc_fwd(const umhs_points *pts, int64_t n, const float *enc, const umhs_mlp_params *p,
                                     const umhs_field_out *out, void *stream) {
    if (n <= 0) return UMHS_OK;
    UMHS_CHECK_ARG(pts && enc && out && out->sigma && out->scalar && out->abund && p, "null argument");
    UMHS_CHECK_ARG(p->num_classes >= 1 && p->num_classes <= UMHS_MAX_CLASSES && p->temperature > 0.f, "bad K / T");
    UMHS_CHECK_ARG(p->base_w0 && p->head_w0 && p->head_w1 && p->head_w2 && p->feat_w0 && p->feat_w1 && p->feat_w2,
                   "null weights");
    UMHS_CHECK_ARG(!p->specular || (p->dir_w0 && p->dir_b0), "specular weights missing");
    UMHS_CHECK_CUDA(cudaFuncSetAttribute(field_mlp_tc_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         TcSmem::f_end));
    field_mlp_tc_fwd_kernel<false><<<tc_grid(n, 2), TM, TcSmem::f_end, as_stream(stream)>>>(*pts, n, enc, *p, *out);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

// =====================================================================================================
// Backward on the tensor cores.
//
// Per tile of 128 samples the activations are recomputed from the hash encoding (bf16 operand tiles kept in
// shared memory), then for every layer one commit group issues
//     wgrad   P[64 x N] += Delta^T[64 x 128 samples] * Act[128 samples x N]     (MN-major operands, M = 64)
//     dgrad   W0[128 x N] = Delta[128 x K] * W[K x N]                            (weights read MN-major)
// The weight-gradient accumulators P_* live in TMEM for the whole kernel (accumulated over all tiles of the
// persistent CTA) and are flushed once at the end with one atomic per element per CTA.  Bias gradients of the
// 64-wide layers are one more N = 8 MMA against a tile of ones; the narrow output layers use warp shuffles.
// Rounds per tile: base0 | base1(+dir0 wgrad) | head0+feat0 | head1 | head2 | dL2 | dL1 | dL0 | feat1 | feat2 |
//                  dL2 | dL1 | dL0 | dbase1 | dbase0 | d_enc.
// =====================================================================================================
namespace umhs {

struct TcBwd {
    // narrow-layer bias gradients: per-thread running column sums (80 rows x 128 threads fp32), reduced once at the end
    static constexpr int sacc = TcSmem::ctrl + 128;          // rows: h2[16] f2[32] b1[16] d0[16]
    static constexpr int act = sacc + 80 * TM * 4;
    // activation tiles; a 2 KB chunk of ones directly behind tA0, tX and tH1 extends those B operands by 8 columns of
    // ones, so the bias gradient of a 64-wide layer is column 64 (32 for the 32-wide inputs) of its weight-gradient MMA
    static constexpr int tA0 = act;                        // [128 x 32] + ones chunk
    static constexpr int tX = tA0 + TM * 32 * 2 + 2048;    // [128 x 32] + ones chunk
    static constexpr int tH0 = tX + TM * 32 * 2 + 2048;    // [128 x 64]
    static constexpr int tX2 = tH0 + TM * 64 * 2;          // [128 x 64]
    static constexpr int tH1 = tX2 + TM * 64 * 2;          // [128 x 64] + ones chunk
    static constexpr int tH2 = tH1 + TM * 64 * 2 + 2048;
    static constexpr int tD1 = tH2 + TM * 64 * 2;
    static constexpr int tD2 = tD1 + TM * 64 * 2;
    static constexpr int end = tD2 + TM * 64 * 2;
    // TMEM columns
    static constexpr int W0 = 0, W1 = 64;
    static constexpr int P_h2t = 128, P_h1 = 144, P_h0 = 216, P_f2t = 256, P_f1 = 288, P_f0 = 360;
    static constexpr int P_b1t = 400, P_b0 = 416, P_d0t = 456;  // .. 472
    static constexpr int ncols = 512;
};

// accumulator columns [col0, col0+64) of this row, masked by relu'(act tile row) -> bf16 delta tile
__device__ __forceinline__ void acc_to_delta_masked(uint32_t trow, int col0, const uint8_t *act_tile, uint8_t *delta_tile,
                                                    int row, bool valid) {
    float v[64];
    tmem_ld64(trow + (uint32_t)col0, v);
#pragma unroll
    for (int c = 0; c < 64; c += 8) {
        float h[8], o[8];
        load8(act_tile, row, c, CH_ACT, h);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = (valid && h[i] > 0.f) ? v[c + i] : 0.f;
        store8(delta_tile, row, c, CH_ACT, o);
    }
}

template <int NMAX>
__device__ __forceinline__ void warp_colsum_to_smem(const float (&v)[NMAX], int n, float *dst, int lane) {
#pragma unroll
    for (int k = 0; k < NMAX; ++k) {
        if (k < n) {  // warp-uniform
            const float s = warp_sum(v[k]);
            if (lane == 0 && s != 0.f) atomicAdd(dst + k, s);
        }
    }
}

// flush one persistent [64 x ncols] accumulator: this thread's row r = 16*warp + lane (lane < 16)
template <class F>
__device__ __forceinline__ void flush_acc(uint32_t trow, int col0, int ncols, int lane, F &&sink) {
    for (int c = 0; c < ncols; c += 16) {
        float v[16];
        tmem_ld16(trow + (uint32_t)(col0 + c), v);  // .sync.aligned: every lane participates
        if (lane < 16) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (c + i < ncols) sink(c + i, v[i]);
        }
    }
}

__global__ void __launch_bounds__(TM)
field_mlp_tc_bwd_kernel(umhs_points P, int64_t n_total, const float *__restrict__ enc, umhs_mlp_params mp,
                        umhs_field_out_grads g, float *__restrict__ d_enc, umhs_mlp_grads dp) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = mp.num_classes, spec = mp.specular;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + TcSmem::ctrl);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + TcSmem::ctrl + 16);

    stage_all_weights(smem, mp, false, tid);
    {
        const float one[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
        store8(smem + TcBwd::tA0, tid, 32, CH_ACT, one);  // chunk 4 of the 32-wide tiles, chunk 8 of tH1
        store8(smem + TcBwd::tX, tid, 32, CH_ACT, one);
        store8(smem + TcBwd::tH1, tid, 64, CH_ACT, one);
        float *sa = reinterpret_cast<float *>(smem + TcBwd::sacc);
        for (int i = tid; i < 80 * TM; i += TM) sa[i] = 0.f;
        // zero the padding columns 32..63 of the specular input tile once (columns 0..31 are rewritten per tile)
        const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 32; c < 64; c += 8) store8(smem + TcBwd::tX2, tid, c, CH_ACT, z);
    }
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<TcBwd::ncols>(tmem_slot);
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tbase = *tmem_slot;
    const uint32_t trow = tbase + ((uint32_t)(warp * 32) << 16);
    RoundSync rs{bar, 0u};

    uint8_t *tA0 = smem + TcBwd::tA0, *tX = smem + TcBwd::tX, *tH0 = smem + TcBwd::tH0, *tX2 = smem + TcBwd::tX2;
    uint8_t *tH1 = smem + TcBwd::tH1, *tH2 = smem + TcBwd::tH2, *tD1 = smem + TcBwd::tD1, *tD2 = smem + TcBwd::tD2;
    const uint32_t sA0 = smem_u32(tA0), sX = smem_u32(tX), sH0 = smem_u32(tH0), sX2 = smem_u32(tX2);
    const uint32_t sH1 = smem_u32(tH1), sH2 = smem_u32(tH2), sD1 = smem_u32(tD1), sD2 = smem_u32(tD2);
    const uint32_t sWb0 = smem_u32(smem + TcSmem::w_base0), sWb1 = smem_u32(smem + TcSmem::w_base1);
    const uint32_t sWhf0 = smem_u32(smem + TcSmem::w_hf0), sWh1 = smem_u32(smem + TcSmem::w_h1);
    const uint32_t sWf1 = smem_u32(smem + TcSmem::w_f1), sWh2 = smem_u32(smem + TcSmem::w_h2);
    const uint32_t sWf2 = smem_u32(smem + TcSmem::w_f2), sWd0 = smem_u32(smem + TcSmem::w_d0);
    const float *Bb0 = reinterpret_cast<const float *>(smem + TcSmem::b_base0);
    const float *Bb1 = reinterpret_cast<const float *>(smem + TcSmem::b_base1);
    const float *Bh0 = reinterpret_cast<const float *>(smem + TcSmem::b_h0);
    const float *Bf0 = reinterpret_cast<const float *>(smem + TcSmem::b_f0);
    const float *Bh1 = reinterpret_cast<const float *>(smem + TcSmem::b_h1);
    const float *Bf1 = reinterpret_cast<const float *>(smem + TcSmem::b_f1);
    const float *Bh2 = reinterpret_cast<const float *>(smem + TcSmem::b_h2);
    const float *Bf2 = reinterpret_cast<const float *>(smem + TcSmem::b_f2);
    const float *Bd0 = reinterpret_cast<const float *>(smem + TcSmem::b_d0);
    // this thread's column of the bias accumulators: row k of a block lives at [k * TM + tid]
    float *sb_h2 = reinterpret_cast<float *>(smem + TcBwd::sacc) + tid, *sb_f2 = sb_h2 + 16 * TM,
          *sb_b1 = sb_f2 + 32 * TM, *sb_d0 = sb_b1 + 16 * TM;
    const float invT = 1.0f / mp.temperature;

    n_total = effective_n(P, n_total);
    const int64_t n_tiles = (n_total + TM - 1) / TM;
    uint32_t acc = 0;  // 0 on this CTA's first tile: the persistent accumulators are initialised, not accumulated
    // the next tile's hash encoding and sample geometry are fetched while the tensor core finishes the current tile
    float4 pre[8];
    Point pt_next;
    auto prefetch = [&](int64_t tile_idx) {
        if (tile_idx >= n_tiles) return;
        const int64_t m = tile_idx * TM + tid;
        const int64_t nn = m < n_total ? m : n_total - 1;
        const float4 *e4 = reinterpret_cast<const float4 *>(enc + enc_row(P, nn) * 32);
#pragma unroll
        for (int c = 0; c < 8; ++c) pre[c] = __ldg(e4 + c);
        pt_next = eval_point(P, nn);
    };
    prefetch(blockIdx.x);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, acc = 1) {
        const int64_t n0 = tile * TM;
        const bool valid = n0 + tid < n_total;
        const int64_t n = valid ? n0 + tid : n_total - 1;
        const Point pt = pt_next;
        float pe[12];
        pos_encoding_tc(pt, pe);
        // ---------------- T0: enc -> tA0 ; (specular) [sh, pe] -> tX2 ; base0 [, dir0]
        {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float4 a = pre[2 * c], b = pre[2 * c + 1];
                const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                store8(tA0, tid, 8 * c, CH_ACT, v);
            }
            if (spec) {
                float x2[32], sh[16];
                sh_deg3_tc((pt.dx + 1.0f) / 2.0f, (pt.dy + 1.0f) / 2.0f, (pt.dz + 1.0f) / 2.0f, sh);
#pragma unroll
                for (int i = 0; i < 16; ++i) x2[i] = sh[i];
#pragma unroll
                for (int i = 0; i < 12; ++i) x2[16 + i] = pe[i];
#pragma unroll
                for (int i = 28; i < 32; ++i) x2[i] = 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c) store8(tX2, tid, 8 * c, CH_ACT, x2 + 8 * c);
            }
        }
        rs.before_issue();
        if (tid == 0) {
            fence_after_sync();
            gemm_kmajor(tbase + TcBwd::W0, sA0, sWb0, 64, 0, 64, 32, 0);
            if (spec) gemm_kmajor(tbase + TcBwd::W1, sX2, sWd0, 16, 0, 16, 32, 0);
            commit(bar);
        }
    if (warp == 0) __syncwarp();  // lanes 1-31 must not spin on the mbarrier while lane 0 is still issuing
        rs.wait_done();
        // ---------------- T1: H0 -> tH0 ; (specular) d spec_hidden -> tD2[:, :16] ; base1 [, dir0 wgrad]
        acc_to_tile_relu(trow, TcBwd::W0, 64, Bb0, tH0, tid, 0);
        if (spec) {
            float h[16], dsh[16];
            tmem_ld16(trow + TcBwd::W1, h);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float4 gq = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid && g.d_spec_hidden) gq = __ldg(reinterpret_cast<const float4 *>(g.d_spec_hidden + n * 16) + q);
                dsh[4 * q] = (h[4 * q] + Bd0[4 * q] > 0.f) ? gq.x : 0.f;
                dsh[4 * q + 1] = (h[4 * q + 1] + Bd0[4 * q + 1] > 0.f) ? gq.y : 0.f;
                dsh[4 * q + 2] = (h[4 * q + 2] + Bd0[4 * q + 2] > 0.f) ? gq.z : 0.f;
                dsh[4 * q + 3] = (h[4 * q + 3] + Bd0[4 * q + 3] > 0.f) ? gq.w : 0.f;
            }
            store8(tD2, tid, 0, CH_ACT, dsh);
            store8(tD2, tid, 8, CH_ACT, dsh + 8);
#pragma unroll
            for (int k = 0; k < 16; ++k) sb_d0[k * TM] += dsh[k];
        }
        rs.before_issue();
        if (tid == 0) {
            fence_after_sync();
            gemm_kmajor(tbase + TcBwd::W0, sH0, sWb1, 16, 0, 16, 64, 0);
            if (spec) gemm_wgrad(tbase + TcBwd::P_d0t, sX2, 0, sD2, 0, 16, acc);
            commit(bar);
        }
        if (warp == 0) __syncwarp();  // lanes 1-31 must not spin on the mbarrier while lane 0 is still issuing
        rs.wait_done();
        // ---------------- T2: o16 ; X = [pe, emb] -> tX ; head0 -> W0, feat0 -> W1
        float o16[16], demb[15];
        tmem_ld16(trow + TcBwd::W0, o16);
#pragma unroll
        for (int i = 0; i < 16; ++i) o16[i] += Bb1[i];
#pragma unroll
        for (int i = 0; i < 15; ++i) demb[i] = (valid && g.d_embedding) ? __ldg(g.d_embedding + n * 15 + i) : 0.f;
        {
            float x[32];
#pragma unroll
            for (int i = 0; i < 12; ++i) x[i] = pe[i];
#pragma unroll
            for (int i = 0; i < 15; ++i) x[12 + i] = o16[1 + i];
#pragma unroll
            for (int i = 27; i < 32; ++i) x[i] = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) store8(tX, tid, 8 * c, CH_ACT, x + 8 * c);
        }
        rs.before_issue();
        if (tid == 0) {
            fence_after_sync();
            gemm_kmajor(tbase + TcBwd::W0, sX, sWhf0, 128, 0, 64, 32, 0);
            gemm_kmajor(tbase + TcBwd::W1, sX, sWhf0, 128, 64, 64, 32, 0);
            commit(bar);
        }
        if (warp == 0) __syncwarp();  // lanes 1-31 must not spin on the mbarrier while lane 0 is still issuing
        rs.wait_done();

        // ================= the two 27->64->64->nout MLPs: m = 0 mlp_head (sigmoid), m = 1 feature_mlp (softmax, s1)
#pragma unroll 1
        for (int m = 0; m < 2; ++m) {
            const float *B0 = m ? Bf0 : Bh0, *B1 = m ? Bf1 : Bh1, *B2 = m ? Bf2 : Bh2;
            const uint32_t sW1 = m ? sWf1 : sWh1, sW2 = m ? sWf2 : sWh2;
            const int w2_rows = m ? 32 : 16;   // rows of the last-layer weight tile == padded nout
            const int nout = m ? K + spec : K;
            // ---- hidden 0 -> tH1 (head: from W0; feat: from W1, computed in T2) ; [feat: also collect head's dX]
            if (m == 1) {
                float dx[16];
                tmem_ld16(trow + TcBwd::W0, dx);       // columns 0..15 of d X (head)
#pragma unroll
                for (int i = 12; i < 16; ++i) demb[i - 12] += dx[i];
                tmem_ld16(trow + TcBwd::W0 + 16, dx);  // columns 16..31
#pragma unroll
                for (int i = 0; i < 11; ++i) demb[4 + i] += dx[i];
            }
            acc_to_tile_relu(trow, m ? TcBwd::W1 : TcBwd::W0, 64, B0, tH1, tid, 0);
            rs.before_issue();
            if (tid == 0) {
                fence_after_sync();
                gemm_kmajor(tbase + TcBwd::W0, sH1, sW1, 64, 0, 64, 64, 0);
                commit(bar);
            }
            if (warp == 0) __syncwarp();  // lanes 1-31 must not spin on the mbarrier while lane 0 is still issuing
            rs.wait_done();
            // ---- hidden 1 -> tH2 ; last layer
            acc_to_tile_relu(trow, TcBwd::W0, 64, B1, tH2, tid, 0);
            rs.before_issue();
            if (tid == 0) {
                fence_after_sync();
                gemm_kmajor(tbase + TcBwd::W0, sH2, sW2, w2_rows, 0, w2_rows, 64, 0);
                commit(bar);
            }
            if (warp == 0) __syncwarp();  // lanes 1-31 must not spin on the mbarrier while lane 0 is still issuing
            // incoming gradients of this MLP's outputs: fetched while the tensor core runs the last layer
            float gin[17];
#pragma unroll
            for (int k = 0; k < 16; ++k)
                gin[k] = (valid && k < K) ? __ldg((m ? g.d_abund : g.d_scalar) + n * K + k) : 0.f;
            gin[16] = (valid && m == 1 && spec && g.d_s1) ? __ldg(g.d_s1 + n) : 0.f;
            rs.wait_done();
            // ---- dz -> tD2[:, :w2_rows] ; wgrad L2 (transposed: rows = in-features) + dgrad L2
            {
                float z[32], dz[32];
                {
                    float t16[16];
                    tmem_ld16(trow + TcBwd::W0, t16);
#pragma unroll
                    for (int k = 0; k < 16; ++k) z[k] = t16[k];
                    if (m == 1) tmem_ld16(trow + TcBwd::W0 + 16, t16);
#pragma unroll
                    for (int k = 0; k < 16; ++k) z[16 + k] = (m == 1) ? t16[k] : 0.f;
                }
#pragma unroll
                for (int k = 0; k < 32; ++k) dz[k] = 0.f;
                if (valid) {
                    // tensor-core path: fast exp / reciprocal (operands are bf16 anyway); one exp per logit
                    if (m == 0) {
#pragma unroll
                        for (int k = 0; k < 16; ++k) {
                            if (k < K) {
                                const float s = __frcp_rn(1.0f + __expf(-(z[k] + B2[k])));
                                dz[k] = gin[k] * s * (1.0f - s);
                            }
                        }
                    } else {
                        float mx = -3.0e38f, e[16], den = 0.f, dot = 0.f, da[16];
#pragma unroll
                        for (int k = 0; k < 16; ++k) {
                            z[k] = (z[k] + B2[k]) * invT;
                            if (k < K) mx = fmaxf(mx, z[k]);
                        }
#pragma unroll
                        for (int k = 0; k < 16; ++k) {
                            e[k] = (k < K) ? __expf(z[k] - mx) : 0.f;
                            da[k] = gin[k];
                            den += e[k];
                            dot += e[k] * da[k];
                        }
                        const float rden = __frcp_rn(den);
                        dot *= rden;
#pragma unroll
                        for (int k = 0; k < 16; ++k) dz[k] = e[k] * rden * invT * (da[k] - dot);
                        if (spec) {
                            // logit K of feature_mlp is the specular gate s1 (raw value: undo the /T applied above
                            // only for columns < 16; column 16 lives in z[16] untouched)
                            float zs = z[16] + B2[16];
#pragma unroll
                            for (int k = 0; k < 16; ++k)
                                if (k == K) zs = z[k] * mp.temperature;
                            const float s = __frcp_rn(1.0f + __expf(-zs));
                            const float ds = gin[16] * s * (1.0f - s);
#pragma unroll
                            for (int k = 0; k < 17; ++k)
                                if (k == K) dz[k] = ds;
                        }
                    }
                }
                store8(tD2, tid, 0, CH_ACT, dz);
                store8(tD2, tid, 8, CH_ACT, dz + 8);
                if (m == 1) {
                    store8(tD2, tid, 16, CH_ACT, dz + 16);
                    store8(tD2, tid, 24, CH_ACT, dz + 24);
                }
                {
                    float *sb = m ? sb_f2 : sb_h2;
#pragma unroll
                    for (int k = 0; k < 17; ++k)
                        if (k < nout) sb[k * TM] += dz[k];
                }
            }
            rs.before_issue();
            if (tid == 0) {
                fence_after_sync();
                gemm_wgrad(tbase + (m ? TcBwd::P_f2t : TcBwd::P_h2t), sH2, 0, sD2, 0, w2_rows, acc);
                gemm_dgrad(tbase + TcBwd::W0, sD2, sW2, w2_rows, 0, 64, w2_rows, 0);
                commit(bar);
            }
            if (warp == 0) __syncwarp();  // lanes 1-31 must not spin on the mbarrier while lane 0 is still issuing
            rs.wait_done();
            // ---- d hidden1 -> tD1 ; wgrad L1 (+bias) + dgrad L1
            acc_to_delta_masked(trow, TcBwd::W0, tH2, tD1, tid, valid);
            rs.before_issue();
            if (tid == 0) {
                fence_after_sync();
                gemm_wgrad(tbase + (m ? TcBwd::P_f1 : TcBwd::P_h1), sD1, 0, sH1, 0, 72, acc);  // 64 + ones
                gemm_dgrad(tbase + TcBwd::W0, sD1, sW1, 64, 0, 64, 64, 0);
                commit(bar);
            }
            if (warp == 0) __syncwarp();  // lanes 1-31 must not spin on the mbarrier while lane 0 is still issuing
            rs.wait_done();
            // ---- d hidden0 -> tD2 ; wgrad L0 (+bias) + dgrad L0 (N = 32: d [pe, emb])
            acc_to_delta_masked(trow, TcBwd::W0, tH1, tD2, tid, valid);
            rs.before_issue();
            if (tid == 0) {
                fence_after_sync();
                gemm_wgrad(tbase + (m ? TcBwd::P_f0 : TcBwd::P_h0), sD2, 0, sX, 0, 40, acc);   // 32 + ones
                gemm_dgrad(tbase + TcBwd::W0, sD2, sWhf0, 128, m ? 64 : 0, 32, 64, 0);
                commit(bar);
            }
            if (warp == 0) __syncwarp();  // lanes 1-31 must not spin on the mbarrier while lane 0 is still issuing
            rs.wait_done();
        }
        // ================= base MLP backward
        {
            float dx[16], dO[16];
            tmem_ld16(trow + TcBwd::W0, dx);  // d X (feature_mlp), columns 0..15
#pragma unroll
            for (int i = 12; i < 16; ++i) demb[i - 12] += dx[i];
            tmem_ld16(trow + TcBwd::W0 + 16, dx);
#pragma unroll
            for (int i = 0; i < 11; ++i) demb[4 + i] += dx[i];
            // density = trunc_exp(o0) * sel  ->  d o0 = d_sigma * sel * exp(clamp(o0, -15, 15))
            const float dsg = (valid && g.d_sigma) ? __ldg(g.d_sigma + n) : 0.f;
            dO[0] = dsg * pt.sel * expf(fminf(fmaxf(o16[0], -15.0f), 15.0f));
#pragma unroll
            for (int i = 0; i < 15; ++i) dO[1 + i] = valid ? demb[i] : 0.f;
            store8(tD2, tid, 0, CH_ACT, dO);
            store8(tD2, tid, 8, CH_ACT, dO + 8);
#pragma unroll
            for (int k = 0; k < 16; ++k) sb_b1[k * TM] += dO[k];
        }
        rs.before_issue();
        if (tid == 0) {
            fence_after_sync();
            gemm_wgrad(tbase + TcBwd::P_b1t, sH0, 0, sD2, 0, 16, acc);
            gemm_dgrad(tbase + TcBwd::W0, sD2, sWb1, 16, 0, 64, 16, 0);
            commit(bar);
        }
        if (warp == 0) __syncwarp();  // lanes 1-31 must not spin on the mbarrier while lane 0 is still issuing
        rs.wait_done();
        acc_to_delta_masked(trow, TcBwd::W0, tH0, tD1, tid, valid);
        rs.before_issue();
        if (tid == 0) {
            fence_after_sync();
            gemm_wgrad(tbase + TcBwd::P_b0, sD1, 0, sA0, 0, 40, acc);  // 32 + ones
            if (d_enc) gemm_dgrad(tbase + TcBwd::W0, sD1, sWb0, 64, 0, 32, 64, 0);
            commit(bar);
        }
        if (warp == 0) __syncwarp();  // lanes 1-31 must not spin on the mbarrier while lane 0 is still issuing
        prefetch(tile + gridDim.x);
        rs.wait_done();
        if (d_enc) {
            float v[16];
            float4 *dst = reinterpret_cast<float4 *>(d_enc + n * 32);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                tmem_ld16(trow + TcBwd::W0 + 16 * c, v);
                if (valid) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) dst[4 * c + q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                }
            }
        }
        fence_before_sync();
    }

    // ================= flush the persistent weight-gradient accumulators (row r = 16*warp + lane, lane < 16)
    __syncthreads();
    fence_after_sync();
    if (acc == 0) {  // this CTA processed no tile (device-side count smaller than the grid): nothing to flush
        if (warp == 0) tmem_dealloc<TcBwd::ncols>(tbase);
        return;
    }
    const int r = warp * 16 + lane;
    flush_acc(trow, TcBwd::P_h2t, K, lane, [&](int c, float v) { if (v != 0.f) atomicAdd(dp.head_w2 + c * 64 + r, v); });
    flush_acc(trow, TcBwd::P_h1, 65, lane, [&](int c, float v) {
        if (v != 0.f) atomicAdd(c < 64 ? dp.head_w1 + r * 64 + c : dp.head_b1 + r, v); });
    flush_acc(trow, TcBwd::P_h0, 33, lane, [&](int c, float v) {
        if (v != 0.f && (c < 27 || c == 32)) atomicAdd(c < 27 ? dp.head_w0 + r * 27 + c : dp.head_b0 + r, v); });
    flush_acc(trow, TcBwd::P_f2t, K + spec, lane, [&](int c, float v) { if (v != 0.f) atomicAdd(dp.feat_w2 + c * 64 + r, v); });
    flush_acc(trow, TcBwd::P_f1, 65, lane, [&](int c, float v) {
        if (v != 0.f) atomicAdd(c < 64 ? dp.feat_w1 + r * 64 + c : dp.feat_b1 + r, v); });
    flush_acc(trow, TcBwd::P_f0, 33, lane, [&](int c, float v) {
        if (v != 0.f && (c < 27 || c == 32)) atomicAdd(c < 27 ? dp.feat_w0 + r * 27 + c : dp.feat_b0 + r, v); });
    flush_acc(trow, TcBwd::P_b1t, 16, lane, [&](int c, float v) { if (v != 0.f) atomicAdd(dp.base_w1 + c * 64 + r, v); });
    flush_acc(trow, TcBwd::P_b0, 33, lane, [&](int c, float v) {
        if (v != 0.f) atomicAdd(c < 32 ? dp.base_w0 + r * 32 + c : dp.base_b0 + r, v); });
    if (spec)
        flush_acc(trow, TcBwd::P_d0t, 16, lane, [&](int c, float v) { if (r < 28 && v != 0.f) atomicAdd(dp.dir_w0 + c * 28 + r, v); });
    {   // narrow-layer biases: sum this CTA's per-thread accumulators (row k = 128 consecutive floats)
        const float *sa = reinterpret_cast<const float *>(smem + TcBwd::sacc);
        if (tid < 80) {
            float a = 0.f;
            for (int i = 0; i < TM; ++i) a += sa[tid * TM + ((i + tid) & (TM - 1))];
            float *dst = nullptr;
            if (tid < 16) dst = (tid < K) ? dp.head_b2 + tid : nullptr;
            else if (tid < 48) dst = (tid - 16 < K + spec) ? dp.feat_b2 + (tid - 16) : nullptr;
            else if (tid < 64) dst = dp.base_b1 + (tid - 48);
            else dst = spec ? dp.dir_b0 + (tid - 64) : nullptr;
            if (dst != nullptr && a != 0.f) atomicAdd(dst, a);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<TcBwd::ncols>(tbase);
}

}  // namespace umhs

extern "C" int umhs_field_mlp_tc_bwd(const umhs_points *pts, int64_t n, const float *enc, const umhs_mlp_params *p,
                                     const umhs_field_out_grads *g, float *d_enc, const umhs_mlp_grads *dp,
                                     void *stream) {
    if (n <= 0) return UMHS_OK;
    UMHS_CHECK_ARG(pts && enc && g && dp && p, "null argument");
    UMHS_CHECK_ARG(p->num_classes >= 1 && p->num_classes <= UMHS_MAX_CLASSES && p->temperature > 0.f, "bad K / T");
    UMHS_CHECK_ARG(g->d_scalar && g->d_abund, "d_scalar/d_abund are required");
    UMHS_CHECK_ARG(dp->base_w0 && dp->base_b0 && dp->base_w1 && dp->base_b1 && dp->head_w0 && dp->head_b0 &&
                       dp->head_w1 && dp->head_b1 && dp->head_w2 && dp->head_b2 && dp->feat_w0 && dp->feat_b0 &&
                       dp->feat_w1 && dp->feat_b1 && dp->feat_w2 && dp->feat_b2,
                   "null parameter-gradient pointer");
    UMHS_CHECK_ARG(!p->specular || (p->dir_w0 && p->dir_b0 && dp->dir_w0 && dp->dir_b0), "specular pointers missing");
    UMHS_CHECK_CUDA(cudaFuncSetAttribute(umhs::field_mlp_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         umhs::TcBwd::end));
    // one persistent CTA per SM (TMEM: 512 columns hold the weight-gradient accumulators); never more CTAs than tiles
    const int64_t n_tiles = (n + umhs::TM - 1) / umhs::TM;
    const int grid = (int)(n_tiles < umhs::kNumSMs ? n_tiles : umhs::kNumSMs);
    umhs::field_mlp_tc_bwd_kernel<<<grid, umhs::TM, umhs::TcBwd::end, umhs::as_stream(stream)>>>(*pts, n, enc, *p, *g,
                                                                                                d_enc, *dp);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

// =====================================================================================================
// Backward, two tiles in flight per SM ("ping-pong").
//
// The single-tile kernel above is bound by its chain of ~16 dependent rounds (epilogue -> CTA barrier -> MMA issue
// -> tensor core -> mbarrier), with 4 warps on the SM and the tensor pipe idle most of the time
// (profiles/r1_ncu_full_summary.md).  Here one CTA of 9 warps runs TWO 128-sample tiles at once:
//   warps 0-3 = group 0, warps 4-7 = group 1 (thread t of a group owns row t & 127 = TMEM lane), warp 8 = MMA issuer.
// A group finishes an epilogue, arrives on its `ready` mbarrier and waits on its `done` mbarrier; the issuer waits
// for `ready`, issues that round's tcgen05.mma program for that group and commits it to `done`.  While one group
// waits for the tensor core the other one runs its epilogue.  All MMAs come from one thread, so the two groups can
// accumulate into the SAME TMEM-resident weight-gradient accumulators, in program order.
// Per group: 64 working TMEM columns; shared: 328 (+16 specular) columns of accumulators.  The base MLP's hidden
// activations stay in registers (packed bf16) between the first and the last rounds of a tile instead of a tile.
// =====================================================================================================
namespace umhs {

struct TcBwd2 {
    static constexpr int sbias = TcSmem::ctrl + 64;       // fp32 [80]: h2[16] f2[32] b1[16] d0[16]  (ctrl holds 5 words)
    static constexpr int grp = (sbias + 80 * 4 + 127) / 128 * 128;
    // per-group activation tiles (a ones chunk behind tA0, tX and tH1: see TcBwd)
    static constexpr int tA0 = 0;
    static constexpr int tX = tA0 + TM * 32 * 2 + 2048;
    static constexpr int tH1 = tX + TM * 32 * 2 + 2048;
    static constexpr int tH2 = tH1 + TM * 64 * 2 + 2048;
    static constexpr int tD1 = tH2 + TM * 64 * 2;
    static constexpr int tD2 = tD1 + TM * 64 * 2;
    static constexpr int tX2 = tD2 + TM * 64 * 2;          // [128 x 32] (specular only)
    static constexpr int grp_bytes = tX2;                  // without specular
    static constexpr int grp_bytes_spec = tX2 + TM * 32 * 2;
    // TMEM columns: working [0,64) group 0, [64,128) group 1; accumulators shared by both groups
    static constexpr int P_h2t = 128, P_h1 = 144, P_h0 = 216, P_f2t = 256, P_f1 = 288, P_f0 = 360;
    static constexpr int P_b1t = 400, P_b0 = 416, P_d0 = 456;  // P_d0: [64 (16 real rows) x 32] .. 488
    static constexpr int ncols = 512;
    static constexpr int kRounds = 16;
    static constexpr int kThreads = 288;
};

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct GroupAddr {  // shared-memory addresses (async-proxy view) of one group's operand tiles
    uint32_t A0, X, H1, H2, D1, D2, X2;
};

// MMA program of round `r` for one group (executed by the single issuing thread)
__device__ __forceinline__ void issue_round(int r, const GroupAddr &g, uint32_t tw, uint32_t tbase, uint32_t acc,
                                            int spec, bool want_denc, uint32_t sWb0, uint32_t sWb1, uint32_t sWhf0,
                                            uint32_t sWh1, uint32_t sWf1, uint32_t sWh2, uint32_t sWf2, uint32_t sWd0) {
    switch (r) {
        case 0:  // base layer 0
            gemm_kmajor(tw, g.A0, sWb0, 64, 0, 64, 32, 0);
            break;
        case 1:  // base layer 1 (N=16) [+ mlp_directional layer 0 (N=16) at working columns 16..31]
            gemm_kmajor(tw, g.H1, sWb1, 16, 0, 16, 64, 0);
            if (spec) gemm_kmajor(tw + 16, g.X2, sWd0, 16, 0, 16, 32, 0);
            break;
        case 2:  // mlp_head layer 0 [+ wgrad of mlp_directional layer 0: d_hidden^T (padded to 64 rows) x X2]
            gemm_kmajor(tw, g.X, sWhf0, 128, 0, 64, 32, 0);
            if (spec) gemm_wgrad(tbase + TcBwd2::P_d0, g.D2, 0, g.X2, 0, 32, acc);
            break;
        case 3:
            gemm_kmajor(tw, g.H1, sWh1, 64, 0, 64, 64, 0);
            break;
        case 4:
            gemm_kmajor(tw, g.H2, sWh2, 16, 0, 16, 64, 0);
            break;
        case 5:
            gemm_dgrad(tw, g.D2, sWh2, 16, 0, 64, 16, 0);
            gemm_wgrad(tbase + TcBwd2::P_h2t, g.H2, 0, g.D2, 0, 16, acc);
            break;
        case 6:
            gemm_dgrad(tw, g.D1, sWh1, 64, 0, 64, 64, 0);
            gemm_wgrad(tbase + TcBwd2::P_h1, g.D1, 0, g.H1, 0, 72, acc);
            break;
        case 7:
            gemm_dgrad(tw, g.D2, sWhf0, 128, 0, 32, 64, 0);
            gemm_wgrad(tbase + TcBwd2::P_h0, g.D2, 0, g.X, 0, 40, acc);
            break;
        case 8:  // feature_mlp layer 0
            gemm_kmajor(tw, g.X, sWhf0, 128, 64, 64, 32, 0);
            break;
        case 9:
            gemm_kmajor(tw, g.H1, sWf1, 64, 0, 64, 64, 0);
            break;
        case 10:
            gemm_kmajor(tw, g.H2, sWf2, 32, 0, 32, 64, 0);
            break;
        case 11:
            gemm_dgrad(tw, g.D2, sWf2, 32, 0, 64, 32, 0);
            gemm_wgrad(tbase + TcBwd2::P_f2t, g.H2, 0, g.D2, 0, 32, acc);
            break;
        case 12:
            gemm_dgrad(tw, g.D1, sWf1, 64, 0, 64, 64, 0);
            gemm_wgrad(tbase + TcBwd2::P_f1, g.D1, 0, g.H1, 0, 72, acc);
            break;
        case 13:
            gemm_dgrad(tw, g.D2, sWhf0, 128, 64, 32, 64, 0);
            gemm_wgrad(tbase + TcBwd2::P_f0, g.D2, 0, g.X, 0, 40, acc);
            break;
        case 14:  // base layer 1 backward (H0 is back in tH1)
            gemm_dgrad(tw, g.D2, sWb1, 16, 0, 64, 16, 0);
            gemm_wgrad(tbase + TcBwd2::P_b1t, g.H1, 0, g.D2, 0, 16, acc);
            break;
        default:  // 15: base layer 0 backward
            if (want_denc) gemm_dgrad(tw, g.D1, sWb0, 64, 0, 32, 64, 0);
            gemm_wgrad(tbase + TcBwd2::P_b0, g.D1, 0, g.A0, 0, 40, acc);
            break;
    }
}

__global__ void __launch_bounds__(TcBwd2::kThreads, 1)
field_mlp_tc_bwd2_kernel(umhs_points P, int64_t n_total, const float *__restrict__ enc, umhs_mlp_params mp,
                         umhs_field_out_grads g, float *__restrict__ d_enc, umhs_mlp_grads dp) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = mp.num_classes, spec = mp.specular;
    const int grp = tid >> 7;          // 0, 1: data groups; 2: the issuer warp
    const int row = tid & 127;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + TcSmem::ctrl);  // ready[0], ready[1], done[0], done[1]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + TcSmem::ctrl + 32);
    float *sbias = reinterpret_cast<float *>(smem + TcBwd2::sbias);
    const int grp_bytes = spec ? TcBwd2::grp_bytes_spec : TcBwd2::grp_bytes;
    uint8_t *gbase = smem + TcBwd2::grp + (grp < 2 ? grp : 0) * grp_bytes;

    // ---- one-time setup: weights, ones chunks, barriers, TMEM
    if (tid < TM) stage_all_weights(smem, mp, false, tid);  // the staging helpers stride by TM threads
    if (tid < 256) {
        const float one[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
        store8(gbase + TcBwd2::tA0, row, 32, CH_ACT, one);
        store8(gbase + TcBwd2::tX, row, 32, CH_ACT, one);
        store8(gbase + TcBwd2::tH1, row, 64, CH_ACT, one);
    }
    if (tid < 80) sbias[tid] = 0.f;
    if (tid == 0) {
        mbar_init(bars + 0, TM);
        mbar_init(bars + 1, TM);
        mbar_init(bars + 2, 1);
        mbar_init(bars + 3, 1);
        fence_mbar_init();
    }
    if (warp == 8) tmem_alloc<TcBwd2::ncols>(tmem_slot);
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tbase = *tmem_slot;

    n_total = effective_n(P, n_total);
    const int64_t n_tiles = (n_total + TM - 1) / TM;
    // tile of (iteration it, group q) of this CTA
    auto tile_of = [&](int64_t it, int q) { return (it * gridDim.x + blockIdx.x) * 2 + q; };

    const uint32_t sWb0 = smem_u32(smem + TcSmem::w_base0), sWb1 = smem_u32(smem + TcSmem::w_base1);
    const uint32_t sWhf0 = smem_u32(smem + TcSmem::w_hf0), sWh1 = smem_u32(smem + TcSmem::w_h1);
    const uint32_t sWf1 = smem_u32(smem + TcSmem::w_f1), sWh2 = smem_u32(smem + TcSmem::w_h2);
    const uint32_t sWf2 = smem_u32(smem + TcSmem::w_f2), sWd0 = smem_u32(smem + TcSmem::w_d0);

    if (grp == 2) {
        // ======================= MMA issuer warp
        GroupAddr ga[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const uint32_t b = smem_u32(smem + TcBwd2::grp + q * grp_bytes);
            ga[q].A0 = b + TcBwd2::tA0; ga[q].X = b + TcBwd2::tX; ga[q].H1 = b + TcBwd2::tH1; ga[q].H2 = b + TcBwd2::tH2;
            ga[q].D1 = b + TcBwd2::tD1; ga[q].D2 = b + TcBwd2::tD2; ga[q].X2 = b + TcBwd2::tX2;
        }
        uint32_t par[2] = {0u, 0u};
        for (int64_t it = 0; tile_of(it, 0) < n_tiles; ++it) {
            const bool act1 = tile_of(it, 1) < n_tiles;
            for (int r = 0; r < TcBwd2::kRounds; ++r) {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    if (q == 1 && !act1) continue;
                    if (lane == 0) {
                        mbar_wait(bars + q, par[q]);  // group q wrote its operand tiles for round r
                        fence_after_sync();
                        const uint32_t acc = (it == 0 && q == 0) ? 0u : 1u;  // the very first tile initialises
                        issue_round(r, ga[q], tbase + 64u * q, tbase, acc, spec, d_enc != nullptr, sWb0, sWb1, sWhf0,
                                    sWh1, sWf1, sWh2, sWf2, sWd0);
                        commit(bars + 2 + q);
                    }
                    par[q] ^= 1u;
                    __syncwarp();
                }
            }
        }
    } else {
        // ======================= data groups: thread `row` of group `grp` owns one sample of the group's tile
        uint8_t *tA0 = gbase + TcBwd2::tA0, *tX = gbase + TcBwd2::tX, *tH1 = gbase + TcBwd2::tH1;
        uint8_t *tH2 = gbase + TcBwd2::tH2, *tD1 = gbase + TcBwd2::tD1, *tD2 = gbase + TcBwd2::tD2;
        uint8_t *tX2 = gbase + TcBwd2::tX2;
        uint64_t *ready = bars + grp, *done = bars + 2 + grp;
        uint32_t par = 0;
        const uint32_t trow = tbase + ((uint32_t)((warp & 3) * 32) << 16) + 64u * grp;
        const float *Bb0 = reinterpret_cast<const float *>(smem + TcSmem::b_base0);
        const float *Bb1 = reinterpret_cast<const float *>(smem + TcSmem::b_base1);
        const float *Bh0 = reinterpret_cast<const float *>(smem + TcSmem::b_h0);
        const float *Bf0 = reinterpret_cast<const float *>(smem + TcSmem::b_f0);
        const float *Bh1 = reinterpret_cast<const float *>(smem + TcSmem::b_h1);
        const float *Bf1 = reinterpret_cast<const float *>(smem + TcSmem::b_f1);
        const float *Bh2 = reinterpret_cast<const float *>(smem + TcSmem::b_h2);
        const float *Bf2 = reinterpret_cast<const float *>(smem + TcSmem::b_f2);
        const float *Bd0 = reinterpret_cast<const float *>(smem + TcSmem::b_d0);
        float *sb_h2 = sbias, *sb_f2 = sbias + 16, *sb_b1 = sbias + 48, *sb_d0 = sbias + 64;
        const float invT = 1.0f / mp.temperature;
        // one round trip with the issuer: operands written -> MMAs of the round done
        auto round_trip = [&]() {
            fence_proxy_async();
            fence_before_sync();
            mbar_arrive(ready);
            mbar_wait(done, par);
            par ^= 1u;
            fence_after_sync();
        };
        auto bias_sum = [&](auto &v, int n, float *dst) {  // column sums of a narrow delta block (n <= 17)
#pragma unroll
            for (int k = 0; k < 17; ++k) {
                if (k < n && k < (int)(sizeof(v) / sizeof(float))) {  // warp-uniform
                    const float s = warp_sum(v[k]);
                    if (lane == 0 && s != 0.f) atomicAdd(dst + k, s);
                }
            }
        };

        // (no cross-tile prefetch here: the other group's work hides this group's global-load latency, and the
        //  registers are needed for the packed H0 activations)
        for (int64_t it = 0; tile_of(it, grp) < n_tiles; ++it) {
            const int64_t tile = tile_of(it, grp);
            const int64_t n0 = tile * TM;
            const bool valid = n0 + row < n_total;
            const int64_t n = valid ? n0 + row : n_total - 1;
            const Point pt = eval_point(P, n);
            float pe[12];
            pos_encoding_tc(pt, pe);
            // ---------------- round 0: enc -> tA0 [, X2 -> tX2]
            {
                const float4 *e4 = reinterpret_cast<const float4 *>(enc + enc_row(P, n) * 32);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float4 a = __ldg(e4 + 2 * c), b = __ldg(e4 + 2 * c + 1);
                    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                    store8(tA0, row, 8 * c, CH_ACT, v);
                }
            }
            if (spec) {
                float x2[32], sh[16];
                sh_deg3_tc((pt.dx + 1.0f) / 2.0f, (pt.dy + 1.0f) / 2.0f, (pt.dz + 1.0f) / 2.0f, sh);
#pragma unroll
                for (int i = 0; i < 16; ++i) x2[i] = sh[i];
#pragma unroll
                for (int i = 0; i < 12; ++i) x2[16 + i] = pe[i];
#pragma unroll
                for (int i = 28; i < 32; ++i) x2[i] = 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c) store8(tX2, row, 8 * c, CH_ACT, x2 + 8 * c);
            }
            round_trip();
            // ---------------- round 1: H0 = relu(. + b) -> tH1, and kept in registers as packed bf16
            uint32_t h0p[32];
            {
                float v[64];
                tmem_ld64(trow, v);
#pragma unroll
                for (int c = 0; c < 64; c += 8) {
                    float o[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) o[i] = fmaxf(v[c + i] + Bb0[c + i], 0.f);
                    store8(tH1, row, c, CH_ACT, o);
#pragma unroll
                    for (int i = 0; i < 8; i += 2) {
                        __nv_bfloat162 p2 = __floats2bfloat162_rn(o[i], o[i + 1]);
                        h0p[(c + i) >> 1] = *reinterpret_cast<uint32_t *>(&p2);
                    }
                }
            }
            round_trip();
            // ---------------- round 2: o16 ; X = [pe, emb] -> tX [; d spec_hidden (padded to 64 columns) -> tD2]
            float o16[16], demb[15];
            tmem_ld16(trow, o16);
#pragma unroll
            for (int i = 0; i < 16; ++i) o16[i] += Bb1[i];
#pragma unroll
            for (int i = 0; i < 15; ++i) demb[i] = (valid && g.d_embedding) ? __ldg(g.d_embedding + n * 15 + i) : 0.f;
            if (spec) {
                float h[16], dsh[16];
                tmem_ld16(trow + 16, h);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float4 gq = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (valid && g.d_spec_hidden)
                        gq = __ldg(reinterpret_cast<const float4 *>(g.d_spec_hidden + n * 16) + q);
                    dsh[4 * q] = (h[4 * q] + Bd0[4 * q] > 0.f) ? gq.x : 0.f;
                    dsh[4 * q + 1] = (h[4 * q + 1] + Bd0[4 * q + 1] > 0.f) ? gq.y : 0.f;
                    dsh[4 * q + 2] = (h[4 * q + 2] + Bd0[4 * q + 2] > 0.f) ? gq.z : 0.f;
                    dsh[4 * q + 3] = (h[4 * q + 3] + Bd0[4 * q + 3] > 0.f) ? gq.w : 0.f;
                }
                const float z8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                store8(tD2, row, 0, CH_ACT, dsh);
                store8(tD2, row, 8, CH_ACT, dsh + 8);
#pragma unroll
                for (int c = 16; c < 64; c += 8) store8(tD2, row, c, CH_ACT, z8);
                bias_sum(dsh, 16, sb_d0);
            }
            {
                float x[32];
#pragma unroll
                for (int i = 0; i < 12; ++i) x[i] = pe[i];
#pragma unroll
                for (int i = 0; i < 15; ++i) x[12 + i] = o16[1 + i];
#pragma unroll
                for (int i = 27; i < 32; ++i) x[i] = 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c) store8(tX, row, 8 * c, CH_ACT, x + 8 * c);
            }
            round_trip();
            // ================= mlp_head (m = 0) then feature_mlp (m = 1): rounds 3..7 and 8..13
#pragma unroll 1
            for (int m = 0; m < 2; ++m) {
                const float *B0 = m ? Bf0 : Bh0, *B1 = m ? Bf1 : Bh1, *B2 = m ? Bf2 : Bh2;
                const int nout = m ? K + spec : K;
                if (m == 1) {
                    // collect d X of mlp_head (working columns 0..31), then run feature_mlp layer 0 (round 8)
                    float dx[16];
                    tmem_ld16(trow, dx);
#pragma unroll
                    for (int i = 12; i < 16; ++i) demb[i - 12] += dx[i];
                    tmem_ld16(trow + 16, dx);
#pragma unroll
                    for (int i = 0; i < 11; ++i) demb[4 + i] += dx[i];
                    round_trip();
                }
                acc_to_tile_relu(trow, 0, 64, B0, tH1, row, 0);
                round_trip();  // layer 1
                acc_to_tile_relu(trow, 0, 64, B1, tH2, row, 0);
                // incoming gradients of this MLP's outputs: fetched while the tensor core runs the last layer
                float gin[17];
#pragma unroll
                for (int k = 0; k < 16; ++k)
                    gin[k] = (valid && k < K) ? __ldg((m ? g.d_abund : g.d_scalar) + n * K + k) : 0.f;
                gin[16] = (valid && m == 1 && spec && g.d_s1) ? __ldg(g.d_s1 + n) : 0.f;
                round_trip();  // layer 2
                {
                    float z[32], dz[32];
                    {
                        float t16[16];
                        tmem_ld16(trow, t16);
#pragma unroll
                        for (int k = 0; k < 16; ++k) z[k] = t16[k];
                        if (m == 1) tmem_ld16(trow + 16, t16);
#pragma unroll
                        for (int k = 0; k < 16; ++k) z[16 + k] = (m == 1) ? t16[k] : 0.f;
                    }
#pragma unroll
                    for (int k = 0; k < 32; ++k) dz[k] = 0.f;
                    if (valid) {
                        if (m == 0) {
#pragma unroll
                            for (int k = 0; k < 16; ++k) {
                                if (k < K) {
                                    const float s = __frcp_rn(1.0f + __expf(-(z[k] + B2[k])));
                                    dz[k] = gin[k] * s * (1.0f - s);
                                }
                            }
                        } else {
                            float mx = -3.0e38f, e[16], den = 0.f, dot = 0.f;
#pragma unroll
                            for (int k = 0; k < 16; ++k) {
                                z[k] = (z[k] + B2[k]) * invT;
                                if (k < K) mx = fmaxf(mx, z[k]);
                            }
#pragma unroll
                            for (int k = 0; k < 16; ++k) {
                                e[k] = (k < K) ? __expf(z[k] - mx) : 0.f;
                                den += e[k];
                                dot += e[k] * gin[k];
                            }
                            const float rden = __frcp_rn(den);
                            dot *= rden;
#pragma unroll
                            for (int k = 0; k < 16; ++k) dz[k] = e[k] * rden * invT * (gin[k] - dot);
                            if (spec) {
                                float zs = z[16] + B2[16];
#pragma unroll
                                for (int k = 0; k < 16; ++k)
                                    if (k == K) zs = z[k] * mp.temperature;
                                const float s = __frcp_rn(1.0f + __expf(-zs));
                                const float ds = gin[16] * s * (1.0f - s);
#pragma unroll
                                for (int k = 0; k < 17; ++k)
                                    if (k == K) dz[k] = ds;
                            }
                        }
                    }
                    store8(tD2, row, 0, CH_ACT, dz);
                    store8(tD2, row, 8, CH_ACT, dz + 8);
                    if (m == 1) {
                        store8(tD2, row, 16, CH_ACT, dz + 16);
                        store8(tD2, row, 24, CH_ACT, dz + 24);
                    }
                    bias_sum(dz, nout, m ? sb_f2 : sb_h2);
                }
                round_trip();  // d layer 2
                acc_to_delta_masked(trow, 0, tH2, tD1, row, valid);
                round_trip();  // d layer 1
                acc_to_delta_masked(trow, 0, tH1, tD2, row, valid);
                round_trip();  // d layer 0
            }
            // ================= base MLP backward (rounds 14, 15)
            {
                float dx[16], dO[16];
                tmem_ld16(trow, dx);
#pragma unroll
                for (int i = 12; i < 16; ++i) demb[i - 12] += dx[i];
                tmem_ld16(trow + 16, dx);
#pragma unroll
                for (int i = 0; i < 11; ++i) demb[4 + i] += dx[i];
                const float dsg = (valid && g.d_sigma) ? __ldg(g.d_sigma + n) : 0.f;
                dO[0] = dsg * pt.sel * __expf(fminf(fmaxf(o16[0], -15.0f), 15.0f));
#pragma unroll
                for (int i = 0; i < 15; ++i) dO[1 + i] = valid ? demb[i] : 0.f;
                store8(tD2, row, 0, CH_ACT, dO);
                store8(tD2, row, 8, CH_ACT, dO + 8);
                bias_sum(dO, 16, sb_b1);
                // H0 back into tH1 (operand of the base layer-1 weight gradient)
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    *reinterpret_cast<uint4 *>(tH1 + tile_off(row, 8 * c, CH_ACT)) =
                        make_uint4(h0p[4 * c], h0p[4 * c + 1], h0p[4 * c + 2], h0p[4 * c + 3]);
            }
            round_trip();
            {
                float v[64];
                tmem_ld64(trow, v);
#pragma unroll
                for (int c = 0; c < 64; c += 8) {
                    float o[8];
#pragma unroll
                    for (int i = 0; i < 8; i += 2) {
                        const float2 h = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&h0p[(c + i) >> 1]));
                        o[i] = (valid && h.x > 0.f) ? v[c + i] : 0.f;
                        o[i + 1] = (valid && h.y > 0.f) ? v[c + i + 1] : 0.f;
                    }
                    store8(tD1, row, c, CH_ACT, o);
                }
            }
            round_trip();
            if (d_enc) {
                float v[16];
                float4 *dst = reinterpret_cast<float4 *>(d_enc + n * 32);
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    tmem_ld16(trow + 16 * c, v);
                    if (valid) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            dst[4 * c + q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                    }
                }
            }
            fence_before_sync();
        }
    }

    // ================= flush the weight-gradient accumulators: row r = 16 * (warp & 3) + lane (lane < 16)
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (grp < 2) {
        const uint32_t tq = tbase + ((uint32_t)((warp & 3) * 32) << 16);
        const int r = (warp & 3) * 16 + lane;
        if (grp == 0) {
            flush_acc(tq, TcBwd2::P_h2t, K, lane, [&](int c, float v) { if (v != 0.f) atomicAdd(dp.head_w2 + c * 64 + r, v); });
            flush_acc(tq, TcBwd2::P_h1, 65, lane, [&](int c, float v) {
                if (v != 0.f) atomicAdd(c < 64 ? dp.head_w1 + r * 64 + c : dp.head_b1 + r, v); });
            flush_acc(tq, TcBwd2::P_h0, 33, lane, [&](int c, float v) {
                if (v != 0.f && (c < 27 || c == 32)) atomicAdd(c < 27 ? dp.head_w0 + r * 27 + c : dp.head_b0 + r, v); });
            flush_acc(tq, TcBwd2::P_b1t, 16, lane, [&](int c, float v) { if (v != 0.f) atomicAdd(dp.base_w1 + c * 64 + r, v); });
            if (spec)
                flush_acc(tq, TcBwd2::P_d0, 28, lane, [&](int c, float v) {
                    if (r < 16 && v != 0.f) atomicAdd(dp.dir_w0 + r * 28 + c, v); });
        } else {
            flush_acc(tq, TcBwd2::P_f2t, K + spec, lane, [&](int c, float v) { if (v != 0.f) atomicAdd(dp.feat_w2 + c * 64 + r, v); });
            flush_acc(tq, TcBwd2::P_f1, 65, lane, [&](int c, float v) {
                if (v != 0.f) atomicAdd(c < 64 ? dp.feat_w1 + r * 64 + c : dp.feat_b1 + r, v); });
            flush_acc(tq, TcBwd2::P_f0, 33, lane, [&](int c, float v) {
                if (v != 0.f && (c < 27 || c == 32)) atomicAdd(c < 27 ? dp.feat_w0 + r * 27 + c : dp.feat_b0 + r, v); });
            flush_acc(tq, TcBwd2::P_b0, 33, lane, [&](int c, float v) {
                if (v != 0.f) atomicAdd(c < 32 ? dp.base_w0 + r * 32 + c : dp.base_b0 + r, v); });
        }
    }
    if (tid < 80) {
        const float a = sbias[tid];
        float *dst = nullptr;
        if (tid < 16) dst = (tid < K) ? dp.head_b2 + tid : nullptr;
        else if (tid < 48) dst = (tid - 16 < K + spec) ? dp.feat_b2 + (tid - 16) : nullptr;
        else if (tid < 64) dst = dp.base_b1 + (tid - 48);
        else dst = spec ? dp.dir_b0 + (tid - 64) : nullptr;
        if (dst != nullptr && a != 0.f) atomicAdd(dst, a);
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 8) tmem_dealloc<TcBwd2::ncols>(tbase);
}

}  // namespace umhs

extern "C" int umhs_field_mlp_tc_bwd2(const umhs_points *pts, int64_t n, const float *enc, const umhs_mlp_params *p,
                                      const umhs_field_out_grads *g, float *d_enc, const umhs_mlp_grads *dp,
                                      void *stream) {
    if (n <= 0) return UMHS_OK;
    UMHS_CHECK_ARG(pts && enc && g && dp && p, "null argument");
    UMHS_CHECK_ARG(p->num_classes >= 1 && p->num_classes <= UMHS_MAX_CLASSES && p->temperature > 0.f, "bad K / T");
    UMHS_CHECK_ARG(g->d_scalar && g->d_abund, "d_scalar/d_abund are required");
    UMHS_CHECK_ARG(dp->base_w0 && dp->base_b0 && dp->base_w1 && dp->base_b1 && dp->head_w0 && dp->head_b0 &&
                       dp->head_w1 && dp->head_b1 && dp->head_w2 && dp->head_b2 && dp->feat_w0 && dp->feat_b0 &&
                       dp->feat_w1 && dp->feat_b1 && dp->feat_w2 && dp->feat_b2,
                   "null parameter-gradient pointer");
    UMHS_CHECK_ARG(!p->specular || (p->dir_w0 && p->dir_b0 && dp->dir_w0 && dp->dir_b0), "specular pointers missing");
    const int bytes = umhs::TcBwd2::grp + 2 * (p->specular ? umhs::TcBwd2::grp_bytes_spec : umhs::TcBwd2::grp_bytes);
    UMHS_CHECK_ARG(bytes <= 232448, "shared-memory plan exceeds 227 KB");
    UMHS_CHECK_CUDA(cudaFuncSetAttribute(umhs::field_mlp_tc_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         bytes));
    // one persistent CTA per SM, two tiles per CTA iteration; never more CTAs than tile pairs
    const int64_t n_pairs = ((n + umhs::TM - 1) / umhs::TM + 1) / 2;
    const int grid = (int)(n_pairs < umhs::kNumSMs ? n_pairs : umhs::kNumSMs);
    umhs::field_mlp_tc_bwd2_kernel<<<grid, umhs::TcBwd2::kThreads, bytes, umhs::as_stream(stream)>>>(*pts, n, enc, *p,
                                                                                                    *g, d_enc, *dp);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}
