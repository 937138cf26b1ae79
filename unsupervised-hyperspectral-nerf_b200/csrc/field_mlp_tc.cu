// field_mlp_tc.cu -- the fused field MLPs on the 5th-generation tensor cores (tcgen05 + TMEM), bf16 operands,
// fp32 accumulation.  Same arithmetic and interface as field_mlp.cu (the fp32 SIMT path); this is the default.
//
// Execution model (all three kernels): ONE persistent CTA per SM holds every layer's weights in shared memory as
// bf16 UMMA operand tiles and runs several 128-sample tiles at once.  A "group" = 128 threads = one tile; thread t
// owns sample row t, which is also TMEM lane t of the group's accumulator columns, so the epilogues (ReLU, softmax,
// sigmoid, trunc_exp) are thread-private.  One extra warp is the MMA issuer: a group finishes an epilogue (writes
// the next operand tile), arrives on its `ready` mbarrier and waits on its `done` mbarrier; the issuer polls the
// `ready` barriers, issues that group's next round of tcgen05.mma and commits it to `done`.  While one group waits
// for the tensor core the others run their epilogues.
//
// Epilogues are kept to the minimum instruction count (they, not the tensor pipe, bound these kernels):
//   * biases are folded into the GEMMs: every weight tile carries its bias as one more K column, multiplied by a
//     column of ones (a constant tile, or a spare padding column of the 27/28-wide inputs);
//   * accumulators become the next bf16 operand with one cvt.rn.relu.bf16x2.f32 per two columns;
//   * ReLU masks of the backward pass are applied on packed bf16x2 words (set.gt mask + and).
//
// Forward rounds per tile:   enc->[base0] -> [base1 | dir0] -> [head0+feat0] -> [head1, feat1] -> [head2, feat2]
#include <utility>

#include "common.cuh"
#include "hash_scatter.cuh"
#include "tc05.cuh"

namespace umhs {
using namespace tc05;

constexpr int TM = 128;          // samples per tile == threads per group
#ifndef UMHS_FWD_ISSUERS
#define UMHS_FWD_ISSUERS 1   // MMA-issuing warps of the forward kernels: 1 = one, 2 = one per data group.  Measured: the
#endif                       // full forward 120.8 (one) against 124.9 us at 825 k samples; the density kernel 45.0 (one
                             // per group) against 47.3 us alone at 825 k samples, but 0.0552 against 0.0525 ms per
                             // training step (two calls: candidates + an occupancy chunk) -- one warp it is
#ifndef UMHS_BWD_ISSUERS
#define UMHS_BWD_ISSUERS 1   // ... of the backward kernel: 2 = one per data group (each group then owns a copy of the
#endif                       // weight-gradient accumulators), 1 = one for both.  Measured equal (430 / 431 us): a round is
                             // bound by the group's own chain epilogue -> hand-off -> MMA -> hand-off, not by queueing
#ifndef UMHS_SCATTER_WARPS
#define UMHS_SCATTER_WARPS (8 - UMHS_BWD_ISSUERS)   // fills the fourth warpgroup
#endif
#ifndef UMHS_ISSUER_SPIN
#define UMHS_ISSUER_SPIN 4096
#endif
#ifndef UMHS_GROUP_SPIN
#define UMHS_GROUP_SPIN 0
#endif
#ifndef UMHS_FLAG_READY
#define UMHS_FLAG_READY 2
#endif
// 1: the groups signal "operands ready" through a plain shared-memory counter (red.release / ld.acquire) that the
// issuer polls, instead of an mbarrier: an mbarrier test by the thread that has just issued tcgen05.commit does not
// return before that commit has arrived (measured: ~230 cycles per round even on a long-completed phase), which
// serialised the groups on the MMA latency.
// 2: ... through a named hardware barrier (bar.arrive by the group's 128 threads, bar.sync by the issuer warp): the
// flag poll is a shared-memory read and queues behind the operand reads of the MMAs the issuer has just launched.
constexpr bool kFlagReady = UMHS_FLAG_READY == 1;
constexpr bool kNamedBarReady = UMHS_FLAG_READY == 2;
constexpr int kReadyBarId0 = 1;            // named barrier of group g = kReadyBarId0 + g (0 is __syncthreads)
constexpr int kReadyBarCount = TM + 32;    // the group's threads arrive, the issuer warp syncs
constexpr int kIssuerSpin = UMHS_ISSUER_SPIN;  // test_wait probes before the suspending try_wait (see tc05::mbar_wait)
constexpr int kGroupSpin = UMHS_GROUP_SPIN;
constexpr int CH_ACT = 2048;     // chunk stride of a 128-row activation tile

// ---- shared-memory plan (bytes).  Weight tiles: [rows x K] bf16 K-major, CH = rows * 16; bias = one more K column.
struct TcSmem {
    static constexpr int w_base0 = 0;                       // [64 x 48]   bias in column 32
    static constexpr int w_base1 = w_base0 + 64 * 48 * 2;   // [16 x 80]   bias in column 64
    static constexpr int w_hf0 = w_base1 + 16 * 80 * 2;     // [128 x 32]  rows 0-63 mlp_head.0, 64-127 feature_mlp.0; bias col 27
    static constexpr int w_h1 = w_hf0 + 128 * 32 * 2;       // [64 x 80]
    static constexpr int w_f1 = w_h1 + 64 * 80 * 2;         // [64 x 80]
    static constexpr int w_h2 = w_f1 + 64 * 80 * 2;         // [16 x 80]
    static constexpr int w_f2 = w_h2 + 16 * 80 * 2;         // [32 x 80]
    static constexpr int w_d0 = w_f2 + 32 * 80 * 2;         // [16 x 32]   bias in column 28
    static constexpr int w_end = w_d0 + 16 * 32 * 2;
    static constexpr int ones = w_end;                      // [128 x 16] of 1.0
    static constexpr int ctrl = ones + 2 * CH_ACT;          // mbarriers ready[4], done[4]; tmem address
    static constexpr int sbias = ctrl + 128;                // backward: fp32 [64] narrow-layer bias gradients
    static constexpr int scat = sbias + 256;                // backward + scatter: uint32 flushed[8], next_unit
    static constexpr int trace = scat + 128;                // developer cycle trace: 2 threads x 256 stamps x 8 B
    static constexpr int groups = trace + 4096;
    static_assert(groups % 128 == 0, "operand tiles must stay 128-byte aligned");
};
__device__ __forceinline__ uint32_t *ready_flags(uint8_t *smem) {  // uint32 [4] behind the mbarriers and the TMEM slot
    return reinterpret_cast<uint32_t *>(smem + TcSmem::ctrl + 96);
}
constexpr int COL_B_BASE0 = 32, COL_B_64 = 64, COL_B_X = 27, COL_B_X2 = 28;

// stage an nn.Linear weight [out x in] (+ bias in K column `bias_col`) as a bf16 K-major operand tile
__device__ __forceinline__ void stage_weight(uint8_t *tile, int rows, int row0, const float *w, const float *bias, int out,
                                             int in, int k_tile, int bias_col, int tid, int nthr) {
    const int ch = rows * 16, cpr = k_tile / 8;
    for (int i = tid; i < out * cpr; i += nthr) {
        const int j = i / cpr, c0 = (i - j * cpr) * 8;
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int c = c0 + q;
            v[q] = (c < in) ? __ldg(w + j * in + c) : ((c == bias_col && bias != nullptr) ? __ldg(bias + j) : 0.f);
        }
        store8(tile, row0 + j, c0, ch, v);
    }
}
__device__ __forceinline__ void zero_rows(uint8_t *tile, int rows, int row_from, int row_to, int k_tile, int tid, int nthr) {
    const int ch = rows * 16, cpr = k_tile / 8;
    const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int i = tid; i < (row_to - row_from) * cpr; i += nthr) {
        const int j = i / cpr, c0 = (i - j * cpr) * 8;
        store8(tile, row_from + j, c0, ch, z);
    }
}

__device__ __forceinline__ void stage_all_weights(uint8_t *smem, const umhs_mlp_params &mp, bool density_only, int tid,
                                                  int nthr) {
    const int K = mp.num_classes, spec = mp.specular;
    stage_weight(smem + TcSmem::w_base0, 64, 0, mp.base_w0, mp.base_b0, 64, 32, 48, COL_B_BASE0, tid, nthr);
    stage_weight(smem + TcSmem::w_base1, 16, 0, mp.base_w1, mp.base_b1, 16, 64, 80, COL_B_64, tid, nthr);
    {
        const float one[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
        for (int i = tid; i < 2 * TM; i += nthr) store8(smem + TcSmem::ones, i & (TM - 1), 8 * (i >> 7), CH_ACT, one);
    }
    if (density_only) return;
    stage_weight(smem + TcSmem::w_hf0, 128, 0, mp.head_w0, mp.head_b0, 64, 27, 32, COL_B_X, tid, nthr);
    stage_weight(smem + TcSmem::w_hf0, 128, 64, mp.feat_w0, mp.feat_b0, 64, 27, 32, COL_B_X, tid, nthr);
    stage_weight(smem + TcSmem::w_h1, 64, 0, mp.head_w1, mp.head_b1, 64, 64, 80, COL_B_64, tid, nthr);
    stage_weight(smem + TcSmem::w_f1, 64, 0, mp.feat_w1, mp.feat_b1, 64, 64, 80, COL_B_64, tid, nthr);
    stage_weight(smem + TcSmem::w_h2, 16, 0, mp.head_w2, mp.head_b2, K, 64, 80, COL_B_64, tid, nthr);
    zero_rows(smem + TcSmem::w_h2, 16, K, 16, 80, tid, nthr);
    stage_weight(smem + TcSmem::w_f2, 32, 0, mp.feat_w2, mp.feat_b2, K + spec, 64, 80, COL_B_64, tid, nthr);
    zero_rows(smem + TcSmem::w_f2, 32, K + spec, 32, 80, tid, nthr);
    if (spec) stage_weight(smem + TcSmem::w_d0, 16, 0, mp.dir_w0, mp.dir_b0, 16, 28, 32, COL_B_X2, tid, nthr);
}

struct WOff {  // byte offsets (from the CTA's dynamic shared memory base) of the weight tiles and of the ones tile
    static constexpr uint32_t b0 = TcSmem::w_base0, b1 = TcSmem::w_base1, hf0 = TcSmem::w_hf0, h1 = TcSmem::w_h1,
                              f1 = TcSmem::w_f1, h2 = TcSmem::w_h2, f2 = TcSmem::w_f2, d0 = TcSmem::w_d0,
                              ones = TcSmem::ones;
};

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

// ---- lean epilogues (bias already inside the accumulator)
// 64 accumulator columns of this thread's row -> ReLU -> bf16 operand tile columns [0, 64); optional packed copy
template <bool KEEP>
__device__ __forceinline__ void epi_relu64(uint32_t taddr, uint8_t *tile, int row, uint32_t *keep) {
    uint32_t r[64];
    tmem_ld64_raw(taddr, r);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        uint4 q;
        q.x = pack_relu_bf16x2(__uint_as_float(r[8 * c]), __uint_as_float(r[8 * c + 1]));
        q.y = pack_relu_bf16x2(__uint_as_float(r[8 * c + 2]), __uint_as_float(r[8 * c + 3]));
        q.z = pack_relu_bf16x2(__uint_as_float(r[8 * c + 4]), __uint_as_float(r[8 * c + 5]));
        q.w = pack_relu_bf16x2(__uint_as_float(r[8 * c + 6]), __uint_as_float(r[8 * c + 7]));
        *reinterpret_cast<uint4 *>(tile + tile_off(row, 8 * c, CH_ACT)) = q;
        if (KEEP) {
            keep[4 * c] = q.x; keep[4 * c + 1] = q.y; keep[4 * c + 2] = q.z; keep[4 * c + 3] = q.w;
        }
    }
}
// 64 accumulator columns, masked by relu'(activation) -> bf16 delta tile.  The activation comes from its operand
// tile (act_tile) or from a packed register copy (act_regs).
__device__ __forceinline__ void epi_mask64(uint32_t taddr, const uint8_t *act_tile, const uint32_t *act_regs,
                                           uint8_t *delta_tile, int row) {
    uint32_t r[64];
    tmem_ld64_raw(taddr, r);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        uint4 h;
        if (act_tile != nullptr) h = *reinterpret_cast<const uint4 *>(act_tile + tile_off(row, 8 * c, CH_ACT));
        else h = make_uint4(act_regs[4 * c], act_regs[4 * c + 1], act_regs[4 * c + 2], act_regs[4 * c + 3]);
        uint4 q;
        q.x = pack_bf16x2(__uint_as_float(r[8 * c]), __uint_as_float(r[8 * c + 1])) & relu_mask_bf16x2(h.x);
        q.y = pack_bf16x2(__uint_as_float(r[8 * c + 2]), __uint_as_float(r[8 * c + 3])) & relu_mask_bf16x2(h.y);
        q.z = pack_bf16x2(__uint_as_float(r[8 * c + 4]), __uint_as_float(r[8 * c + 5])) & relu_mask_bf16x2(h.z);
        q.w = pack_bf16x2(__uint_as_float(r[8 * c + 6]), __uint_as_float(r[8 * c + 7])) & relu_mask_bf16x2(h.w);
        *reinterpret_cast<uint4 *>(delta_tile + tile_off(row, 8 * c, CH_ACT)) = q;
    }
}

// Optional cycle trace (developer aid, off unless umhs_debug_trace() installs a buffer): in CTA 0, thread 0 of
// group 0 (who = 0) and the issuer (who = 1) stamp (tag, clock) pairs.  Stamps go to SHARED memory and are copied out
// when the kernel ends: a stamp that stores to global memory perturbs what it measures (the acquire operations of the
// handshakes then wait for those stores to drain: ~230 phantom cycles per round).
__device__ long long *g_trace = nullptr;
__device__ int g_trace_cap = 0;
constexpr int kTraceCap = 256;  // stamps kept per traced thread
#ifndef UMHS_TRACE
#define UMHS_TRACE 0   // build with -DUMHS_TRACE=1 (tools/build_variant.py) for tools/trace_mlp.py; stamps cost cycles
#endif
struct Tracer {
    uint2 *sbuf;
    int slot, who, skip;
    __device__ __forceinline__ Tracer(bool on, int who_, uint8_t *smem, int skip_)
        : sbuf(nullptr), slot(0), who(who_), skip(skip_) {
        if (UMHS_TRACE && on && blockIdx.x == 0 && g_trace != nullptr)
            sbuf = reinterpret_cast<uint2 *>(smem + TcSmem::trace) + who_ * kTraceCap;
    }
    __device__ __forceinline__ void stamp(int tag) {
        if (UMHS_TRACE && sbuf != nullptr) {
            if (skip > 0) --skip;
            else if (slot < kTraceCap) {
                sbuf[slot] = make_uint2((uint32_t)tag, (uint32_t)clock64());
                ++slot;
            }
        }
    }
    __device__ __forceinline__ void flush() {
        if (!UMHS_TRACE || sbuf == nullptr) return;
        for (int i = 0; i < slot && i < g_trace_cap; ++i) {
            g_trace[2 * (i * 2 + who)] = sbuf[i].x;
            g_trace[2 * (i * 2 + who) + 1] = sbuf[i].y;
        }
    }
};

// ---- coalesced movement of [128 x 32] fp32 row blocks (hash encodings in, their gradients out).
// A thread-per-row access (128 B per lane at a 128 B stride) costs 32 L1 wavefronts per warp instruction; here a warp
// instruction covers 4 whole rows: lane l handles the 16-byte piece (l & 7) of row 4 * i + (l >> 3), i = 0..7.
// load: rows come from this warp's own lanes (row index held by lane 4 * i + (l >> 3), shuffled)
__device__ __forceinline__ void load_enc_rows(const float *__restrict__ enc, int64_t my_row, int lane, float4 (&e)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t r = __shfl_sync(kFull, my_row, 4 * i + (lane >> 3));
        e[i] = __ldg(reinterpret_cast<const float4 *>(enc + r * 32) + (lane & 7));
    }
}
// ... and go to the bf16 operand tile as 8-byte pieces (4 columns) of rows warp_row0 + 4 * i + (l >> 3)
__device__ __forceinline__ void store_enc_rows(uint8_t *tile, int warp_row0, int lane, const float4 (&e)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = warp_row0 + 4 * i + (lane >> 3);
        uint2 q;
        q.x = pack_bf16x2(e[i].x, e[i].y);
        q.y = pack_bf16x2(e[i].z, e[i].w);
        *reinterpret_cast<uint2 *>(tile + tile_off(r, 4 * (lane & 7), CH_ACT)) = q;
    }
}
// ---- bf16 tile layout (common.cuh: enc_tile_piece): the row is already in operand order -- four 16-byte pieces per
// thread, straight from global memory into the tile, no conversion, no cross-lane traffic
__device__ __forceinline__ void load_enc_pieces(const float *__restrict__ enc_tiles, int64_t my_row, float4 (&e)[8]) {
    const uint8_t *base = reinterpret_cast<const uint8_t *>(enc_tiles);
#pragma unroll
    for (int c = 0; c < 4; ++c) e[c] = __ldg(reinterpret_cast<const float4 *>(base + enc_tile_piece(my_row, c)));
}
__device__ __forceinline__ void store_enc_pieces(uint8_t *tile, int row, const float4 (&e)[8]) {
#pragma unroll
    for (int c = 0; c < 4; ++c) *reinterpret_cast<float4 *>(tile + c * CH_ACT + (row >> 3) * 128 + (row & 7) * 16) = e[c];
}
// gradient rows in the tile layout: packed to bf16 and parked in the thread's own slots of a shared-memory tile, then
// (behind the next operand hand-off, see stage_rows) copied out as 16-byte pieces -- a warp writes 512 contiguous bytes
__device__ __forceinline__ void stage_pieces(const float (&v)[32], uint8_t *tile, int row) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
        *reinterpret_cast<uint4 *>(tile + c * CH_ACT + (row >> 3) * 128 + (row & 7) * 16) =
            make_uint4(pack_bf16x2(v[8 * c], v[8 * c + 1]), pack_bf16x2(v[8 * c + 2], v[8 * c + 3]),
                       pack_bf16x2(v[8 * c + 4], v[8 * c + 5]), pack_bf16x2(v[8 * c + 6], v[8 * c + 7]));
}
__device__ __forceinline__ void flush_staged_pieces(float *__restrict__ d_enc_tiles, int64_t n, const uint8_t *tile, int row) {
    if (n < 0) return;
    uint8_t *base = reinterpret_cast<uint8_t *>(d_enc_tiles);
#pragma unroll
    for (int c = 0; c < 4; ++c)
        *reinterpret_cast<uint4 *>(base + enc_tile_piece(n, c)) =
            *reinterpret_cast<const uint4 *>(tile + c * CH_ACT + (row >> 3) * 128 + (row & 7) * 16);
}

// store: every lane holds its own row v[32]; the warp transposes through a private 4 KB staging block (XOR-swizzled
// 16-byte pieces: conflict-free both ways) and writes whole rows.  row_of_lane < 0: that row is not written.
// Two halves, so that the global stores can be issued BEHIND the next operand hand-off (fence.proxy.async also waits
// for the thread's outstanding global stores: stores issued right before it stall the group until they have drained).
__device__ __forceinline__ void stage_rows(const float (&v)[32], uint8_t *stage, int lane) {
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; ++q)
        *reinterpret_cast<float4 *>(stage + lane * 128 + ((q ^ (lane & 7)) << 4)) =
            make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    __syncwarp();
}
__device__ __forceinline__ void flush_staged_rows(float *__restrict__ dst, int64_t my_row, const uint8_t *stage, int lane) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int rl = 4 * i + (lane >> 3);
        const int64_t r = __shfl_sync(kFull, my_row, rl);
        const float4 x = *reinterpret_cast<const float4 *>(stage + rl * 128 + (((lane & 7) ^ (rl & 7)) << 4));
        if (r >= 0) reinterpret_cast<float4 *>(dst + r * 32)[lane & 7] = x;
    }
    __syncwarp();
}
__device__ __forceinline__ void store_rows_coalesced(float *__restrict__ dst, int64_t my_row, const float (&v)[32],
                                                     uint8_t *stage, int lane) {
    stage_rows(v, stage, lane);
    flush_staged_rows(dst, my_row, stage, lane);
}

// ---- data warp -> scatter warps: "the d_enc rows of my first `count` tiles are visible" (SCATTER form of the backward)
__device__ __forceinline__ void publish_rows(uint32_t *flag, uint32_t count, int lane) {
    // CTA scope is enough (publisher and readers share the CTA) and it has to be: a GPU-scope fence here waits behind
    // the thousands of reductions the scatter warps of this SM have in flight (measured: the kernel 3 x slower)
    __threadfence_block();
    __syncwarp();
    if (lane == 0) asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(flag)), "r"(count) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_shared(const uint32_t *flag) {
    uint32_t v;
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(flag)) : "memory");
    return v;
}
template <int R>
__device__ __forceinline__ void setmaxnreg_inc() {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R));
}
template <int R>
__device__ __forceinline__ void setmaxnreg_dec() {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R));
}

// ---- group <-> issuer handshake
struct GroupSync {
    uint64_t *ready, *done;
    uint32_t parity;
    int lane;
    Tracer tr;
    uint32_t *flag = nullptr;  // kFlagReady: this group's "operands ready" counter (4 arrivals per round)
    int bar_id = 0;            // kNamedBarReady: this group's named barrier
    int round = 0;
    // operand tiles written & accumulators consumed -> issuer may run this group's next round; wait for its MMAs
    __device__ __forceinline__ void arrive() {
        tr.stamp(2 * round);
        fence_proxy_async();  // NB: also waits for this thread's outstanding GLOBAL stores -- keep those behind it
        fence_before_sync();
        if (UMHS_TRACE > 1) tr.stamp(200 + round);
        if (kNamedBarReady) {
            named_bar_arrive(bar_id, kReadyBarCount);  // every thread of the group; nobody waits here
            return;
        }
        __syncwarp();
        if (lane == 0) {
            if (kFlagReady) flag_release_add(flag);
            else mbar_arrive(ready);
        }
    }
    __device__ __forceinline__ void wait() {
        mbar_wait<kGroupSpin>(done, parity);
        parity ^= 1u;
        fence_after_sync();
        tr.stamp(2 * round + 1);
        ++round;
    }
    __device__ __forceinline__ void round_trip() {
        arrive();
        wait();
    }
};

// compile-time loop: f(std::integral_constant<int, 0>{}), ..., f(std::integral_constant<int, N - 1>{})
template <class F, int... I>
__device__ __forceinline__ void static_for_impl(F &&f, std::integer_sequence<int, I...>) {
    (f(std::integral_constant<int, I>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F &&f) {
    static_for_impl(f, std::make_integer_sequence<int, N>{});
}

// Issuer: ALL 32 lanes of the issuer warp run this in lock step.  Strict rotation -- per tile iteration, round
// r = 0..ROUNDS-1, group g = 0..G-1 -- fully unrolled, so that every round is straight-line code whose descriptors
// are compile-time constants added to warp-uniform bases: the issuing thread paces the whole SM, and anything it
// computes per MMA (runtime descriptor arithmetic, dispatch, divergence handling) is serialised into every round.
// One lane is elected per round (elect_one) and issues that round's MMAs and their commit.
// issue(round, group, acc): acc = 0 on the CTA's very first tile (persistent accumulators are initialised).
// An issuer warp serves the groups [G0, G1) of the CTA's G groups (one warp for all of them, or one warp per group:
// the MMAs of different groups touch disjoint accumulators, so nothing orders them against each other, and a group no
// longer queues behind the other groups' visits -- ~450 cycles each -- on every one of its rounds).
template <int G0, int G1, int G, int ROUNDS, class F>
__device__ __forceinline__ void issuer_rotation(uint64_t *bars, const int64_t (&n_iter)[G], Tracer tr, F &&issue) {
    uint32_t par[G];
    const uint32_t *flags = reinterpret_cast<const uint32_t *>(bars) + 24;  // ctrl + 96, see ready_flags()
#pragma unroll
    for (int g = 0; g < G; ++g) par[g] = 0u;
#pragma unroll 1
    for (int64_t it = 0; it < n_iter[G0]; ++it) {  // a group never has fewer tiles than the groups behind it
        static_for<ROUNDS>([&](auto rc) {
            static_for<G1 - G0>([&](auto gi) {
                constexpr int g = G0 + decltype(gi)::value;
                constexpr std::integral_constant<int, g> gc{};
                if (it < n_iter[g]) {  // warp-uniform
                    tr.stamp(6000 + 100 * g + decltype(rc)::value);
                    if (kNamedBarReady) {
                        named_bar_sync(kReadyBarId0 + g, kReadyBarCount);
                    } else if (kFlagReady) {
                        par[g] += 4u;  // running arrival count of this group (one per data warp and round)
                        flag_wait(flags + g, par[g]);
                    } else {
                        mbar_wait<kIssuerSpin>(bars + g, par[g]);
                        par[g] ^= 1u;
                    }
                    tr.stamp(3000 + 100 * g + decltype(rc)::value);
                    fence_after_sync();
                    tr.stamp(2000 + 100 * g + decltype(rc)::value);
                    if (elect_one()) {
                        // `it >> 40` is zero, but not provably so: the descriptors of a round are then recomputed in
                        // the uniform datapath (a couple of UIADD3 per MMA) instead of being hoisted out of the tile
                        // loop into ~50 vector registers that need an R2UR each before every tcgen05.mma
                        // Phase 0: the MMAs the group's next epilogue waits for (forward / dgrad), committed at once;
                        // phase 1: MMAs nobody waits for (weight gradients into the persistent accumulators) go out
                        // BEHIND the commit, while the group already runs its epilogue.  The next commit of this thread
                        // covers them, so an operand tile they read may be overwritten one round later, not earlier.
                        issue(rc, gc, (it == 0 && g == G0) ? 0u : 1u, (uint32_t)(it >> 40), std::integral_constant<int, 0>{});
                        commit(bars + G + g);
                        tr.stamp(4000 + 100 * g + decltype(rc)::value);
                        issue(rc, gc, (it == 0 && g == G0) ? 0u : 1u, (uint32_t)(it >> 40), std::integral_constant<int, 1>{});
                        tr.stamp(5000 + 100 * g + decltype(rc)::value);
                    }
                    __syncwarp();
                    tr.stamp(1000 + 100 * g + decltype(rc)::value);
                }
            });
        });
    }
    tr.flush();
}

// ---- saved activations (training forward -> backward without recomputation).  Per 128-sample tile one block of
// UMHS_ACT_TILE_BYTES: the operand tiles exactly as they sit in shared memory (chunk c of 8 columns at c * 2048, row r
// inside a chunk at (r >> 3) * 128 + (r & 7) * 16): the backward brings a whole tile back with one bulk copy.
struct ActTile {
    static constexpr int H0 = 0, X = 16384, HH1 = 24576, HH2 = 40960, HF1 = 57344, HF2 = 73728;
    static constexpr int bytes = UMHS_ACT_TILE_BYTES;
};
static_assert(ActTile::HF2 + 16384 == ActTile::bytes, "saved-activation block");
// Each thread copies its own row (one 16-byte piece per chunk) from the shared-memory tile it has just written to the
// block in global memory, BEHIND the hand-off of the round (the values are read back from shared memory, so nothing
// stays in registers, and the streaming stores have a whole round to drain before the next hand-off's
// fence.proxy.async, which waits for a thread's outstanding global stores).  A warp's 32 rows are 512 contiguous bytes
// per chunk.  (Measured first: one bulk store, cp.async.bulk shared -> global, of those 512 bytes per warp and chunk --
// 1.1 M copy-engine operations per forward, ~40 cycles each: 3.6 TB/s, +70 us.)
__device__ __forceinline__ void save_row(uint8_t *dst_block, int dst_off, const uint8_t *smem_tile, int chunks, int row) {
    const int off = (row >> 3) * 128 + (row & 7) * 16;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        if (c < chunks) {
            const float4 v = *reinterpret_cast<const float4 *>(smem_tile + c * CH_ACT + off);
            __stcs(reinterpret_cast<float4 *>(dst_block + dst_off + c * CH_ACT + off), v);
        }
    }
}

// issuer warp `ii` of `I` serves groups [ii * G / I, (ii + 1) * G / I)
template <int I, int G, int ROUNDS, class F>
__device__ __forceinline__ void issuer_dispatch(int ii, uint64_t *bars, const int64_t (&n_iter)[G], uint8_t *smem, int lane,
                                                F &&issue) {
    static_for<I>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        if (ii == i)  // warp-uniform
            issuer_rotation<i * G / I, (i + 1) * G / I, G, ROUNDS>(
                bars, n_iter, Tracer(lane == 0 && i == 0, 1, smem, 6 * (G / I) * ROUNDS), issue);
    });
}

// =====================================================================================================
// Forward: G_FWD tiles in flight per SM.
// Group tiles: tX [128 x 32] (enc, then [pe, emb, 1]), tH [128 x 64], tF [128 x 64] (its first 32 columns hold
// the specular input [sh16, pe12, 1] during round 1).  TMEM: 128 working columns per group.
// =====================================================================================================
// Groups in flight: the MMA issuer saturates with three (it needs ~400 cycles per group and round, a group's own
// chain is ~900), and 13 warps leave 128 registers per thread -- enough to fetch the NEXT tile's inputs two rounds
// ahead.  With four groups (17 warps, 96 registers) the tile's inputs were loaded at the top of the tile: every group
// stalled there together (the strict issue rotation keeps the groups in lock step), ~3.3 k of 13 k cycles per
// iteration.  The density pass (2 rounds, small epilogues) keeps four groups.
#ifndef UMHS_DENSITY_G
#define UMHS_DENSITY_G 4
#endif
#ifndef UMHS_DENSITY_PREFETCH
#define UMHS_DENSITY_PREFETCH 0
#endif
template <bool DENSITY_ONLY>
struct TcFwdCfg {
    static constexpr int G = DENSITY_ONLY ? UMHS_DENSITY_G : 3;
    static constexpr int I = UMHS_FWD_ISSUERS == 2 ? G : 1;  // MMA-issuing warps
    static constexpr int kThreads = G * TM + 32 * I;
};
struct TcFwd {
    static constexpr int tX = 0;
    static constexpr int tH = tX + TM * 32 * 2;
    static constexpr int tF = tH + TM * 64 * 2;
    static constexpr int grp_bytes = tF + TM * 64 * 2;     // 40 KB
    static constexpr int grp_bytes_density = tF;           // 24 KB
    static constexpr int kRounds = 5, kRoundsDensity = 2;
};

// KM: compile-time bound on the number of classes K (4, 8 or 16; the density kernel ignores it)
template <bool DENSITY_ONLY, int KM>
__global__ void __launch_bounds__(TcFwdCfg<DENSITY_ONLY>::kThreads, 1)
field_mlp_tc_fwd_kernel(umhs_points P, int64_t n_total, const float *__restrict__ enc, umhs_mlp_params mp,
                        umhs_field_out out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int G = TcFwdCfg<DENSITY_ONLY>::G;
    constexpr int ROUNDS = DENSITY_ONLY ? TcFwd::kRoundsDensity : TcFwd::kRounds;
    constexpr int GRP_BYTES = DENSITY_ONLY ? TcFwd::grp_bytes_density : TcFwd::grp_bytes;
    constexpr int WCOLS = DENSITY_ONLY ? 64 : 128;   // working TMEM columns per group
    constexpr int NCOLS = (G * WCOLS <= 256) ? 256 : 512;  // tcgen05.alloc takes a power of two
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = mp.num_classes, spec = DENSITY_ONLY ? 0 : mp.specular;
    const int grp = tid >> 7, row = tid & 127;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + TcSmem::ctrl);  // ready[G], done[G]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + TcSmem::ctrl + 64);

    stage_all_weights(smem, mp, DENSITY_ONLY, tid, TcFwdCfg<DENSITY_ONLY>::kThreads);
    if (tid == 0) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
            mbar_init(bars + g, 4);
            mbar_init(bars + G + g, 1);
            mbar_init(reinterpret_cast<uint64_t *>(smem + TcSmem::sbias) + g, 1);  // `full`: bulk-copied tiles
            ready_flags(smem)[g] = 0u;
        }
        fence_mbar_init();
    }
    if (warp == G * 4) tmem_alloc<NCOLS>(tmem_slot);
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tbase = *tmem_slot;
    n_total = effective_n(P, n_total);
    const int64_t n_tiles = (n_total + TM - 1) / TM;
    auto tile_of = [&](int64_t it, int q) { return (it * gridDim.x + blockIdx.x) * G + q; };

    if (grp >= G) {
        // ======================= MMA issuer warps (all lanes, warp-uniform)
        int64_t n_iter[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int64_t first = (int64_t)blockIdx.x * G + g, stride = (int64_t)gridDim.x * G;
            n_iter[g] = first < n_tiles ? (n_tiles - 1 - first) / stride + 1 : 0;
        }
        const uint32_t sbase = smem_u32(smem);
        issuer_dispatch<TcFwdCfg<DENSITY_ONLY>::I, G, ROUNDS>(warp - G * 4, bars, n_iter, smem, lane, [&](auto rc, auto gc, uint32_t, uint32_t zero, auto ph) {
            constexpr int r = decltype(rc)::value, g = decltype(gc)::value;
            if constexpr (decltype(ph)::value != 0) return;  // the forward has no deferred MMAs
            const uint32_t B16 = desc_base16(sbase) + zero;   // descriptors = B16 + compile-time constants
            using w = WOff;
            constexpr uint32_t b = (uint32_t)(TcSmem::groups + g * GRP_BYTES);
            constexpr uint32_t sX = b + TcFwd::tX, sH = b + TcFwd::tH, sF = b + TcFwd::tF;
            const uint32_t tw = tbase + (uint32_t)(WCOLS * g);
            if constexpr (r == 0) {
                gemm_kmajor(tw, B16, sX, w::b0, 64, 0, 64, 32, 0);
                gemm_bias_step(tw, B16, w::ones, w::b0, 64, 0, 64, COL_B_BASE0);
            } else if constexpr (r == 1) {
                gemm_kmajor(tw, B16, sH, w::b1, 16, 0, 16, 64, 0);
                gemm_bias_step(tw, B16, w::ones, w::b1, 16, 0, 16, COL_B_64);
                if (spec) gemm_kmajor(tw + 64, B16, sF, w::d0, 16, 0, 16, 32, 0);
            } else if constexpr (r == 2) {
                gemm_kmajor(tw, B16, sX, w::hf0, 128, 0, 128, 32, 0);
            } else if constexpr (r == 3) {
                gemm_kmajor(tw, B16, sH, w::h1, 64, 0, 64, 64, 0);
                gemm_kmajor(tw + 64, B16, sF, w::f1, 64, 0, 64, 64, 0);
                gemm_bias_step(tw, B16, w::ones, w::h1, 64, 0, 64, COL_B_64);
                gemm_bias_step(tw + 64, B16, w::ones, w::f1, 64, 0, 64, COL_B_64);
            } else {
                gemm_kmajor(tw, B16, sH, w::h2, 16, 0, 16, 64, 0);
                gemm_kmajor(tw + 64, B16, sF, w::f2, 32, 0, 32, 64, 0);
                gemm_bias_step(tw, B16, w::ones, w::h2, 16, 0, 16, COL_B_64);
                gemm_bias_step(tw + 64, B16, w::ones, w::f2, 32, 0, 32, COL_B_64);
            }
        });
        __syncwarp();
    } else {
        // ======================= data groups
        uint8_t *gbase = smem + TcSmem::groups + grp * GRP_BYTES;
        uint8_t *tX = gbase + TcFwd::tX, *tH = gbase + TcFwd::tH, *tF = gbase + TcFwd::tF;
        GroupSync gs{bars + grp, bars + G + grp, 0u, lane, Tracer(tid == 0, 0, smem, 2 * ROUNDS), ready_flags(smem) + grp, kReadyBarId0 + grp};
        const uint32_t trow = tbase + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(WCOLS * grp);
        const float invT = DENSITY_ONLY ? 1.f : 1.0f / mp.temperature;
        const bool vec_out = !DENSITY_ONLY && ((reinterpret_cast<uintptr_t>(out.scalar) |
                                                reinterpret_cast<uintptr_t>(out.abund)) & 15u) == 0;

        // Inputs of a tile (hash-encoding rows + sample geometry) sit behind a chain of dependent loads (ray index ->
        // origin / direction, enc_index -> row): the NEXT tile's are issued while the tensor core works on this one --
        // the index level behind round 0's / 1's arrive, the rows one round later -- and consumed at the next tile's top.
        float4 e_in[8];
        PointRaw raw_in;
        PointIdx pi_in;
        int64_t erow_in;
        auto sample_of = [&](int64_t tile) {
            const int64_t m = tile * TM + row;
            return m < n_total ? m : n_total - 1;  // tail threads shadow the last sample, write nothing
        };
        auto load_idx = [&](int64_t tile) {
            if (tile >= n_tiles) return;  // group-uniform
            const int64_t nn = sample_of(tile);
            pi_in = load_point_idx(P, nn);
            erow_in = enc_row(P, nn);
        };
        // Contiguous bf16 tiles (no enc_index gather: the sampler's candidate pass, the occupancy update, evaluation
        // rendering) do not go through registers at all: ONE bulk copy (cp.async.bulk, the TMA engine) per tile brings
        // the 8 KB operand tile from global memory straight into tX, a whole tile ahead, completing on an mbarrier.
        const bool save = !DENSITY_ONLY && out.act_tiles != nullptr;  // training: keep the activation tiles (ActTile)
        const bool bulk = mp.enc_tiles && P.enc_index == nullptr && !save;
        uint64_t *full = reinterpret_cast<uint64_t *>(smem + TcSmem::sbias) + grp;  // (sbias is a backward-only region)
        auto bulk_issue = [&](int64_t tile) {
            if (!bulk || row != 0 || tile >= n_tiles) return;
            mbar_expect_tx(full, 8192u);
            bulk_load(tX, reinterpret_cast<const uint8_t *>(enc) + tile * 8192, 8192u, full);
        };
        auto load_rest = [&](int64_t tile) {
            if (tile >= n_tiles) return;
            if (!bulk) {
                if (mp.enc_tiles) load_enc_pieces(enc, erow_in, e_in);
                else load_enc_rows(enc, erow_in, lane, e_in);
            }
            raw_in = load_point_raw(P, sample_of(tile), pi_in);
        };
        if (!(DENSITY_ONLY && !UMHS_DENSITY_PREFETCH)) {
            load_idx(tile_of(0, grp));
            load_rest(tile_of(0, grp));
        }
        bulk_issue(tile_of(0, grp));

        for (int64_t it = 0; tile_of(it, grp) < n_tiles; ++it) {
            gs.round = 0;
            const int64_t n0 = tile_of(it, grp) * TM;
            const bool valid = n0 + row < n_total;
            const int64_t n = valid ? n0 + row : n_total - 1;
            if (DENSITY_ONLY && !UMHS_DENSITY_PREFETCH) {
                load_idx(tile_of(it, grp));
                load_rest(tile_of(it, grp));
            }
            const Point pt = make_point(P, raw_in);
            float pe[12];
            // ---------------- round 0: enc -> tX (whole rows, coalesced) ; base0
            uint8_t *act = save ? reinterpret_cast<uint8_t *>(out.act_tiles) + tile_of(it, grp) * (int64_t)ActTile::bytes : nullptr;
            if (bulk) mbar_wait(full, (uint32_t)(it & 1));  // the copy engine has delivered this tile
            else if (mp.enc_tiles) store_enc_pieces(tX, row, e_in);
            else store_enc_rows(tX, (warp & 3) * 32, lane, e_in);
            if (!DENSITY_ONLY) pos_encoding_tc(pt, pe);
            gs.arrive();
            if (DENSITY_ONLY && UMHS_DENSITY_PREFETCH) load_idx(tile_of(it + 1, grp));
            gs.wait();
            if (DENSITY_ONLY) bulk_issue(tile_of(it + 1, grp));  // tX is free again: the next tile starts to arrive
            // ---------------- round 1: H0 = relu(.) -> tH ; (specular) [sh, pe, 1] -> tF[:, :32] ; base1 [, dir0]
            epi_relu64<false>(trow, tH, row, nullptr);
            if (spec) {
                float x2[32], sh[16];
                sh_deg3_tc((pt.dx + 1.0f) / 2.0f, (pt.dy + 1.0f) / 2.0f, (pt.dz + 1.0f) / 2.0f, sh);
#pragma unroll
                for (int i = 0; i < 16; ++i) x2[i] = sh[i];
#pragma unroll
                for (int i = 0; i < 12; ++i) x2[16 + i] = pe[i];
                x2[28] = 1.0f;
                x2[29] = x2[30] = x2[31] = 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c) store8(tF, row, 8 * c, CH_ACT, x2 + 8 * c);
            }
            gs.arrive();
            if (save) save_row(act, ActTile::H0, tH, 8, row);
            if (DENSITY_ONLY && UMHS_DENSITY_PREFETCH) load_rest(tile_of(it + 1, grp));
            else if (!DENSITY_ONLY) load_idx(tile_of(it + 1, grp));
            gs.wait();
            // ---------------- round 2: sigma, emb ; X = [pe, emb, 1] -> tX ; head0 | feat0 (N = 128)
            float o16[16];
            tmem_ld16(trow, o16);
            if (valid) {
                out.sigma[n] = expf(o16[0]) * pt.sel;  // trunc_exp forward == exp; density * selector
                if (!DENSITY_ONLY && out.embedding) {
#pragma unroll
                    for (int i = 0; i < 15; ++i) out.embedding[n * 15 + i] = o16[i + 1];
                }
            }
            if (DENSITY_ONLY) continue;  // the next round_trip orders these TMEM reads before the next tile's MMA
            if (spec) {
                float h[16];
                tmem_ld16(trow + 64, h);
                if (valid && out.spec_hidden) {
                    float4 *dst = reinterpret_cast<float4 *>(out.spec_hidden + n * 16);
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        dst[q] = make_float4(fmaxf(h[4 * q], 0.f), fmaxf(h[4 * q + 1], 0.f), fmaxf(h[4 * q + 2], 0.f),
                                             fmaxf(h[4 * q + 3], 0.f));
                }
            }
            {
                float x[32];
#pragma unroll
                for (int i = 0; i < 12; ++i) x[i] = pe[i];
#pragma unroll
                for (int i = 0; i < 15; ++i) x[12 + i] = o16[1 + i];
                x[27] = 1.0f;
                x[28] = x[29] = x[30] = x[31] = 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c) store8(tX, row, 8 * c, CH_ACT, x + 8 * c);
            }
            gs.arrive();
            if (save) save_row(act, ActTile::X, tX, 4, row);
            load_rest(tile_of(it + 1, grp));
            gs.wait();
            bulk_issue(tile_of(it + 1, grp));  // X in tX has been consumed: the next tile's encodings may land
            // ---------------- round 3: hidden 0 of both MLPs -> tH, tF ; head1, feat1
            epi_relu64<false>(trow, tH, row, nullptr);
            epi_relu64<false>(trow + 64, tF, row, nullptr);
            gs.arrive();
            if (save) {
                save_row(act, ActTile::HH1, tH, 8, row);
                save_row(act, ActTile::HF1, tF, 8, row);
            }
            gs.wait();
            // ---------------- round 4: hidden 1 -> tH, tF ; head2 (N = 16), feat2 (N = 32)
            epi_relu64<false>(trow, tH, row, nullptr);
            epi_relu64<false>(trow + 64, tF, row, nullptr);
            gs.arrive();
            if (save) {
                save_row(act, ActTile::HH2, tH, 8, row);
                save_row(act, ActTile::HF2, tF, 8, row);
            }
            gs.wait();
            // ---------------- outputs (KM bounds the class loops at compile time: with a run-time K every one of the
            // 16 guarded iterations still issued its ex2 / rcp -- 50 MUFU instructions per thread where K = 4 needs
            // 13, ~2.5 k of a tile's 14 k cycles with three groups sharing the SM's 16 MUFU lanes)
            {
                if (UMHS_TRACE > 1) gs.tr.stamp(300);
                float zh[16], zf[16], zf2[16];
                tmem_ld16(trow, zh);
                tmem_ld16(trow + 64, zf);
                if (K + spec > 16) tmem_ld16(trow + 80, zf2);
                if (valid) {
                    float mx = -3.0e38f;
#pragma unroll
                    for (int k = 0; k < KM; ++k)
                        if (k < K) mx = fmaxf(mx, zf[k] * invT);
                    float den = 0.f, e[KM], sg[KM];
#pragma unroll
                    for (int k = 0; k < KM; ++k) {
                        e[k] = (k < K) ? __expf(zf[k] * invT - mx) : 0.f;
                        den += e[k];
                        sg[k] = __frcp_rn(1.0f + __expf(-zh[k]));
                    }
                    const float rden = __frcp_rn(den);
                    if (KM == 4 && K == 4 && vec_out) {
                        *reinterpret_cast<float4 *>(out.scalar + n * 4) = make_float4(sg[0], sg[1], sg[2], sg[3]);
                        *reinterpret_cast<float4 *>(out.abund + n * 4) =
                            make_float4(e[0] * rden, e[1] * rden, e[2] * rden, e[3] * rden);
                    } else {
#pragma unroll
                        for (int k = 0; k < KM; ++k) {
                            if (k < K) {
                                out.scalar[n * K + k] = sg[k];
                                out.abund[n * K + k] = e[k] * rden;
                            }
                        }
                    }
                    if (spec && out.s1) {
                        float z = zf2[0];  // K = 16: the gate logit is column 16
#pragma unroll
                        for (int k = 0; k < 16; ++k)
                            if (k <= KM && k == K) z = zf[k];
                        out.s1[n] = __frcp_rn(1.0f + __expf(-z));
                    }
                }
                if (UMHS_TRACE > 1) gs.tr.stamp(301);
            }
        }
        gs.tr.flush();
    }
    fence_before_sync();
    __syncthreads();
    if (warp == G * 4) tmem_dealloc<NCOLS>(tbase);
}

// =====================================================================================================
// Backward on the tensor cores, two 128-sample tiles in flight per SM.
//
// Per tile the activations are recomputed from the hash encoding (bf16 operand tiles kept in shared memory),
// then for every layer one commit group issues
//     dgrad   W[128 x N]  = Delta[128 x K] * W[K x N]                           (weights read MN-major)
//     wgrad   P[64 x N] += Delta^T[64 x 128 samples] * Act[128 samples x N]     (MN-major operands, M = 64)
// The weight-gradient accumulators P_* live in TMEM for the whole kernel (accumulated over all tiles of the
// persistent CTA, by both groups: all MMAs come from one thread, in program order) and are flushed once at the end
// with one atomic per element per CTA.  Bias gradients of the wide layers are one more accumulator column: the
// activation operand of their wgrad carries a column of ones (a chunk of ones behind tA0 / tH1, column 27 / 28 of
// the 27/28-wide inputs); the narrow output layers use warp shuffles.
// Rounds per tile: base0 | base1 [+dir0] | head0 [+dir0 wgrad] | head1 | head2 | dL2 | dL1 | dL0 | feat0 | feat1 |
//                  feat2 | dL2 | dL1 | dL0 | dbase1 | dbase0.
// The base MLP's hidden activations stay in registers (packed bf16) between the first and the last rounds.
// =====================================================================================================
struct TcBwd {
    static constexpr int G = 2;
    // per-group activation tiles (a chunk of ones behind tA0 and tH1)
    static constexpr int tA0 = 0;                           // [128 x 32] + ones chunk
    static constexpr int tX = tA0 + TM * 32 * 2 + 2048;     // [128 x 32]  [pe, emb, 1]
    static constexpr int tH1 = tX + TM * 32 * 2;            // [128 x 64] + ones chunk
    static constexpr int tH2 = tH1 + TM * 64 * 2 + 2048;    // [128 x 64]
    static constexpr int tD1 = tH2 + TM * 64 * 2;           // [128 x 64]  (its first 32 columns: specular input, rounds 1-2)
    static constexpr int tD2 = tD1 + TM * 64 * 2;           // [128 x 64]
    static constexpr int grp_bytes = tD2 + TM * 64 * 2;     // 84 KB
    // TMEM columns: working [0,64) group 0, [64,128) group 1; accumulators shared by both groups
    static constexpr int P_h2t = 128, P_h1 = 144, P_h0 = 216, P_f2t = 248, P_f1 = 280, P_f0 = 352;
    static constexpr int P_b1t = 384, P_b0 = 400, P_d0 = 440;  // P_d0: [64 (16 real rows) x 32] .. 472
    static constexpr int ncols = 512;
    static constexpr int kRounds = 16;
    static constexpr int I = UMHS_BWD_ISSUERS;              // MMA-issuing warps: one per group, or one for both
    static constexpr int kThreads = G * TM + 32 * I;
    // with the hash-grid scatter on board: 4 warpgroups -- data, data, [issuer + 3 scatter warps], 4 scatter warps
    static constexpr int kScatterWarps = UMHS_SCATTER_WARPS;
    static constexpr int kThreadsScatter = kThreads + 32 * kScatterWarps;  // 512
    static constexpr int kRegsData = 168, kRegsOther = I + UMHS_SCATTER_WARPS == 8 ? 88 : 168;  // setmaxnreg: 256 * 168 + 256 * 88 = 512 * 128
    static constexpr int smem_bytes = TcSmem::groups + G * grp_bytes;
};
static_assert(TcBwd::kThreadsScatter % 128 == 0, "setmaxnreg works on whole warpgroups");
constexpr bool kScatterSetRegs = TcBwd::kRegsOther != TcBwd::kRegsData;
static_assert(2 * TM * TcBwd::kRegsData + (TcBwd::kThreadsScatter - 2 * TM) * TcBwd::kRegsOther <= 65536,
              "register pool of the CTA");

// what the scatter warps need of the hash grid (see umhs_field_mlp_tc_bwd_scatter)
struct ScatterArgs {
    const float *scalings;
    float2 *d_table;
    int L, log2T, agg_levels;
};
static_assert(TcBwd::smem_bytes <= 232448, "shared-memory plan exceeds 227 KB");

// Column sums of a [32 rows x KM] block held one row per lane (KM = 4, 8, 16), by a butterfly transpose-reduce
// (KM - 1 + log2(32 / KM) shuffles instead of 5 * KM): returns, in every lane, the sum over the warp of column
// colsum_col<KM>(lane).
template <int KM>
__device__ __forceinline__ int colsum_col(int lane) { return lane / (32 / KM); }
template <int KM>
__device__ __forceinline__ bool colsum_owner(int lane) { return (lane & (32 / KM - 1)) == 0; }
template <int KM>
__device__ __forceinline__ float warp_colsum(const float *v, int lane) {
    float a[KM];
#pragma unroll
    for (int i = 0; i < KM; ++i) a[i] = v[i];
    int off = 16;
#pragma unroll
    for (int width = KM; width > 1; width >>= 1, off >>= 1) {
        const bool hi = lane & off;  // lanes with this bit keep the upper half of the remaining columns
#pragma unroll
        for (int i = 0; i < width / 2; ++i) {
            const float keep = hi ? a[i + width / 2] : a[i], send = hi ? a[i] : a[i + width / 2];
            a[i] = keep + __shfl_xor_sync(kFull, send, off);
        }
    }
#pragma unroll
    for (; off > 0; off >>= 1) a[0] += __shfl_xor_sync(kFull, a[0], off);
    return a[0];
}

// flush one persistent [64 x ncols] accumulator: this thread's row r = 16*warp + lane (lane < 16)
// (both: lanes 16-31 hold group 1's copy of the same rows; it is added when that group ran at all)
template <class F>
__device__ __forceinline__ void flush_acc(uint32_t trow, int col0, int ncols, int lane, bool both, bool hi_valid, F &&sink) {
    for (int c = 0; c < ncols; c += 16) {
        float v[16];
        tmem_ld16(trow + (uint32_t)(col0 + c), v);  // .sync.aligned: every lane participates
        if (both) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float o = __shfl_down_sync(kFull, v[i], 16);
                if (hi_valid) v[i] += o;
            }
        }
        if (lane < 16) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (c + i < ncols) sink(c + i, v[i]);
        }
    }
}

// Flush of the persistent weight-gradient accumulators + narrow-layer bias sums (both backward kernels; every thread of
// the CTA calls it after the role branches): row r = 16 * (warp & 3) + lane (lane < 16) of every [64 x ncols] accumulator.
__device__ __forceinline__ void flush_weight_grads(uint32_t tbase, int tid, int grp, int G, int K, int spec,
                                                   const umhs_mlp_grads &dp, bool any_tile, bool two_copies,
                                                   bool hi_valid, const float *sbias) {
    const int warp = tid >> 5, lane = tid & 31;
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (grp < G && any_tile) {
        const uint32_t tq = tbase + ((uint32_t)((warp & 3) * 32) << 16);
        const int r = (warp & 3) * 16 + lane;
        if (grp == 0) {
            flush_acc(tq, TcBwd::P_h2t, K, lane, two_copies, hi_valid, [&](int c, float v) { if (v != 0.f) red_add(dp.head_w2 + c * 64 + r, v); });
            flush_acc(tq, TcBwd::P_h1, 65, lane, two_copies, hi_valid, [&](int c, float v) {
                if (v != 0.f) red_add(c < 64 ? dp.head_w1 + r * 64 + c : dp.head_b1 + r, v); });
            flush_acc(tq, TcBwd::P_h0, 28, lane, two_copies, hi_valid, [&](int c, float v) {
                if (v != 0.f) red_add(c < 27 ? dp.head_w0 + r * 27 + c : dp.head_b0 + r, v); });
            flush_acc(tq, TcBwd::P_b1t, 16, lane, two_copies, hi_valid, [&](int c, float v) { if (v != 0.f) red_add(dp.base_w1 + c * 64 + r, v); });
            if (spec)
                flush_acc(tq, TcBwd::P_d0, 29, lane, two_copies, hi_valid, [&](int c, float v) {
                    if (r < 16 && v != 0.f) red_add(c < 28 ? dp.dir_w0 + r * 28 + c : dp.dir_b0 + r, v); });
        } else {
            flush_acc(tq, TcBwd::P_f2t, K + spec, lane, two_copies, hi_valid, [&](int c, float v) { if (v != 0.f) red_add(dp.feat_w2 + c * 64 + r, v); });
            flush_acc(tq, TcBwd::P_f1, 65, lane, two_copies, hi_valid, [&](int c, float v) {
                if (v != 0.f) red_add(c < 64 ? dp.feat_w1 + r * 64 + c : dp.feat_b1 + r, v); });
            flush_acc(tq, TcBwd::P_f0, 28, lane, two_copies, hi_valid, [&](int c, float v) {
                if (v != 0.f) red_add(c < 27 ? dp.feat_w0 + r * 27 + c : dp.feat_b0 + r, v); });
            flush_acc(tq, TcBwd::P_b0, 33, lane, two_copies, hi_valid, [&](int c, float v) {
                if (v != 0.f) red_add(c < 32 ? dp.base_w0 + r * 32 + c : dp.base_b0 + r, v); });
        }
    }
    if (tid < 64) {
        const float a = sbias[tid];
        float *dst = nullptr;
        if (tid < 16) dst = (tid < K) ? dp.head_b2 + tid : nullptr;
        else if (tid < 48) dst = (tid - 16 < K + spec) ? dp.feat_b2 + (tid - 16) : nullptr;
        else dst = dp.base_b1 + (tid - 48);
        if (dst != nullptr && a != 0.f) red_add(dst, a);
    }
    fence_before_sync();
    __syncthreads();
}

// KM: compile-time bound on the number of classes K (4, 8 or 16): sizes the softmax / sigmoid epilogues
//
// SCATTER (opt-in, measured slower than the two kernels back to back: 0.66 against 0.63 ms at 819 k samples): the
// hash-grid scatter of d_enc (hash_scatter.cuh; otherwise the separate kernel of hash_grid.cu) runs INSIDE this kernel,
// on 7 more warps per CTA.  On paper the two kernels need different halves of an SM -- this one is a chain of
// dependent tensor-core rounds that leaves ~80 % of the issue slots idle, the scatter is bound by the reduction path
// (40 M red sectors) -- but the reductions and the shared-memory stores of the epilogues share the SM's load/store
// unit: with the reductions switched off the scatter warps cost 0.06 ms, with them 0.30 ms (the whole stand-alone
// scatter), i.e. nothing overlaps.  With 3 scatter warps (no setmaxnreg) 1.06 ms: a warp waits ~250 cycles per
// reduction it issues, so few warps cannot keep the reduction path busy either.  A data warp publishes "tile t of my 32 rows is in d_enc" (fence + a counter in shared memory) after its
// deferred row flush; scatter warps take (tile, quarter) units in that order from a shared ticket, wait for the counter,
// read the rows back (L2) and issue the reductions.  d_enc must be the bf16 tile buffer.  Registers: the CTA is
// launched with 512 x 128; the data warpgroups raise themselves to 168, the other two drop to 88 (setmaxnreg).
template <int KM, bool SCATTER>
__global__ void __launch_bounds__(SCATTER ? TcBwd::kThreadsScatter : TcBwd::kThreads, 1)  // 9 warps: 3 on one SM sub-partition (16K registers) -> 168
field_mlp_tc_bwd_kernel(umhs_points P, int64_t n_total, const float *__restrict__ enc, umhs_mlp_params mp,
                        umhs_field_out_grads g, float *__restrict__ d_enc, umhs_mlp_grads dp, ScatterArgs sc) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int G = TcBwd::G;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = mp.num_classes, spec = mp.specular;
    const int grp = warp < G * 4 ? (tid >> 7) : G, row = tid & 127;  // grp == G: not a data thread
    uint32_t *scat = reinterpret_cast<uint32_t *>(smem + TcSmem::scat);  // flushed[G * 4], next_unit
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + TcSmem::ctrl);  // ready[G], done[G]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + TcSmem::ctrl + 64);
    float *sbias = reinterpret_cast<float *>(smem + TcSmem::sbias);      // h2[16] f2[17 of 32]... see below
    uint8_t *gbase = smem + TcSmem::groups + (grp < G ? grp : 0) * TcBwd::grp_bytes;

    // ---- one-time setup: weights, ones chunks, barriers, TMEM
    stage_all_weights(smem, mp, false, tid, (int)blockDim.x);
    if (tid < 16) scat[tid] = 0u;
    if (tid < G * TM) {
        const float one[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
        store8(gbase + TcBwd::tA0, row, 32, CH_ACT, one);
        store8(gbase + TcBwd::tH1, row, 64, CH_ACT, one);
    }
    if (tid < 64) sbias[tid] = 0.f;
    if (tid == 0) {
#pragma unroll
        for (int q = 0; q < G; ++q) {
            mbar_init(bars + q, 4);
            mbar_init(bars + G + q, 1);
            ready_flags(smem)[q] = 0u;
        }
        fence_mbar_init();
    }
    if (warp == G * 4) tmem_alloc<TcBwd::ncols>(tmem_slot);
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tbase = *tmem_slot;
    n_total = effective_n(P, n_total);
    const int64_t n_tiles = (n_total + TM - 1) / TM;
    auto tile_of = [&](int64_t it, int q) { return (it * gridDim.x + blockIdx.x) * G + q; };

    int64_t n_iter[G];
#pragma unroll
    for (int q = 0; q < G; ++q) {
        const int64_t first = (int64_t)blockIdx.x * G + q, stride = (int64_t)gridDim.x * G;
        n_iter[q] = first < n_tiles ? (n_tiles - 1 - first) / stride + 1 : 0;
    }
    // (setmaxnreg sits at the top of the two role branches: the register budget of a region is what dominates it)
    if (grp == G) {
      if (SCATTER && kScatterSetRegs) setmaxnreg_dec<TcBwd::kRegsOther>();
      if (SCATTER && warp >= G * 4 + TcBwd::I) {
        // ======================= scatter warps: units (iteration, group, quarter) in the order the data warps publish
        const uint32_t total = (uint32_t)(n_iter[0] * (G * 4));  // group 0 never has fewer tiles than the others
        const uint8_t *gt = reinterpret_cast<const uint8_t *>(d_enc);
        for (;;) {
            uint32_t u = 0;
            if (lane == 0) u = atomicAdd(scat + G * 4, 1u);
            u = __shfl_sync(kFull, u, 0);
            if (u >= total) break;
            const int64_t it = u / (G * 4);
            const int q = (u >> 2) % G, quarter = u & 3;
            if (tile_of(it, q) >= n_tiles) continue;
            while (ld_acquire_shared(scat + q * 4 + quarter) <= (uint32_t)it) __nanosleep(64);
            const int64_t n = tile_of(it, q) * TM + quarter * 32 + lane;
            const bool valid = n < n_total;
            float x = 0.f, y = 0.f, z = 0.f;
            if (valid) {
                const Point pt = eval_point(P, n);
                x = pt.x;
                y = pt.y;
                z = pt.z;
            }
            scatter_sample_runs(x, y, z, valid, n, gt, sc.scalings, sc.log2T, sc.d_table, sc.agg_levels, lane);
        }
      } else {
        // ======================= MMA issuer warp (all lanes, warp-uniform)
        const uint32_t sbase = smem_u32(smem);
        const bool want_denc = d_enc != nullptr;
        issuer_dispatch<TcBwd::I, G, TcBwd::kRounds>(warp - G * 4, bars, n_iter, smem, lane, [&](auto rc, auto gc, uint32_t acc,
                                                                                  uint32_t zero, auto ph) {
            constexpr int r = decltype(rc)::value, q = decltype(gc)::value;
            constexpr bool crit = decltype(ph)::value == 0, late = !crit;  // before / behind the round's commit
            const uint32_t B16 = desc_base16(sbase) + zero;   // descriptors = B16 + compile-time constants
            using w = WOff;
            constexpr uint32_t b = (uint32_t)(TcSmem::groups + q * TcBwd::grp_bytes);
            constexpr uint32_t sA0 = b + TcBwd::tA0, sX = b + TcBwd::tX, sH1 = b + TcBwd::tH1, sH2 = b + TcBwd::tH2;
            constexpr uint32_t sD1 = b + TcBwd::tD1, sD2 = b + TcBwd::tD2, sX2 = sD1;
            const uint32_t tw = tbase + 64u * q;
            // persistent weight-gradient accumulators ([64 x N], M = 64: TMEM lanes 0-15 of every quarter): with one
            // issuer per group each group owns a copy -- group 1's in lanes 16-31 of the same columns
            const uint32_t ta = tbase + ((TcBwd::I == 2 && q == 1) ? (16u << 16) : 0u);
            if constexpr (r == 0) {         // base layer 0
                if (crit) {
                    gemm_kmajor(tw, B16, sA0, w::b0, 64, 0, 64, 32, 0);
                    gemm_bias_step(tw, B16, w::ones, w::b0, 64, 0, 64, COL_B_BASE0);
                }
            } else if constexpr (r == 1) {  // base layer 1 (N=16) [+ mlp_directional layer 0 at working columns 16..31]
                if (crit) {
                    gemm_kmajor(tw, B16, sH1, w::b1, 16, 0, 16, 64, 0);
                    gemm_bias_step(tw, B16, w::ones, w::b1, 16, 0, 16, COL_B_64);
                    if (spec) gemm_kmajor(tw + 16, B16, sX2, w::d0, 16, 0, 16, 32, 0);
                }
            } else if constexpr (r == 2) {  // mlp_head layer 0 [+ wgrad of mlp_directional layer 0]
                if (crit) gemm_kmajor(tw, B16, sX, w::hf0, 128, 0, 64, 32, 0);
                if (late && spec) gemm_wgrad(ta + TcBwd::P_d0, B16, sD2, 0, sX2, 0, 32, acc);
            } else if constexpr (r == 3) {
                if (crit) {
                    gemm_kmajor(tw, B16, sH1, w::h1, 64, 0, 64, 64, 0);
                    gemm_bias_step(tw, B16, w::ones, w::h1, 64, 0, 64, COL_B_64);
                }
            } else if constexpr (r == 4) {
                if (crit) {
                    gemm_kmajor(tw, B16, sH2, w::h2, 16, 0, 16, 64, 0);
                    gemm_bias_step(tw, B16, w::ones, w::h2, 16, 0, 16, COL_B_64);
                }
            } else if constexpr (r == 5) {
                if (crit) gemm_dgrad(tw, B16, sD2, w::h2, 16, 0, 64, 16, 0);
                if (late) gemm_wgrad(ta + TcBwd::P_h2t, B16, sH2, 0, sD2, 0, 16, acc);
            } else if constexpr (r == 6) {
                if (crit) gemm_dgrad(tw, B16, sD1, w::h1, 64, 0, 64, 64, 0);
                if (late) gemm_wgrad(ta + TcBwd::P_h1, B16, sD1, 0, sH1, 0, 72, acc);  // 64 + ones
            } else if constexpr (r == 7) {
                if (crit) gemm_dgrad(tw, B16, sD2, w::hf0, 128, 0, 32, 64, 0);
                if (late) gemm_wgrad(ta + TcBwd::P_h0, B16, sD2, 0, sX, 0, 32, acc);   // column 27 = ones
            } else if constexpr (r == 8) {  // feature_mlp layer 0
                if (crit) gemm_kmajor(tw, B16, sX, w::hf0, 128, 64, 64, 32, 0);
            } else if constexpr (r == 9) {
                if (crit) {
                    gemm_kmajor(tw, B16, sH1, w::f1, 64, 0, 64, 64, 0);
                    gemm_bias_step(tw, B16, w::ones, w::f1, 64, 0, 64, COL_B_64);
                }
            } else if constexpr (r == 10) {
                if (crit) {
                    gemm_kmajor(tw, B16, sH2, w::f2, 32, 0, 32, 64, 0);
                    gemm_bias_step(tw, B16, w::ones, w::f2, 32, 0, 32, COL_B_64);
                }
            } else if constexpr (r == 11) {
                if (crit) gemm_dgrad(tw, B16, sD2, w::f2, 32, 0, 64, 32, 0);
                if (late) gemm_wgrad(ta + TcBwd::P_f2t, B16, sH2, 0, sD2, 0, 32, acc);
            } else if constexpr (r == 12) {
                if (crit) gemm_dgrad(tw, B16, sD1, w::f1, 64, 0, 64, 64, 0);
                if (late) gemm_wgrad(ta + TcBwd::P_f1, B16, sD1, 0, sH1, 0, 72, acc);
            } else if constexpr (r == 13) {
                if (crit) gemm_dgrad(tw, B16, sD2, w::hf0, 128, 64, 32, 64, 0);
                if (late) gemm_wgrad(ta + TcBwd::P_f0, B16, sD2, 0, sX, 0, 32, acc);
            } else if constexpr (r == 14) {  // base layer 1 backward (H0 is back in tH1; d [o0, emb] sits in tD1)
                if (crit) gemm_dgrad(tw, B16, sD1, w::b1, 16, 0, 64, 16, 0);
                if (late) gemm_wgrad(ta + TcBwd::P_b1t, B16, sH1, 0, sD1, 0, 16, acc);
            } else {                         // 15: base layer 0 backward (wgrad N = 32 + ones chunk), delta in tD2
                // its weight gradient reads tA0, which the NEXT tile's first epilogue overwrites: not deferred
                if (crit) {
                    if (want_denc) gemm_dgrad(tw, B16, sD2, w::b0, 64, 0, 32, 64, 0);
                    gemm_wgrad(ta + TcBwd::P_b0, B16, sD2, 0, sA0, 0, 40, acc);
                }
            }
        });
        __syncwarp();
      }
    } else {
        // ======================= data groups: thread `row` of group `grp` owns one sample of the group's tile
        if (SCATTER && kScatterSetRegs) setmaxnreg_inc<TcBwd::kRegsData>();
        uint8_t *tA0 = gbase + TcBwd::tA0, *tX = gbase + TcBwd::tX, *tH1 = gbase + TcBwd::tH1;
        uint8_t *tH2 = gbase + TcBwd::tH2, *tD1 = gbase + TcBwd::tD1, *tD2 = gbase + TcBwd::tD2;
        uint8_t *tX2 = tD1;
        GroupSync gs{bars + grp, bars + G + grp, 0u, lane, Tracer(tid == 0, 0, smem, 2 * TcBwd::kRounds), ready_flags(smem) + grp, kReadyBarId0 + grp};
        const uint32_t trow = tbase + ((uint32_t)((warp & 3) * 32) << 16) + 64u * grp;
        float *sb_h2 = sbias, *sb_f2 = sbias + 16, *sb_b1 = sbias + 48;   // [16], [32], [16]
        const float invT = 1.0f / mp.temperature;
        // narrow-layer bias gradients: per-lane running sums of column colsum_col<KM>(lane) (see warp_colsum)
        float bs_h2 = 0.f, bs_f2 = 0.f, bs_f2s = 0.f, bs_b1 = 0.f;
        // Inputs of a tile (hash encoding row + sample geometry).  They sit behind a chain of dependent loads
        // (ray index -> origin / direction, enc_index -> encoding row), so the NEXT tile's are fetched during the
        // feature_mlp backward rounds of the current one (a DRAM latency is ~2k cycles, about two rounds).
        float4 e_in[8];
        PointRaw raw_in;   // raw loads only: the arithmetic (contraction, selector) waits for the values -> next tile's top
        PointIdx pi_in;
        int64_t erow_in;
        int64_t pending_row = -1;  // d_enc row of the PREVIOUS tile, staged in shared memory, not yet stored
        auto sample_of = [&](int64_t tile) {
            const int64_t m = tile * TM + row;
            return m < n_total ? m : n_total - 1;
        };
        auto load_idx = [&](int64_t tile) {
            if (tile >= n_tiles) return;
            const int64_t nn = sample_of(tile);
            pi_in = load_point_idx(P, nn);
            erow_in = enc_row(P, nn);
        };
        auto load_rest = [&](int64_t tile) {
            if (tile >= n_tiles) return;  // group-uniform
            if (mp.enc_tiles) load_enc_pieces(enc, erow_in, e_in);
            else load_enc_rows(enc, erow_in, lane, e_in);
            raw_in = load_point_raw(P, sample_of(tile), pi_in);
        };
        load_idx(tile_of(0, grp));
        load_rest(tile_of(0, grp));

        for (int64_t it = 0; tile_of(it, grp) < n_tiles; ++it) {
            gs.round = 0;
            const int64_t n0 = tile_of(it, grp) * TM;
            const bool valid = n0 + row < n_total;
            const int64_t n = valid ? n0 + row : n_total - 1;
            // Invalid (tail) rows shadow the last sample and run the whole chain with ZERO incoming gradients: every
            // delta they produce is exactly zero, so they add nothing to the weight gradients and write nothing.
            const Point pt = make_point(P, raw_in);
            float pe[12];
            pos_encoding_tc(pt, pe);
            // ---------------- round 0: enc -> tA0 [, X2 -> tX2]
            if (mp.enc_tiles) store_enc_pieces(tA0, row, e_in);
            else store_enc_rows(tA0, (warp & 3) * 32, lane, e_in);
            if (spec) {
                float x2[32], sh[16];
                sh_deg3_tc((pt.dx + 1.0f) / 2.0f, (pt.dy + 1.0f) / 2.0f, (pt.dz + 1.0f) / 2.0f, sh);
#pragma unroll
                for (int i = 0; i < 16; ++i) x2[i] = sh[i];
#pragma unroll
                for (int i = 0; i < 12; ++i) x2[16 + i] = pe[i];
                x2[28] = 1.0f;
                x2[29] = x2[30] = x2[31] = 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c) store8(tX2, row, 8 * c, CH_ACT, x2 + 8 * c);
            }
            gs.arrive();
            // the previous tile's d_enc rows (staged in this warp's 4 KB of tD2, which nothing touches before round 2)
            if (d_enc && it > 0) {
                if (mp.enc_tiles) flush_staged_pieces(d_enc, pending_row, tD2, row);
                else flush_staged_rows(d_enc, pending_row, tD2 + (warp & 3) * 4096, lane);
                if (SCATTER) publish_rows(scat + grp * 4 + (warp & 3), (uint32_t)it, lane);
            }
            gs.wait();
            // ---------------- round 1: H0 = relu(.) -> tH1, and kept in registers as packed bf16
            uint32_t h0p[32];
            epi_relu64<true>(trow, tH1, row, h0p);
            gs.round_trip();
            // ---------------- round 2: o16 ; X = [pe, emb, 1] -> tX [; d spec_hidden (padded to 64 columns) -> tD2]
            float o16[16], demb[15];
            tmem_ld16(trow, o16);
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
                    dsh[4 * q] = (h[4 * q] > 0.f) ? gq.x : 0.f;
                    dsh[4 * q + 1] = (h[4 * q + 1] > 0.f) ? gq.y : 0.f;
                    dsh[4 * q + 2] = (h[4 * q + 2] > 0.f) ? gq.z : 0.f;
                    dsh[4 * q + 3] = (h[4 * q + 3] > 0.f) ? gq.w : 0.f;
                }
                const float z8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                store8(tD2, row, 0, CH_ACT, dsh);
                store8(tD2, row, 8, CH_ACT, dsh + 8);
#pragma unroll
                for (int c = 16; c < 64; c += 8) store8(tD2, row, c, CH_ACT, z8);
            }
            {
                float x[32];
#pragma unroll
                for (int i = 0; i < 12; ++i) x[i] = pe[i];
#pragma unroll
                for (int i = 0; i < 15; ++i) x[12 + i] = o16[1 + i];
                x[27] = 1.0f;
                x[28] = x[29] = x[30] = x[31] = 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c) store8(tX, row, 8 * c, CH_ACT, x + 8 * c);
            }
            gs.round_trip();
            // ================= mlp_head (m = 0) then feature_mlp (m = 1): rounds 3..7 and 8..13
            float dsg = 0.f;  // d loss / d sigma of this sample (fetched with the second MLP's gradients)
#pragma unroll 1
            for (int m = 0; m < 2; ++m) {
                const int nout = m ? K + spec : K;
                // incoming gradients of this MLP's outputs (and, with the second MLP, of the density): issued three
                // rounds before their use -- they come from DRAM
                float gin[KM + 1];
#pragma unroll
                for (int k = 0; k < KM; ++k)
                    gin[k] = (valid && k < K) ? __ldg((m ? g.d_abund : g.d_scalar) + n * K + k) : 0.f;
                gin[KM] = (valid && m == 1 && spec && g.d_s1) ? __ldg(g.d_s1 + n) : 0.f;
                if (m == 1) dsg = (valid && g.d_sigma) ? __ldg(g.d_sigma + n) : 0.f;
                if (m == 1) {
                    // collect d X of mlp_head (working columns 12..26), then run feature_mlp layer 0 (round 8)
                    float dx[32];
                    tmem_ld32(trow, dx);
#pragma unroll
                    for (int i = 0; i < 15; ++i) demb[i] += dx[12 + i];
                    gs.round_trip();
                }
                epi_relu64<false>(trow, tH1, row, nullptr);
                gs.round_trip();  // layer 1
                epi_relu64<false>(trow, tH2, row, nullptr);
                gs.round_trip();  // layer 2
                {
                    // z: logits 0..KM-1 and (feature_mlp with specular) the gate logit, which sits at column K
                    float z[KM + 1], dz[32];
                    {
                        float t16[16];
                        tmem_ld16(trow, t16);
#pragma unroll
                        for (int k = 0; k < KM; ++k) z[k] = t16[k];
                        if (KM == 16) {
                            if (m == 1) tmem_ld16(trow + 16, t16);  // warp-uniform
                            z[KM] = t16[0];
                        } else {
                            z[KM] = t16[KM];
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 32; ++k) dz[k] = 0.f;
                    if (m == 0) {
#pragma unroll
                        for (int k = 0; k < KM; ++k) {
                            if (k < K) {
                                const float s = __frcp_rn(1.0f + __expf(-z[k]));
                                dz[k] = gin[k] * s * (1.0f - s);
                            }
                        }
                    } else {
                        float zs = z[KM];  // raw gate logit when K == KM
#pragma unroll
                        for (int k = 0; k < KM; ++k)
                            if (k == K) zs = z[k];
                        float mx = -3.0e38f, e[KM], den = 0.f, dot = 0.f;
#pragma unroll
                        for (int k = 0; k < KM; ++k) {
                            z[k] *= invT;
                            if (k < K) mx = fmaxf(mx, z[k]);
                        }
#pragma unroll
                        for (int k = 0; k < KM; ++k) {
                            e[k] = (k < K) ? __expf(z[k] - mx) : 0.f;
                            den += e[k];
                            dot += e[k] * gin[k];
                        }
                        const float rden = __frcp_rn(den);
                        dot *= rden;
#pragma unroll
                        for (int k = 0; k < KM; ++k) dz[k] = e[k] * rden * invT * (gin[k] - dot);
                        if (spec) {
                            const float s = __frcp_rn(1.0f + __expf(-zs));
                            const float ds = gin[KM] * s * (1.0f - s);
#pragma unroll
                            for (int k = 0; k <= KM; ++k)
                                if (k == K) dz[k] = ds;
                        }
                    }
                    store8(tD2, row, 0, CH_ACT, dz);
                    store8(tD2, row, 8, CH_ACT, dz + 8);
                    if (m == 1) {
                        store8(tD2, row, 16, CH_ACT, dz + 16);
                        store8(tD2, row, 24, CH_ACT, dz + 24);
                    }
                    if (m == 0) bs_h2 += warp_colsum<KM>(dz, lane);
                    else {
                        bs_f2 += warp_colsum<KM>(dz, lane);
                        if (nout > KM) bs_f2s += warp_sum(dz[KM]);
                    }
                }
                if (m == 1) load_idx(tile_of(it + 1, grp));   // next tile: index level of the dependent loads ...
                gs.round_trip();  // d layer 2
                epi_mask64(trow, tH2, nullptr, tD1, row);
                gs.round_trip();  // d layer 1
                if (m == 1) load_rest(tile_of(it + 1, grp));  // ... then the rows: a DRAM latency ahead of their use
                epi_mask64(trow, tH1, nullptr, tD2, row);
                gs.round_trip();  // d layer 0
            }
            // ================= base MLP backward (rounds 14, 15)
            {
                float dx[32], dO[16];
                tmem_ld32(trow, dx);  // d X (feature_mlp)
#pragma unroll
                for (int i = 0; i < 15; ++i) demb[i] += dx[12 + i];
                // density = trunc_exp(o0) * sel  ->  d o0 = d_sigma * sel * exp(clamp(o0, -15, 15))
                dO[0] = dsg * pt.sel * __expf(fminf(fmaxf(o16[0], -15.0f), 15.0f));
#pragma unroll
                for (int i = 0; i < 15; ++i) dO[1 + i] = demb[i];
                // (tD1, not tD2: the deferred weight-gradient MMAs of round 13 may still be reading tD2)
                store8(tD1, row, 0, CH_ACT, dO);
                store8(tD1, row, 8, CH_ACT, dO + 8);
                bs_b1 += warp_colsum<16>(dO, lane);
                // H0 back into tH1 (operand of the base layer-1 weight gradient)
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    *reinterpret_cast<uint4 *>(tH1 + tile_off(row, 8 * c, CH_ACT)) =
                        make_uint4(h0p[4 * c], h0p[4 * c + 1], h0p[4 * c + 2], h0p[4 * c + 3]);
            }
            gs.round_trip();
            epi_mask64(trow, nullptr, h0p, tD2, row);
            gs.round_trip();
            if (d_enc) {
                // whole rows, coalesced, through this warp's 4 KB of the (now idle) tD2 tile; the global stores go out
                // behind the next tile's first hand-off (or after the loop)
                float v[32];
                tmem_ld32(trow, v);
                if (mp.enc_tiles) stage_pieces(v, tD2, row);
                else stage_rows(v, tD2 + (warp & 3) * 4096, lane);
                pending_row = valid ? n : -1;
            }
        }
        if (d_enc && tile_of(0, grp) < n_tiles) {
            if (mp.enc_tiles) flush_staged_pieces(d_enc, pending_row, tD2, row);
            else flush_staged_rows(d_enc, pending_row, tD2 + (warp & 3) * 4096, lane);
            if (SCATTER) publish_rows(scat + grp * 4 + (warp & 3), (uint32_t)n_iter[grp], lane);
        }
        gs.tr.flush();
        if (colsum_owner<KM>(lane)) {  // one lane of each run owns column colsum_col<KM>(lane)
            const int col = colsum_col<KM>(lane);
            if (bs_h2 != 0.f) atomicAdd(sb_h2 + col, bs_h2);
            if (bs_f2 != 0.f) atomicAdd(sb_f2 + col, bs_f2);
        }
        if (colsum_owner<16>(lane) && bs_b1 != 0.f) atomicAdd(sb_b1 + colsum_col<16>(lane), bs_b1);
        if (lane == 0 && bs_f2s != 0.f) atomicAdd(sb_f2 + KM, bs_f2s);
    }

    // ================= flush the weight-gradient accumulators
    flush_weight_grads(tbase, tid, grp, G, K, spec, dp, tile_of(0, 0) < n_tiles, TcBwd::I == 2, tile_of(0, 1) < n_tiles, sbias);
    if (warp == G * 4) tmem_dealloc<TcBwd::ncols>(tbase);
}

// =====================================================================================================
// Backward from SAVED activations (umhs_field_mlp_tc_bwd_saved): 8 rounds per tile instead of 16.
//
// Half of the rounds of the kernel above recompute the forward (base0, base1, head0-2, feat0-2).  The training forward
// can keep its bf16 operand tiles instead (umhs_field_out.act_tiles, 704 B per sample, streamed out by its threads
// behind each round's hand-off: save_row); here they come back by ONE bulk copy per tile and layer (TMA, mbarrier
// transaction count) into two 16 KB slots, each two rounds ahead of the MMAs that read it, and the output-layer
// deltas are formed from the forward's outputs (sigmoid' = s (1 - s), softmax Jacobian from the abundances,
// trunc_exp' = the density itself).  Same MMAs, same accumulators, same flush as the kernel above.
//   round  critical (dgrad)                      deferred (wgrad)                operands arriving for
//   1      dz_h  [128x16] W_h2  -> d h2 (head)   --                              slot A: Hh2   slot B: Hh1   tX: X
//   2      d2_h  W_h1 ; P_h2t += Hh2^T dz_h -> d h1   P_h1 += d2_h^T [Hh1 | 1]
//   3      d1_h  W_h0           -> d X (head)    P_h0  += d1_h^T X               slot A <- Hf2
//   4      dz_f  [128x32] W_f2  -> d h2 (feat)   P_f2t += Hf2^T dz_f             slot B <- Hf1
//   5      d2_f  W_f1           -> d h1          P_f1  += d2_f^T [Hf1 | 1]
//   6      d1_f  W_f0           -> d X (feat)    P_f0  += d1_f^T X               slot A <- H0
//   7      dO    [128x16] W_b1  -> d h0          P_b1t += H0^T dO                slot B[8K:] <- enc rows (threads)
//   8      d0    W_b0           -> d enc         P_b0  += d0^T [enc | 1] (not deferred)   tX <- next X
// A slot is refilled as soon as the commit that covers its last reader has arrived (deferred MMAs of round r are
// covered by the commit of round r + 1).
// =====================================================================================================
struct TcBwd2 {
    static constexpr int G = 2;
    static constexpr int sA = 0;                              // [128 x 64]
    static constexpr int sB = sA + TM * 64 * 2;               // [128 x 64] + ones chunk
    static constexpr int tX = sB + TM * 64 * 2 + 2048;        // [128 x 32]
    static constexpr int tD1 = tX + TM * 32 * 2;              // [128 x 64]
    static constexpr int tD2 = tD1 + TM * 64 * 2;             // [128 x 64]
    static constexpr int tZ = tD2 + TM * 64 * 2;              // [128 x 32] output-layer deltas
    static constexpr int grp_bytes = tZ + TM * 32 * 2;        // 82 KB
    static constexpr int kRounds = 8;
    static constexpr int kThreads = G * TM + 32;
    static constexpr int smem_bytes = TcSmem::groups + G * grp_bytes;
};
static_assert(TcBwd2::smem_bytes <= 232448, "shared-memory plan exceeds 227 KB");

struct SavedFwd {  // what the training forward left behind (all read-only here)
    const uint8_t *act_tiles;
    const float *sigma, *scalar, *abund, *s1, *spec_hidden;
};

template <int KM>
__global__ void __launch_bounds__(TcBwd2::kThreads, 1)
field_mlp_tc_bwd_saved_kernel(umhs_points P, int64_t n_total, const float *__restrict__ enc, umhs_mlp_params mp,
                              umhs_field_out_grads g, SavedFwd sv, float *__restrict__ d_enc, umhs_mlp_grads dp) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int G = TcBwd2::G;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = mp.num_classes, spec = mp.specular;
    const int grp = warp < G * 4 ? (tid >> 7) : G, row = tid & 127;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + TcSmem::ctrl);  // ready[G], done[G]
    uint64_t *fulls = reinterpret_cast<uint64_t *>(smem + TcSmem::scat);  // [G][3]: slot A, slot B, tX filled
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + TcSmem::ctrl + 64);
    float *sbias = reinterpret_cast<float *>(smem + TcSmem::sbias);
    uint8_t *gbase = smem + TcSmem::groups + (grp < G ? grp : 0) * TcBwd2::grp_bytes;

    stage_all_weights(smem, mp, false, tid, (int)blockDim.x);
    if (tid < G * TM) {
        const float one[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
        store8(gbase + TcBwd2::sB, row, 64, CH_ACT, one);  // the ones chunk behind slot B
    }
    if (tid < 64) sbias[tid] = 0.f;
    if (tid == 0) {
#pragma unroll
        for (int q = 0; q < G; ++q) {
            mbar_init(bars + q, 4);
            mbar_init(bars + G + q, 1);
            for (int i = 0; i < 3; ++i) mbar_init(fulls + 3 * q + i, 1);
            ready_flags(smem)[q] = 0u;
        }
        fence_mbar_init();
    }
    if (warp == G * 4) tmem_alloc<TcBwd::ncols>(tmem_slot);
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tbase = *tmem_slot;
    n_total = effective_n(P, n_total);
    const int64_t n_tiles = (n_total + TM - 1) / TM;
    auto tile_of = [&](int64_t it, int q) { return (it * gridDim.x + blockIdx.x) * G + q; };
    int64_t n_iter[G];
#pragma unroll
    for (int q = 0; q < G; ++q) {
        const int64_t first = (int64_t)blockIdx.x * G + q, stride = (int64_t)gridDim.x * G;
        n_iter[q] = first < n_tiles ? (n_tiles - 1 - first) / stride + 1 : 0;
    }

    if (grp == G) {
        // ======================= MMA issuer warp
        const uint32_t sbase = smem_u32(smem);
        const bool want_denc = d_enc != nullptr;
        issuer_dispatch<1, G, TcBwd2::kRounds>(0, bars, n_iter, smem, lane, [&](auto rc, auto gc, uint32_t acc,
                                                                                uint32_t zero, auto ph) {
            constexpr int r = decltype(rc)::value, q = decltype(gc)::value;
            constexpr bool crit = decltype(ph)::value == 0, late = !crit;
            const uint32_t B16 = desc_base16(sbase) + zero;
            using w = WOff;
            constexpr uint32_t b = (uint32_t)(TcSmem::groups + q * TcBwd2::grp_bytes);
            constexpr uint32_t sA = b + TcBwd2::sA, sB = b + TcBwd2::sB, sX = b + TcBwd2::tX;
            constexpr uint32_t sD1 = b + TcBwd2::tD1, sD2 = b + TcBwd2::tD2, sZ = b + TcBwd2::tZ;
            const uint32_t tw = tbase + 64u * q;
            if constexpr (r == 0) {
                if (crit) {
                    gemm_dgrad(tw, B16, sZ, w::h2, 16, 0, 64, 16, 0);
                    // mlp_directional layer 0: d h (padded to 64 columns) sits in tD1, its input [sh, pe, 1] in tD2;
                    // not deferred: the next epilogue writes tD1
                    if (spec) gemm_wgrad(tbase + TcBwd::P_d0, B16, sD1, 0, sD2, 0, 32, acc);
                }
                // (the weight gradient of this layer needs Hh2, which has only just been requested when the tile
                // starts: it goes out with round 2, so that round 1 waits for nothing but the group's own dz)
            } else if constexpr (r == 1) {
                if (crit) {
                    gemm_dgrad(tw, B16, sD1, w::h1, 64, 0, 64, 64, 0);
                    gemm_wgrad(tbase + TcBwd::P_h2t, B16, sA, 0, sZ, 0, 16, acc);  // not deferred: tZ is rewritten next
                }
                if (late) gemm_wgrad(tbase + TcBwd::P_h1, B16, sD1, 0, sB, 0, 72, acc);
            } else if constexpr (r == 2) {
                if (crit) gemm_dgrad(tw, B16, sD2, w::hf0, 128, 0, 32, 64, 0);
                if (late) gemm_wgrad(tbase + TcBwd::P_h0, B16, sD2, 0, sX, 0, 32, acc);
            } else if constexpr (r == 3) {
                if (crit) gemm_dgrad(tw, B16, sZ, w::f2, 32, 0, 64, 32, 0);
                if (late) gemm_wgrad(tbase + TcBwd::P_f2t, B16, sA, 0, sZ, 0, 32, acc);
            } else if constexpr (r == 4) {
                if (crit) gemm_dgrad(tw, B16, sD1, w::f1, 64, 0, 64, 64, 0);
                if (late) gemm_wgrad(tbase + TcBwd::P_f1, B16, sD1, 0, sB, 0, 72, acc);
            } else if constexpr (r == 5) {
                if (crit) gemm_dgrad(tw, B16, sD2, w::hf0, 128, 64, 32, 64, 0);
                if (late) gemm_wgrad(tbase + TcBwd::P_f0, B16, sD2, 0, sX, 0, 32, acc);
            } else if constexpr (r == 6) {
                if (crit) gemm_dgrad(tw, B16, sZ, w::b1, 16, 0, 64, 16, 0);
                if (late) gemm_wgrad(tbase + TcBwd::P_b1t, B16, sA, 0, sZ, 0, 16, acc);
            } else {
                if (crit) {
                    if (want_denc) gemm_dgrad(tw, B16, sD1, w::b0, 64, 0, 32, 64, 0);
                    gemm_wgrad(tbase + TcBwd::P_b0, B16, sD1, 0, sB + 8192u, 0, 40, acc);  // [enc | ones]
                }
            }
        });
        __syncwarp();
    } else {
        // ======================= data groups
        uint8_t *sA = gbase + TcBwd2::sA, *sB = gbase + TcBwd2::sB, *tX = gbase + TcBwd2::tX;
        uint8_t *tD1 = gbase + TcBwd2::tD1, *tD2 = gbase + TcBwd2::tD2, *tZ = gbase + TcBwd2::tZ;
        uint64_t *fullA = fulls + 3 * grp, *fullB = fullA + 1, *fullX = fullA + 2;
        uint32_t parA = 0u, parB = 0u, parX = 0u;
        GroupSync gs{bars + grp, bars + G + grp, 0u, lane, Tracer(tid == 0, 0, smem, 2 * TcBwd2::kRounds), ready_flags(smem) + grp, kReadyBarId0 + grp};
        const uint32_t trow = tbase + ((uint32_t)((warp & 3) * 32) << 16) + 64u * grp;
        float *sb_h2 = sbias, *sb_f2 = sbias + 16, *sb_b1 = sbias + 48;
        const float invT = 1.0f / mp.temperature;
        float bs_h2 = 0.f, bs_f2 = 0.f, bs_f2s = 0.f, bs_b1 = 0.f;
        const uint64_t pol = l2_evict_first_policy();
        auto fetch = [&](uint8_t *dst, uint64_t *bar, int64_t tile, int off, uint32_t bytes) {  // row 0 of the group
            mbar_expect_tx(bar, bytes);
            bulk_load_hint(dst, sv.act_tiles + tile * (int64_t)ActTile::bytes + off, bytes, bar, pol);
        };
        auto wait_full = [&](uint64_t *bar, uint32_t &par) {
            mbar_wait(bar, par);
            par ^= 1u;
        };
        auto sample_of = [&](int64_t tile) {
            const int64_t m = tile * TM + row;
            return m < n_total ? m : n_total - 1;
        };
        PointRaw raw_in;
        PointIdx pi_in;
        int64_t pending_row = -1;
        // what the FIRST epilogue of a tile needs (gradient and value of the mlp_head outputs) is fetched during the
        // last rounds of the tile before: a DRAM latency right behind the tile's first hand-off otherwise
        float gh_in[KM], sh_in[KM];
        auto load_head = [&](int64_t tile) {
            const int64_t m = tile * TM + row;
            const bool ok = m < n_total;
#pragma unroll
            for (int k = 0; k < KM; ++k) {
                gh_in[k] = (ok && k < K) ? __ldg(g.d_scalar + m * K + k) : 0.f;
                sh_in[k] = (ok && k < K) ? __ldg(sv.scalar + m * K + k) : 0.f;
            }
        };
        if (tile_of(0, grp) < n_tiles) {
            load_head(tile_of(0, grp));
            if (row == 0) {
                fetch(sA, fullA, tile_of(0, grp), ActTile::HH2, 16384u);
                fetch(sB, fullB, tile_of(0, grp), ActTile::HH1, 16384u);
                fetch(tX, fullX, tile_of(0, grp), ActTile::X, 8192u);
            }
            pi_in = load_point_idx(P, sample_of(tile_of(0, grp)));
            raw_in = load_point_raw(P, sample_of(tile_of(0, grp)), pi_in);
        }

        for (int64_t it = 0; tile_of(it, grp) < n_tiles; ++it) {
            gs.round = 0;
            const int64_t tile = tile_of(it, grp), next = tile_of(it + 1, grp);
            const int64_t n0 = tile * TM;
            const bool valid = n0 + row < n_total;
            const int64_t n = valid ? n0 + row : n_total - 1;
            // invalid (tail) rows run the chain with ZERO incoming gradients: every delta they produce is exactly zero
            const Point pt = make_point(P, raw_in);
            const int64_t erow = enc_row(P, n);
            // ---------------- before round 1: dz of mlp_head -> tZ [; specular: d h -> tD1 (64 wide), X2 -> tD2]
            {
                float dz[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) dz[k] = 0.f;
#pragma unroll
                for (int k = 0; k < KM; ++k) dz[k] = gh_in[k] * sh_in[k] * (1.0f - sh_in[k]);  // (zero for tail rows)
                store8(tZ, row, 0, CH_ACT, dz);
                store8(tZ, row, 8, CH_ACT, dz + 8);
                bs_h2 += warp_colsum<KM>(dz, lane);
            }
            if (spec) {
                float pe[12], x2[32], sh[16], dsh[16];
                pos_encoding_tc(pt, pe);
                sh_deg3_tc((pt.dx + 1.0f) / 2.0f, (pt.dy + 1.0f) / 2.0f, (pt.dz + 1.0f) / 2.0f, sh);
#pragma unroll
                for (int i = 0; i < 16; ++i) x2[i] = sh[i];
#pragma unroll
                for (int i = 0; i < 12; ++i) x2[16 + i] = pe[i];
                x2[28] = 1.0f;
                x2[29] = x2[30] = x2[31] = 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c) store8(tD2, row, 8 * c, CH_ACT, x2 + 8 * c);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float4 gq = make_float4(0.f, 0.f, 0.f, 0.f), h = gq;
                    if (valid && g.d_spec_hidden) {
                        gq = __ldg(reinterpret_cast<const float4 *>(g.d_spec_hidden + n * 16) + q);
                        h = __ldg(reinterpret_cast<const float4 *>(sv.spec_hidden + n * 16) + q);
                    }
                    dsh[4 * q] = (h.x > 0.f) ? gq.x : 0.f;
                    dsh[4 * q + 1] = (h.y > 0.f) ? gq.y : 0.f;
                    dsh[4 * q + 2] = (h.z > 0.f) ? gq.z : 0.f;
                    dsh[4 * q + 3] = (h.w > 0.f) ? gq.w : 0.f;
                }
                const float z8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                store8(tD1, row, 0, CH_ACT, dsh);
                store8(tD1, row, 8, CH_ACT, dsh + 8);
#pragma unroll
                for (int c = 16; c < 64; c += 8) store8(tD1, row, c, CH_ACT, z8);
            }
            // gradients of the second MLP's outputs and of the density: used two / six rounds from here
            float ga[KM + 1], pa[KM + 1];
#pragma unroll
            for (int k = 0; k < KM; ++k) {
                ga[k] = (valid && k < K) ? __ldg(g.d_abund + n * K + k) : 0.f;
                pa[k] = (valid && k < K) ? __ldg(sv.abund + n * K + k) : 0.f;
            }
            ga[KM] = (valid && spec && g.d_s1) ? __ldg(g.d_s1 + n) : 0.f;
            pa[KM] = (valid && spec && g.d_s1) ? __ldg(sv.s1 + n) : 0.f;
            const float dsg = (valid && g.d_sigma) ? __ldg(g.d_sigma + n) : 0.f;
            const float sig = valid ? __ldg(sv.sigma + n) : 0.f;
            float demb[15];
#pragma unroll
            for (int i = 0; i < 15; ++i) demb[i] = (valid && g.d_embedding) ? __ldg(g.d_embedding + n * 15 + i) : 0.f;
            gs.arrive();
            if (d_enc && it > 0) flush_staged_pieces(d_enc, pending_row, tD2 + 8192, row);  // (tD2's upper half: idle until round 3's epilogue... see below)
            gs.wait();
            // ---------------- round 2: d h2 (head) masked by Hh2 -> tD1
            wait_full(fullA, parA);  // Hh2 (requested when the previous tile finished: a round of slack by now)
            epi_mask64(trow, sA, nullptr, tD1, row);
            wait_full(fullB, parB);  // Hh1
            gs.round_trip();
            // ---------------- round 3: d h1 masked by Hh1 -> tD2 ; dz of feature_mlp -> tZ ; slot A <- Hf2
            if (row == 0) fetch(sA, fullA, tile, ActTile::HF2, 16384u);
            epi_mask64(trow, sB, nullptr, tD2, row);
            {
                float dz[32];
#pragma unroll
                for (int k = 0; k < 32; ++k) dz[k] = 0.f;
                float dot = 0.f;
#pragma unroll
                for (int k = 0; k < KM; ++k) dot += pa[k] * ga[k];
#pragma unroll
                for (int k = 0; k < KM; ++k) dz[k] = pa[k] * invT * (ga[k] - dot);
                if (spec) {
                    const float ds = ga[KM] * pa[KM] * (1.0f - pa[KM]);
#pragma unroll
                    for (int k = 0; k <= KM; ++k)
                        if (k == K) dz[k] = ds;
                }
                store8(tZ, row, 0, CH_ACT, dz);
                store8(tZ, row, 8, CH_ACT, dz + 8);
                store8(tZ, row, 16, CH_ACT, dz + 16);
                store8(tZ, row, 24, CH_ACT, dz + 24);
                bs_f2 += warp_colsum<KM>(dz, lane);
                if (K + spec > KM) bs_f2s += warp_sum(dz[KM]);
            }
            wait_full(fullX, parX);  // X (first use in this tile)
            gs.round_trip();
            // ---------------- round 4: d X of mlp_head ; slot B <- Hf1
            if (row == 0) fetch(sB, fullB, tile, ActTile::HF1, 16384u);
            {
                float dx[32];
                tmem_ld32(trow, dx);
#pragma unroll
                for (int i = 0; i < 15; ++i) demb[i] += dx[12 + i];
            }
            float4 e_in[8];
            load_enc_pieces(enc, erow, e_in);  // this tile's encoding rows: operand of the last weight gradient
            wait_full(fullA, parA);  // Hf2
            gs.round_trip();
            // ---------------- round 5: d h2 (feat) masked by Hf2 -> tD1
            epi_mask64(trow, sA, nullptr, tD1, row);
            wait_full(fullB, parB);  // Hf1
            if (next < n_tiles) pi_in = load_point_idx(P, sample_of(next));
            gs.round_trip();
            // ---------------- round 6: d h1 masked by Hf1 -> tD2 ; slot A <- H0
            if (row == 0) fetch(sA, fullA, tile, ActTile::H0, 16384u);
            epi_mask64(trow, sB, nullptr, tD2, row);
            gs.round_trip();
            // ---------------- round 7: d X of feature_mlp ; dO = [d o0, d emb] -> tZ ; encoding rows -> slot B[8K:]
            {
                float dx[32], dO[16];
                tmem_ld32(trow, dx);
#pragma unroll
                for (int i = 0; i < 15; ++i) demb[i] += dx[12 + i];
                // density = trunc_exp(o0) * sel: d o0 = d_sigma * sel * exp(clamp(o0, -15, 15)) = d_sigma * clamp(density)
                dO[0] = dsg * pt.sel * fminf(fmaxf(sig, 3.0590232e-7f), 3269017.4f);
#pragma unroll
                for (int i = 0; i < 15; ++i) dO[1 + i] = demb[i];
                store8(tZ, row, 0, CH_ACT, dO);
                store8(tZ, row, 8, CH_ACT, dO + 8);
                bs_b1 += warp_colsum<16>(dO, lane);
            }
            store_enc_pieces(sB + 8192, row, e_in);
            if (next < n_tiles) {
                raw_in = load_point_raw(P, sample_of(next), pi_in);
                load_head(next);
            }
            wait_full(fullA, parA);  // H0
            gs.round_trip();
            // ---------------- round 8: d h0 masked by H0 -> tD1 ; tX <- next X
            if (row == 0 && next < n_tiles) fetch(tX, fullX, next, ActTile::X, 8192u);
            epi_mask64(trow, sA, nullptr, tD1, row);
            gs.round_trip();
            // ---------------- d enc ; slots A, B <- the next tile's Hh2, Hh1
            if (row == 0 && next < n_tiles) {
                fetch(sA, fullA, next, ActTile::HH2, 16384u);
                fetch(sB, fullB, next, ActTile::HH1, 16384u);
            }
            if (d_enc) {
                float v[32];
                tmem_ld32(trow, v);
                stage_pieces(v, tD2 + 8192, row);  // tD2's upper 8 KB (idle until round 3's epilogue of the next tile)
                pending_row = valid ? n : -1;
            }
        }
        if (d_enc && tile_of(0, grp) < n_tiles) flush_staged_pieces(d_enc, pending_row, tD2 + 8192, row);
        gs.tr.flush();
        if (colsum_owner<KM>(lane)) {
            const int col = colsum_col<KM>(lane);
            if (bs_h2 != 0.f) atomicAdd(sb_h2 + col, bs_h2);
            if (bs_f2 != 0.f) atomicAdd(sb_f2 + col, bs_f2);
        }
        if (colsum_owner<16>(lane) && bs_b1 != 0.f) atomicAdd(sb_b1 + colsum_col<16>(lane), bs_b1);
        if (lane == 0 && bs_f2s != 0.f) atomicAdd(sb_f2 + KM, bs_f2s);
    }
    flush_weight_grads(tbase, tid, grp, G, K, spec, dp, tile_of(0, 0) < n_tiles, false, false, sbias);
    if (warp == G * 4) tmem_dealloc<TcBwd::ncols>(tbase);
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
    const int acols = (mode >= 2) ? 64 : K;
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
    if (warp == 0) {
        if (elect_one()) {
            const uint32_t b16 = desc_base16(smem_u32(smem));
            if (mode == 0) gemm_kmajor(tbase, b16, 0u, 32768u, N, 0, N, K, 0);
            else if (mode == 1) gemm_dgrad(tbase, b16, 0u, 32768u, K, 0, N, K, 0);
            else if (mode == 2) gemm_wgrad(tbase, b16, 0u, 0, 32768u, 0, N, 0);
            else gemm_wgrad(tbase + (16u << 16), b16, 0u, 0, 32768u, 0, N, 0);  // mode 3: D at TMEM lanes 16..31 of each quarter
            commit(bar);
        }
        __syncwarp();
    }
    mbar_wait(bar, 0);
    fence_after_sync();
    const uint32_t trow = tbase + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < N; c += 16) {
        float v[16];
        tmem_ld16(trow + (uint32_t)c, v);
        if (mode == 2) {
            if (lane < 16)
                for (int i = 0; i < 16; ++i) D[(warp * 16 + lane) * N + c + i] = v[i];
        } else if (mode == 3) {
            if (lane >= 16)
                for (int i = 0; i < 16; ++i) D[(warp * 16 + lane - 16) * N + c + i] = v[i];
        } else {
            for (int i = 0; i < 16; ++i) D[tid * N + c + i] = v[i];
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<128>(tbase);
}

// ------------------------------------------------------------------ tcgen05 issue / execution micro-benchmark (one CTA)
// `reps` back-to-back GEMMs of one form on zeroed tiles: out[0] = cycles to ISSUE them (one elected thread),
// out[1] = cycles until their commit arrives.  mode as in the self-test; mode 3 = bias step (one K = 16 MMA).
__global__ void __launch_bounds__(TM)
tc05_microbench_kernel(int mode, int N, int K, int reps, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *tA = smem, *tB = smem + 32768;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 65536);
    uint32_t *slot = reinterpret_cast<uint32_t *>(smem + 65536 + 16);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 65536 / 16; i += TM) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<128>(slot);
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tbase = *slot;
    if (warp == 0) {
        if (elect_one()) {
            const long long t0 = clock64();
            for (int i = 0; i < reps; ++i) {
                const uint32_t b16 = desc_base16(smem_u32(smem));
                if (mode == 0) gemm_kmajor(tbase, b16, 0u, 32768u, N, 0, N, K, 1);
                else if (mode == 1) gemm_dgrad(tbase, b16, 0u, 32768u, K, 0, N, K, 1);
                else if (mode == 2) gemm_wgrad(tbase, b16, 0u, 0, 32768u, 0, N, 1);
                else gemm_bias_step(tbase, b16, 0u, 32768u, N, 0, N, 0);
            }
            const long long t1 = clock64();
            commit(bar);
            mbar_wait(bar, 0);
            const long long t2 = clock64();
            out[0] = t1 - t0;
            out[1] = t2 - t0;
        }
        __syncwarp();
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<128>(tbase);
}

__global__ void set_trace_kernel(long long *buf, int cap) {
    g_trace = buf;
    g_trace_cap = cap;
}

// persistent grid: one CTA per SM, never more CTAs than tile groups
static int tc_grid(int64_t n, int tiles_per_cta) {
    const int64_t n_tiles = (n + TM - 1) / TM;
    const int64_t want = (n_tiles + tiles_per_cta - 1) / tiles_per_cta;
    return (int)(want < kNumSMs ? (want < 1 ? 1 : want) : kNumSMs);
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
    UMHS_CHECK_ARG(mode >= 0 && mode <= 3 && N >= 16 && N <= 128 && N % 16 == 0 && K >= 16 && K <= 128 && K % 16 == 0,
                   "bad shape");
    const int bytes = 65536 + 128;
    UMHS_CHECK_CUDA(cudaFuncSetAttribute(tc05_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    tc05_selftest_kernel<<<1, TM, bytes, as_stream(stream)>>>(mode, A, B, D, N, K);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

extern "C" int umhs_tc05_microbench(int32_t mode, int32_t N, int32_t K, int32_t reps, int64_t *cycles, void *stream) {
    UMHS_CHECK_ARG(cycles && mode >= 0 && mode <= 3 && N >= 8 && N <= 128 && N % 8 == 0 && K >= 16 && K <= 128 &&
                       K % 16 == 0 && reps >= 1, "bad argument");
    const int bytes = 65536 + 128;
    UMHS_CHECK_CUDA(cudaFuncSetAttribute(tc05_microbench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    tc05_microbench_kernel<<<1, TM, bytes, as_stream(stream)>>>(mode, N, K, reps, reinterpret_cast<long long *>(cycles));
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

extern "C" int umhs_density_tc_fwd(const umhs_points *pts, int64_t n, const float *enc, const umhs_mlp_params *p,
                                   float *sigma, void *stream) {
    if (n <= 0) return UMHS_OK;
    UMHS_CHECK_ARG(pts && enc && sigma && p && p->base_w0 && p->base_b0 && p->base_w1 && p->base_b1, "null argument");
    umhs_field_out out = {};
    out.sigma = sigma;
    const int bytes = TcSmem::groups + TcFwdCfg<true>::G * TcFwd::grp_bytes_density;
    UMHS_CHECK_CUDA(cudaFuncSetAttribute(field_mlp_tc_fwd_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         bytes));
    field_mlp_tc_fwd_kernel<true, 4><<<tc_grid(n, TcFwdCfg<true>::G), TcFwdCfg<true>::kThreads, bytes, as_stream(stream)>>>(
        *pts, n, enc, *p, out);
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
    const int bytes = TcSmem::groups + TcFwdCfg<false>::G * TcFwd::grp_bytes;
    auto kernel = p->num_classes <= 4 ? field_mlp_tc_fwd_kernel<false, 4>
                : p->num_classes <= 8 ? field_mlp_tc_fwd_kernel<false, 8> : field_mlp_tc_fwd_kernel<false, 16>;
    UMHS_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    kernel<<<tc_grid(n, TcFwdCfg<false>::G), TcFwdCfg<false>::kThreads, bytes, as_stream(stream)>>>(*pts, n, enc, *p,
                                                                                                  *out);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

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
    auto kernel = p->num_classes <= 4 ? field_mlp_tc_bwd_kernel<4, false>
                : p->num_classes <= 8 ? field_mlp_tc_bwd_kernel<8, false> : field_mlp_tc_bwd_kernel<16, false>;
    UMHS_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TcBwd::smem_bytes));
    kernel<<<tc_grid(n, TcBwd::G), TcBwd::kThreads, TcBwd::smem_bytes, as_stream(stream)>>>(*pts, n, enc, *p, *g, d_enc,
                                                                                          *dp, ScatterArgs{});
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

extern "C" int umhs_field_mlp_tc_bwd_saved(const umhs_points *pts, int64_t n, const float *enc, const umhs_mlp_params *p,
                                           const umhs_field_out_grads *g, const umhs_field_saved *saved,
                                           void *d_enc_tiles, const umhs_mlp_grads *dp, void *stream) {
    if (n <= 0) return UMHS_OK;
    UMHS_CHECK_ARG(pts && enc && g && dp && p && saved, "null argument");
    UMHS_CHECK_ARG(p->enc_tiles == 1, "the saved-activation backward works on the bf16 tile buffers (enc_tiles = 1)");
    UMHS_CHECK_ARG(saved->act_tiles && saved->sigma && saved->scalar && saved->abund, "saved forward state missing");
    UMHS_CHECK_ARG((reinterpret_cast<uintptr_t>(saved->act_tiles) & 15u) == 0, "act_tiles must be 16-byte aligned");
    UMHS_CHECK_ARG(!p->specular || (saved->s1 && saved->spec_hidden), "saved specular outputs missing");
    UMHS_CHECK_ARG(p->num_classes >= 1 && p->num_classes <= UMHS_MAX_CLASSES && p->temperature > 0.f, "bad K / T");
    UMHS_CHECK_ARG(g->d_scalar && g->d_abund, "d_scalar/d_abund are required");
    UMHS_CHECK_ARG(dp->base_w0 && dp->base_b0 && dp->base_w1 && dp->base_b1 && dp->head_w0 && dp->head_b0 &&
                       dp->head_w1 && dp->head_b1 && dp->head_w2 && dp->head_b2 && dp->feat_w0 && dp->feat_b0 &&
                       dp->feat_w1 && dp->feat_b1 && dp->feat_w2 && dp->feat_b2,
                   "null parameter-gradient pointer");
    UMHS_CHECK_ARG(!p->specular || (p->dir_w0 && p->dir_b0 && dp->dir_w0 && dp->dir_b0), "specular pointers missing");
    auto kernel = p->num_classes <= 4 ? field_mlp_tc_bwd_saved_kernel<4>
                : p->num_classes <= 8 ? field_mlp_tc_bwd_saved_kernel<8> : field_mlp_tc_bwd_saved_kernel<16>;
    UMHS_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TcBwd2::smem_bytes));
    const SavedFwd sv{reinterpret_cast<const uint8_t *>(saved->act_tiles), saved->sigma, saved->scalar, saved->abund,
                      saved->s1, saved->spec_hidden};
    kernel<<<tc_grid(n, TcBwd2::G), TcBwd2::kThreads, TcBwd2::smem_bytes, as_stream(stream)>>>(
        *pts, n, enc, *p, *g, sv, reinterpret_cast<float *>(d_enc_tiles), *dp);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

extern "C" int umhs_field_mlp_tc_bwd_scatter(const umhs_points *pts, int64_t n, const float *enc,
                                             const umhs_mlp_params *p, const umhs_field_out_grads *g, void *d_enc_tiles,
                                             const umhs_mlp_grads *dp, const umhs_hash_grid *grid, float *d_table,
                                             void *stream) {
    if (n <= 0) return UMHS_OK;
    UMHS_CHECK_ARG(pts && enc && g && dp && p && grid && d_table && d_enc_tiles, "null argument");
    UMHS_CHECK_ARG(p->enc_tiles == 1, "the fused scatter reads d_enc as the bf16 tile buffer (enc_tiles = 1)");
    UMHS_CHECK_ARG(grid->scalings && grid->num_levels == 16 && grid->log2_table_size >= 1 && grid->log2_table_size <= 28,
                   "bad hash grid (the tile layout is 16 levels x 2 features wide)");
    UMHS_CHECK_ARG(p->num_classes >= 1 && p->num_classes <= UMHS_MAX_CLASSES && p->temperature > 0.f, "bad K / T");
    UMHS_CHECK_ARG(g->d_scalar && g->d_abund, "d_scalar/d_abund are required");
    UMHS_CHECK_ARG(dp->base_w0 && dp->base_b0 && dp->base_w1 && dp->base_b1 && dp->head_w0 && dp->head_b0 &&
                       dp->head_w1 && dp->head_b1 && dp->head_w2 && dp->head_b2 && dp->feat_w0 && dp->feat_b0 &&
                       dp->feat_w1 && dp->feat_b1 && dp->feat_w2 && dp->feat_b2,
                   "null parameter-gradient pointer");
    UMHS_CHECK_ARG(!p->specular || (p->dir_w0 && p->dir_b0 && dp->dir_w0 && dp->dir_b0), "specular pointers missing");
    auto kernel = p->num_classes <= 4 ? field_mlp_tc_bwd_kernel<4, true>
                : p->num_classes <= 8 ? field_mlp_tc_bwd_kernel<8, true> : field_mlp_tc_bwd_kernel<16, true>;
    UMHS_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TcBwd::smem_bytes));
    const ScatterArgs sc{grid->scalings, reinterpret_cast<float2 *>(d_table), grid->num_levels, grid->log2_table_size,
                         grid->num_levels < 10 ? grid->num_levels : 10};
    kernel<<<tc_grid(n, TcBwd::G), TcBwd::kThreadsScatter, TcBwd::smem_bytes, as_stream(stream)>>>(
        *pts, n, enc, *p, *g, reinterpret_cast<float *>(d_enc_tiles), *dp, sc);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}
