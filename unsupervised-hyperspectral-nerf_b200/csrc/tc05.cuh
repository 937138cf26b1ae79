// tc05.cuh -- hand-written Blackwell (sm_100a) tensor-core building blocks: tcgen05.mma with
// shared-memory operands and TMEM accumulators, tcgen05.ld epilogue loads, mbarrier completion.
//
// Operand storage ("chunk-slab", the canonical no-swizzle UMMA layout): a [rows x cols] bf16 matrix
// lives at   byte(row, col) = (col/8)*CH + (row/8)*128 + (row%8)*16 + (col%8)*2,  CH = (rows/8)*128,
// i.e. 8x8 core matrices of 128 contiguous bytes.  The SAME buffer is
//   * a K-major operand  (MN = rows, K = cols): SBO = 128 (next 8-row group), LBO = CH (next 8 K);
//   * an MN-major operand (MN = cols, K = rows): SBO = CH (next 8 MN),        LBO = 128 (next 8 K),
// so an activation tile [samples x features] feeds forward / dgrad GEMMs (K-major, K = features) and
// weight-gradient GEMMs (MN-major, K = samples) without any transpose, and an nn.Linear weight
// [out x in] is the K-major B operand of the forward GEMM and the MN-major B operand of dgrad.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace tc05 {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- shared-memory matrix descriptor (SWIZZLE_NONE, version 1)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;  // descriptor version for sm_100
    return d;
}

// ---- instruction descriptor: bf16 x bf16 -> f32, dense
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
    uint32_t d = 0;
    d |= 1u << 4;   // c_format = F32
    d |= 1u << 7;   // a_format = BF16
    d |= 1u << 10;  // b_format = BF16
    d |= (uint32_t)(a_mn_major & 1) << 15;
    d |= (uint32_t)(b_mn_major & 1) << 16;
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}

// One lane of a CONVERGED warp (all 32 lanes call this together with warp-uniform arguments).  The MMA-issuing
// code must stay warp-uniform: inside a divergent `if (lane == 0)` the compiler wraps every tcgen05 instruction
// (they live on the uniform datapath) in an elect/branch loop with register->uniform-register moves, which costs
// ~100+ cycles per MMA on the single thread that paces the whole SM.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// Issued by ONE thread: call inside `if (elect_one()) { ... }` of a converged warp (one elect per MMA round).
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                       uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// all previously issued MMAs of this thread arrive on `bar` when complete (same elected thread as the MMAs)
__device__ __forceinline__ void commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// make generic-proxy smem writes visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// bounded wait: a wrong descriptor must fail loudly (trap), never hang the GPU.
// mbarrier.try_wait suspends the thread in hardware; measured on B200 it costs ~230 cycles even when the phase has
// long completed, and about as much to wake up.  SPIN polls first: `SPIN` non-blocking mbarrier.test_wait probes (a
// probe is a shared-memory round trip, ~30 cycles) before falling back to the suspending form.
template <int SPIN = 0>
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    if (SPIN > 0) {
#pragma unroll 1
        for (int it = 0; it < SPIN; ++it) {
            uint32_t ok;
            asm volatile(
                "{\n\t"
                ".reg .pred p;\n\t"
                "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t"
                "}\n"
                : "=r"(ok)
                : "r"(a), "r"(parity)
                : "memory");
            if (ok) return;
        }
    }
#pragma unroll 1
    for (uint32_t it = 0; it < (1u << 24); ++it) {
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
        if (ok) return;
    }
    __trap();
}

// ---- TMEM allocation (one full warp executes these)
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "n"(NCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}

// ---- TMEM -> registers: this warp's 32 lanes x 16 consecutive fp32 columns (lane i <-> TMEM lane 32*(warp%4)+i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- operand tile addressing
__device__ __forceinline__ uint32_t tile_off(int row, int col, int ch_bytes) {
    return (uint32_t)((col >> 3) * ch_bytes + (row >> 3) * 128 + (row & 7) * 16 + (col & 7) * 2);
}

// store 8 consecutive columns [c0, c0+8) of `row` (c0 multiple of 8) as bf16: one 16-byte st.shared
__device__ __forceinline__ void store8(uint8_t *buf, int row, int c0, int ch_bytes, const float *v) {
    __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]);
    __nv_bfloat162 p1 = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]);
    __nv_bfloat162 p3 = __floats2bfloat162_rn(v[6], v[7]);
    uint4 q;
    q.x = *reinterpret_cast<uint32_t *>(&p0);
    q.y = *reinterpret_cast<uint32_t *>(&p1);
    q.z = *reinterpret_cast<uint32_t *>(&p2);
    q.w = *reinterpret_cast<uint32_t *>(&p3);
    *reinterpret_cast<uint4 *>(buf + tile_off(row, c0, ch_bytes)) = q;
}

__device__ __forceinline__ void load8(const uint8_t *buf, int row, int c0, int ch_bytes, float *v) {
    const uint4 q = *reinterpret_cast<const uint4 *>(buf + tile_off(row, c0, ch_bytes));
    const __nv_bfloat162 *p = reinterpret_cast<const __nv_bfloat162 *>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(p[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}

// Descriptors as "base + compile-time constant".  The issuing thread paces the SM, and every descriptor it has to
// assemble at run time (shift, mask, or-in the strides: 4-5 uniform-datapath instructions each) sits on that path.
// With b16 = (shared-window address of the CTA's dynamic smem) >> 4 -- one warp-uniform value -- the descriptor of
// the operand at byte offset `off` from that base is
//     lo = b16 + ((off >> 4) | (LBO >> 4) << 16)         (start address: bits 0..13, LBO: bits 16..29)
//     hi = (SBO >> 4) | 1 << 14                          (SBO: bits 32..45, descriptor version: bit 46)
// i.e. ONE add per descriptor once `off` and the strides are constants after inlining (shared memory ends below
// 256 KB, so the 14-bit address field never carries into the LBO field).
__device__ __forceinline__ uint32_t desc_base16(uint32_t smem_base_saddr) { return (smem_base_saddr & 0x3FFFFu) >> 4; }
__device__ __forceinline__ uint64_t desc_at(uint32_t b16, uint32_t off_bytes, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    const uint32_t lo = b16 + ((off_bytes >> 4) | ((lbo_bytes >> 4) << 16));
    const uint32_t hi = ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14);
    return ((uint64_t)hi << 32) | (uint64_t)lo;
}

// D[128 x N] (+)= A[128 x K] * B[N x K]^T, both operands K-major tiles at byte offsets a_off / b_off from the base.
//   a: tile of 128 rows (CH = 2048); b: tile of `b_rows` rows (CH = b_rows*16), rows [b_row0, b_row0+N) used.
__device__ __forceinline__ void gemm_kmajor(uint32_t d_tmem, uint32_t b16, uint32_t a_off, uint32_t b_off, int b_rows,
                                            int b_row0, int N, int K, uint32_t accumulate) {
    const uint32_t idesc = make_idesc(128, N, 0, 0);
    const uint32_t a_ch = 2048, b_ch = (uint32_t)b_rows * 16;
    const uint32_t b0 = b_off + (uint32_t)(b_row0 >> 3) * 128;
#pragma unroll
    for (int k = 0; k < K; k += 16)
        mma_ss(d_tmem, desc_at(b16, a_off + (uint32_t)(k >> 3) * a_ch, a_ch, 128),
               desc_at(b16, b0 + (uint32_t)(k >> 3) * b_ch, b_ch, 128), idesc, (k > 0) ? 1u : accumulate);
}

// dgrad: D[128 x N] = A[128 x K] * W[K x N], A K-major (delta tile), W stored as nn.Linear [K(out) x N(in)]
// K-major-for-forward, read here as an MN-major B operand (MN = in, K = out).  Rows (out) [w_row0, w_row0+K).
__device__ __forceinline__ void gemm_dgrad(uint32_t d_tmem, uint32_t b16, uint32_t a_off, uint32_t w_off, int w_rows,
                                           int w_row0, int N, int K, uint32_t accumulate) {
    const uint32_t idesc = make_idesc(128, N, 0, 1);
    const uint32_t a_ch = 2048, w_ch = (uint32_t)w_rows * 16;
    const uint32_t w0 = w_off + (uint32_t)(w_row0 >> 3) * 128;
    // MN-major: SBO = stride between 8-wide MN chunks (= w_ch), LBO = stride between 8-deep K groups (= 128)
#pragma unroll
    for (int k = 0; k < K; k += 16)
        mma_ss(d_tmem, desc_at(b16, a_off + (uint32_t)(k >> 3) * a_ch, a_ch, 128),
               desc_at(b16, w0 + (uint32_t)(k >> 3) * 128, 128, w_ch), idesc, (k > 0) ? 1u : accumulate);
}

// wgrad: D[64 x N] += P[128 samples x 64]^T * Q[128 samples x N]: both are activation tiles (CH = 2048) read as
// MN-major operands with K = samples.  P columns [p_col0, p_col0+64), Q columns [q_col0, q_col0+N).
__device__ __forceinline__ void gemm_wgrad(uint32_t d_tmem, uint32_t b16, uint32_t p_off, int p_col0, uint32_t q_off,
                                           int q_col0, int N, uint32_t accumulate) {
    const uint32_t idesc = make_idesc(64, N, 1, 1);
    const uint32_t ch = 2048;
    const uint32_t p0 = p_off + (uint32_t)(p_col0 >> 3) * ch, q0 = q_off + (uint32_t)(q_col0 >> 3) * ch;
#pragma unroll
    for (int s = 0; s < 128; s += 16)
        mma_ss(d_tmem, desc_at(b16, p0 + (uint32_t)(s >> 3) * 128, 128, ch),
               desc_at(b16, q0 + (uint32_t)(s >> 3) * 128, 128, ch), idesc, (s > 0) ? 1u : accumulate);
}

// ---- TMEM -> registers, 64 consecutive columns with a single wait (four loads in flight)
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, float (&v)[64]) {
    uint32_t r[64];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[16 * q + 0]), "=r"(r[16 * q + 1]), "=r"(r[16 * q + 2]), "=r"(r[16 * q + 3]),
              "=r"(r[16 * q + 4]), "=r"(r[16 * q + 5]), "=r"(r[16 * q + 6]), "=r"(r[16 * q + 7]),
              "=r"(r[16 * q + 8]), "=r"(r[16 * q + 9]), "=r"(r[16 * q + 10]), "=r"(r[16 * q + 11]),
              "=r"(r[16 * q + 12]), "=r"(r[16 * q + 13]), "=r"(r[16 * q + 14]), "=r"(r[16 * q + 15])
            : "r"(taddr + 16u * q)
            : "memory");
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- plain shared-memory signal (release add / acquire load): a handshake that stays out of the mbarrier unit
__device__ __forceinline__ void flag_release_add(uint32_t *flag) {
    asm volatile("red.release.cta.shared::cta.add.u32 [%0], 1;" ::"r"(smem_u32(flag)) : "memory");
}
__device__ __forceinline__ uint32_t flag_acquire_load(const uint32_t *flag) {
    uint32_t v;
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(flag)) : "memory");
    return v;
}
// wait until the counter has reached `want` (wrap-safe); bounded like mbar_wait
__device__ __forceinline__ void flag_wait(const uint32_t *flag, uint32_t want) {
#pragma unroll 1
    for (uint32_t it = 0; it < (1u << 26); ++it)
        if ((int32_t)(flag_acquire_load(flag) - want) >= 0) return;
    __trap();
}

// ---- TMA, 1-D bulk form: `bytes` (multiple of 16, both addresses 16-byte aligned) from global to shared memory by
// the copy engine; completion is counted in bytes on an mbarrier (arm it with mbar_expect_tx first).  One thread issues.
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// The saved-activation stream (~0.6 GB per step, written once by streaming stores and read once by these bulk copies)
// carries an evict-first L2 policy: it should not push the hash table and its gradient out of the L2.
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_load_hint(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar,
                                               uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// ---- named CTA barriers (the hardware barrier unit: no shared-memory access, unlike mbarriers / flags, whose
// probes queue behind the tensor core's operand reads).  Producer / consumer form: the producers bar.arrive (they do
// not wait), the consumer bar.sync; `count` = all participating threads of both sides (a multiple of 32).
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// ---- mbarrier: arrive / non-blocking test (issuer polling several groups)
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// ---- packed bf16x2 epilogue arithmetic: one instruction converts two fp32 accumulators (and applies ReLU)
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
__device__ __forceinline__ uint32_t pack_relu_bf16x2(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
// 0xffff per half where the (post-ReLU, hence >= +0) bf16 activation is positive
__device__ __forceinline__ uint32_t relu_mask_bf16x2(uint32_t h) {
    uint32_t m;
    asm("set.gt.u32.bf16x2 %0, %1, %2;" : "=r"(m) : "r"(h), "r"(0u));
    return m;
}

// ---- TMEM -> registers, raw 32-bit words: 32 / 64 consecutive columns, one wait
__device__ __forceinline__ void tmem_ld32_raw_nowait(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld64_raw(uint32_t taddr, uint32_t (&r)[64]) {
    tmem_ld32_raw_nowait(taddr, r);
    tmem_ld32_raw_nowait(taddr + 32u, r + 32);
    tmem_wait_ld();
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    tmem_ld32_raw_nowait(taddr, r);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// One more K = 16 step whose A operand is the constant tile of ones: adds the bias column stored in the weight
// tile at K columns [kcol, kcol + 16) (bias in column kcol, zeros behind it) to D[128 x N].
__device__ __forceinline__ void gemm_bias_step(uint32_t d_tmem, uint32_t b16, uint32_t ones_off, uint32_t b_off,
                                               int b_rows, int b_row0, int N, int kcol) {
    const uint32_t b_ch = (uint32_t)b_rows * 16;
    mma_ss(d_tmem, desc_at(b16, ones_off, 2048, 128),
           desc_at(b16, b_off + (uint32_t)(b_row0 >> 3) * 128 + (uint32_t)(kcol >> 3) * b_ch, b_ch, 128),
           make_idesc(128, N, 0, 0), 1u);
}


}  // namespace tc05
