// composite_group.cu -- compositing x unmixing x spec->sRGB x cluster probe, lane-group-per-ray form (no specular head).
//
// Same arithmetic and the same reference lines as composite.cu (ref umhsnerf/umhs_model.py:245-313,
// umhs_renderer.py:15-30, utils/spec_to_rgb.py:112-127, utils/clusterprobe.py:17-38, umhs_field.py:229-238), mapped
// differently.  The warp-per-ray form is instruction bound: for a 50-sample ray most of its ~1200 warp instructions
// are 5-step shuffle reductions of 2K+2 per-ray sums, a single-lane output tail and scalar [N,K] loads, all paid per
// ray.  Here a ray belongs to a GROUP of G lanes (G = 8 for B <= 128 bands, 16 for <= 256, 32 for <= 512), so a warp
// carries 32/G rays side by side:
//   * sample phase: lane g of a group walks samples g, g+G, ...; the transmittance scan is a log2(G)-step segmented
//     shuffle scan with a carry; [N,K] rows move as float4; the next iteration's loads are issued before this
//     iteration's math;
//   * the 2K+2 per-ray sums are reduced over log2(G) steps for 32/G rays at once;
//   * band phase: lane g owns float4 band groups g, g+G, ... (E and the colour matrix are staged band-padded in
//     shared memory and read as conflict-free float4, broadcast across the groups of a warp);
//   * backward: the forward's transmittance is an input (4 B/sample instead of a recomputation pass), the reverse
//     pass is a single sweep, and dE = sum_r mix_r (x) dS_r is a register-blocked product over the rays of a CTA
//     iteration staged in shared memory (no per-ray atomics); one atomic per element per CTA at the end.
#include <math_constants.h>

#include "composite.cuh"

namespace umhs {

constexpr int kGrpWarps = 4;
constexpr int kGrpThreads = kGrpWarps * 32;

template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
template <int G>
__device__ __forceinline__ float group_incl_scan(float v, int g) {
#pragma unroll
    for (int o = 1; o < G; o <<= 1) {
        const float t = __shfl_up_sync(kFull, v, o, G);
        if (g >= o) v += t;
    }
    return v;
}
template <int G>
__device__ __forceinline__ float group_incl_rscan(float v, int g) {
#pragma unroll
    for (int o = 1; o < G; o <<= 1) {
        const float t = __shfl_down_sync(kFull, v, o, G);
        if (g + o < G) v += t;
    }
    return v;
}

__device__ __forceinline__ float dot4(const float4 a, const float4 b) {
    return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
}
__device__ __forceinline__ void fma4(float4 &d, float s, const float4 a) {
    d.x += s * a.x; d.y += s * a.y; d.z += s * a.z; d.w += s * a.w;
}

// one [.,K] row (K floats at p): float4 pieces when the rows are 16-byte aligned, scalars otherwise
template <int KMAX>
__device__ __forceinline__ void load_row(const float *p, int K, bool vec, float (&v)[KMAX]) {
    if (vec) {
#pragma unroll
        for (int q = 0; q < KMAX / 4; ++q) {
            if (4 * q < K) {
                const float4 t = __ldg(reinterpret_cast<const float4 *>(p) + q);
                v[4 * q + 0] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
            if (k < K) v[k] = __ldg(p + k);
    }
}
template <int KMAX>
__device__ __forceinline__ void store_row(float *p, int K, bool vec, const float (&v)[KMAX]) {
    if (vec) {
#pragma unroll
        for (int q = 0; q < KMAX / 4; ++q)
            if (4 * q < K)
                reinterpret_cast<float4 *>(p)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    } else {
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
            if (k < K) p[k] = v[k];
    }
}
// four consecutive bands 4*b4 .. 4*b4+3 of a [.,B] row
__device__ __forceinline__ float4 load_bands(const float *row, int b4, int B, bool vec) {
    if (vec) return __ldg(reinterpret_cast<const float4 *>(row) + b4);
    const int b = 4 * b4;
    float4 t;
    t.x = b + 0 < B ? __ldg(row + b + 0) : 0.f;
    t.y = b + 1 < B ? __ldg(row + b + 1) : 0.f;
    t.z = b + 2 < B ? __ldg(row + b + 2) : 0.f;
    t.w = b + 3 < B ? __ldg(row + b + 3) : 0.f;
    return t;
}
__device__ __forceinline__ void store_bands(float *row, int b4, int B, bool vec, const float4 t) {
    if (vec) {
        reinterpret_cast<float4 *>(row)[b4] = t;
        return;
    }
    const int b = 4 * b4;
    if (b + 0 < B) row[b + 0] = t.x;
    if (b + 1 < B) row[b + 1] = t.y;
    if (b + 2 < B) row[b + 2] = t.z;
    if (b + 3 < B) row[b + 3] = t.w;
}

// E [K,B] -> E_s [K,Bp] and M [B,3] -> M_s [3,Bp], zero-padded to Bp = 4 * nb4 bands
__device__ __forceinline__ void stage_tables(const umhs_composite_in &in, int K, int B, int Bp, float *E_s, float *M_s) {
    for (int i = threadIdx.x; i < K * Bp; i += kGrpThreads) {
        const int k = i / Bp, b = i - k * Bp;
        E_s[i] = b < B ? __ldg(in.endmembers + k * B + b) : 0.f;
    }
    for (int i = threadIdx.x; i < 3 * Bp; i += kGrpThreads) {
        const int c = i / Bp, b = i - c * Bp;
        M_s[i] = b < B ? __ldg(in.rgb_matrix + 3 * b + c) : 0.f;
    }
}

template <int KMAX>
struct SampleRegs {
    float t0, t1, sg, T;
    float sv[KMAX], av[KMAX];
};

template <int KMAX, bool WITH_T>
__device__ __forceinline__ void load_sample(SampleRegs<KMAX> &s, bool v, int idx, int K, bool vecK, const float *ts,
                                            const float *te, const float *sgp, const float *trp, const float *scp,
                                            const float *abp) {
    s.t0 = s.t1 = s.sg = s.T = 0.f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) s.sv[k] = s.av[k] = 0.f;
    if (v) {
        s.t0 = __ldg(ts + idx);
        s.t1 = __ldg(te + idx);
        s.sg = __ldg(sgp + idx);
        if (WITH_T) s.T = __ldg(trp + idx);
        load_row<KMAX>(scp + (int64_t)idx * K, K, vecK, s.sv);
        load_row<KMAX>(abp + (int64_t)idx * K, K, vecK, s.av);
    }
}

// ------------------------------------------------------------------------------------------------ forward
template <int KMAX, int G, int NI>
__global__ void __launch_bounds__(kGrpThreads, KMAX <= 4 ? 8 : 4)
composite_fwd_group_kernel(umhs_composite_in in, umhs_composite_out out, int vecK, int vecB) {
    extern __shared__ float4 smem4[];
    float *smem = reinterpret_cast<float *>(smem4);
    constexpr int RPW = 32 / G;              // rays per warp
    constexpr int RPC = kGrpWarps * RPW;     // rays per CTA iteration
    constexpr bool PREFETCH = KMAX <= 8;
    const int K = in.num_classes, B = in.num_bands;
    const int Bp = (B + 3) & ~3, nb4 = Bp >> 2;
    float *E_s = smem, *M_s = E_s + K * Bp, *En_s = M_s + 3 * Bp;
    const float4 *E4 = reinterpret_cast<const float4 *>(E_s), *M4 = reinterpret_cast<const float4 *>(M_s);
    stage_tables(in, K, B, Bp, E_s, M_s);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int k = warp; k < K; k += kGrpWarps) {  // clusterprobe.py:27 F.normalize(clusters): ||E_k||, eps 1e-12
        float s = 0.f;
        for (int b = lane; b < Bp; b += 32) s += E_s[k * Bp + b] * E_s[k * Bp + b];
        s = warp_sum(s);
        if (lane == 0) En_s[k] = fmaxf(sqrtf(s), 1e-12f);
    }
    __syncthreads();

    const int g = lane & (G - 1), q = lane / G;
    const bool vK = vecK != 0, vB = vecB != 0;
    float tmin = CUDART_INF_F, tmax = -CUDART_INF_F;
    for (int64_t base = (int64_t)blockIdx.x * RPC; base < in.n_rays; base += (int64_t)gridDim.x * RPC) {
        const int64_t r = base + warp * RPW + q;
        const bool rv = r < in.n_rays;
        int64_t start = 0;
        int cnt = 0;
        if (rv) {
            const longlong2 pi = __ldg(reinterpret_cast<const longlong2 *>(in.packed_info) + r);
            start = pi.x;
            cnt = (int)pi.y;
        }
        const float *ts = in.t_starts + start, *te = in.t_ends + start, *sgp = in.sigma + start;
        const float *scp = in.scalar + start * K, *abp = in.abund + start * K;
        float *wout = out.weights ? out.weights + start : nullptr;
        float *tout = out.trans ? out.trans + start : nullptr;
        const int iters = (__reduce_max_sync(kFull, cnt) + G - 1) / G;

        float carry = 0.f, acc = 0.f, dep = 0.f, mix[KMAX], ab[KMAX];
#pragma unroll
        for (int k = 0; k < KMAX; ++k) mix[k] = ab[k] = 0.f;
        SampleRegs<KMAX> cur, nxt;
        if (PREFETCH) load_sample<KMAX, false>(cur, g < cnt, g, K, vK, ts, te, sgp, nullptr, scp, abp);
        for (int it = 0; it < iters; ++it) {
            const int idx = it * G + g;
            const bool v = idx < cnt;
            if (PREFETCH)
                load_sample<KMAX, false>(nxt, idx + G < cnt, idx + G, K, vK, ts, te, sgp, nullptr, scp, abp);
            else
                load_sample<KMAX, false>(cur, v, idx, K, vK, ts, te, sgp, nullptr, scp, abp);
            const float sd = cur.sg * (cur.t1 - cur.t0);
            const float incl = group_incl_scan<G>(sd, g);
            const float T = expf(-(carry + (incl - sd)));
            const float alpha = 1.0f - expf(-sd);
            const float w = v ? T * alpha : 0.f;
            carry += __shfl_sync(kFull, incl, G - 1, G);
            const float tm = (cur.t0 + cur.t1) / 2.0f;
            if (v) {
                if (wout) wout[idx] = w;
                if (tout) tout[idx] = T;
                tmin = fminf(tmin, tm);
                tmax = fmaxf(tmax, tm);
            }
            acc += w;
            dep += w * tm;
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
                if (k < K) {
                    mix[k] += w * cur.sv[k] * cur.av[k];
                    ab[k] += w * cur.av[k];
                }
            }
            if (PREFETCH) cur = nxt;
        }
        acc = group_sum<G>(acc);
        dep = group_sum<G>(dep);
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            if (k < K) {
                mix[k] = group_sum<G>(mix[k]);
                ab[k] = group_sum<G>(ab[k]);
            }
        }
        // ---- band phase: unmix once per ray, colour projection, probe inner products
        float lin0 = 0.f, lin1 = 0.f, lin2 = 0.f, nrm = 0.f, dot[KMAX];
#pragma unroll
        for (int k = 0; k < KMAX; ++k) dot[k] = 0.f;
        float *srow = out.spectral ? out.spectral + r * B : nullptr;
        float *s2row = (in.specular_in && out.spectral2) ? out.spectral2 + r * B : nullptr;
        const float *sprow = in.specular_in ? in.specular_in + r * B : nullptr;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const int b4 = g + G * i;
            if (b4 < nb4) {
                float4 sp = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int k = 0; k < KMAX; ++k)
                    if (k < K) fma4(sp, mix[k], E4[k * nb4 + b4]);
                if (sprow) {  // specular term composited beforehand (tensor-core kernel): spectral = unmixed + specular
                    if (rv && s2row) store_bands(s2row, b4, B, vB, sp);
                    if (rv) {
                        const float4 a = load_bands(sprow, b4, B, vB);
                        sp.x += a.x; sp.y += a.y; sp.z += a.z; sp.w += a.w;
                    }
                }
                if (rv && srow) store_bands(srow, b4, B, vB, sp);
                lin0 += dot4(sp, M4[0 * nb4 + b4]);
                lin1 += dot4(sp, M4[1 * nb4 + b4]);
                lin2 += dot4(sp, M4[2 * nb4 + b4]);
                nrm += dot4(sp, sp);
#pragma unroll
                for (int k = 0; k < KMAX; ++k)
                    if (k < K) dot[k] += dot4(sp, E4[k * nb4 + b4]);
            }
        }
        lin0 = group_sum<G>(lin0);
        lin1 = group_sum<G>(lin1);
        lin2 = group_sum<G>(lin2);
        nrm = fmaxf(sqrtf(group_sum<G>(nrm)), 1e-12f);
        float mx = -CUDART_INF_F;
        int arg = 0;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            if (k < K) {
                dot[k] = group_sum<G>(dot[k]) / (nrm * En_s[k]);
                if (dot[k] > mx) {
                    mx = dot[k];
                    arg = k;
                }
            }
        }
        // ---- per-ray tail: lanes 0..2 of the group take one colour channel each, lane 0 the rest
        const float gate = acc > 0.5f ? 1.f : 0.f;
        if (rv && g < 3) {
            const float x = g == 0 ? lin0 : (g == 1 ? lin1 : lin2);
            if (out.rgb) out.rgb[r * 3 + g] = fminf(fmaxf(gamma_fwd(x), 0.f), 1.f);
            if (out.seg_pred && in.palette) out.seg_pred[r * 3 + g] = __ldg(in.palette + 3 * arg + g) * gate;
        }
        if (rv && g == 0) {
            if (out.accumulation) out.accumulation[r] = acc;
            if (out.depth) out.depth[r] = dep / (acc + 1e-10f);
            if (out.seg_raw) out.seg_raw[r] = (float)arg * gate;  // argmax of the probabilities == first argmax of dot
            if (out.abundances) store_row<KMAX>(out.abundances + r * K, K, vK, ab);
            if (out.seg_probs) {
                float den = 0.f, pr[KMAX];
#pragma unroll
                for (int k = 0; k < KMAX; ++k) {
                    pr[k] = (k < K) ? expf(0.2f * (dot[k] - mx)) : 0.f;
                    den += pr[k];
                }
#pragma unroll
                for (int k = 0; k < KMAX; ++k) pr[k] = pr[k] / den;
                store_row<KMAX>(out.seg_probs + r * K, K, vK, pr);
            }
        }
    }
    if (out.t_minmax) {
        tmin = -warp_max(-tmin);
        tmax = warp_max(tmax);
        if (lane == 0 && tmin <= tmax) {  // non-negative floats order like their bit patterns
            atomicMin(reinterpret_cast<int *>(out.t_minmax), __float_as_int(tmin));
            atomicMax(reinterpret_cast<int *>(out.t_minmax) + 1, __float_as_int(tmax));
        }
    }
}

// ------------------------------------------------------------------------------------------------ backward
template <int KMAX, int G, int NI>
__global__ void __launch_bounds__(kGrpThreads, KMAX <= 4 ? 6 : 3)
composite_bwd_group_kernel(umhs_composite_in in, umhs_composite_grads_in gin, umhs_composite_grads_out gout, int vecK,
                           int vecB) {
    extern __shared__ float4 smem4[];
    float *smem = reinterpret_cast<float *>(smem4);
    constexpr int RPW = 32 / G;
    constexpr int RPC = kGrpWarps * RPW;
    constexpr bool PREFETCH = KMAX <= 4;
    constexpr int SLOTS = (KMAX * G * NI + kGrpThreads - 1) / kGrpThreads;  // float4 entries of dE per thread
    const int K = in.num_classes, B = in.num_bands;
    const int Bp = (B + 3) & ~3, nb4 = Bp >> 2;
    float *E_s = smem, *M_s = E_s + K * Bp, *dS_s = M_s + 3 * Bp, *mk_s = dS_s + RPC * Bp, *dE_s = mk_s + RPC * KMAX;
    const float4 *E4 = reinterpret_cast<const float4 *>(E_s), *M4 = reinterpret_cast<const float4 *>(M_s);
    float4 *dS4 = reinterpret_cast<float4 *>(dS_s);
    stage_tables(in, K, B, Bp, E_s, M_s);
    for (int i = threadIdx.x; i < K * Bp; i += kGrpThreads) dE_s[i] = 0.f;
    __syncthreads();

    // dE entries (k, float4 band group) owned by this thread; when there are fewer entries than threads the rays of
    // a CTA iteration are split across `jsplit` threads per entry
    const int NE = K * nb4;
    const int jsplit = NE < kGrpThreads ? kGrpThreads / NE : 1;
    int ek[SLOTS], eb[SLOTS];
    float4 dEacc[SLOTS];
    int j0 = 0;
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
        int e = threadIdx.x + kGrpThreads * s;
        bool on = e < NE;
        if (jsplit > 1) {
            on = (s == 0) && (threadIdx.x < NE * jsplit);
            e = threadIdx.x % NE;
            j0 = threadIdx.x / NE;
        }
        ek[s] = on ? e / nb4 : -1;
        eb[s] = on ? e - (e / nb4) * nb4 : 0;
        dEacc[s] = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane & (G - 1), q = lane / G;
    const int rl = warp * RPW + q;  // ray slot within the CTA iteration
    const bool vK = vecK != 0, vB = vecB != 0;
    for (int64_t base = (int64_t)blockIdx.x * RPC; base < in.n_rays; base += (int64_t)gridDim.x * RPC) {
        const int64_t r = base + rl;
        const bool rv = r < in.n_rays;
        int64_t start = 0;
        int cnt = 0;
        if (rv) {
            const longlong2 pi = __ldg(reinterpret_cast<const longlong2 *>(in.packed_info) + r);
            start = pi.x;
            cnt = (int)pi.y;
        }
        // ---- band phase: gradient with respect to the composited spectrum of this ray
        float dl = 0.f;
        if (gin.d_rgb) {
            float lin0 = 0.f, lin1 = 0.f, lin2 = 0.f;
            const float *sprow = gin.spectral + r * B;
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                const int b4 = g + G * i;
                if (b4 < nb4 && rv) {
                    const float4 sp = load_bands(sprow, b4, B, vB);
                    lin0 += dot4(sp, M4[0 * nb4 + b4]);
                    lin1 += dot4(sp, M4[1 * nb4 + b4]);
                    lin2 += dot4(sp, M4[2 * nb4 + b4]);
                }
            }
            lin0 = group_sum<G>(lin0);
            lin1 = group_sum<G>(lin1);
            lin2 = group_sum<G>(lin2);
            if (rv && g < 3) {  // one colour channel per lane
                const float x = g == 0 ? lin0 : (g == 1 ? lin1 : lin2);
                const float gx = gamma_fwd(x);
                const float pass = (gx >= 0.f && gx <= 1.f) ? 1.f : 0.f;
                dl = __ldg(gin.d_rgb + r * 3 + g) * gamma_grad(x) * pass;
            }
        }
        const float dlin0 = __shfl_sync(kFull, dl, 0, G), dlin1 = __shfl_sync(kFull, dl, 1, G),
                    dlin2 = __shfl_sync(kFull, dl, 2, G);
        float4 dS[NI];
        const float *dsrow = gin.d_spectral ? gin.d_spectral + r * B : nullptr;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const int b4 = g + G * i;
            dS[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (b4 < nb4) {
                if (rv && dsrow) dS[i] = load_bands(dsrow, b4, B, vB);
                fma4(dS[i], dlin0, M4[0 * nb4 + b4]);
                fma4(dS[i], dlin1, M4[1 * nb4 + b4]);
                fma4(dS[i], dlin2, M4[2 * nb4 + b4]);
                dS4[rl * nb4 + b4] = dS[i];
            }
        }
        float dmix[KMAX], mixg[KMAX];
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            dmix[k] = mixg[k] = 0.f;
            if (k < K) {
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < NI; ++i) {
                    const int b4 = g + G * i;
                    if (b4 < nb4) s += dot4(dS[i], E4[k * nb4 + b4]);
                }
                dmix[k] = group_sum<G>(s);
            }
        }
        const float dacc = (rv && gin.d_accumulation) ? __ldg(gin.d_accumulation + r) : 0.f;

        // ---- sample phase: one sweep in reverse, suffix sum of dw * w carried backwards
        const float *ts = in.t_starts + start, *te = in.t_ends + start, *sgp = in.sigma + start;
        const float *trp = gin.trans + start;
        const float *dwx = gin.d_weights ? gin.d_weights + start : nullptr;
        const float *scp = in.scalar + start * K, *abp = in.abund + start * K;
        float *dsg = gout.d_sigma + start, *dsc = gout.d_scalar + start * K, *dab = gout.d_abund + start * K;
        const int iters = (__reduce_max_sync(kFull, cnt) + G - 1) / G;
        float rcarry = 0.f;
        SampleRegs<KMAX> cur, nxt;
        if (PREFETCH) {
            const int idx = (iters - 1) * G + g;
            load_sample<KMAX, true>(cur, iters > 0 && idx < cnt, idx, K, vK, ts, te, sgp, trp, scp, abp);
        }
        for (int it = iters - 1; it >= 0; --it) {
            const int idx = it * G + g;
            const bool v = idx < cnt;
            if (PREFETCH)
                load_sample<KMAX, true>(nxt, it > 0 && idx - G < cnt, idx - G, K, vK, ts, te, sgp, trp, scp, abp);
            else
                load_sample<KMAX, true>(cur, v, idx, K, vK, ts, te, sgp, trp, scp, abp);
            const float delta = cur.t1 - cur.t0;
            const float sd = cur.sg * delta;
            const float e = expf(-sd);
            const float T = cur.T;
            const float w = v ? T * (1.0f - e) : 0.f;
            const float tm = (cur.t0 + cur.t1) / 2.0f;
            // scale_gradients_by_distance_squared (ref umhs_model.py:241-242): every field-output gradient x clamp(t^2)
            const float gs = in.grad_scaling ? fminf(fmaxf(tm * tm, 0.f), 1.f) : 1.f;
            float dw = dacc + ((dwx && v) ? __ldg(dwx + idx) : 0.f);
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
                if (k < K) {
                    const float p = cur.sv[k] * cur.av[k];
                    dw += dmix[k] * p;
                    mixg[k] += gs * w * p;
                }
            }
            const float x = dw * w;
            const float rs = group_incl_rscan<G>(x, g);
            const float S = rcarry + (rs - x);
            rcarry += __shfl_sync(kFull, rs, 0, G);
            if (v) {
                dsg[idx] = (dw * T * e - S) * delta * gs;
                float o1[KMAX], o2[KMAX];
                const float wg = w * gs;
#pragma unroll
                for (int k = 0; k < KMAX; ++k) {
                    o1[k] = wg * dmix[k] * cur.av[k];
                    o2[k] = wg * dmix[k] * cur.sv[k];
                }
                store_row<KMAX>(dsc + (int64_t)idx * K, K, vK, o1);
                store_row<KMAX>(dab + (int64_t)idx * K, K, vK, o2);
            }
            if (PREFETCH) cur = nxt;
        }
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            const float mk = (k < K) ? group_sum<G>(mixg[k]) : 0.f;
            if (g == 0) mk_s[rl * KMAX + k] = mk;
        }
        __syncthreads();
        // ---- dE += sum over the rays of this CTA iteration of (sum_n gs w s a)_r (x) dS_r
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            if (ek[s] >= 0) {
                float4 a = dEacc[s];
                for (int j = j0; j < RPC; j += jsplit) fma4(a, mk_s[j * KMAX + ek[s]], dS4[j * nb4 + eb[s]]);
                dEacc[s] = a;
            }
        }
        __syncthreads();
    }
    if (gout.d_endmembers) {
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            if (ek[s] >= 0) {
                float *d = dE_s + ek[s] * Bp + 4 * eb[s];
                if (jsplit > 1) {
                    atomicAdd(d + 0, dEacc[s].x); atomicAdd(d + 1, dEacc[s].y);
                    atomicAdd(d + 2, dEacc[s].z); atomicAdd(d + 3, dEacc[s].w);
                } else {
                    d[0] = dEacc[s].x; d[1] = dEacc[s].y; d[2] = dEacc[s].z; d[3] = dEacc[s].w;
                }
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < K * Bp; i += kGrpThreads) {
            const int k = i / Bp, b = i - k * Bp;
            if (b < B && dE_s[i] != 0.f) atomicAdd(gout.d_endmembers + k * B + b, dE_s[i]);
        }
    }
}

// ------------------------------------------------------------------------------------------------ launchers
static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename Kern>
static int grid_for(Kern kern, size_t smem_bytes, int64_t n_rays, int rays_per_cta, int *grid) {
    int per_sm = 0;
    UMHS_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    UMHS_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kGrpThreads, smem_bytes));
    if (per_sm < 1) per_sm = 1;
    const int64_t want = (n_rays + rays_per_cta - 1) / rays_per_cta;
    const int64_t cap = (int64_t)kNumSMs * per_sm;
    *grid = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
    return UMHS_OK;
}

template <int KMAX, int G, int NI>
static int launch_fwd_group(const umhs_composite_in *in, const umhs_composite_out *out, cudaStream_t st) {
    const int K = in->num_classes, B = in->num_bands, Bp = (B + 3) & ~3;
    const size_t bytes = (size_t)(K * Bp + 3 * Bp + K + 4) * sizeof(float);
    const int vecK = (K % 4 == 0) && aligned16(in->scalar) && aligned16(in->abund) &&
                     (!out->abundances || aligned16(out->abundances)) && (!out->seg_probs || aligned16(out->seg_probs));
    const int vecB = (B % 4 == 0) && (!out->spectral || aligned16(out->spectral)) &&
                     (!in->specular_in || (aligned16(in->specular_in) && (!out->spectral2 || aligned16(out->spectral2))));
    int grid = 1;
    const int rc = grid_for(composite_fwd_group_kernel<KMAX, G, NI>, bytes, in->n_rays, kGrpWarps * (32 / G), &grid);
    if (rc != UMHS_OK) return rc;
    composite_fwd_group_kernel<KMAX, G, NI><<<grid, kGrpThreads, bytes, st>>>(*in, *out, vecK, vecB);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

template <int KMAX, int G, int NI>
static int launch_bwd_group(const umhs_composite_in *in, const umhs_composite_grads_in *gin,
                            const umhs_composite_grads_out *gout, cudaStream_t st) {
    const int K = in->num_classes, B = in->num_bands, Bp = (B + 3) & ~3;
    constexpr int RPC = kGrpWarps * (32 / G);
    const size_t bytes = (size_t)(K * Bp + 3 * Bp + RPC * Bp + RPC * KMAX + K * Bp) * sizeof(float);
    const int vecK = (K % 4 == 0) && aligned16(in->scalar) && aligned16(in->abund) && aligned16(gout->d_scalar) &&
                     aligned16(gout->d_abund);
    const int vecB = (B % 4 == 0) && (!gin->d_spectral || aligned16(gin->d_spectral)) &&
                     (!gin->spectral || aligned16(gin->spectral));
    int grid = 1;
    const int rc = grid_for(composite_bwd_group_kernel<KMAX, G, NI>, bytes, in->n_rays, RPC, &grid);
    if (rc != UMHS_OK) return rc;
    composite_bwd_group_kernel<KMAX, G, NI><<<grid, kGrpThreads, bytes, st>>>(*in, *gin, *gout, vecK, vecB);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

// G lanes per ray x NI float4 band groups per lane cover B <= 16 G NI bands
#define UMHS_DISPATCH_GROUP(FN, ...)                                            \
    do {                                                                        \
        const int K_ = in->num_classes, B_ = in->num_bands;                     \
        if (B_ <= 64) {                                                         \
            if (K_ <= 4) return FN<4, 8, 2>(__VA_ARGS__);                       \
            if (K_ <= 8) return FN<8, 8, 2>(__VA_ARGS__);                       \
            return FN<16, 8, 2>(__VA_ARGS__);                                   \
        } else if (B_ <= 128) {                                                 \
            if (K_ <= 4) return FN<4, 8, 4>(__VA_ARGS__);                       \
            if (K_ <= 8) return FN<8, 8, 4>(__VA_ARGS__);                       \
            return FN<16, 8, 4>(__VA_ARGS__);                                   \
        } else if (B_ <= 256) {                                                 \
            if (K_ <= 4) return FN<4, 16, 4>(__VA_ARGS__);                      \
            if (K_ <= 8) return FN<8, 16, 4>(__VA_ARGS__);                      \
            return FN<16, 16, 4>(__VA_ARGS__);                                  \
        } else {                                                                \
            if (K_ <= 4) return FN<4, 32, 4>(__VA_ARGS__);                      \
            if (K_ <= 8) return FN<8, 32, 4>(__VA_ARGS__);                      \
            return FN<16, 32, 4>(__VA_ARGS__);                                  \
        }                                                                       \
    } while (0)

int composite_group_fwd(const umhs_composite_in *in, const umhs_composite_out *out, cudaStream_t st) {
    UMHS_DISPATCH_GROUP(launch_fwd_group, in, out, st);
}

int composite_group_bwd(const umhs_composite_in *in, const umhs_composite_grads_in *gin,
                        const umhs_composite_grads_out *gout, cudaStream_t st) {
    UMHS_DISPATCH_GROUP(launch_bwd_group, in, gin, gout, st);
}

}  // namespace umhs
