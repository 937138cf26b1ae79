// composite.cu -- per-ray alpha compositing over packed samples fused with abundance x endmember
// unmixing, the specular head's last layer, spec->XYZ->sRGB and the cluster probe.
//
// Replaces ref umhsnerf/umhs_model.py:245-313 (nerfacc pack weights, depth / accumulation
// renderers, SpectralRenderer.forward x4 -> ref umhsnerf/umhs_renderer.py:15-30, ColourSystem
// -> ref umhsnerf/utils/spec_to_rgb.py:112-127, ClusterLookup -> ref
// umhsnerf/utils/clusterprobe.py:17-38) and the linear mixing model of ref
// umhsnerf/umhs_field.py:229-251.
//
// Key identity: the mixing model is linear in the endmembers, so
//     sum_n w_n sum_k s_nk a_nk E[k,b]  ==  sum_k (sum_n w_n s_nk a_nk) E[k,b]
// i.e. composite the K products first, unmix once per RAY.  The reference's [N,B,K] and [N,B]
// intermediates never exist; per-sample traffic is 12 + 8K bytes.
//
// One warp per ray.  Lanes run over the ray's samples for the transmittance (warp-shuffle
// segmented scan with a carry across 32-sample chunks; the backward is the reverse scan) and
// over the bands for the per-ray epilogue.  E, the colour matrix and the specular last layer
// are staged in shared memory per CTA; dE / dW accumulate in shared memory and are flushed with
// one atomic per element per CTA.
#include <math_constants.h>

#include "composite.cuh"

namespace umhs {

constexpr int kCompWarps = 8;
constexpr int kCompThreads = kCompWarps * 32;
constexpr int kMaxChunks = 128;  // chunk carries kept in smem: rays up to 4096 samples (longer: recomputed)

struct SmemLayout {
    int E, M, W, Bv, En, dE, dW, dB, carry, hbuf, total;
};

__host__ __device__ inline SmemLayout make_layout(int K, int B, bool specular, bool bwd) {
    SmemLayout s;
    int o = 0;
    s.E = o; o += K * B;
    s.M = o; o += 3 * B;
    s.W = o; o += specular ? B * 17 : 0;
    s.Bv = o; o += specular ? B : 0;
    s.En = o; o += K;
    s.dE = o; o += bwd ? K * B : 0;
    s.dW = o; o += (bwd && specular) ? B * 17 : 0;
    s.dB = o; o += (bwd && specular) ? B : 0;
    s.carry = o; o += bwd ? kCompWarps * kMaxChunks : 0;
    o = (o + 3) & ~3;  // 16-byte aligned: float4 rows
    s.hbuf = o; o += specular ? kCompWarps * 32 * 16 : 0;  // per warp: the spec_hidden rows of one 32-sample chunk
    s.total = o;
    return s;
}

// The spec_hidden rows [m,16] of a chunk of m <= 32 consecutive samples -> this warp's shared-memory buffer: one
// coalesced pass (<= 4 float4 per lane, all in flight) instead of one exposed 64-byte global load per sample.
__device__ __forceinline__ void stage_chunk_h(const float *rows, int m, float4 *hb, int lane) {
    const float4 *src = reinterpret_cast<const float4 *>(rows);
    __syncwarp();  // the previous chunk's readers are done
    for (int q = lane; q < 4 * m; q += 32) hb[q] = __ldg(src + q);
    __syncwarp();
}
__device__ __forceinline__ void read_h16(const float4 *hb, int j, float (&hv)[16]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 v = hb[4 * j + q];
        hv[4 * q + 0] = v.x; hv[4 * q + 1] = v.y; hv[4 * q + 2] = v.z; hv[4 * q + 3] = v.w;
    }
}

template <int KMAX, int BPL>
__global__ void __launch_bounds__(kCompThreads, (BPL == 4 && KMAX <= 8) ? 2 : 1)
composite_fwd_kernel(umhs_composite_in in, umhs_composite_out out) {
    extern __shared__ float smem[];
    const int K = in.num_classes, B = in.num_bands;
    const bool specular = in.specular != 0;
    const SmemLayout L = make_layout(K, B, specular, false);
    float *E_s = smem + L.E, *M_s = smem + L.M, *W_s = smem + L.W, *b_s = smem + L.Bv, *En_s = smem + L.En;
    for (int i = threadIdx.x; i < K * B; i += kCompThreads) E_s[i] = __ldg(in.endmembers + i);
    for (int i = threadIdx.x; i < 3 * B; i += kCompThreads) M_s[i] = __ldg(in.rgb_matrix + i);
    if (specular) {
        for (int i = threadIdx.x; i < 16 * B; i += kCompThreads) W_s[(i >> 4) * 17 + (i & 15)] = __ldg(in.dir_w1 + i);
        for (int i = threadIdx.x; i < B; i += kCompThreads) b_s[i] = __ldg(in.dir_b1 + i);
    }
    __syncthreads();
    if (threadIdx.x < K) {
        float s = 0.f;
        for (int b = 0; b < B; ++b) s += E_s[threadIdx.x * B + b] * E_s[threadIdx.x * B + b];
        En_s[threadIdx.x] = fmaxf(sqrtf(s), 1e-12f);
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4 *hbuf = reinterpret_cast<float4 *>(smem + L.hbuf) + warp * 128;
    // B <= 128: the 16 weights (+ bias) of each of this lane's 4 bands live in registers for the whole kernel -- the
    // per-sample sigmoid(W h + b) term then needs no shared-memory reads at all (it was one LDS per FMA)
    constexpr bool WREG = (BPL == 4);
    float W_r[WREG ? 4 : 1][16], b_r[WREG ? 4 : 1];
    if (WREG && specular) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int b = lane + 32 * i;
            b_r[i] = b < B ? b_s[b] : 0.f;
#pragma unroll
            for (int q = 0; q < 16; ++q) W_r[i][q] = b < B ? W_s[b * 17 + q] : 0.f;
        }
    }
    float tmin = CUDART_INF_F, tmax = -CUDART_INF_F;
    for (int64_t r = (int64_t)blockIdx.x * kCompWarps + warp; r < in.n_rays; r += (int64_t)gridDim.x * kCompWarps) {
        const int64_t start = __ldg(in.packed_info + 2 * r), cnt = __ldg(in.packed_info + 2 * r + 1);
        float carry = 0.f, acc = 0.f, dep = 0.f;
        float mix[KMAX], ab[KMAX], sacc[BPL];
#pragma unroll
        for (int k = 0; k < KMAX; ++k) mix[k] = ab[k] = 0.f;
#pragma unroll
        for (int i = 0; i < BPL; ++i) sacc[i] = 0.f;
        for (int64_t c0 = 0; c0 < cnt; c0 += 32) {
            const int64_t n = start + c0 + lane;
            const bool v = c0 + lane < cnt;
            float t0 = 0.f, t1 = 0.f, sg = 0.f;
            if (v) {
                t0 = __ldg(in.t_starts + n);
                t1 = __ldg(in.t_ends + n);
                sg = __ldg(in.sigma + n);
            }
            const float sd = sg * (t1 - t0);
            const float incl = warp_incl_scan(sd, lane);
            const float T = expf(-(carry + (incl - sd)));
            const float alpha = 1.0f - expf(-sd);
            const float w = v ? T * alpha : 0.f;
            carry += __shfl_sync(kFull, incl, 31);
            if (v) {
                if (out.weights) out.weights[n] = w;
                if (out.trans) out.trans[n] = T;
                const float tm = (t0 + t1) / 2.0f;
                tmin = fminf(tmin, tm);
                tmax = fmaxf(tmax, tm);
                acc += w;
                dep += w * tm;
#pragma unroll
                for (int k = 0; k < KMAX; ++k) {
                    if (k < K) {
                        const float s = __ldg(in.scalar + n * K + k), a = __ldg(in.abund + n * K + k);
                        mix[k] += w * s * a;
                        ab[k] += w * a;
                    }
                }
            }
            if (specular) {
                const float ws = v ? w * __ldg(in.s1 + n) : 0.f;
                const int m = (int)((cnt - c0) < 32 ? (cnt - c0) : 32);
                stage_chunk_h(in.spec_hidden + (start + c0) * 16, m, hbuf, lane);
                for (int j = 0; j < m; ++j) {
                    const float wj = __shfl_sync(kFull, ws, j);
                    if (wj == 0.f) continue;  // warp-uniform
                    float hv[16];
                    read_h16(hbuf, j, hv);
#pragma unroll
                    for (int i = 0; i < BPL; ++i) {
                        const int b = lane + 32 * i;
                        if (b < B) {
                            float z = WREG ? b_r[i < 4 ? i : 0] : b_s[b];
#pragma unroll
                            for (int q = 0; q < 16; ++q)
                                z += (WREG ? W_r[i < 4 ? i : 0][q] : W_s[b * 17 + q]) * hv[q];
                            sacc[i] += wj * sigmoidf(z);
                        }
                    }
                }
            }
        }
        acc = warp_sum(acc);
        dep = warp_sum(dep);
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            if (k < K) {
                mix[k] = warp_sum(mix[k]);
                ab[k] = warp_sum(ab[k]);
            }
        }
        // ---- per-ray epilogue, lanes over bands
        float lin[3] = {0.f, 0.f, 0.f}, nrm = 0.f, dot[KMAX];
#pragma unroll
        for (int k = 0; k < KMAX; ++k) dot[k] = 0.f;
#pragma unroll
        for (int i = 0; i < BPL; ++i) {
            const int b = lane + 32 * i;
            if (b < B) {
                float s2 = 0.f;
#pragma unroll
                for (int k = 0; k < KMAX; ++k)
                    if (k < K) s2 += mix[k] * E_s[k * B + b];
                const float sp = specular ? s2 + sacc[i] : s2;
                if (out.spectral) out.spectral[r * B + b] = sp;
                if (specular) {
                    if (out.spectral2) out.spectral2[r * B + b] = s2;
                    if (out.specular) out.specular[r * B + b] = sacc[i];
                }
                lin[0] += sp * M_s[3 * b + 0];
                lin[1] += sp * M_s[3 * b + 1];
                lin[2] += sp * M_s[3 * b + 2];
                nrm += sp * sp;
#pragma unroll
                for (int k = 0; k < KMAX; ++k)
                    if (k < K) dot[k] += sp * E_s[k * B + b];
            }
        }
        lin[0] = warp_sum(lin[0]);
        lin[1] = warp_sum(lin[1]);
        lin[2] = warp_sum(lin[2]);
        nrm = fmaxf(sqrtf(warp_sum(nrm)), 1e-12f);
        float mx = -CUDART_INF_F;
        int arg = 0;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            if (k < K) {
                dot[k] = warp_sum(dot[k]) / (nrm * En_s[k]);
                if (dot[k] > mx) {
                    mx = dot[k];
                    arg = k;
                }
            }
        }
        if (lane == 0) {
            if (out.accumulation) out.accumulation[r] = acc;
            if (out.depth) out.depth[r] = dep / (acc + 1e-10f);
            if (out.rgb) {
#pragma unroll
                for (int c = 0; c < 3; ++c) out.rgb[r * 3 + c] = fminf(fmaxf(gamma_fwd(lin[c]), 0.f), 1.f);
            }
            float den = 0.f, pr[KMAX];
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
                pr[k] = (k < K) ? expf(0.2f * (dot[k] - mx)) : 0.f;
                den += pr[k];
            }
            // argmax of the probabilities == first argmax of the inner products
            const float gate = acc > 0.5f ? 1.f : 0.f;
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
                if (k < K) {
                    if (out.abundances) out.abundances[r * K + k] = ab[k];
                    if (out.seg_probs) out.seg_probs[r * K + k] = pr[k] / den;
                }
            }
            if (out.seg_raw) out.seg_raw[r] = (float)arg * gate;
            if (out.seg_pred && in.palette) {
#pragma unroll
                for (int c = 0; c < 3; ++c) out.seg_pred[r * 3 + c] = __ldg(in.palette + 3 * arg + c) * gate;
            }
        }
    }
    if (out.t_minmax) {
        tmin = -warp_max(-tmin);
        tmax = warp_max(tmax);
        if (lane == 0 && tmin <= tmax) {
            // non-negative floats order like their bit patterns
            atomicMin(reinterpret_cast<int *>(out.t_minmax), __float_as_int(tmin));
            atomicMax(reinterpret_cast<int *>(out.t_minmax) + 1, __float_as_int(tmax));
        }
    }
}

__global__ void depth_clip_kernel(float *depth, int64_t n, const float *t_minmax) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) depth[i] = fminf(fmaxf(depth[i], t_minmax[0]), t_minmax[1]);
}

// exclusive prefix of sigma*delta at the start of chunk c (from smem carries or recomputed)
__device__ __forceinline__ float chunk_prefix(const umhs_composite_in &in, int64_t start, int64_t cnt, int c,
                                              const float *carries, int lane) {
    if (c < kMaxChunks) return carries[c];
    float run = carries[kMaxChunks - 1];
    for (int cc = kMaxChunks - 1; cc < c; ++cc) {
        const int64_t n = start + (int64_t)cc * 32 + lane;
        float sd = 0.f;
        if ((int64_t)cc * 32 + lane < cnt) sd = __ldg(in.sigma + n) * (__ldg(in.t_ends + n) - __ldg(in.t_starts + n));
        run += warp_sum(sd);
    }
    return run;
}

template <int KMAX, int BPL>
__global__ void __launch_bounds__(kCompThreads)
composite_bwd_kernel(umhs_composite_in in, umhs_composite_grads_in gin, umhs_composite_grads_out gout) {
    extern __shared__ float smem[];
    const int K = in.num_classes, B = in.num_bands;
    const bool specular = in.specular != 0;
    const SmemLayout L = make_layout(K, B, specular, true);
    float *E_s = smem + L.E, *M_s = smem + L.M, *W_s = smem + L.W, *b_s = smem + L.Bv;
    float *dE_s = smem + L.dE, *dW_s = smem + L.dW, *dB_s = smem + L.dB;
    for (int i = threadIdx.x; i < K * B; i += kCompThreads) {
        E_s[i] = __ldg(in.endmembers + i);
        dE_s[i] = 0.f;
    }
    for (int i = threadIdx.x; i < 3 * B; i += kCompThreads) M_s[i] = __ldg(in.rgb_matrix + i);
    if (specular) {
        for (int i = threadIdx.x; i < 16 * B; i += kCompThreads) W_s[(i >> 4) * 17 + (i & 15)] = __ldg(in.dir_w1 + i);
        for (int i = threadIdx.x; i < 17 * B; i += kCompThreads) dW_s[i] = 0.f;
        for (int i = threadIdx.x; i < B; i += kCompThreads) {
            b_s[i] = __ldg(in.dir_b1 + i);
            dB_s[i] = 0.f;
        }
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *carries = smem + L.carry + warp * kMaxChunks;
    float4 *hbuf = reinterpret_cast<float4 *>(smem + L.hbuf) + warp * 128;
    // B <= 128 (4 bands per lane): dW / db of this lane's bands accumulate in registers for the whole kernel (DREG);
    // with K <= 4 there is room for the weights themselves too (WREG), so the per-sample term reads no shared memory
    constexpr bool DREG = (BPL == 4);
    constexpr bool WREG = (BPL == 4) && (KMAX <= 4);
    float W_r[WREG ? 4 : 1][16], b_r[WREG ? 4 : 1], dW_r[DREG ? 4 : 1][16], dB_r[DREG ? 4 : 1];
    if (DREG) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            dB_r[i] = 0.f;
#pragma unroll
            for (int t = 0; t < 16; ++t) dW_r[i][t] = 0.f;
        }
    }
    if (WREG && specular) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int b = lane + 32 * i;
            b_r[i] = b < B ? b_s[b] : 0.f;
#pragma unroll
            for (int t = 0; t < 16; ++t) W_r[i][t] = b < B ? W_s[b * 17 + t] : 0.f;
        }
    }
    for (int64_t r = (int64_t)blockIdx.x * kCompWarps + warp; r < in.n_rays; r += (int64_t)gridDim.x * kCompWarps) {
        const int64_t start = __ldg(in.packed_info + 2 * r), cnt = __ldg(in.packed_info + 2 * r + 1);
        if (cnt == 0) continue;
        // ---- band phase: gradient w.r.t. the composited spectrum
        float dS[BPL], dS2[BPL];
        {
            float lin[3] = {0.f, 0.f, 0.f};
            if (gin.d_rgb) {
#pragma unroll
                for (int i = 0; i < BPL; ++i) {
                    const int b = lane + 32 * i;
                    if (b < B) {
                        const float sp = __ldg(gin.spectral + r * B + b);
                        lin[0] += sp * M_s[3 * b + 0];
                        lin[1] += sp * M_s[3 * b + 1];
                        lin[2] += sp * M_s[3 * b + 2];
                    }
                }
            }
            float dlin[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                dlin[c] = 0.f;
                if (gin.d_rgb) {
                    const float x = warp_sum(lin[c]);
                    const float g = gamma_fwd(x);
                    const float pass = (g >= 0.f && g <= 1.f) ? 1.f : 0.f;
                    dlin[c] = __ldg(gin.d_rgb + r * 3 + c) * gamma_grad(x) * pass;
                }
            }
#pragma unroll
            for (int i = 0; i < BPL; ++i) {
                const int b = lane + 32 * i;
                dS[i] = dS2[i] = 0.f;
                if (b < B) {
                    float g = gin.d_spectral ? __ldg(gin.d_spectral + r * B + b) : 0.f;
                    g += dlin[0] * M_s[3 * b + 0] + dlin[1] * M_s[3 * b + 1] + dlin[2] * M_s[3 * b + 2];
                    dS[i] = g;
                    dS2[i] = g + ((specular && gin.d_spectral2) ? __ldg(gin.d_spectral2 + r * B + b) : 0.f);
                }
            }
        }
        float dmix[KMAX], mixg[KMAX];
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            dmix[k] = mixg[k] = 0.f;
            if (k < K) {
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < BPL; ++i) {
                    const int b = lane + 32 * i;
                    if (b < B) s += dS2[i] * E_s[k * B + b];
                }
                dmix[k] = warp_sum(s);
            }
        }
        const float dacc = gin.d_accumulation ? __ldg(gin.d_accumulation + r) : 0.f;

        // ---- pass A: exclusive prefix of sigma*delta at each chunk start
        const int nchunks = (int)((cnt + 31) / 32);
        {
            float run = 0.f;
            for (int c = 0; c < nchunks; ++c) {
                if (c < kMaxChunks && lane == 0) carries[c] = run;
                if (c >= kMaxChunks - 1) break;
                const int64_t n = start + (int64_t)c * 32 + lane;
                float sd = 0.f;
                if ((int64_t)c * 32 + lane < cnt)
                    sd = __ldg(in.sigma + n) * (__ldg(in.t_ends + n) - __ldg(in.t_starts + n));
                const float incl = warp_incl_scan(sd, lane);
                run += __shfl_sync(kFull, incl, 31);
            }
            __syncwarp();
        }
        // ---- pass B: chunks in reverse, suffix sum of dw*w carried backwards
        float rcarry = 0.f;
        for (int c = nchunks - 1; c >= 0; --c) {
            const int64_t c0 = (int64_t)c * 32;
            const int64_t n = start + c0 + lane;
            const bool v = c0 + lane < cnt;
            float t0 = 0.f, t1 = 0.f, sg = 0.f;
            if (v) {
                t0 = __ldg(in.t_starts + n);
                t1 = __ldg(in.t_ends + n);
                sg = __ldg(in.sigma + n);
            }
            const float sd = sg * (t1 - t0);
            const float incl = warp_incl_scan(sd, lane);
            const float pre = chunk_prefix(in, start, cnt, c, carries, lane);
            const float T = expf(-(pre + (incl - sd)));
            const float e = expf(-sd);
            const float w = v ? T * (1.0f - e) : 0.f;
            const float tm = (t0 + t1) / 2.0f;
            const float gs = in.grad_scaling ? fminf(fmaxf(tm * tm, 0.f), 1.f) : 1.f;
            float dw = dacc;
            float sv[KMAX], av[KMAX];
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
                sv[k] = av[k] = 0.f;
                if (k < K && v) {
                    sv[k] = __ldg(in.scalar + n * K + k);
                    av[k] = __ldg(in.abund + n * K + k);
                    dw += dmix[k] * sv[k] * av[k];
                    mixg[k] += gs * w * sv[k] * av[k];
                }
            }
            if (specular) {
                const float s1v = v ? __ldg(in.s1 + n) : 0.f;
                const float wsg = w * s1v * gs;  // weight of dz accumulation for this lane's sample
                float q_mine = 0.f;
                const int m = (int)((cnt - c0) < 32 ? (cnt - c0) : 32);
                stage_chunk_h(in.spec_hidden + (start + c0) * 16, m, hbuf, lane);
                for (int j = 0; j < m; ++j) {
                    const float wj = __shfl_sync(kFull, wsg, j);
                    float hv[16];
                    read_h16(hbuf, j, hv);
                    float q = 0.f, dh[16];
#pragma unroll
                    for (int t = 0; t < 16; ++t) dh[t] = 0.f;
#pragma unroll
                    for (int i = 0; i < BPL; ++i) {
                        const int b = lane + 32 * i;
                        if (b < B) {
                            float z = WREG ? b_r[i < 4 ? i : 0] : b_s[b];
#pragma unroll
                            for (int t = 0; t < 16; ++t)
                                z += (WREG ? W_r[i < 4 ? i : 0][t] : W_s[b * 17 + t]) * hv[t];
                            const float sp = sigmoidf(z);
                            q += dS[i] * sp;
                            const float u = wj * dS[i] * sp * (1.0f - sp);  // d loss / d z_b (grad-scaled)
                            if (DREG) {
                                // this lane's 4 x 17 entries of dW / db accumulate in registers over all samples of
                                // all rays of the warp (they were 17 shared-memory atomics per band per sample)
                                dB_r[i < 4 ? i : 0] += u;
#pragma unroll
                                for (int t = 0; t < 16; ++t) {
                                    dW_r[i < 4 ? i : 0][t] += u * hv[t];
                                    dh[t] += u * (WREG ? W_r[i < 4 ? i : 0][t] : W_s[b * 17 + t]);
                                }
                            } else if (u != 0.f) {
                                atomicAdd(dB_s + b, u);
#pragma unroll
                                for (int t = 0; t < 16; ++t) {
                                    atomicAdd(dW_s + b * 17 + t, u * hv[t]);
                                    dh[t] += u * W_s[b * 17 + t];
                                }
                            }
                        }
                    }
                    q = warp_sum(q);
                    // d h[t] = sum over bands (lanes): butterfly transpose-reduce, 16 shuffles instead of 80; lane L ends
                    // up with t = bits (L4 L3 L2 L1) of its index, both lanes of a pair hold the sum
                    {
#pragma unroll
                        for (int half = 8, off = 16; half >= 1; half >>= 1, off >>= 1) {
                            const bool up = (lane & off) != 0;
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                if (k < half) {
                                    const float send = up ? dh[k] : dh[k + half];
                                    const float keep = up ? dh[k + half] : dh[k];
                                    dh[k] = keep + __shfl_xor_sync(kFull, send, off);
                                }
                            }
                        }
                        dh[0] += __shfl_xor_sync(kFull, dh[0], 1);
                        const int t_mine = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 +
                                           ((lane >> 1) & 1);
                        if ((lane & 1) == 0 && gout.d_spec_hidden)
                            gout.d_spec_hidden[(start + c0 + j) * 16 + t_mine] = dh[0];
                    }
                    if (lane == j) q_mine = q;
                }
                dw += s1v * q_mine;
                if (v && gout.d_s1) gout.d_s1[n] = w * q_mine * gs;
            }
            const float x = dw * w;
            const float rs = warp_incl_rscan(x, lane);
            const float S = rcarry + (rs - x);
            rcarry += __shfl_sync(kFull, rs, 0);
            if (v) {
                const float dsd = dw * T * e - S;
                gout.d_sigma[n] = dsd * (t1 - t0) * gs;
#pragma unroll
                for (int k = 0; k < KMAX; ++k) {
                    if (k < K) {
                        gout.d_scalar[n * K + k] = w * dmix[k] * av[k] * gs;
                        gout.d_abund[n * K + k] = w * dmix[k] * sv[k] * gs;
                    }
                }
            }
        }
        // ---- dE += (sum_n gs w s a) (x) dS2
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            if (k < K) {
                const float mk = warp_sum(mixg[k]);
#pragma unroll
                for (int i = 0; i < BPL; ++i) {
                    const int b = lane + 32 * i;
                    if (b < B) atomicAdd(dE_s + k * B + b, mk * dS2[i]);
                }
            }
        }
    }
    if (DREG && specular) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int b = lane + 32 * i;
            if (b < B) {
                if (dB_r[i] != 0.f) atomicAdd(dB_s + b, dB_r[i]);
#pragma unroll
                for (int t = 0; t < 16; ++t)
                    if (dW_r[i][t] != 0.f) atomicAdd(dW_s + b * 17 + t, dW_r[i][t]);
            }
        }
    }
    __syncthreads();
    if (gout.d_endmembers)
        for (int i = threadIdx.x; i < K * B; i += kCompThreads)
            if (dE_s[i] != 0.f) atomicAdd(gout.d_endmembers + i, dE_s[i]);
    if (specular) {
        if (gout.d_dir_w1)
            for (int i = threadIdx.x; i < 16 * B; i += kCompThreads) {
                const float v = dW_s[(i >> 4) * 17 + (i & 15)];
                if (v != 0.f) atomicAdd(gout.d_dir_w1 + i, v);
            }
        if (gout.d_dir_b1)
            for (int i = threadIdx.x; i < B; i += kCompThreads)
                if (dB_s[i] != 0.f) atomicAdd(gout.d_dir_b1 + i, dB_s[i]);
    }
}

static int check_in(const umhs_composite_in *in) {
    if (!in) return -1;
    if (in->num_classes < 1 || in->num_classes > UMHS_MAX_CLASSES) return -1;
    if (in->num_bands < 1 || in->num_bands > UMHS_MAX_BANDS) return -1;
    if (!in->packed_info || !in->t_starts || !in->t_ends || !in->sigma || !in->scalar || !in->abund) return -1;
    if (!in->endmembers || !in->rgb_matrix) return -1;
    if (in->specular && (!in->s1 || !in->spec_hidden || !in->dir_w1 || !in->dir_b1)) return -1;
    if (in->specular && (reinterpret_cast<uintptr_t>(in->spec_hidden) & 15u)) return -1;  // rows are read as float4
    return 0;
}

// persistent grid: as many CTAs as fit the device at once (the specular kernels hold their weights / gradient
// accumulators in registers: 1-2 CTAs per SM), fewer when there are fewer rays
template <typename Kern>
static int comp_grid(Kern kern, size_t smem_bytes, int64_t n_rays, int max_per_sm, int *grid) {
    int per_sm = 0;
    UMHS_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    UMHS_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kCompThreads, smem_bytes));
    per_sm = per_sm < 1 ? 1 : (per_sm > max_per_sm ? max_per_sm : per_sm);
    const int64_t want = (n_rays + kCompWarps - 1) / kCompWarps;
    const int64_t cap = (int64_t)kNumSMs * per_sm;
    *grid = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
    return UMHS_OK;
}

template <int KMAX, int BPL>
static int launch_fwd(const umhs_composite_in *in, const umhs_composite_out *out, cudaStream_t st) {
    const SmemLayout L = make_layout(in->num_classes, in->num_bands, in->specular != 0, false);
    const size_t bytes = (size_t)L.total * sizeof(float);
    int grid = 1;
    const int rc = comp_grid(composite_fwd_kernel<KMAX, BPL>, bytes, in->n_rays, 8, &grid);
    if (rc != UMHS_OK) return rc;
    composite_fwd_kernel<KMAX, BPL><<<grid, kCompThreads, bytes, st>>>(*in, *out);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

template <int KMAX, int BPL>
static int launch_bwd(const umhs_composite_in *in, const umhs_composite_grads_in *gin,
                      const umhs_composite_grads_out *gout, cudaStream_t st) {
    const SmemLayout L = make_layout(in->num_classes, in->num_bands, in->specular != 0, true);
    const size_t bytes = (size_t)L.total * sizeof(float);
    // fewer, fatter CTAs: each flushes K*B (+17B) atomics at the end
    int grid = 1;
    const int rc = comp_grid(composite_bwd_kernel<KMAX, BPL>, bytes, in->n_rays, 2, &grid);
    if (rc != UMHS_OK) return rc;
    composite_bwd_kernel<KMAX, BPL><<<grid, kCompThreads, bytes, st>>>(*in, *gin, *gout);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

}  // namespace umhs

using namespace umhs;

#define UMHS_DISPATCH_KB(FN, ...)                                  \
    do {                                                           \
        const int K_ = in->num_classes, B_ = in->num_bands;        \
        if (B_ <= 128) {                                           \
            if (K_ <= 4) return FN<4, 4>(__VA_ARGS__);             \
            if (K_ <= 8) return FN<8, 4>(__VA_ARGS__);             \
            return FN<16, 4>(__VA_ARGS__);                         \
        } else {                                                   \
            if (K_ <= 4) return FN<4, 16>(__VA_ARGS__);            \
            if (K_ <= 8) return FN<8, 16>(__VA_ARGS__);            \
            return FN<16, 16>(__VA_ARGS__);                        \
        }                                                          \
    } while (0)

extern "C" int umhs_composite_fwd(const umhs_composite_in *in, const umhs_composite_out *out, void *stream) {
    UMHS_CHECK_ARG(check_in(in) == 0, "bad composite input");
    UMHS_CHECK_ARG(out != nullptr, "null output");
    if (in->n_rays <= 0) return UMHS_OK;
    if (!in->specular) return composite_group_fwd(in, out, as_stream(stream));  // lane-group form (composite_group.cu)
    UMHS_DISPATCH_KB(launch_fwd, in, out, as_stream(stream));
}

extern "C" int umhs_depth_clip(float *depth, int64_t n_rays, const float *t_minmax, void *stream) {
    UMHS_CHECK_ARG(depth && t_minmax, "null argument");
    if (n_rays <= 0) return UMHS_OK;
    depth_clip_kernel<<<div_up(n_rays, 256), 256, 0, as_stream(stream)>>>(depth, n_rays, t_minmax);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

// ---- chunk-wise depth clip (evaluation).  The reference renders an image as ceil(H W / 512) separate forward calls
// (ref umhs_model.py:605-616, eval_num_rays_per_chunk: ref umhs_config.py:50) and nerfstudio's DepthRenderer clips each
// call's expected depth to that CALL's [min, max] of the sample mid-points; a fused whole-image pass has to reproduce the
// per-chunk bounds.  Samples are sorted along their ray, so a ray's extremes are its first and last sample.
// chunk(ray r) = (ray_ids ? ray_ids[r] : r) / chunk_rays  (ray_ids: row-major pixel ids when rays arrive in tile order)
__global__ void depth_chunk_bounds_kernel(int64_t R, const int64_t *__restrict__ packed_info, const float *__restrict__ t0,
                                          const float *__restrict__ t1, const int64_t *__restrict__ ray_ids,
                                          int64_t chunk_rays, float *__restrict__ bounds) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const int64_t start = packed_info[2 * r], cnt = packed_info[2 * r + 1];
    if (cnt <= 0) return;
    const int64_t c = (ray_ids ? ray_ids[r] : r) / chunk_rays;
    const float lo = (t0[start] + t1[start]) / 2.0f, hi = (t0[start + cnt - 1] + t1[start + cnt - 1]) / 2.0f;
    // non-negative floats order like their bit patterns
    atomicMin(reinterpret_cast<int *>(bounds) + 2 * c, __float_as_int(lo));
    atomicMax(reinterpret_cast<int *>(bounds) + 2 * c + 1, __float_as_int(hi));
}
__global__ void depth_chunk_clip_kernel(int64_t R, const int64_t *__restrict__ ray_ids, int64_t chunk_rays,
                                        const float *__restrict__ bounds, float *__restrict__ depth) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const int64_t c = (ray_ids ? ray_ids[r] : r) / chunk_rays;
    float lo = bounds[2 * c], hi = bounds[2 * c + 1];
    // a call without any sample: nerfstudio's VolumetricSampler fabricates one sample at t0 = t1 = 1
    if (lo > hi) lo = hi = 1.0f;
    depth[r] = fminf(fmaxf(depth[r], lo), hi);
}

extern "C" int umhs_depth_clip_chunks(float *depth, int64_t n_rays, const int64_t *packed_info, const float *t_starts,
                                      const float *t_ends, const int64_t *ray_ids, int64_t chunk_rays, float *bounds,
                                      void *stream) {
    UMHS_CHECK_ARG(depth && packed_info && t_starts && t_ends && bounds && chunk_rays > 0, "bad argument");
    if (n_rays <= 0) return UMHS_OK;
    depth_chunk_bounds_kernel<<<div_up(n_rays, 256), 256, 0, as_stream(stream)>>>(n_rays, packed_info, t_starts, t_ends,
                                                                                 ray_ids, chunk_rays, bounds);
    UMHS_CHECK_LAUNCH();
    depth_chunk_clip_kernel<<<div_up(n_rays, 256), 256, 0, as_stream(stream)>>>(n_rays, ray_ids, chunk_rays, bounds, depth);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

extern "C" int umhs_composite_bwd(const umhs_composite_in *in, const umhs_composite_grads_in *gin,
                                  const umhs_composite_grads_out *gout, void *stream) {
    UMHS_CHECK_ARG(check_in(in) == 0, "bad composite input");
    UMHS_CHECK_ARG(gin && gout, "null argument");
    UMHS_CHECK_ARG(gout->d_sigma && gout->d_scalar && gout->d_abund, "d_sigma/d_scalar/d_abund are required");
    UMHS_CHECK_ARG(!gin->d_rgb || gin->spectral, "d_rgb needs the forward spectral output");
    UMHS_CHECK_ARG(!in->specular || (gout->d_s1 && gout->d_spec_hidden), "specular needs d_s1/d_spec_hidden");
    if (in->n_rays <= 0) return UMHS_OK;
    // lane-group form when the forward's transmittance is at hand; otherwise the warp-per-ray kernel recomputes it
    if (!in->specular && gin->trans) return composite_group_bwd(in, gin, gout, as_stream(stream));
    UMHS_DISPATCH_KB(launch_bwd, in, gin, gout, as_stream(stream));
}
