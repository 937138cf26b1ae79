// unfused.cu -- the same arithmetic as composite.cu cut at the reference's MODULE boundaries, so that
// UMHSField.get_outputs (per-sample spectra, ref umhsnerf/umhs_field.py:229-251),
// nerfacc.render_weight_from_density (ref umhsnerf/umhs_model.py:246-252) and
// SpectralRenderer.forward (ref umhsnerf/umhs_renderer.py:15-30) remain callable on their own.
// The training / eval step of UMHSModel does NOT use these (it uses the fused kernels).
#include "common.cuh"

namespace umhs {

constexpr int kUWarps = 8;
constexpr int kUThreads = kUWarps * 32;

__device__ __forceinline__ void load16(const float *h, float (&hv)[16]) {
    const float4 *p = reinterpret_cast<const float4 *>(h);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 v = __ldg(p + q);
        hv[4 * q] = v.x; hv[4 * q + 1] = v.y; hv[4 * q + 2] = v.z; hv[4 * q + 3] = v.w;
    }
}

// one warp per sample, lanes over bands
__global__ void __launch_bounds__(kUThreads)
unmix_fwd_kernel(int64_t n_total, int K, int B, int specular, const float *__restrict__ scalar,
                 const float *__restrict__ abund, const float *__restrict__ s1, const float *__restrict__ spec_hidden,
                 const float *__restrict__ E, const float *__restrict__ W, const float *__restrict__ bias,
                 float *__restrict__ spec, float *__restrict__ specular_out) {
    const int lane = threadIdx.x & 31;
    for (int64_t n = (int64_t)blockIdx.x * kUWarps + (threadIdx.x >> 5); n < n_total;
         n += (int64_t)gridDim.x * kUWarps) {
        float c[UMHS_MAX_CLASSES];
#pragma unroll
        for (int k = 0; k < UMHS_MAX_CLASSES; ++k)
            c[k] = k < K ? __ldg(scalar + n * K + k) * __ldg(abund + n * K + k) : 0.f;
        float hv[16], s1v = 0.f;
        if (specular) {
            load16(spec_hidden + n * 16, hv);
            s1v = __ldg(s1 + n);
        }
        for (int b = lane; b < B; b += 32) {
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < UMHS_MAX_CLASSES; ++k)
                if (k < K) a += c[k] * __ldg(E + k * B + b);
            spec[n * B + b] = a;
            if (specular && specular_out) {
                float z = __ldg(bias + b);
#pragma unroll
                for (int t = 0; t < 16; ++t) z += __ldg(W + b * 16 + t) * hv[t];
                specular_out[n * B + b] = s1v * (1.0f / (1.0f + expf(-z)));
            }
        }
    }
}

__global__ void __launch_bounds__(kUThreads)
unmix_bwd_kernel(int64_t n_total, int K, int B, int specular, const float *__restrict__ scalar,
                 const float *__restrict__ abund, const float *__restrict__ s1, const float *__restrict__ spec_hidden,
                 const float *__restrict__ E, const float *__restrict__ W, const float *__restrict__ bias,
                 const float *__restrict__ d_spec, const float *__restrict__ d_specular, float *__restrict__ d_scalar,
                 float *__restrict__ d_abund, float *__restrict__ d_s1, float *__restrict__ d_spec_hidden,
                 float *__restrict__ d_E, float *__restrict__ d_W, float *__restrict__ d_bias) {
    extern __shared__ float smem[];
    float *dE_s = smem;               // K*B
    float *dW_s = dE_s + K * B;       // B*17 (specular)
    float *dB_s = dW_s + (specular ? B * 17 : 0);
    for (int i = threadIdx.x; i < K * B; i += kUThreads) dE_s[i] = 0.f;
    if (specular) {
        for (int i = threadIdx.x; i < 17 * B; i += kUThreads) dW_s[i] = 0.f;
        for (int i = threadIdx.x; i < B; i += kUThreads) dB_s[i] = 0.f;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (int64_t n = (int64_t)blockIdx.x * kUWarps + (threadIdx.x >> 5); n < n_total;
         n += (int64_t)gridDim.x * kUWarps) {
        float sv[UMHS_MAX_CLASSES], av[UMHS_MAX_CLASSES], c[UMHS_MAX_CLASSES];
#pragma unroll
        for (int k = 0; k < UMHS_MAX_CLASSES; ++k) {
            sv[k] = k < K ? __ldg(scalar + n * K + k) : 0.f;
            av[k] = k < K ? __ldg(abund + n * K + k) : 0.f;
            c[k] = 0.f;
        }
        float hv[16], dh[16], s1v = 0.f, q = 0.f;
#pragma unroll
        for (int t = 0; t < 16; ++t) dh[t] = 0.f;
        if (specular) {
            load16(spec_hidden + n * 16, hv);
            s1v = __ldg(s1 + n);
        }
        for (int b = lane; b < B; b += 32) {
            const float g = __ldg(d_spec + n * B + b);
#pragma unroll
            for (int k = 0; k < UMHS_MAX_CLASSES; ++k) {
                if (k < K) {
                    c[k] += g * __ldg(E + k * B + b);
                    const float v = sv[k] * av[k] * g;
                    if (v != 0.f) atomicAdd(dE_s + k * B + b, v);
                }
            }
            if (specular && d_specular) {
                const float gs = __ldg(d_specular + n * B + b);
                float z = __ldg(bias + b);
#pragma unroll
                for (int t = 0; t < 16; ++t) z += __ldg(W + b * 16 + t) * hv[t];
                const float sp = 1.0f / (1.0f + expf(-z));
                q += gs * sp;
                const float u = s1v * gs * sp * (1.0f - sp);
                if (u != 0.f) {
                    atomicAdd(dB_s + b, u);
#pragma unroll
                    for (int t = 0; t < 16; ++t) {
                        atomicAdd(dW_s + b * 17 + t, u * hv[t]);
                        dh[t] += u * __ldg(W + b * 16 + t);
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < UMHS_MAX_CLASSES; ++k) {
            if (k < K) {
                const float ck = warp_sum(c[k]);
                if (lane == 0) {
                    d_scalar[n * K + k] = av[k] * ck;
                    d_abund[n * K + k] = sv[k] * ck;
                }
            }
        }
        if (specular) {
            q = warp_sum(q);
            float mine = 0.f;
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                const float s = warp_sum(dh[t]);
                if (lane == t) mine = s;
            }
            if (lane == 0 && d_s1) d_s1[n] = q;
            if (lane < 16 && d_spec_hidden) d_spec_hidden[n * 16 + lane] = mine;
        }
    }
    __syncthreads();
    if (d_E)
        for (int i = threadIdx.x; i < K * B; i += kUThreads)
            if (dE_s[i] != 0.f) atomicAdd(d_E + i, dE_s[i]);
    if (specular) {
        if (d_W)
            for (int i = threadIdx.x; i < 16 * B; i += kUThreads) {
                const float v = dW_s[(i >> 4) * 17 + (i & 15)];
                if (v != 0.f) atomicAdd(d_W + i, v);
            }
        if (d_bias)
            for (int i = threadIdx.x; i < B; i += kUThreads)
                if (dB_s[i] != 0.f) atomicAdd(d_bias + i, dB_s[i]);
    }
}

// one warp per ray: w = T * alpha, T = exp(-exclusive_sum(sigma * delta))
__global__ void __launch_bounds__(kUThreads)
weights_fwd_kernel(const int64_t *__restrict__ packed_info, int64_t n_rays, const float *__restrict__ t0,
                   const float *__restrict__ t1, const float *__restrict__ sigma, float *__restrict__ weights,
                   float *__restrict__ trans) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * kUWarps + (threadIdx.x >> 5);
    if (r >= n_rays) return;
    const int64_t start = __ldg(packed_info + 2 * r), cnt = __ldg(packed_info + 2 * r + 1);
    float carry = 0.f;
    for (int64_t c0 = 0; c0 < cnt; c0 += 32) {
        const int64_t n = start + c0 + lane;
        const bool v = c0 + lane < cnt;
        const float sd = v ? __ldg(sigma + n) * (__ldg(t1 + n) - __ldg(t0 + n)) : 0.f;
        const float incl = warp_incl_scan(sd, lane);
        const float T = expf(-(carry + (incl - sd)));
        carry += __shfl_sync(kFull, incl, 31);
        if (v) {
            if (weights) weights[n] = T * (1.0f - expf(-sd));
            if (trans) trans[n] = T;
        }
    }
}

// d_sigma from d_weights: reverse segmented scan (chunks walked backwards; T from the saved forward output)
__global__ void __launch_bounds__(kUThreads)
weights_bwd_kernel(const int64_t *__restrict__ packed_info, int64_t n_rays, const float *__restrict__ t0,
                   const float *__restrict__ t1, const float *__restrict__ sigma, const float *__restrict__ trans,
                   const float *__restrict__ d_weights, float *__restrict__ d_sigma) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * kUWarps + (threadIdx.x >> 5);
    if (r >= n_rays) return;
    const int64_t start = __ldg(packed_info + 2 * r), cnt = __ldg(packed_info + 2 * r + 1);
    float rcarry = 0.f;
    for (int64_t c0 = ((cnt - 1) / 32) * 32; c0 >= 0 && cnt > 0; c0 -= 32) {
        const int64_t n = start + c0 + lane;
        const bool v = c0 + lane < cnt;
        float sd = 0.f, T = 0.f, dw = 0.f, dt = 0.f;
        if (v) {
            dt = __ldg(t1 + n) - __ldg(t0 + n);
            sd = __ldg(sigma + n) * dt;
            T = __ldg(trans + n);
            dw = __ldg(d_weights + n);
        }
        const float e = expf(-sd);
        const float x = dw * T * (1.0f - e);
        const float rs = warp_incl_rscan(x, lane);
        const float S = rcarry + (rs - x);
        rcarry += __shfl_sync(kFull, rs, 0);
        if (v) d_sigma[n] = (dw * T * e - S) * dt;
    }
}

// out[ray[n], c] += w[n] * v[n, c]  (index_add_ semantics; one warp per sample, lanes over channels)
__global__ void __launch_bounds__(kUThreads)
accumulate_fwd_kernel(const float *__restrict__ w, const float *__restrict__ v, const int64_t *__restrict__ ray,
                      int64_t n_total, int C, float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    for (int64_t n = (int64_t)blockIdx.x * kUWarps + (threadIdx.x >> 5); n < n_total;
         n += (int64_t)gridDim.x * kUWarps) {
        const int64_t r = __ldg(ray + n);
        const float wn = __ldg(w + n);
        for (int c = lane; c < C; c += 32) {
            const float x = v ? wn * __ldg(v + n * C + c) : wn;
            if (x != 0.f) atomicAdd(out + r * C + c, x);
        }
    }
}

__global__ void __launch_bounds__(kUThreads)
accumulate_bwd_kernel(const float *__restrict__ w, const float *__restrict__ v, const int64_t *__restrict__ ray,
                      int64_t n_total, int C, const float *__restrict__ d_out, float *__restrict__ d_w,
                      float *__restrict__ d_v) {
    const int lane = threadIdx.x & 31;
    for (int64_t n = (int64_t)blockIdx.x * kUWarps + (threadIdx.x >> 5); n < n_total;
         n += (int64_t)gridDim.x * kUWarps) {
        const int64_t r = __ldg(ray + n);
        const float wn = __ldg(w + n);
        float acc = 0.f;
        for (int c = lane; c < C; c += 32) {
            const float g = __ldg(d_out + r * C + c);
            acc += v ? g * __ldg(v + n * C + c) : g;
            if (d_v) d_v[n * C + c] = wn * g;
        }
        if (d_w) {
            acc = warp_sum(acc);
            if (lane == 0) d_w[n] = acc;
        }
    }
}

static int ugrid(int64_t n) {
    const int64_t want = (n + kUWarps - 1) / kUWarps;
    const int64_t cap = (int64_t)kNumSMs * 16;
    return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

}  // namespace umhs

using namespace umhs;

extern "C" int umhs_unmix_fwd(int64_t n, int32_t K, int32_t B, int32_t specular, const float *scalar,
                              const float *abund, const float *s1, const float *spec_hidden, const float *endmembers,
                              const float *dir_w1, const float *dir_b1, float *spec, float *specular_out,
                              void *stream) {
    UMHS_CHECK_ARG(scalar && abund && endmembers && spec, "null argument");
    UMHS_CHECK_ARG(K >= 1 && K <= UMHS_MAX_CLASSES && B >= 1 && B <= UMHS_MAX_BANDS, "K or B out of range");
    UMHS_CHECK_ARG(!specular || (s1 && spec_hidden && dir_w1 && dir_b1), "specular inputs missing");
    if (n <= 0) return UMHS_OK;
    unmix_fwd_kernel<<<ugrid(n), kUThreads, 0, as_stream(stream)>>>(n, K, B, specular, scalar, abund, s1, spec_hidden,
                                                                    endmembers, dir_w1, dir_b1, spec, specular_out);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

extern "C" int umhs_unmix_bwd(int64_t n, int32_t K, int32_t B, int32_t specular, const float *scalar,
                              const float *abund, const float *s1, const float *spec_hidden, const float *endmembers,
                              const float *dir_w1, const float *dir_b1, const float *d_spec, const float *d_specular,
                              float *d_scalar, float *d_abund, float *d_s1, float *d_spec_hidden, float *d_endmembers,
                              float *d_dir_w1, float *d_dir_b1, void *stream) {
    UMHS_CHECK_ARG(scalar && abund && endmembers && d_spec && d_scalar && d_abund, "null argument");
    UMHS_CHECK_ARG(K >= 1 && K <= UMHS_MAX_CLASSES && B >= 1 && B <= UMHS_MAX_BANDS, "K or B out of range");
    UMHS_CHECK_ARG(!specular || (s1 && spec_hidden && dir_w1 && dir_b1), "specular inputs missing");
    if (n <= 0) return UMHS_OK;
    const size_t bytes = sizeof(float) * ((size_t)K * B + (specular ? 18 * (size_t)B : 0));
    UMHS_CHECK_CUDA(cudaFuncSetAttribute(unmix_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    int grid = ugrid(n);
    if (grid > kNumSMs * 2) grid = kNumSMs * 2;
    unmix_bwd_kernel<<<grid, kUThreads, bytes, as_stream(stream)>>>(n, K, B, specular, scalar, abund, s1, spec_hidden,
                                                                    endmembers, dir_w1, dir_b1, d_spec, d_specular,
                                                                    d_scalar, d_abund, d_s1, d_spec_hidden,
                                                                    d_endmembers, d_dir_w1, d_dir_b1);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

extern "C" int umhs_render_weights_fwd(const int64_t *packed_info, int64_t n_rays, const float *t_starts,
                                       const float *t_ends, const float *sigma, float *weights, float *trans,
                                       void *stream) {
    UMHS_CHECK_ARG(packed_info && t_starts && t_ends && sigma, "null argument");
    if (n_rays <= 0) return UMHS_OK;
    weights_fwd_kernel<<<div_up(n_rays, kUWarps), kUThreads, 0, as_stream(stream)>>>(packed_info, n_rays, t_starts,
                                                                                     t_ends, sigma, weights, trans);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

extern "C" int umhs_render_weights_bwd(const int64_t *packed_info, int64_t n_rays, const float *t_starts,
                                       const float *t_ends, const float *sigma, const float *trans,
                                       const float *d_weights, float *d_sigma, void *stream) {
    UMHS_CHECK_ARG(packed_info && t_starts && t_ends && sigma && trans && d_weights && d_sigma, "null argument");
    if (n_rays <= 0) return UMHS_OK;
    weights_bwd_kernel<<<div_up(n_rays, kUWarps), kUThreads, 0, as_stream(stream)>>>(
        packed_info, n_rays, t_starts, t_ends, sigma, trans, d_weights, d_sigma);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

extern "C" int umhs_accumulate_fwd(const float *weights, const float *values, const int64_t *ray_indices, int64_t n,
                                   int32_t C, float *out, void *stream) {
    UMHS_CHECK_ARG(weights && ray_indices && out && C >= 1, "bad argument");
    if (n <= 0) return UMHS_OK;
    accumulate_fwd_kernel<<<ugrid(n), kUThreads, 0, as_stream(stream)>>>(weights, values, ray_indices, n, C, out);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

extern "C" int umhs_accumulate_bwd(const float *weights, const float *values, const int64_t *ray_indices, int64_t n,
                                   int32_t C, const float *d_out, float *d_weights, float *d_values, void *stream) {
    UMHS_CHECK_ARG(weights && ray_indices && d_out && C >= 1, "bad argument");
    if (n <= 0) return UMHS_OK;
    accumulate_bwd_kernel<<<ugrid(n), kUThreads, 0, as_stream(stream)>>>(weights, values, ray_indices, n, C, d_out,
                                                                        d_weights, d_values);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}
