// optim.cu -- fused dense Adam over one flat fp32 parameter buffer (+ endmember clamp), slab reduce.
//
// torch.optim.Adam semantics as configured at ref umhsnerf/umhs_config.py:59-64 (lr 2e-2,
// eps 1e-15, betas (0.9, 0.999), no weight decay); the endmember clamp of ref
// umhsnerf/umhs_model.py:568-570 is folded in for [clamp_begin, clamp_end).
// HBM-bound streaming: 16 B read (p,g,m,v) + 12-16 B written per parameter, float4 vectorised.
#include "common.cuh"

namespace umhs {

struct AdamArgs {
    float *p, *g, *m, *v;
    int64_t n;
    float lr, beta1, beta2, eps, bc1, bc2_sqrt;
    const float *inv_scale, *found_inf;
    const float *hyper;  // optional device [lr, 1-beta1^t, sqrt(1-beta2^t)]: per-step scalars of a replayed CUDA graph
    int64_t clamp_begin, clamp_end;
    int zero_grad;
};

__device__ __forceinline__ void adam_one(float &p, float &g, float &m, float &v, const AdamArgs &a, float gs,
                                         int64_t i) {
    const float gr = g * gs;
    m = a.beta1 * m + (1.0f - a.beta1) * gr;
    v = a.beta2 * v + (1.0f - a.beta2) * gr * gr;
    const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
    p -= (a.lr / a.bc1) * (m / denom);
    if (i >= a.clamp_begin && i < a.clamp_end) p = fminf(fmaxf(p, 0.0f), 1.0f);
    if (a.zero_grad) g = 0.0f;
}

__global__ void __launch_bounds__(256) adam_kernel(AdamArgs a) {
    if (a.hyper) {
        a.lr = __ldg(a.hyper);
        a.bc1 = __ldg(a.hyper + 1);
        a.bc2_sqrt = __ldg(a.hyper + 2);
    }
    if (a.found_inf && *a.found_inf != 0.0f) {
        // GradScaler semantics: skip the step; grads are still cleared for the next iteration
        if (a.zero_grad)
            for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n;
                 i += (int64_t)gridDim.x * blockDim.x)
                a.g[i] = 0.0f;
        return;
    }
    const float gs = a.inv_scale ? *a.inv_scale : 1.0f;
    const int64_t n4 = a.n / 4;
    float4 *p4 = reinterpret_cast<float4 *>(a.p), *g4 = reinterpret_cast<float4 *>(a.g);
    float4 *m4 = reinterpret_cast<float4 *>(a.m), *v4 = reinterpret_cast<float4 *>(a.v);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 p = p4[i], g = g4[i], m = m4[i], v = v4[i];
        adam_one(p.x, g.x, m.x, v.x, a, gs, 4 * i + 0);
        adam_one(p.y, g.y, m.y, v.y, a, gs, 4 * i + 1);
        adam_one(p.z, g.z, m.z, v.z, a, gs, 4 * i + 2);
        adam_one(p.w, g.w, m.w, v.w, a, gs, 4 * i + 3);
        p4[i] = p;
        m4[i] = m;
        v4[i] = v;
        if (a.zero_grad) g4[i] = g;
    }
    for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n;
         i += (int64_t)gridDim.x * blockDim.x)
        adam_one(a.p[i], a.g[i], a.m[i], a.v[i], a, gs, i);
}

__global__ void reduce_slabs_kernel(const float *__restrict__ slabs, int64_t n_slabs, int64_t stride, int64_t n,
                                    float *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int64_t k = 0; k < n_slabs; ++k) s += slabs[k * stride + i];
        out[i] += s;
    }
}

}  // namespace umhs

using namespace umhs;

extern "C" int umhs_adam_step(float *param, float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float lr,
                              float beta1, float beta2, float eps, int32_t step, const float *inv_scale,
                              const float *found_inf, int64_t clamp_begin, int64_t clamp_end, int32_t zero_grad,
                              void *stream) {
    UMHS_CHECK_ARG(param && grad && exp_avg && exp_avg_sq, "null argument");
    UMHS_CHECK_ARG(step >= 1, "step counts from 1");
    UMHS_CHECK_ARG(((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16 == 0,
                   "buffers must be 16-byte aligned");
    if (n <= 0) return UMHS_OK;
    AdamArgs a;
    a.p = param; a.g = grad; a.m = exp_avg; a.v = exp_avg_sq; a.n = n;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps;
    a.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
    a.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
    a.inv_scale = inv_scale; a.found_inf = found_inf; a.hyper = nullptr;
    a.clamp_begin = clamp_begin; a.clamp_end = clamp_end; a.zero_grad = zero_grad;
    const int64_t want = (n / 4 + 255) / 256;
    const int grid = (int)(want < kNumSMs * 8 ? (want < 1 ? 1 : want) : kNumSMs * 8);
    adam_kernel<<<grid, 256, 0, as_stream(stream)>>>(a);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

extern "C" int umhs_adam_step_dev(float *param, float *grad, float *exp_avg, float *exp_avg_sq, int64_t n,
                                  float beta1, float beta2, float eps, const float *hyper, const float *inv_scale,
                                  const float *found_inf, int64_t clamp_begin, int64_t clamp_end, int32_t zero_grad,
                                  void *stream) {
    UMHS_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && hyper, "null argument");
    UMHS_CHECK_ARG(((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16 == 0,
                   "buffers must be 16-byte aligned");
    if (n <= 0) return UMHS_OK;
    AdamArgs a;
    a.p = param; a.g = grad; a.m = exp_avg; a.v = exp_avg_sq; a.n = n;
    a.lr = 0.f; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.bc1 = 1.f; a.bc2_sqrt = 1.f;
    a.inv_scale = inv_scale; a.found_inf = found_inf; a.hyper = hyper;
    a.clamp_begin = clamp_begin; a.clamp_end = clamp_end; a.zero_grad = zero_grad;
    const int64_t want = (n / 4 + 255) / 256;
    const int grid = (int)(want < kNumSMs * 8 ? (want < 1 ? 1 : want) : kNumSMs * 8);
    adam_kernel<<<grid, 256, 0, as_stream(stream)>>>(a);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

extern "C" int umhs_reduce_slabs(const float *slabs, int64_t n_slabs, int64_t stride, int64_t n, float *out,
                                 void *stream) {
    UMHS_CHECK_ARG(slabs && out, "null argument");
    if (n <= 0 || n_slabs <= 0) return UMHS_OK;
    reduce_slabs_kernel<<<div_up(n, 256) < kNumSMs * 8 ? div_up(n, 256) : kNumSMs * 8, 256, 0, as_stream(stream)>>>(
        slabs, n_slabs, stride, n, out);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}
