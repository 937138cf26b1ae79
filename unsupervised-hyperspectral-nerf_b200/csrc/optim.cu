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
    // Two float4 groups per thread and iteration (six independent 16-byte loads in flight before the first use: the
    // kernel was latency-bound at 47 % of the DRAM read peak with three).  The moments are streamed (ld/st .cs,
    // evict-first): 128 MiB that nothing reads before the next optimiser step should not push the hash table and its
    // gradient (read again by the next forward / scatter) out of the 126 MB L2.
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    auto untouched = [&](const float4 &g, const float4 &m, const float4 &v, int64_t i) {
        // Entries that never received a gradient (g = m = v = 0: most of the coarse, spatially dense hash levels
        // outside the scene) are left untouched by Adam -- bit-identical to running the update (m, v stay 0, the step
        // is 0 / (0 + eps) = 0) -- so their parameter read and all four writes are skipped: 48 instead of 128 bytes.
        // (clamped entries still need the clamp applied)
        return g.x == 0.f && g.y == 0.f && g.z == 0.f && g.w == 0.f && m.x == 0.f && m.y == 0.f && m.z == 0.f &&
               m.w == 0.f && v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f &&
               !(4 * i + 3 >= a.clamp_begin && 4 * i < a.clamp_end);
    };
    auto update = [&](float4 p, float4 g, float4 m, float4 v, int64_t i) {
        adam_one(p.x, g.x, m.x, v.x, a, gs, 4 * i + 0);
        adam_one(p.y, g.y, m.y, v.y, a, gs, 4 * i + 1);
        adam_one(p.z, g.z, m.z, v.z, a, gs, 4 * i + 2);
        adam_one(p.w, g.w, m.w, v.w, a, gs, 4 * i + 3);
        p4[i] = p;
        __stcs(m4 + i, m);
        __stcs(v4 + i, v);
        if (a.zero_grad) g4[i] = g;
    };
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += 2 * stride) {
        const int64_t j = i + stride;
        const bool has_j = j < n4;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 g0 = g4[i], m0 = __ldcs(m4 + i), v0 = __ldcs(v4 + i);
        const float4 g1 = has_j ? g4[j] : z, m1 = has_j ? __ldcs(m4 + j) : z, v1 = has_j ? __ldcs(v4 + j) : z;
        const bool t0 = !untouched(g0, m0, v0, i), t1 = has_j && !untouched(g1, m1, v1, j);
        float4 p0 = z, p1 = z;
        if (t0) p0 = p4[i];
        if (t1) p1 = p4[j];
        if (t0) update(p0, g0, m0, v0, i);
        if (t1) update(p1, g1, m1, v1, j);
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

// =====================================================================================================
// Data-parallel optimiser step fused with its gradient exchange over NVLink peer memory.
//
// Every rank's flat gradient and flat parameter buffers are peer-mapped (symmetric memory).  Rank r owns the shard
// [r n / W, (r + 1) n / W): it READS that shard of every rank's gradient through NVLink (the reduce-scatter), runs
// Adam on it with its own slice of the optimiser state, and WRITES the updated parameters into every rank's
// parameter buffer (the all-gather) -- one kernel, no intermediate HBM round trip, 1/W of the optimiser traffic per
// GPU, and the same 2 (W - 1) / W n bytes on the wire as an all-reduce.  Two cross-GPU barriers bracket it
// (umhs_dp_barrier): gradients complete before, parameters landed / gradients consumed after.
// =====================================================================================================
struct DpAdamArgs {
    const float *peer_grad[UMHS_MAX_RANKS];
    float *peer_param[UMHS_MAX_RANKS];
    float *m, *v;
    const float *mc_grad;  // optional NVLS multicast addresses of the same buffers: one multimem.ld_reduce returns the
    float *mc_param;       // sum over all replicas (reduced in the switch), one multimem.st writes all replicas
    int64_t begin, end;  // this rank's shard (multiples of 4)
    int world, rank;
    float beta1, beta2, eps, inv_world;
    const float *hyper;
    int64_t clamp_begin, clamp_end;
};

__global__ void __launch_bounds__(256) dp_adam_kernel(DpAdamArgs a) {
    const float lr = __ldg(a.hyper), bc1 = __ldg(a.hyper + 1), bc2_sqrt = __ldg(a.hyper + 2);
    const float4 *own_p = reinterpret_cast<const float4 *>(a.peer_param[a.rank]);
    float4 *m4 = reinterpret_cast<float4 *>(a.m), *v4 = reinterpret_cast<float4 *>(a.v);
    for (int64_t i = a.begin / 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.end / 4;
         i += (int64_t)gridDim.x * blockDim.x) {
        // reduce-scatter: this shard of every rank's gradient (peer loads first: they have the longest latency)
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.mc_grad != nullptr) {
            asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                         : "=f"(g.x), "=f"(g.y), "=f"(g.z), "=f"(g.w)
                         : "l"(a.mc_grad + 4 * i)
                         : "memory");
        } else {
#pragma unroll
            for (int w = 0; w < UMHS_MAX_RANKS; ++w) {
                if (w < a.world) {
                    const float4 x = reinterpret_cast<const float4 *>(a.peer_grad[w])[i];
                    g.x += x.x; g.y += x.y; g.z += x.z; g.w += x.w;
                }
            }
        }
        float4 m = m4[i], v = v4[i];
        // untouched entries (no gradient on any rank, no optimiser state): Adam leaves them bit-identical, every
        // replica already holds them -- skip the update and, above all, the all-gather writes (see adam_kernel)
        if (g.x == 0.f && g.y == 0.f && g.z == 0.f && g.w == 0.f && m.x == 0.f && m.y == 0.f && m.z == 0.f &&
            m.w == 0.f && v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f &&
            !(4 * i + 3 >= a.clamp_begin && 4 * i < a.clamp_end))
            continue;
        float4 p = own_p[i];
        float *pp = &p.x, *gg = &g.x, *mm = &m.x, *vv = &v.x;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float gr = gg[c] * a.inv_world;
            mm[c] = a.beta1 * mm[c] + (1.0f - a.beta1) * gr;
            vv[c] = a.beta2 * vv[c] + (1.0f - a.beta2) * gr * gr;
            const float denom = sqrtf(vv[c]) / bc2_sqrt + a.eps;
            pp[c] -= (lr / bc1) * (mm[c] / denom);
            const int64_t idx = 4 * i + c;
            if (idx >= a.clamp_begin && idx < a.clamp_end) pp[c] = fminf(fmaxf(pp[c], 0.0f), 1.0f);
        }
        m4[i] = m;
        v4[i] = v;
        // all-gather: the updated shard goes to every rank's parameter buffer
        if (a.mc_param != nullptr) {
            asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(a.mc_param + 4 * i),
                         "f"(p.x), "f"(p.y), "f"(p.z), "f"(p.w)
                         : "memory");
        } else {
#pragma unroll
            for (int w = 0; w < UMHS_MAX_RANKS; ++w)
                if (w < a.world) reinterpret_cast<float4 *>(a.peer_param[w])[i] = p;
        }
    }
}

// Cross-GPU barrier on peer-mapped signal pads: pad[channel][src] of rank dst counts the barriers src has entered.
// Epochs only grow, so nothing is ever reset.  One CTA; thread w < world talks to rank w.
__global__ void dp_barrier_kernel(DpAdamArgs a_unused, uint32_t *const *pads, int rank, int world, int channel,
                                  uint32_t *epoch) {
    (void)a_unused;
    __shared__ uint32_t e_s;
    if (threadIdx.x == 0) e_s = ++(*epoch);
    __syncthreads();
    const uint32_t e = e_s;
    __threadfence_system();  // this GPU's earlier kernels' writes are ordered before the signal
    if ((int)threadIdx.x < world) {
        const int w = threadIdx.x;
        uint32_t *remote = pads[w] + channel * UMHS_MAX_RANKS + rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(e) : "memory");
        const uint32_t *local = pads[rank] + channel * UMHS_MAX_RANKS + w;
        uint32_t seen;
        unsigned long long spins = 0;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(local) : "memory");
            if (++spins > (1ull << 34)) __trap();  // a missing peer must fail loudly, not hang forever
        } while ((int32_t)(seen - e) < 0);
    }
    __syncthreads();
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

extern "C" int umhs_dp_barrier(uint32_t *const *signal_pads, int32_t rank, int32_t world, int32_t channel,
                               uint32_t *epoch, void *stream) {
    UMHS_CHECK_ARG(signal_pads && epoch && world >= 1 && world <= UMHS_MAX_RANKS && rank >= 0 && rank < world &&
                       channel >= 0 && channel < 8, "bad argument");
    dp_barrier_kernel<<<1, 32, 0, as_stream(stream)>>>(DpAdamArgs{}, signal_pads, rank, world, channel, epoch);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

extern "C" int umhs_dp_adam_step(const float *const *peer_grads, float *const *peer_params, float *exp_avg,
                                 float *exp_avg_sq, int64_t n, int32_t rank, int32_t world, float beta1, float beta2,
                                 float eps, const float *hyper, int64_t clamp_begin, int64_t clamp_end,
                                 const float *mc_grad, float *mc_param, void *stream) {
    return umhs_dp_adam_step_range(peer_grads, peer_params, exp_avg, exp_avg_sq, n, 0, n, rank, world, beta1, beta2, eps,
                                   hyper, clamp_begin, clamp_end, mc_grad, mc_param, stream);
}

extern "C" int umhs_dp_adam_step_range(const float *const *peer_grads, float *const *peer_params, float *exp_avg,
                                       float *exp_avg_sq, int64_t n, int64_t range_begin, int64_t range_end,
                                       int32_t rank, int32_t world, float beta1, float beta2, float eps,
                                       const float *hyper, int64_t clamp_begin, int64_t clamp_end, const float *mc_grad,
                                       float *mc_param, void *stream) {
    UMHS_CHECK_ARG(peer_grads && peer_params && exp_avg && exp_avg_sq && hyper, "null argument");
    UMHS_CHECK_ARG(range_begin >= 0 && range_begin <= range_end && range_end <= n && range_begin % 4 == 0 &&
                       range_end % 4 == 0, "the range must be inside the buffer, on multiples of 4");
    UMHS_CHECK_ARG(((uintptr_t)mc_grad | (uintptr_t)mc_param) % 16 == 0, "multicast addresses must be 16-byte aligned");
    UMHS_CHECK_ARG(world >= 1 && world <= UMHS_MAX_RANKS && rank >= 0 && rank < world, "bad rank / world");
    UMHS_CHECK_ARG(n % 4 == 0, "the flat buffer length must be a multiple of 4");
    if (n <= 0) return UMHS_OK;
    DpAdamArgs a;
    for (int w = 0; w < UMHS_MAX_RANKS; ++w) {
        a.peer_grad[w] = w < world ? peer_grads[w] : nullptr;   // HOST arrays of device pointers
        a.peer_param[w] = w < world ? peer_params[w] : nullptr;
        UMHS_CHECK_ARG(w >= world || (a.peer_grad[w] && a.peer_param[w] && ((uintptr_t)a.peer_grad[w] % 16 == 0) &&
                                      ((uintptr_t)a.peer_param[w] % 16 == 0)), "bad peer pointer");
    }
    // this rank's shard of the range: ceil(range / 4 / world) float4 groups
    const int64_t n4 = (range_end - range_begin) / 4, per = (n4 + world - 1) / world;
    a.begin = range_begin + 4 * (per * rank < n4 ? per * rank : n4);
    a.end = range_begin + 4 * (per * (rank + 1) < n4 ? per * (rank + 1) : n4);
    if (a.begin == a.end) return UMHS_OK;
    a.m = exp_avg; a.v = exp_avg_sq; a.world = world; a.rank = rank;
    a.mc_grad = mc_grad; a.mc_param = mc_param;
    a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.inv_world = 1.0f / (float)world; a.hyper = hyper;
    a.clamp_begin = clamp_begin; a.clamp_end = clamp_end;
    const int64_t want = ((a.end - a.begin) / 4 + 255) / 256;
    const int grid = (int)(want < kNumSMs * 8 ? (want < 1 ? 1 : want) : kNumSMs * 8);
    dp_adam_kernel<<<grid, 256, 0, as_stream(stream)>>>(a);
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
