// hash_scatter.cuh -- per-sample scatter of the hash-encoding gradient into the table gradient (all levels), shared by
// the stand-alone scatter kernel (hash_grid.cu) and the scatter warps of the tensor-core MLP backward (field_mlp_tc.cu).
#pragma once
#include "common.cuh"

namespace umhs {

// Reductions without a return value, spelled out: nvcc emits atom.global.add and leaves the atom -> red conversion to
// ptxas, which does NOT do it in a kernel that contains setmaxnreg (the fused backward: every reduction became a
// returning scalar ATOMG and the scatter warps waited a round trip on each -- 3 x slower).
__device__ __forceinline__ void red_add(float *p, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void red_add(float2 *p, float2 v) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void red_add(float4 *p, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

// Corner order follows the reference: f0=(c,c,c) f1=(c,f,c) f2=(f,f,c) f3=(f,c,c)
//                                     f4=(c,c,f) f5=(c,f,f) f6=(f,f,f) f7=(f,c,f)
// returns true when the x-neighbours of this sample form aligned row pairs (see the header): (f3,f0) (f2,f1) (f7,f4) (f6,f5)
__device__ __forceinline__ bool level_setup(float x, float y, float z, float scale, uint32_t tmask, uint32_t base,
                                            uint32_t (&idx)[8], float &ox, float &oy, float &oz, uint32_t (&cell)[3]) {
    const float sx = x * scale, sy = y * scale, sz = z * scale;
    const float flx = floorf(sx), fly = floorf(sy), flz = floorf(sz);
    const uint32_t cx = (uint32_t)(int)ceilf(sx), cy = (uint32_t)(int)ceilf(sy), cz = (uint32_t)(int)ceilf(sz);
    const uint32_t fx = (uint32_t)(int)flx, fy = (uint32_t)(int)fly, fz = (uint32_t)(int)flz;
    ox = sx - flx;
    oy = sy - fly;
    oz = sz - flz;
    // low log2(T) bits of the int64 product == low bits of the uint32-wrapped product
    const uint32_t hcy = cy * 2654435761u, hfy = fy * 2654435761u;
    const uint32_t hcz = cz * 805459861u, hfz = fz * 805459861u;
    idx[0] = ((cx ^ hcy ^ hcz) & tmask) + base;
    idx[1] = ((cx ^ hfy ^ hcz) & tmask) + base;
    idx[2] = ((fx ^ hfy ^ hcz) & tmask) + base;
    idx[3] = ((fx ^ hcy ^ hcz) & tmask) + base;
    idx[4] = ((cx ^ hcy ^ hfz) & tmask) + base;
    idx[5] = ((cx ^ hfy ^ hfz) & tmask) + base;
    idx[6] = ((fx ^ hfy ^ hfz) & tmask) + base;
    idx[7] = ((fx ^ hcy ^ hfz) & tmask) + base;
    // floor AND ceil: two samples with equal floors, one of them exactly on a cell face (ceil == floor), differ in
    // corners.  Injective for |coordinate| < 2^30.
    cell[0] = (fx << 1) | (uint32_t)(cx != fx);
    cell[1] = (fy << 1) | (uint32_t)(cy != fy);
    cell[2] = (fz << 1) | (uint32_t)(cz != fz);
    return ((fx & 1u) == 0u) && (cx != fx);
}
__device__ __forceinline__ bool level_setup(float x, float y, float z, float scale, uint32_t tmask, uint32_t base,
                                            uint32_t (&idx)[8], float &ox, float &oy, float &oz) {
    uint32_t cell[3];
    return level_setup(x, y, z, scale, tmask, base, idx, ox, oy, oz, cell);
}

// One float2 reduction into the table.  Lanes of a warp hold consecutive samples of a ray, so at
// coarse levels runs of adjacent lanes hit the same entry: a segmented warp scan sums each run and
// only its last lane issues the red.global.add.v2.f32 (warp-aggregated atomics).
template <bool AGGREGATE>
__device__ __forceinline__ void scatter_add(float2 *table, uint32_t idx, float2 v, int lane) {
    if (AGGREGATE) {
        const uint32_t prev = __shfl_up_sync(kFull, idx, 1);
        const bool head = (lane == 0) || (prev != idx);
        const unsigned heads = __ballot_sync(kFull, head);
        // start lane of my run = highest head bit at or below my lane
        const unsigned below = heads & (0xffffffffu >> (31 - lane));
        const int start = 31 - __clz(below);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float tx = __shfl_up_sync(kFull, v.x, o);
            const float ty = __shfl_up_sync(kFull, v.y, o);
            if (lane - o >= start) {
                v.x += tx;
                v.y += ty;
            }
        }
        const bool tail = (lane == 31) || ((heads >> (lane + 1)) & 1u);
        if (tail && (v.x != 0.0f || v.y != 0.0f)) red_add(table + idx, v);
    } else {
        if (v.x != 0.0f || v.y != 0.0f) red_add(table + idx, v);
    }
}

// Levels [l_begin, l_end) of one sample (one lane).  Every lane of the warp must call this (the coarse levels shuffle across the
// warp); lanes without a sample pass valid = false.  load_d(l) returns the gradient of encoding columns 2 l, 2 l + 1.
template <bool PAIRS, class LoadD>
__device__ __forceinline__ void scatter_sample(float x, float y, float z, bool valid, LoadD &&load_d,
                                               const float *__restrict__ scalings, int l_begin, int l_end, int log2T,
                                               float2 *__restrict__ d_table, int agg_levels, int lane) {
    const uint32_t tmask = (log2T >= 32) ? 0xffffffffu : ((1u << log2T) - 1u);
    for (int l = l_begin; l < l_end; ++l) {
        uint32_t idx[8];
        float ox, oy, oz;
        const bool pair = level_setup(x, y, z, __ldg(scalings + l), tmask, (uint32_t)l << log2T, idx, ox, oy, oz);
        float2 d = make_float2(0.f, 0.f);
        if (valid) d = load_d(l);
        const float ux = 1.0f - ox, uy = 1.0f - oy, uz = 1.0f - oz;
        // weight of corner (X,Y,Z): (X==c ? o : 1-o) per axis; cz side carries oz
        const float wz[2] = {oz, uz};
        const float wxy[4] = {ox * oy, ox * uy, ux * uy, ux * oy};  // (c,c) (c,f) (f,f) (f,c)
        if (l < agg_levels) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float w = wxy[c & 3] * wz[c >> 2];
                scatter_add<true>(d_table, idx[c], make_float2(w * d.x, w * d.y), lane);
            }
        } else if (PAIRS && valid && pair) {
            // fine levels: no two lanes share an entry; aligned x-pairs go out as one 16-byte reduction
            constexpr int kFloor[4] = {3, 2, 7, 6}, kCeil[4] = {0, 1, 4, 5};
            float4 *t4 = reinterpret_cast<float4 *>(d_table);
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const int cf = kFloor[p], cc = kCeil[p];
                const float wf = wxy[cf & 3] * wz[cf >> 2], wc = wxy[cc & 3] * wz[cc >> 2];
                const uint32_t fi = idx[cf];
                const bool odd = (fi & 1u) != 0u;
                const float wa = odd ? wc : wf, wb = odd ? wf : wc;  // first / second row of the pair
                const float4 v = make_float4(wa * d.x, wa * d.y, wb * d.x, wb * d.y);
                if (v.x != 0.0f || v.y != 0.0f || v.z != 0.0f || v.w != 0.0f) red_add(t4 + (fi >> 1), v);
            }
        } else if (valid) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float w = wxy[c & 3] * wz[c >> 2];
                scatter_add<false>(d_table, idx[c], make_float2(w * d.x, w * d.y), lane);
            }
        }
    }
}

// ---- the same scatter for FEW warps with many registers (the scatter warps of the tensor-core MLP backward: 7 per SM).
// There a warp's latency is what counts, not the SM's throughput: runs of lanes in the same cell are found ONCE per
// level (cell of the previous lane, one ballot), the 16 products (8 corners x 2 features) of a run are summed by one
// segmented Hillis-Steele scan whose 16 shuffles per step are independent (the per-corner form above is a chain of
// 8 x 5 dependent steps) and which stops at the longest run of the warp; only the last lane of each run issues
// reductions.  In the stand-alone kernel (43 resident warps, bound by the reduction path) this form runs 40 % fewer
// instructions at the same speed, so that kernel keeps the 40-register form.
struct Runs {
    int start;      // first lane of my run
    int longest;    // longest run of the warp (warp-uniform)
    bool tail;      // I am the last lane of my run
};
__device__ __forceinline__ Runs find_runs(const uint32_t (&cell)[3], int lane) {
    const uint32_t p0 = __shfl_up_sync(kFull, cell[0], 1), p1 = __shfl_up_sync(kFull, cell[1], 1);
    const uint32_t p2 = __shfl_up_sync(kFull, cell[2], 1);
    const bool head = (lane == 0) || p0 != cell[0] || p1 != cell[1] || p2 != cell[2];
    const unsigned heads = __ballot_sync(kFull, head);
    Runs r;
    r.start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
    const unsigned above = (lane == 31) ? 0u : (heads >> (lane + 1));  // heads strictly above my lane
    r.tail = (lane == 31) || (above & 1u);
    const int next = above ? (lane + __ffs(above)) : 32;               // first lane of the next run
    r.longest = (int)__reduce_max_sync(kFull, head ? (unsigned)(next - lane) : 0u);
    return r;
}
__device__ __forceinline__ void sum_runs(float2 (&v)[8], const Runs &r, int lane) {
#pragma unroll 1
    for (int o = 1; o < r.longest; o <<= 1) {
        const bool take = lane - o >= r.start;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float tx = __shfl_up_sync(kFull, v[c].x, o), ty = __shfl_up_sync(kFull, v[c].y, o);
            if (take) {
                v[c].x += tx;
                v[c].y += ty;
            }
        }
    }
}
// All 16 levels of one sample whose gradient row sits in the bf16 tile buffer `gt` (read through L2: the rows were
// written by other warps of this kernel).  Every lane of the warp calls this; lanes without a sample pass valid = false.
__device__ __forceinline__ void scatter_sample_runs(float x, float y, float z, bool valid, int64_t n,
                                                    const uint8_t *__restrict__ gt, const float *__restrict__ scalings,
                                                    int log2T, float2 *__restrict__ d_table, int agg_levels, int lane) {
    const uint32_t tmask = (log2T >= 32) ? 0xffffffffu : ((1u << log2T) - 1u);
    uint4 piece[4];  // the whole row: columns 8 i .. 8 i + 7 = levels 4 i .. 4 i + 3
#pragma unroll
    for (int i = 0; i < 4; ++i)
        piece[i] = valid ? __ldcg(reinterpret_cast<const uint4 *>(gt + enc_tile_piece(n, i))) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int lb = 0; lb < 4; ++lb) {
        const uint32_t words[4] = {piece[lb].x, piece[lb].y, piece[lb].z, piece[lb].w};
#pragma unroll 1
        for (int li = 0; li < 4; ++li) {
            const int l = 4 * lb + li;
            const uint32_t wd = li == 0 ? words[0] : li == 1 ? words[1] : li == 2 ? words[2] : words[3];
            const float2 d = make_float2(__uint_as_float(wd << 16), __uint_as_float(wd & 0xffff0000u));
            uint32_t idx[8], cell[3];
            float ox, oy, oz;
            const bool pair =
                level_setup(x, y, z, __ldg(scalings + l), tmask, (uint32_t)l << log2T, idx, ox, oy, oz, cell);
            const float ux = 1.0f - ox, uy = 1.0f - oy, uz = 1.0f - oz;
            const float wz[2] = {oz, uz};
            const float wxy[4] = {ox * oy, ox * uy, ux * uy, ux * oy};  // (c,c) (c,f) (f,f) (f,c)
            float2 v[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float w = wxy[c & 3] * wz[c >> 2];
                v[c] = make_float2(w * d.x, w * d.y);
            }
            bool emit = valid;
            if (l < agg_levels) {  // warp-uniform
                const Runs r = find_runs(cell, lane);
                sum_runs(v, r, lane);
                emit = r.tail;     // (a run of lanes without samples sums to zero and is skipped below)
            }
            if (emit && pair) {
                // aligned x-pairs go out as one 16-byte reduction
                constexpr int kFloor[4] = {3, 2, 7, 6}, kCeil[4] = {0, 1, 4, 5};
                float4 *t4 = reinterpret_cast<float4 *>(d_table);
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const float2 vf = v[kFloor[p]], vc = v[kCeil[p]];
                    const uint32_t fi = idx[kFloor[p]];
                    const bool odd = (fi & 1u) != 0u;
                    const float2 va = odd ? vc : vf, vb = odd ? vf : vc;  // first / second row of the pair
                    if (va.x != 0.0f || va.y != 0.0f || vb.x != 0.0f || vb.y != 0.0f)
                        red_add(t4 + (fi >> 1), make_float4(va.x, va.y, vb.x, vb.y));
                }
            } else if (emit) {
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (v[c].x != 0.0f || v[c].y != 0.0f) red_add(d_table + idx[c], v[c]);
            }
        }
    }
}

// gradient of encoding columns 2 l, 2 l + 1 of sample n in the bf16 tile buffer (common.cuh: enc_tile_piece)
__device__ __forceinline__ float2 unpack_tile_pair(uint32_t w) {
    return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}

}  // namespace umhs
