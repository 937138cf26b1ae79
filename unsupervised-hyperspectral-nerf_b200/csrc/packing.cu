// packing.cu -- occupancy-grid ray marching, sample packing and ray-index bookkeeping.
//
// Replaces nerfstudio VolumetricSampler / nerfacc 0.5.2 OccGridEstimator.sampling
// (traverse_grids, render_visibility_from_density, boolean-mask compaction) reached from
// ref umhsnerf/umhs_model.py:229-237, nerfacc.pack_info (ref umhsnerf/umhs_model.py:245) and
// the occupancy-grid update callback (ref umhsnerf/umhs_model.py:549-554).
//
// THIS TRANSLATION UNIT IS COMPILED WITH -fmad=false: march_ray() executes the same IEEE fp32
// operation sequence as oracle/pack_oracle.c (gcc -ffp-contract=off), so ray_indices, t_starts,
// t_ends and packed_info are bit-exact against it.
//
// HBM-bound integer/byte work: one thread per ray for the DDA (divergent by nature; rays of a
// batch are independent), one warp per ray for the visibility scan and the compaction, a
// single-CTA scan for the offsets (R is 1e4..5e5).
#include <math_constants.h>
#include <stdlib.h>

#include "common.cuh"

namespace umhs {

constexpr int kMaxGridLevels = 8;
constexpr float kMissT = 1e10f;

__device__ __forceinline__ bool ray_aabb(const float o[3], const float inv[3], const float *aabb, float &tmin_o,
                                         float &tmax_o) {
    float tmin, tmax, t0, t1;
    if (inv[0] >= 0) { tmin = (aabb[0] - o[0]) * inv[0]; tmax = (aabb[3] - o[0]) * inv[0]; }
    else             { tmin = (aabb[3] - o[0]) * inv[0]; tmax = (aabb[0] - o[0]) * inv[0]; }
    if (inv[1] >= 0) { t0 = (aabb[1] - o[1]) * inv[1]; t1 = (aabb[4] - o[1]) * inv[1]; }
    else             { t0 = (aabb[4] - o[1]) * inv[1]; t1 = (aabb[1] - o[1]) * inv[1]; }
    if (tmin > t1 || t0 > tmax) return false;
    if (t0 > tmin) tmin = t0;
    if (t1 < tmax) tmax = t1;
    if (inv[2] >= 0) { t0 = (aabb[2] - o[2]) * inv[2]; t1 = (aabb[5] - o[2]) * inv[2]; }
    else             { t0 = (aabb[5] - o[2]) * inv[2]; t1 = (aabb[2] - o[2]) * inv[2]; }
    if (tmin > t1 || t0 > tmax) return false;
    if (t0 > tmin) tmin = t0;
    if (t1 < tmax) tmax = t1;
    if (tmax <= 0) return false;
    tmin_o = tmin;
    tmax_o = tmax;
    return true;
}

__device__ __forceinline__ float calc_dt(float t, float cone_angle, float dt_min, float dt_max) {
    float v = t * cone_angle;
    v = v < dt_min ? dt_min : v;
    v = v > dt_max ? dt_max : v;
    return v;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__device__ __forceinline__ int voxel_index(float v, float fres, int res) {
    v = v < -1.0f ? -1.0f : (v > fres ? fres : v);
    return clampi((int)v, 0, res - 1);
}

// First point of the orbit t -> t + dt(t), dt(t) = clamp(t * cone_angle, step, 1e10), with t + dt/2 >= target:
// what the oracle reaches by stepping (`for (;;) { dt = calc_dt(t); if (t + dt * 0.5f >= target) break; t += dt; }`).
//
// Bit-exact shortcut for the stretch where dt == step (t * cone_angle < step: all of a bounded scene at the
// reference's cone angle).  Inside one binade [2^e, 2^(e+1)) every orbit point is a multiple of u = 2^(e-23), and
// fl(t + step) - t is the SAME multiple of u at every point (step rounded to the grid; the only exception, step
// exactly half-way between two grid points, is left to the plain loop).  So k steps are the integer update
// m += k * inc on the mantissa, k is found from the target by one division and corrected with the oracle's own fp32
// test, and only the step that leaves the binade is taken with a real fp32 add.  A few dozen instructions instead of
// hundreds of dependent iterations per gap.
__device__ __forceinline__ float advance_orbit(float t, float target, float cone_angle, float step) {
    const float half = step * 0.5f;  // dt * 0.5f while dt == step
    for (;;) {
        float dt = calc_dt(t, cone_angle, step, 1e10f);
        if (t + dt * 0.5f >= target) return t;
        const uint32_t tb = __float_as_uint(t), sb = __float_as_uint(step);
        const int e = (int)(tb >> 23) - 127, es = (int)(sb >> 23) - 127;  // t > 0, step > 0: sign bits clear
        const int sh = e - es;                                            // step = ms * 2^(es-23), u = 2^(e-23)
        const bool normal = t > 0.0f && (tb >> 23) != 0 && (tb >> 23) != 255 && (sb >> 23) != 0 && dt == step;
        bool fast = normal && sh < 24;
        uint32_t inc = 0;
        if (fast) {
            const uint32_t ms = (sb & 0x7fffffu) | 0x800000u;
            if (sh <= 0) {
                fast = sh > -8;  // step is a multiple of u: no rounding at all
                inc = fast ? (ms << (-sh)) : 0u;
            } else {
                const uint32_t rem = ms & ((1u << sh) - 1u), halfbit = 1u << (sh - 1);
                fast = rem != halfbit;                   // a tie rounds to even: alternating increments
                inc = (ms >> sh) + (rem > halfbit ? 1u : 0u);
            }
            fast = fast && inc != 0u;
        }
        if (fast) {
            const float u = __uint_as_float((uint32_t)(e - 23 + 127) << 23);  // 2^(e-23), e >= -100 here in practice
            const uint32_t m = (tb & 0x7fffffu) | 0x800000u;
            // points of this binade: m + k * inc < 2^24
            const uint32_t k_bin = (0xffffffu - m) / inc;
            // first k with (t_k + half >= target), t_k = (m + k inc) u: estimate, then correct with the exact test
            const float gap = (target - half) - t;
            float kf = ceilf(gap / ((float)inc * u));
            kf = kf < 0.0f ? 0.0f : kf;
            uint32_t k = kf >= (float)k_bin ? k_bin : (uint32_t)kf;
            auto at = [&](uint32_t kk) { return (float)(m + kk * inc) * u; };  // exact: integer < 2^24 times 2^n
            // dt == step must hold at every point used: t * cone_angle is monotone in t, so test the far end
            while (k > 0 && !(calc_dt(at(k), cone_angle, step, 1e10f) == step)) k >>= 1;
            while (k > 0 && at(k - 1) + half >= target) --k;
            while (k < k_bin && !(at(k) + half >= target) && calc_dt(at(k + 1), cone_angle, step, 1e10f) == step) ++k;
            if (k > 0) {
                t = at(k);
                continue;  // re-test at the new point (stop, next binade, or varying dt)
            }
        }
        t += dt;  // plain step: leaves the binade, tie binade, or dt != step
    }
}

// March one ray (twin of oracle/pack_oracle.c::march_ray).  WRITE=false only counts.
//
// Same fp32 operation sequence on t_last as the oracle, hence bit-exact outputs, with two shortcuts that cannot
// change the result:
//  * lazy empty-space skipping.  t_last walks the orbit t -> t + dt(t); in an EMPTY cell the oracle advances it to
//    the first orbit point with t + dt/2 >= t_traverse(cell).  The thresholds grow monotonically along the ray and
//    so does t + dt(t)/2 along the orbit, so advancing once to the LAST pending threshold lands on the same orbit
//    point as advancing cell by cell.  The advance is therefore deferred until an occupied cell needs t_last, and
//    never happens for the empty tail of the ray (most of the oracle's iterations);
//  * in grid levels that hold no occupied cell (level_any[l] == 0) the DDA still runs (its fp32 thresholds are part
//    of the orbit bookkeeping) but the occupancy bytes are not fetched.
template <bool WRITE, bool BITS>
__device__ int64_t march_ray(const float o[3], const float d[3], const uint8_t *__restrict__ binaries,
                             const uint32_t *__restrict__ bits, const float *__restrict__ aabbs,
                             const int *__restrict__ level_any, int levels, int res,
                             float near_plane, float far_plane, float step_size, float cone_angle, float *t0_out,
                             float *t1_out, int64_t limit) {
    const float inv[3] = {1.0f / d[0], 1.0f / d[1], 1.0f / d[2]};
    float tmins[kMaxGridLevels], tmaxs[kMaxGridLevels], ts[2 * kMaxGridLevels];
    bool hits[kMaxGridLevels];
    for (int l = 0; l < levels; ++l) {
        float a = kMissT, b = kMissT;
        hits[l] = ray_aabb(o, inv, aabbs + 6 * l, a, b);
        if (!hits[l]) { a = kMissT; b = kMissT; }
        tmins[l] = a; tmaxs[l] = b;
        ts[l] = a; ts[levels + l] = b;
    }
    for (int i = 1; i < 2 * levels; ++i) {  // insertion sort
        const float v = ts[i];
        int j = i - 1;
        while (j >= 0 && ts[j] > v) { ts[j + 1] = ts[j]; --j; }
        ts[j + 1] = v;
    }
    const float fres = (float)res;
    const int64_t cells_per_level = (int64_t)res * res * res;
    int64_t n = 0;
    float t_last = near_plane;
    // Intervals behind the last one whose grid level holds any occupied cell can only produce empty cells: nothing
    // is emitted there and nothing later reads the orbit state they would leave behind, so the march stops before
    // them (with level_any == nullptr every level counts as possibly occupied).
    int last_useful = -1;
    for (int i = 0; i < 2 * levels - 1; ++i) {
        const float a = ts[i], b = ts[i + 1];
        if (!(a < b)) continue;
        for (int l = 0; l < levels; ++l)
            if (hits[l] && tmins[l] <= a && b <= tmaxs[l]) {
                if (level_any == nullptr || level_any[l] != 0) last_useful = i;
                break;
            }
    }
    // occupied runs recorded by phase 1, consumed by phase 2: (last empty threshold before the run or -1, run end)
    constexpr int kRunCap = 24;
    float rec_pend[kRunCap], rec_end[kRunCap];
    int n_rec = 0;
    float pend = -1.0f;  // threshold of the empty cells passed since the last occupied run (< 0: none)
    // Phase 2: walk the orbit through the recorded runs.  Returns true when a capacity-truncated ray is full.
    auto flush_runs = [&]() -> bool {
        for (int k = 0; k < n_rec; ++k) {
            if (rec_pend[k] >= 0.0f) t_last = advance_orbit(t_last, rec_pend[k], cone_angle, step_size);
            const float t_run = rec_end[k];
            for (;;) {
                const float dt = calc_dt(t_last, cone_angle, step_size, 1e10f);
                if (t_last + dt * 0.5f >= t_run) break;
                if (WRITE) {
                    if (n >= limit) return true;  // capacity-truncated ray
                    t0_out[n] = t_last;
                    t1_out[n] = t_last + dt;
                } else if (n < limit) {
                    t0_out[n] = t_last;  // count pass: stage the interval starts (t1 = t0 + calc_dt(t0))
                }
                ++n;
                t_last += dt;
            }
        }
        n_rec = 0;
        return false;
    };
    for (int i = 0; i <= last_useful; ++i) {
        const float a = ts[i], b = ts[i + 1];
        if (!(a < b)) continue;
        int level = -1;
        for (int l = 0; l < levels; ++l)
            if (hits[l] && tmins[l] <= a && b <= tmaxs[l]) { level = l; break; }
        if (level < 0) continue;
        const float this_tmin = fmaxf(a, near_plane);
        const float this_tmax = fminf(b, far_plane);
        if (!(this_tmin < this_tmax)) continue;
        const bool lookup = (level_any == nullptr) || (level_any[level] != 0);
        const float *aabb = aabbs + 6 * level;
        const float eps = 1e-6f;
        float delta[3], tdist[3];
        int cur[3], over[3], stepi[3];
        for (int k = 0; k < 3; ++k) {
            const float ext = aabb[3 + k] - aabb[k];
            const float voxel = ext / fres;
            const float rs = o[k] + d[k] * (this_tmin + eps);
            const float re = o[k] + d[k] * (this_tmax - eps);
            cur[k] = voxel_index(((rs - aabb[k]) / ext) * fres, fres, res);
            const int fin = voxel_index(((re - aabb[k]) / ext) * fres, fres, res);
            const int start_index = cur[k] + (d[k] > 0.0f ? 1 : 0);
            const float tmax_k = ((aabb[k] + (((float)start_index * voxel) - rs)) * inv[k]) + this_tmin;
            const float sf = (d[k] == 0.0f) ? 0.0f : (d[k] > 0.0f ? 1.0f : -1.0f);
            stepi[k] = (int)sf;
            tdist[k] = (d[k] == 0.0f) ? this_tmax : tmax_k;
            delta[k] = (d[k] == 0.0f) ? this_tmax : (voxel * inv[k]) * sf;
            over[k] = fin + stepi[k];
        }
        const uint8_t *lvl_bin = binaries + (int64_t)level * cells_per_level;
        // bit-packed copy of the level (1 bit per cell, 256 KB at 128^3: L1-resident) when the caller built one
        const uint32_t *lvl_bits = BITS ? bits + (int64_t)level * ((cells_per_level + 31) / 32) : nullptr;
        auto fetch = [&](int c) -> uint32_t {
            if (BITS) return (__ldg(lvl_bits + (c >> 5)) >> (c & 31)) & 1u;
            return (uint32_t)__ldg(lvl_bin + c);
        };
        // steps each axis may still take before it passes the interval's last cell or leaves the grid
        int left[3];
        for (int k = 0; k < 3; ++k) {
            // d == 0: the oracle's `cur == over` test fires on the first selection of this axis iff fin == cur
            if (stepi[k] == 0) { left[k] = (cur[k] == over[k]) ? 1 : 0x7fffffff; continue; }
            const int to_over = (over[k] - cur[k]) * stepi[k];                       // >= 1 when the end is ahead
            const int to_edge = stepi[k] > 0 ? (res - cur[k]) : (cur[k] + 1);        // steps until outside [0,res)
            left[k] = (to_over >= 1 && to_over < to_edge) ? to_over : to_edge;
        }
        int cell = (cur[0] * res + cur[1]) * res + cur[2];
        const int s0 = stepi[0] * res * res, s1 = stepi[1] * res, s2 = stepi[2];
        const float dl0 = delta[0], dl1 = delta[1], dl2 = delta[2];
        float td0 = tdist[0], td1 = tdist[1], td2 = tdist[2];
        int l0 = left[0], l1 = left[1], l2 = left[2];
        uint32_t occ = lookup ? fetch(cell) : 0u;
        bool last = false;
        int guard = 4 * res + 8;
        // Phase 1 (this loop): the DDA alone, one cell per iteration for every lane of the warp, no inner loops.  Runs
        // of cells with equal occupancy collapse to their last threshold: for an OCCUPIED run c1..cm with thresholds
        // T1 < .. < Tm the oracle emits, cell by cell, the orbit points with t + dt/2 < Ti -- together exactly the
        // points with t + dt/2 < Tm (same monotonicity as the lazy skip); an empty run leaves `pend` = its last
        // threshold.  Every occupied run is RECORDED as (pend before it, Tm); the orbit work happens in phase 2
        // (flush_runs), where the k-th occupied run of every ray of the warp is processed side by side.  With the orbit
        // loops inside the DDA loop the lanes of a warp serialised on each other's runs (ncu: 12 of 32 lanes active).
        // The step itself is written with selects, not branches: the lanes of a warp pick different axes in almost
        // every iteration, and a three-way divergent branch made the warp execute all three bodies (ncu: ~85
        // instructions per iteration, 8 cycles each at one to three warps per scheduler).
        // Measured and dropped: a one-deep software pipeline (issue the fetch of cell i+1, then do the run bookkeeping
        // of cell i with the word loaded an iteration earlier): 0.127 -> 0.61 ms -- the carried state and the
        // two-exit loop cost far more than the L1 / L2 latency they were meant to hide.
        while (!last && guard > 0) {
            const float t_traverse = fminf(fminf(td0, fminf(td1, td2)), this_tmax);
            // single DDA step: x if strictly smallest, else y if < z, else z
            const bool ax0 = (td0 < td1) && (td0 < td2);
            const bool ax1 = !ax0 && (td1 < td2);
            const bool ax2 = !ax0 && !ax1;
            td0 = ax0 ? td0 + dl0 : td0;
            td1 = ax1 ? td1 + dl1 : td1;
            td2 = ax2 ? td2 + dl2 : td2;
            cell += ax0 ? s0 : (ax1 ? s1 : s2);
            l0 -= ax0 ? 1 : 0;
            l1 -= ax1 ? 1 : 0;
            l2 -= ax2 ? 1 : 0;
            last = (l0 == 0) | (l1 == 0) | (l2 == 0);  // only the stepped axis can have reached 0 (all start >= 1)
            const uint32_t occ_next = (lookup && !last) ? fetch(cell) : 0u;
            --guard;
            if (last || guard == 0 || occ_next != occ) {  // the cell just left ends a run
                if (!occ) {
                    pend = t_traverse;  // monotone: later thresholds dominate earlier ones
                } else {
                    rec_pend[n_rec] = pend;
                    rec_end[n_rec] = t_traverse;
                    pend = -1.0f;
                    if (++n_rec == kRunCap && flush_runs()) return n;
                }
            }
            occ = occ_next;
        }
    }
    if (flush_runs()) return n;
    return n;
}

// rays_per_warp < 32 leaves the upper lanes of every warp idle ON PURPOSE: the DDA is a long dependent instruction
// stream with data-dependent trip counts, so a full warp runs at the pace of its slowest ray with ~10 of 32 lanes
// active, and 16 384 rays are only 3.5 warps per SM.  Narrow warps cost nothing (those lanes were idle) and give
// the schedulers 4x the warps to interleave, each as long as the slowest of 8 rays instead of 32.
template <bool WRITE, bool BITS>
__global__ void __launch_bounds__(128)
march_kernel(umhs_march_args a, const int *__restrict__ level_any, int64_t *counts, const int64_t *packed_info,
             int64_t *ray_indices, float *t_starts, float *t_ends, int rays_per_warp) {
    const int lane = threadIdx.x & 31;
    if (lane >= rays_per_warp) return;
    const int64_t r = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * rays_per_warp + lane;
    if (r >= a.n_rays) return;
    const float o[3] = {__ldg(a.rays_o + 3 * r), __ldg(a.rays_o + 3 * r + 1), __ldg(a.rays_o + 3 * r + 2)};
    const float d[3] = {__ldg(a.rays_d + 3 * r), __ldg(a.rays_d + 3 * r + 1), __ldg(a.rays_d + 3 * r + 2)};
    if (!WRITE) {
        // with a staging buffer the count pass also records the first stage_stride interval starts of the ray, and
        // the write pass turns into a coalesced copy (rays longer than the stride are re-marched there)
        float *stage = a.stage ? a.stage + r * a.stage_stride : nullptr;
        counts[r] = march_ray<false, BITS>(o, d, a.binaries, a.occ_bits, a.aabbs, level_any, a.levels, a.resolution, __ldg(a.near_planes + r),
                                     __ldg(a.far_planes + r), a.step_size, a.cone_angle, stage, nullptr,
                                     stage ? a.stage_stride : 0);
    } else {
        const int64_t off = __ldg(packed_info + 2 * r), lim = __ldg(packed_info + 2 * r + 1);
        if (lim <= 0) return;
        const int64_t n = march_ray<true, BITS>(o, d, a.binaries, a.occ_bits, a.aabbs, level_any, a.levels, a.resolution, __ldg(a.near_planes + r),
                                          __ldg(a.far_planes + r), a.step_size, a.cone_angle, t_starts + off,
                                          t_ends + off, lim);
        for (int64_t i = 0; i < n; ++i) ray_indices[off + i] = r;
    }
}

// Write pass from the staging buffer: one warp per ray copies the staged interval starts to the packed arrays
// (t_end = t_start + calc_dt(t_start): the same fp32 expression the march evaluates).  Rays with more samples
// than the staging stride are re-marched by lane 0.
template <bool BITS>
__global__ void __launch_bounds__(256)
march_write_staged_kernel(umhs_march_args a, const int *__restrict__ level_any, const int64_t *__restrict__ counts,
                          const int64_t *__restrict__ packed_info, int64_t *__restrict__ ray_indices,
                          float *__restrict__ t_starts, float *__restrict__ t_ends) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= a.n_rays) return;
    const int64_t off = __ldg(packed_info + 2 * r), lim = __ldg(packed_info + 2 * r + 1);
    if (lim <= 0) return;
    const int64_t full = __ldg(counts + r);
    if (full <= a.stage_stride) {
        const float *src = a.stage + r * a.stage_stride;
        for (int64_t i = lane; i < lim; i += 32) {
            const float t0 = src[i];
            const float dt = calc_dt(t0, a.cone_angle, a.step_size, 1e10f);
            t_starts[off + i] = t0;
            t_ends[off + i] = t0 + dt;
            ray_indices[off + i] = r;
        }
    } else {
        if (lane == 0) {
            const float o[3] = {__ldg(a.rays_o + 3 * r), __ldg(a.rays_o + 3 * r + 1), __ldg(a.rays_o + 3 * r + 2)};
            const float d[3] = {__ldg(a.rays_d + 3 * r), __ldg(a.rays_d + 3 * r + 1), __ldg(a.rays_d + 3 * r + 2)};
            march_ray<true, BITS>(o, d, a.binaries, a.occ_bits, a.aabbs, level_any, a.levels, a.resolution, __ldg(a.near_planes + r),
                            __ldg(a.far_planes + r), a.step_size, a.cone_angle, t_starts + off, t_ends + off, lim);
        }
        for (int64_t i = lane; i < lim; i += 32) ray_indices[off + i] = r;
    }
}

// level_any[l] = 1 iff grid level l holds an occupied cell (8 levels max; caller zero-fills)
__global__ void level_any_kernel(const uint8_t *__restrict__ binaries, int64_t cells_per_level, int levels,
                                 int *__restrict__ level_any) {
    const int l = blockIdx.y;
    const uint4 *p = reinterpret_cast<const uint4 *>(binaries + (int64_t)l * cells_per_level);
    // 16-byte loads only when this level starts on a 16-byte boundary (res^3 need not be a multiple of 16)
    const int64_t n16 = (reinterpret_cast<uintptr_t>(p) & 15) == 0 ? cells_per_level / 16 : 0;
    bool any = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16 && !any; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldg(p + i);
        any = (v.x | v.y | v.z | v.w) != 0u;
    }
    for (int64_t i = n16 * 16 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells_per_level;
         i += (int64_t)gridDim.x * blockDim.x)
        any = any || binaries[(int64_t)l * cells_per_level + i] != 0;
    if (__syncthreads_or(any) && threadIdx.x == 0) atomicOr(level_any + l, 1);
}

// bits[l][w] bit b = binaries[l][32 w + b]; level_any[l] |= any.  One thread per 32 cells.
__global__ void __launch_bounds__(256)
occ_pack_bits_kernel(const uint8_t *__restrict__ binaries, int64_t cells_per_level, int levels,
                     uint32_t *__restrict__ bits, int *__restrict__ level_any) {
    const int l = blockIdx.y;
    const int64_t words = (cells_per_level + 31) / 32;
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t word = 0;
    if (w < words) {
        const uint8_t *src = binaries + (int64_t)l * cells_per_level + w * 32;
        if (w * 32 + 32 <= cells_per_level && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
            const uint4 a = __ldg(reinterpret_cast<const uint4 *>(src)), b = __ldg(reinterpret_cast<const uint4 *>(src) + 1);
            const uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
            for (int q = 0; q < 8; ++q) {
#pragma unroll
                for (int j = 0; j < 4; ++j) word |= (((v[q] >> (8 * j)) & 0xffu) ? 1u : 0u) << (4 * q + j);
            }
        } else {
            for (int b = 0; b < 32 && w * 32 + b < cells_per_level; ++b) word |= (src[b] ? 1u : 0u) << b;
        }
        bits[(int64_t)l * words + w] = word;
    }
    if (__syncthreads_or(word != 0u) && threadIdx.x == 0) atomicOr(level_any + l, 1);
}

// Single-CTA exclusive scan of int64 counts -> packed_info[r] = (offset, count); *total = sum.
// Every access is coalesced (lane = consecutive element): a warp owns 1024 consecutive elements of a 16 384-element
// tile and walks them 32 at a time.  Pass 1 only sums (32 independent loads per lane, one shuffle reduction per warp),
// the 16 warp totals are scanned by warp 0, pass 2 re-reads the counts (L1 hits), scans each group of 32 with
// shuffles plus a running carry and writes (offset, count) as 16-byte pairs.  The previous form gave every thread 16
// CONSECUTIVE elements: each of its 16 loads and 16 stores touched 32 different lines per warp, and one SM's LSU
// serialised ~32 k wavefronts per tile (20 us per 16 384 rays, 0.55 ms per 480 k-ray eval image; now ~5 us).
constexpr int kScanThreads = 512;
constexpr int kScanWarps = kScanThreads / 32;
constexpr int kScanIters = 32;                                  // groups of 32 elements per warp per tile
constexpr int kScanTile = kScanWarps * kScanIters * 32;         // 16 384 elements
__device__ __forceinline__ int64_t warp_incl_scan64(int64_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int64_t t = __shfl_up_sync(kFull, v, o);
        if (lane >= o) v += t;
    }
    return v;
}
__global__ void __launch_bounds__(kScanThreads)
pack_scan_kernel(const int64_t *__restrict__ counts, int64_t n, int64_t *__restrict__ packed_info, int64_t *total,
                 int64_t capacity) {
    __shared__ int64_t warp_sums[32];
    __shared__ int64_t base_s, tile_total_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) base_s = 0;
    __syncthreads();
    for (int64_t t0 = 0; t0 < n; t0 += kScanTile) {
        const int64_t w0 = t0 + (int64_t)warp * (kScanIters * 32) + lane;
        int64_t sum = 0;
#pragma unroll
        for (int it = 0; it < kScanIters; ++it) {
            const int64_t i = w0 + it * 32;
            sum += (i < n) ? counts[i] : 0;
        }
        sum = warp_incl_scan64(sum, lane);
        if (lane == 31) warp_sums[warp] = sum;
        __syncthreads();
        if (warp == 0) {
            const int64_t w = lane < kScanWarps ? warp_sums[lane] : 0;
            const int64_t wi = warp_incl_scan64(w, lane);
            warp_sums[lane] = wi - w;  // exclusive warp offsets
            if (lane == 31) tile_total_s = wi;
        }
        __syncthreads();
        int64_t run = base_s + warp_sums[warp];
#pragma unroll 4
        for (int it = 0; it < kScanIters; ++it) {
            const int64_t i = w0 + it * 32;
            const int64_t v = (i < n) ? counts[i] : 0;
            const int64_t incl = warp_incl_scan64(v, lane);
            const int64_t excl = run + (incl - v);
            run += __shfl_sync(kFull, incl, 31);
            if (i < n) {
                int64_t start = excl, cnt = v;
                if (capacity > 0) {  // truncate the packing at `capacity` samples
                    start = excl < capacity ? excl : capacity;
                    cnt = (excl + v <= capacity) ? v : (capacity - start);
                }
                reinterpret_cast<longlong2 *>(packed_info)[i] = make_longlong2(start, cnt);
            }
        }
        __syncthreads();
        if (tid == 0) base_s += tile_total_s;
        __syncthreads();
    }
    if (tid == 0 && total) {
        total[0] = (capacity > 0 && base_s > capacity) ? capacity : base_s;
        total[1] = base_s;
    }
}

// keep = (T >= early_stop_eps) & (alpha >= alpha_thre), one warp per ray.
constexpr int kVisWarps = 8;
__global__ void __launch_bounds__(kVisWarps * 32)
visibility_kernel(const int64_t *__restrict__ packed_info, int64_t n_rays, const float *__restrict__ t_starts,
                  const float *__restrict__ t_ends, const float *__restrict__ sigma, float early_stop_eps,
                  float alpha_thre, const float *__restrict__ occ_mean, uint8_t *__restrict__ keep,
                  int64_t *__restrict__ keep_counts) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * kVisWarps + (threadIdx.x >> 5);
    if (r >= n_rays) return;
    float thre = alpha_thre;
    if (occ_mean) thre = fminf(thre, __ldg(occ_mean));
    const int64_t start = __ldg(packed_info + 2 * r), cnt = __ldg(packed_info + 2 * r + 1);
    float carry = 0.f;
    int64_t kept = 0;
    for (int64_t c0 = 0; c0 < cnt; c0 += 32) {
        const int64_t n = start + c0 + lane;
        const bool v = c0 + lane < cnt;
        float sd = 0.f;
        if (v) sd = __ldg(sigma + n) * (__ldg(t_ends + n) - __ldg(t_starts + n));
        const float incl = warp_incl_scan(sd, lane);
        const float T = expf(-(carry + (incl - sd)));
        const float alpha = 1.0f - expf(-sd);
        carry += __shfl_sync(kFull, incl, 31);
        bool k = v && (T >= early_stop_eps);
        if (thre > 0.0f) k = k && (alpha >= thre);
        if (v) keep[n] = k ? 1 : 0;
        kept += __popc(__ballot_sync(kFull, k));
    }
    if (lane == 0) keep_counts[r] = kept;
}

__global__ void __launch_bounds__(kVisWarps * 32)
compact_kernel(const int64_t *__restrict__ pin, const int64_t *__restrict__ pout, int64_t n_rays,
               const uint8_t *__restrict__ keep, const float *__restrict__ t0_in, const float *__restrict__ t1_in,
               int64_t *__restrict__ ray_indices, float *__restrict__ t0, float *__restrict__ t1,
               int32_t *__restrict__ src_index) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * kVisWarps + (threadIdx.x >> 5);
    if (r >= n_rays) return;
    const int64_t start = __ldg(pin + 2 * r), cnt = __ldg(pin + 2 * r + 1);
    int64_t dst = __ldg(pout + 2 * r);
    for (int64_t c0 = 0; c0 < cnt; c0 += 32) {
        const int64_t n = start + c0 + lane;
        const bool k = (c0 + lane < cnt) && keep[n];
        const unsigned m = __ballot_sync(kFull, k);
        if (k) {
            const int64_t j = dst + __popc(m & ((1u << lane) - 1u));
            ray_indices[j] = r;
            t0[j] = t0_in[n];
            t1[j] = t1_in[n];
            if (src_index) src_index[j] = (int32_t)n;
        }
        dst += __popc(m);
    }
}

__global__ void count_rays_kernel(const int64_t *__restrict__ ray_indices, int64_t n, int64_t n_rays,
                                  unsigned long long *counts) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t r = ray_indices[i];
    if (r >= 0 && r < n_rays) atomicAdd(counts + r, 1ull);
}

// ------------------------------------------------------------------ occupancy grid maintenance
// counter-based uniform draws for the grid update (splitmix64 of seed + counter): deterministic in (seed, index), so
// data-parallel replicas that share the seed pick the same cells and the same in-cell offsets without a collective
__device__ __forceinline__ float occ_uniform(uint64_t seed, uint64_t counter) {
    uint64_t z = seed + (counter + 1ull) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (float)(z >> 40) * (1.0f / 16777216.0f);  // 24 bits -> [0, 1)
}

// cell -> point inside the cell: x = (coord + jitter) / res mapped into the level's aabb.  jitter NULL: drawn from
// (seed, 3 i + k).  id < 0 (an unused slot of umhs_occ_select_cells): skipped.
__global__ void occ_cell_points_kernel(const int64_t *__restrict__ cell_ids, const float *__restrict__ jitter,
                                       uint64_t seed, const uint64_t *__restrict__ seed_dev, int64_t n, int res,
                                       const float *__restrict__ aabb, float *__restrict__ pos) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (seed_dev) seed += *seed_dev * 0xD1342543DE82EF95ull;  // device-side epoch: advances across CUDA-graph replays
    const int64_t id = cell_ids ? cell_ids[i] : i;
    if (id < 0) {
        pos[3 * i] = pos[3 * i + 1] = pos[3 * i + 2] = 0.f;
        return;
    }
    const float fres = (float)res;
    const float c[3] = {(float)(id / ((int64_t)res * res)), (float)((id / res) % res), (float)(id % res)};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float j = jitter ? jitter[3 * i + k] : occ_uniform(seed, (uint64_t)(3 * i + k));
        const float x = (c[k] + j) / fres;
        pos[3 * i + k] = aabb[k] + x * (aabb[3 + k] - aabb[k]);
    }
}

// ---- steady-state cell selection (nerfacc OccGridEstimator._sample_uniform_and_occupied_cells for one level):
// n_u uniform cells, then the occupied cells -- all of them, in increasing order, when there are at most n_u, else n_u
// random picks with replacement.  The reference uses nonzero() + randint(); here the occupied cells come from the
// bit-packed grid by rank: prefix[w] = set bits in words [0, w), rank r -> binary search for its word, then the
// (r - prefix[w] + 1)-th set bit of that word.  No sort, no host sync; unused slots are -1.
constexpr int kPrefixThreads = 1024;
// two-level prefix: block b scans its 1024 words (prefix[w] = set bits in the block's words before w) and publishes its
// total in prefix[n_words + 1 + b]; the select kernel turns the totals into block bases (a handful of blocks: 64 at
// 128^3 cells) and searches block, then word
__global__ void __launch_bounds__(kPrefixThreads)
occ_popc_prefix_kernel(const uint32_t *__restrict__ bits, int64_t n_words, int32_t *__restrict__ prefix) {
    __shared__ int32_t warp_tot[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t w = (int64_t)blockIdx.x * kPrefixThreads + tid;
    const int32_t mine = w < n_words ? __popc(bits[w]) : 0;
    int32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int32_t t = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int32_t v = warp_tot[lane];
        int32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t t = __shfl_up_sync(kFull, inc, o);
            if (lane >= o) inc += t;
        }
        warp_tot[lane] = inc - v;
        if (lane == 31) prefix[n_words + 1 + blockIdx.x] = inc;  // this block's total
    }
    __syncthreads();
    if (w < n_words) prefix[w] = warp_tot[warp] + incl - mine;
}

constexpr int kMaxPrefixBlocks = 2048;  // 2 M words = 64 M cells per level
__global__ void occ_select_kernel(const uint32_t *__restrict__ bits, const int32_t *__restrict__ prefix, int64_t n_words,
                                  int64_t n_cells, int64_t n_u, const float *__restrict__ u_uniform,
                                  const float *__restrict__ u_pick, uint64_t seed,
                                  const uint64_t *__restrict__ seed_dev, int64_t *__restrict__ cell_ids,
                                  int64_t *__restrict__ n_valid) {
    extern __shared__ int32_t block_base[];  // [n_blocks + 1] exclusive prefix of the block totals
    const int n_blocks = (int)((n_words + kPrefixThreads - 1) / kPrefixThreads);
    if (threadIdx.x == 0) {
        int32_t run = 0;
        for (int b = 0; b < n_blocks; ++b) {
            block_base[b] = run;
            run += prefix[n_words + 1 + b];
        }
        block_base[n_blocks] = run;
    }
    __syncthreads();
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (seed_dev) seed += *seed_dev * 0xD1342543DE82EF95ull;
    const int64_t count = block_base[n_blocks];
    if (j == 0 && n_valid) *n_valid = n_u + (count < n_u ? count : n_u);
    if (j >= 2 * n_u) return;
    if (j < n_u) {  // torch.randint(n_cells, (n_u,))
        const float u = u_uniform ? u_uniform[j] : occ_uniform(seed, (uint64_t)j);
        int64_t c = (int64_t)(u * (float)n_cells);
        cell_ids[j] = c < n_cells ? c : n_cells - 1;
        return;
    }
    const int64_t s = j - n_u;
    int64_t rank;
    if (count <= n_u) {
        if (s >= count) {
            cell_ids[j] = -1;
            return;
        }
        rank = s;  // all occupied cells, in increasing order (nonzero())
    } else {       // occ_idx[torch.randint(len(occ_idx), (n_u,))]
        const float u = u_pick ? u_pick[s] : occ_uniform(seed, (uint64_t)(n_u + s));
        rank = (int64_t)(u * (float)count);
        if (rank >= count) rank = count - 1;
    }
    int lo = 0, hi = n_blocks;  // last block whose base <= rank
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (block_base[mid] <= rank) lo = mid;
        else hi = mid;
    }
    const int32_t r_in = (int32_t)(rank - block_base[lo]);
    int64_t wl = (int64_t)lo * kPrefixThreads, wh = wl + kPrefixThreads < n_words ? wl + kPrefixThreads : n_words;
    while (wh - wl > 1) {  // last word of the block whose in-block prefix <= r_in
        const int64_t mid = (wl + wh) >> 1;
        if (prefix[mid] <= r_in) wl = mid;
        else wh = mid;
    }
    const int bit = __fns(bits[wl], 0, r_in - prefix[wl] + 1);
    cell_ids[j] = wl * 32 + bit;
}

// occs[id] = max(occs[id] * decay, sigma * scale): phase 1 reads, phase 2 scatters (torch index-assign
// semantics: the right-hand side sees only old values; duplicate ids: any one of them wins).  id < 0: skipped.
__global__ void occ_ema_read_kernel(const float *__restrict__ occs, const int64_t *__restrict__ cell_ids,
                                    int64_t level_offset, const float *__restrict__ sigma, float scale, float decay,
                                    int64_t n, float *__restrict__ tmp) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t c = cell_ids ? cell_ids[i] : i;
    if (c < 0) return;
    tmp[i] = fmaxf(occs[c + level_offset] * decay, sigma[i] * scale);
}
__global__ void occ_ema_write_kernel(float *__restrict__ occs, const int64_t *__restrict__ cell_ids,
                                     int64_t level_offset, int64_t n, const float *__restrict__ tmp) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t c = cell_ids ? cell_ids[i] : i;
    if (c < 0) return;
    occs[c + level_offset] = tmp[i];
}

__global__ void occ_sum_kernel(const float *__restrict__ occs, int64_t n, double *scratch) {
    double s = 0.0, c = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = occs[i];
        if (v >= 0.f) { s += (double)v; c += 1.0; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(kFull, s, o);
        c += __shfl_xor_sync(kFull, c, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(scratch, s);
        atomicAdd(scratch + 1, c);
    }
}

__global__ void occ_binarize_kernel(const float *__restrict__ occs, int64_t n, float occ_thre, const double *scratch,
                                    uint8_t *__restrict__ binaries, float *mean_out) {
    const float mean = (float)(scratch[0] / fmax(scratch[1], 1.0));
    const float thre = fminf(mean, occ_thre);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && mean_out) *mean_out = mean;
    if (i < n) binaries[i] = occs[i] > thre ? 1 : 0;
}

// rays per warp of the march kernels (see march_kernel); UMHS_MARCH_RPW overrides it for A/B timing
static int march_rays_per_warp(int64_t n_rays) {
    int rpw = n_rays >= 131072 ? 32 : (n_rays >= 65536 ? 16 : 8);  // measured: 16 k rays 0.158 / 0.147 / 0.139 ms at
    if (const char *e = getenv("UMHS_MARCH_RPW")) rpw = atoi(e);   // 32 / 16 / 8; 480 k rays (eval) fastest at 32
    return rpw < 1 ? 1 : (rpw > 32 ? 32 : rpw);
}

static int check_march(const umhs_march_args *a) {
    if (!a || !a->rays_o || !a->rays_d || !a->near_planes || !a->far_planes || !a->binaries || !a->aabbs) return -1;
    if (a->levels < 1 || a->levels > kMaxGridLevels || a->resolution < 1 || a->resolution > 1024) return -1;
    return 0;
}

}  // namespace umhs

using namespace umhs;

extern "C" int umhs_march_level_any(const umhs_march_args *a, int32_t *level_any, void *stream) {
    UMHS_CHECK_ARG(check_march(a) == 0 && level_any, "bad march arguments");
    const int64_t cells = (int64_t)a->resolution * a->resolution * a->resolution;
    UMHS_CHECK_CUDA(cudaMemsetAsync(level_any, 0, sizeof(int32_t) * 8, as_stream(stream)));
    level_any_kernel<<<dim3(64, a->levels), 256, 0, as_stream(stream)>>>(a->binaries, cells, a->levels, level_any);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

extern "C" int umhs_occ_pack_bits(const umhs_march_args *a, uint32_t *bits, int32_t *level_any, void *stream) {
    // only the grid fields of the march arguments are read (a trainer repacks after an occupancy update, without rays)
    UMHS_CHECK_ARG(a && a->binaries && a->levels >= 1 && a->levels <= kMaxGridLevels && a->resolution >= 1 &&
                       a->resolution <= 1024 && bits && level_any,
                   "bad grid arguments");
    const int64_t cells = (int64_t)a->resolution * a->resolution * a->resolution;
    UMHS_CHECK_CUDA(cudaMemsetAsync(level_any, 0, sizeof(int32_t) * 8, as_stream(stream)));
    occ_pack_bits_kernel<<<dim3(div_up((cells + 31) / 32, 256), a->levels), 256, 0, as_stream(stream)>>>(
        a->binaries, cells, a->levels, bits, level_any);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

// Shared-memory carve-out of the march / scan kernels (percent of the SM's unified L1 + shared memory; -1 = the
// driver's default, i.e. all L1 for these kernels, which use none).  An SM runs ONE carve-out configuration at a time:
// with the default, a march CTA cannot share an SM with the persistent tensor-core MLP kernels (~222 KB of shared
// memory) -- set to 100 when the march of the next batch is forked beside them (UMHSTrainer, pipelined mode).
static int g_march_carveout = -2;
static void apply_march_carveout() {
    if (g_march_carveout == -2) {
        const char *e = getenv("UMHS_MARCH_CARVEOUT");
        g_march_carveout = e ? atoi(e) : -1;
        if (g_march_carveout >= 0) {
            cudaFuncSetAttribute(march_kernel<false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, g_march_carveout);
            cudaFuncSetAttribute(march_kernel<false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, g_march_carveout);
            cudaFuncSetAttribute(march_kernel<true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, g_march_carveout);
            cudaFuncSetAttribute(march_kernel<true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, g_march_carveout);
            cudaFuncSetAttribute(march_write_staged_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, g_march_carveout);
            cudaFuncSetAttribute(march_write_staged_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, g_march_carveout);
            cudaFuncSetAttribute(pack_scan_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, g_march_carveout);
        }
    }
}

extern "C" int umhs_march_count(const umhs_march_args *a, int64_t *counts, void *stream) {
    UMHS_CHECK_ARG(check_march(a) == 0 && counts, "bad march arguments");
    if (a->n_rays <= 0) return UMHS_OK;
    apply_march_carveout();
    // The trainer runs this kernel beside the hash scatter and the optimiser of the previous step (kernels without
    // shared memory, like this one: all of the SM's 256 KB stays L1 and the bit-packed level-0 grid stays resident).
    const int rpw = march_rays_per_warp(a->n_rays);
    const dim3 grid(div_up(div_up(a->n_rays, rpw), 2));
    if (a->occ_bits)
        march_kernel<false, true><<<grid, 64, 0, as_stream(stream)>>>(*a, a->level_any, counts, nullptr, nullptr,
                                                                      nullptr, nullptr, rpw);
    else
        march_kernel<false, false><<<grid, 64, 0, as_stream(stream)>>>(*a, a->level_any, counts, nullptr, nullptr,
                                                                       nullptr, nullptr, rpw);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

extern "C" int umhs_pack_scan(const int64_t *counts, int64_t n_rays, int64_t *packed_info, int64_t *total,
                              int64_t capacity, void *stream) {
    UMHS_CHECK_ARG(counts && packed_info, "null argument");
    UMHS_CHECK_ARG(n_rays >= 0, "negative n_rays");
    pack_scan_kernel<<<1, kScanThreads, 0, as_stream(stream)>>>(counts, n_rays, packed_info, total, capacity);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

extern "C" int umhs_march_write(const umhs_march_args *a, const int64_t *packed_info, int64_t *ray_indices,
                                float *t_starts, float *t_ends, void *stream) {
    UMHS_CHECK_ARG(check_march(a) == 0 && packed_info && ray_indices && t_starts && t_ends, "bad march arguments");
    if (a->n_rays <= 0) return UMHS_OK;
    if (a->stage != nullptr && a->stage_stride > 0 && a->stage_counts != nullptr) {
        if (a->occ_bits)
            march_write_staged_kernel<true><<<div_up(a->n_rays, 8), 256, 0, as_stream(stream)>>>(
                *a, a->level_any, a->stage_counts, packed_info, ray_indices, t_starts, t_ends);
        else
            march_write_staged_kernel<false><<<div_up(a->n_rays, 8), 256, 0, as_stream(stream)>>>(
                *a, a->level_any, a->stage_counts, packed_info, ray_indices, t_starts, t_ends);
        UMHS_CHECK_LAUNCH();
        return UMHS_OK;
    }
    const int rpw = march_rays_per_warp(a->n_rays);
    const dim3 grid(div_up(div_up(a->n_rays, rpw), 2));
    if (a->occ_bits)
        march_kernel<true, true><<<grid, 64, 0, as_stream(stream)>>>(*a, a->level_any, nullptr, packed_info,
                                                                     ray_indices, t_starts, t_ends, rpw);
    else
        march_kernel<true, false><<<grid, 64, 0, as_stream(stream)>>>(*a, a->level_any, nullptr, packed_info,
                                                                      ray_indices, t_starts, t_ends, rpw);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

extern "C" int umhs_visibility_mask(const int64_t *packed_info, int64_t n_rays, const float *t_starts,
                                    const float *t_ends, const float *sigma, float early_stop_eps, float alpha_thre,
                                    const float *occ_mean, uint8_t *keep, int64_t *keep_counts, void *stream) {
    UMHS_CHECK_ARG(packed_info && t_starts && t_ends && sigma && keep && keep_counts, "null argument");
    if (n_rays <= 0) return UMHS_OK;
    visibility_kernel<<<div_up(n_rays, kVisWarps), kVisWarps * 32, 0, as_stream(stream)>>>(
        packed_info, n_rays, t_starts, t_ends, sigma, early_stop_eps, alpha_thre, occ_mean, keep, keep_counts);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

extern "C" int umhs_compact_samples(const int64_t *packed_info_in, const int64_t *packed_info_out, int64_t n_rays,
                                    const uint8_t *keep, const float *t_starts_in, const float *t_ends_in,
                                    int64_t *ray_indices, float *t_starts, float *t_ends, int32_t *src_index,
                                    void *stream) {
    UMHS_CHECK_ARG(packed_info_in && packed_info_out && keep && t_starts_in && t_ends_in && ray_indices && t_starts &&
                       t_ends,
                   "null argument");
    if (n_rays <= 0) return UMHS_OK;
    compact_kernel<<<div_up(n_rays, kVisWarps), kVisWarps * 32, 0, as_stream(stream)>>>(
        packed_info_in, packed_info_out, n_rays, keep, t_starts_in, t_ends_in, ray_indices, t_starts, t_ends, src_index);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

extern "C" int umhs_pack_info(const int64_t *ray_indices, int64_t n, int64_t n_rays, int64_t *counts,
                              int64_t *packed_info, void *stream) {
    UMHS_CHECK_ARG(counts && packed_info && (ray_indices || n == 0), "null argument");
    UMHS_CHECK_ARG(n_rays >= 0 && n >= 0, "negative size");
    cudaStream_t st = as_stream(stream);
    if (n_rays > 0) UMHS_CHECK_CUDA(cudaMemsetAsync(counts, 0, sizeof(int64_t) * n_rays, st));
    if (n > 0) {
        count_rays_kernel<<<div_up(n, 256), 256, 0, st>>>(ray_indices, n, n_rays,
                                                         reinterpret_cast<unsigned long long *>(counts));
        UMHS_CHECK_LAUNCH();
    }
    pack_scan_kernel<<<1, kScanThreads, 0, st>>>(counts, n_rays, packed_info, nullptr, 0);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

extern "C" int umhs_occ_cell_points(const int64_t *cell_ids, const float *jitter, uint64_t seed,
                                    const uint64_t *seed_dev, int64_t n, int32_t resolution, const float *aabb6,
                                    float *positions, void *stream) {
    UMHS_CHECK_ARG(aabb6 && positions && resolution > 0, "bad argument");
    if (n <= 0) return UMHS_OK;
    occ_cell_points_kernel<<<div_up(n, 256), 256, 0, as_stream(stream)>>>(cell_ids, jitter, seed, seed_dev, n,
                                                                          resolution, aabb6, positions);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

extern "C" int umhs_occ_select_cells(const uint32_t *bits, int64_t n_cells, int64_t n_uniform, const float *u_uniform,
                                     const float *u_pick, uint64_t seed, const uint64_t *seed_dev, int32_t *prefix,
                                     int64_t *cell_ids, int64_t *n_valid, void *stream) {
    UMHS_CHECK_ARG(bits && prefix && cell_ids && n_cells > 0 && n_uniform > 0, "bad argument");
    UMHS_CHECK_ARG(n_cells < ((int64_t)1 << 31), "grid too large");
    const int64_t n_words = (n_cells + 31) / 32;
    const int n_blocks = div_up(n_words, kPrefixThreads);
    UMHS_CHECK_ARG(n_blocks <= kMaxPrefixBlocks, "grid too large");
    occ_popc_prefix_kernel<<<n_blocks, kPrefixThreads, 0, as_stream(stream)>>>(bits, n_words, prefix);
    UMHS_CHECK_LAUNCH();
    occ_select_kernel<<<div_up(2 * n_uniform, 256), 256, (n_blocks + 1) * sizeof(int32_t), as_stream(stream)>>>(
        bits, prefix, n_words, n_cells, n_uniform, u_uniform, u_pick, seed, seed_dev, cell_ids, n_valid);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

extern "C" int umhs_occ_ema_update(float *occs, const int64_t *cell_ids, int64_t level_offset, const float *sigma,
                                   float scale, float decay, int64_t n, float *tmp, void *stream) {
    UMHS_CHECK_ARG(occs && sigma && tmp, "null argument");
    if (n <= 0) return UMHS_OK;
    occ_ema_read_kernel<<<div_up(n, 256), 256, 0, as_stream(stream)>>>(occs, cell_ids, level_offset, sigma, scale,
                                                                       decay, n, tmp);
    UMHS_CHECK_LAUNCH();
    occ_ema_write_kernel<<<div_up(n, 256), 256, 0, as_stream(stream)>>>(occs, cell_ids, level_offset, n, tmp);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

extern "C" int umhs_occ_binarize(const float *occs, int64_t n_cells, float occ_thre, double *scratch,
                                 uint8_t *binaries, float *mean_out, void *stream) {
    UMHS_CHECK_ARG(occs && scratch && binaries, "null argument");
    if (n_cells <= 0) return UMHS_OK;
    cudaStream_t st = as_stream(stream);
    UMHS_CHECK_CUDA(cudaMemsetAsync(scratch, 0, 2 * sizeof(double), st));
    occ_sum_kernel<<<kNumSMs * 4, 256, 0, st>>>(occs, n_cells, scratch);
    UMHS_CHECK_LAUNCH();
    occ_binarize_kernel<<<div_up(n_cells, 256), 256, 0, st>>>(occs, n_cells, occ_thre, scratch, binaries, mean_out);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}
