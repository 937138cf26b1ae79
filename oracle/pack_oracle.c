/*
 * pack_oracle.c -- CPU restatement of the occupancy-grid ray march ("packing").
 *
 * TEST INFRASTRUCTURE ONLY (checker for tests/, smoke() and bench.py's CPU legs).
 *
 * Reference call site: ref umhsnerf/umhs_model.py:229-237 (self.sampler(...)), grid
 * built at ref umhsnerf/umhs_model.py:198-205.  The arithmetic lives in the
 * third-party package nerfacc == 0.5.2 (pinned: ref nerfstudioa100_environment.yml:188;
 * OccGridEstimator.sampling -> traverse_grids, csrc/grid.cu), which is NOT under
 * /root/reference and not installed here.  This file restates its published
 * algorithm as described in SURVEY.md Appendix A.4.  PARITY UNPINNED: the
 * reference holds no golden vector for this path.
 *
 * Restatement choices (documented, identical in the CUDA kernel packing.cu):
 *   - an interval between two consecutive sorted box-hit t values is marched in the
 *     FINEST grid level whose [tmin,tmax] contains it (equivalent to nerfacc's
 *     enter/leave bookkeeping for nested concentric boxes, independent of how
 *     torch.sort breaks ties);
 *   - eps = 1e-6 when locating the start / end voxel;
 *   - a miss is t = +1e10 for both ends.
 *
 * Compile with -O2 -ffp-contract=off (no FMA contraction): the CUDA twin is built
 * with -fmad=false, so both sides execute the same IEEE fp32 operation sequence
 * and the outputs (ray_indices, t_starts, t_ends, packed_info) are BIT-EXACT.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#define MAX_LEVELS 8
#define MISS_T 1e10f

static inline int ray_aabb(const float o[3], const float inv[3], const float *aabb, float *tmin_o, float *tmax_o)
{
    float tmin, tmax, t0, t1;
    if (inv[0] >= 0) { tmin = (aabb[0] - o[0]) * inv[0]; tmax = (aabb[3] - o[0]) * inv[0]; }
    else             { tmin = (aabb[3] - o[0]) * inv[0]; tmax = (aabb[0] - o[0]) * inv[0]; }
    if (inv[1] >= 0) { t0 = (aabb[1] - o[1]) * inv[1]; t1 = (aabb[4] - o[1]) * inv[1]; }
    else             { t0 = (aabb[4] - o[1]) * inv[1]; t1 = (aabb[1] - o[1]) * inv[1]; }
    if (tmin > t1 || t0 > tmax) return 0;
    if (t0 > tmin) tmin = t0;
    if (t1 < tmax) tmax = t1;
    if (inv[2] >= 0) { t0 = (aabb[2] - o[2]) * inv[2]; t1 = (aabb[5] - o[2]) * inv[2]; }
    else             { t0 = (aabb[5] - o[2]) * inv[2]; t1 = (aabb[2] - o[2]) * inv[2]; }
    if (tmin > t1 || t0 > tmax) return 0;
    if (t0 > tmin) tmin = t0;
    if (t1 < tmax) tmax = t1;
    if (tmax <= 0) return 0;
    *tmin_o = tmin; *tmax_o = tmax;
    return 1;
}

static inline float calc_dt(float t, float cone_angle, float dt_min, float dt_max)
{
    float v = t * cone_angle;
    v = v < dt_min ? dt_min : v;
    v = v > dt_max ? dt_max : v;
    return v;
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* voxel index of a normalised coordinate: clamp in the float domain first so that the
 * float->int conversion is defined for far-away points (C: UB, CUDA: saturating). */
static inline int voxel_index(float v, float fres, int res)
{
    v = v < -1.0f ? -1.0f : (v > fres ? fres : v);
    return clampi((int)v, 0, res - 1);
}

/* March one ray.  If out arrays are NULL only counts.  Returns number of samples. */
static int64_t march_ray(const float o[3], const float d[3], const uint8_t *binaries, const float *aabbs,
                         int levels, int res, float near_plane, float far_plane, float step_size,
                         float cone_angle, float *t0_out, float *t1_out)
{
    float inv[3] = {1.0f / d[0], 1.0f / d[1], 1.0f / d[2]};
    float tmins[MAX_LEVELS], tmaxs[MAX_LEVELS], ts[2 * MAX_LEVELS];
    int hits[MAX_LEVELS];
    for (int l = 0; l < levels; ++l) {
        float a = MISS_T, b = MISS_T;
        hits[l] = ray_aabb(o, inv, aabbs + 6 * l, &a, &b);
        if (!hits[l]) { a = MISS_T; b = MISS_T; }
        tmins[l] = a; tmaxs[l] = b;
        ts[l] = a; ts[levels + l] = b;
    }
    /* insertion sort of the 2*levels t values */
    for (int i = 1; i < 2 * levels; ++i) {
        float v = ts[i]; int j = i - 1;
        while (j >= 0 && ts[j] > v) { ts[j + 1] = ts[j]; --j; }
        ts[j + 1] = v;
    }
    const float fres = (float)res;
    const int64_t cells_per_level = (int64_t)res * res * res;
    int64_t n = 0;
    float t_last = near_plane;
    for (int i = 0; i < 2 * levels - 1; ++i) {
        const float a = ts[i], b = ts[i + 1];
        if (!(a < b)) continue;
        int level = -1;
        for (int l = 0; l < levels; ++l)
            if (hits[l] && tmins[l] <= a && b <= tmaxs[l]) { level = l; break; }
        if (level < 0) continue;
        const float this_tmin = fmaxf(a, near_plane);
        const float this_tmax = fminf(b, far_plane);
        if (!(this_tmin < this_tmax)) continue;
        const float *aabb = aabbs + 6 * level;
        /* DDA setup */
        const float eps = 1e-6f;
        float voxel[3], delta[3], tdist[3];
        int cur[3], fin[3], stepi[3];
        for (int k = 0; k < 3; ++k) {
            const float ext = aabb[3 + k] - aabb[k];
            voxel[k] = ext / fres;
            const float rs = o[k] + d[k] * (this_tmin + eps);
            const float re = o[k] + d[k] * (this_tmax - eps);
            cur[k] = voxel_index(((rs - aabb[k]) / ext) * fres, fres, res);
            fin[k] = voxel_index(((re - aabb[k]) / ext) * fres, fres, res);
            const int start_index = cur[k] + (d[k] > 0.0f ? 1 : 0);
            const float tmax_k = ((aabb[k] + (((float)start_index * voxel[k]) - rs)) * inv[k]) + this_tmin;
            const float sf = (d[k] == 0.0f) ? 0.0f : (d[k] > 0.0f ? 1.0f : -1.0f);
            stepi[k] = (int)sf;
            tdist[k] = (d[k] == 0.0f) ? this_tmax : tmax_k;
            delta[k] = (d[k] == 0.0f) ? this_tmax : (voxel[k] * inv[k]) * sf;
        }
        const int over[3] = {fin[0] + stepi[0], fin[1] + stepi[1], fin[2] + stepi[2]};
        for (int guard = 0; guard < 4 * res + 8; ++guard) {
            float t_traverse = fminf(tdist[0], fminf(tdist[1], tdist[2]));
            t_traverse = fminf(t_traverse, this_tmax);
            const int64_t cell = (int64_t)level * cells_per_level + ((int64_t)cur[0] * res + cur[1]) * res + cur[2];
            if (!binaries[cell]) {
                /* empty cell: skip ahead in whole steps */
                for (;;) {
                    const float dt = calc_dt(t_last, cone_angle, step_size, 1e10f);
                    if (t_last + dt * 0.5f >= t_traverse) break;
                    t_last += dt;
                }
            } else {
                for (;;) {
                    const float dt = calc_dt(t_last, cone_angle, step_size, 1e10f);
                    if (t_last + dt * 0.5f >= t_traverse) break;
                    if (t0_out) { t0_out[n] = t_last; t1_out[n] = t_last + dt; }
                    ++n;
                    t_last += dt;
                }
            }
            /* single DDA step: x if strictly smallest, else y if < z, else z */
            int ax;
            if (tdist[0] < tdist[1] && tdist[0] < tdist[2]) ax = 0;
            else if (tdist[1] < tdist[2]) ax = 1;
            else ax = 2;
            cur[ax] += stepi[ax];
            tdist[ax] += delta[ax];
            if (cur[ax] == over[ax]) break;
            if (cur[ax] < 0 || cur[ax] >= res) break;
        }
    }
    return n;
}

/* Pass 1: per-ray sample counts. */
void umhs_oracle_march_count(const float *rays_o, const float *rays_d, const uint8_t *binaries, const float *aabbs,
                             int levels, int res, const float *near_planes, const float *far_planes,
                             float step_size, float cone_angle, int64_t n_rays, int64_t *counts)
{
    for (int64_t r = 0; r < n_rays; ++r)
        counts[r] = march_ray(rays_o + 3 * r, rays_d + 3 * r, binaries, aabbs, levels, res, near_planes[r],
                              far_planes[r], step_size, cone_angle, NULL, NULL);
}

/* Pass 2: write samples at offsets[r] (exclusive scan of counts). */
void umhs_oracle_march_write(const float *rays_o, const float *rays_d, const uint8_t *binaries, const float *aabbs,
                             int levels, int res, const float *near_planes, const float *far_planes,
                             float step_size, float cone_angle, int64_t n_rays, const int64_t *offsets,
                             int64_t *ray_indices, float *t_starts, float *t_ends)
{
    for (int64_t r = 0; r < n_rays; ++r) {
        const int64_t off = offsets[r];
        const int64_t n = march_ray(rays_o + 3 * r, rays_d + 3 * r, binaries, aabbs, levels, res, near_planes[r],
                                    far_planes[r], step_size, cone_angle, t_starts + off, t_ends + off);
        for (int64_t i = 0; i < n; ++i) ray_indices[off + i] = r;
    }
}
