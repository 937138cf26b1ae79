// common.cuh -- shared device helpers for libumhs_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/umhs.h"

namespace umhs {

void set_error(const char *fmt, ...);

#define UMHS_CHECK_ARG(cond, msg)                                   \
    do {                                                            \
        if (!(cond)) {                                              \
            umhs::set_error("%s: %s", __func__, msg);               \
            return UMHS_ERR_INVALID;                                \
        }                                                           \
    } while (0)

#define UMHS_CHECK_LAUNCH()                                                          \
    do {                                                                             \
        cudaError_t e__ = cudaGetLastError();                                        \
        if (e__ != cudaSuccess) {                                                    \
            umhs::set_error("%s: CUDA error %s", __func__, cudaGetErrorString(e__)); \
            return UMHS_ERR_CUDA;                                                    \
        }                                                                            \
    } while (0)

#define UMHS_CHECK_CUDA(expr)                                                                  \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            umhs::set_error("%s: %s failed: %s", __func__, #expr, cudaGetErrorString(e__));    \
            return UMHS_ERR_CUDA;                                                              \
        }                                                                                      \
    } while (0)

constexpr int kNumSMs = 148;  // B200
constexpr unsigned kFull = 0xffffffffu;

static inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
    return v;
}

// inclusive prefix sum across the warp (lane order)
__device__ __forceinline__ float warp_incl_scan(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float t = __shfl_up_sync(kFull, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// inclusive suffix sum across the warp (lane i gets sum over lanes >= i)
__device__ __forceinline__ float warp_incl_rscan(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float t = __shfl_down_sync(kFull, v, o);
        if (lane + o < 32) v += t;
    }
    return v;
}

__device__ __forceinline__ float sigmoidf(float x) { return 1.0f / (1.0f + __expf(-x)); }

// A sample point: world position (for the NeRF position encoding), grid coordinate in [0,1]
// already multiplied by the selector (ref umhs_field.py:302-310), selector, and direction.
struct Point {
    float wx, wy, wz;
    float x, y, z;
    float sel;
    float dx, dy, dz;
    float tmid;  // (t0 + t1) / 2 for the ray form, 0 for explicit positions
};

// number of points a kernel processes: the host-side n, clamped by the optional device-side count
__device__ __forceinline__ int64_t effective_n(const umhs_points &P, int64_t n) {
    if (P.n_dev == nullptr) return n;
    const int64_t m = __ldg(P.n_dev);
    return m < n ? (m < 0 ? 0 : m) : n;
}

// row of point n in the hash-encoding array given to the MLP kernels
__device__ __forceinline__ int64_t enc_row(const umhs_points &P, int64_t n) {
    return P.enc_index ? (int64_t)__ldg(P.enc_index + n) : n;
}

// The index-level loads of a point (first level of the dependent load chain), so that a kernel can issue them
// early: ray index and interval of the ray form; nothing for explicit positions.
struct PointIdx {
    int64_t r;
    float t0, t1;
};
__device__ __forceinline__ PointIdx load_point_idx(const umhs_points &P, int64_t n) {
    PointIdx i;
    i.r = 0;
    i.t0 = i.t1 = 0.f;
    if (P.positions == nullptr) {
        i.r = __ldg(P.ray_indices + n);
        i.t0 = __ldg(P.t_starts + n);
        i.t1 = __ldg(P.t_ends + n);
    }
    return i;
}

// The data-level loads of a point (second level of the chain), kept raw so that a kernel can issue them a tile
// ahead and do the arithmetic when the values are needed: position (explicit form) or ray origin (ray form) in a*,
// direction in d*, t0 + t1 in tm.
struct PointRaw {
    float a0, a1, a2, d0, d1, d2, tm;
};
__device__ __forceinline__ PointRaw load_point_raw(const umhs_points &P, int64_t n, const PointIdx &pi) {
    PointRaw q;
    if (P.positions != nullptr) {
        q.a0 = __ldg(P.positions + 3 * n + 0);
        q.a1 = __ldg(P.positions + 3 * n + 1);
        q.a2 = __ldg(P.positions + 3 * n + 2);
        if (P.directions != nullptr) {
            q.d0 = __ldg(P.directions + 3 * n + 0);
            q.d1 = __ldg(P.directions + 3 * n + 1);
            q.d2 = __ldg(P.directions + 3 * n + 2);
        } else {
            q.d0 = q.d1 = q.d2 = 1.0f;
        }
        q.tm = 0.0f;
    } else {
        const int64_t r = pi.r;
        q.tm = (pi.t0 + pi.t1);
        q.d0 = __ldg(P.rays_d + 3 * r + 0);
        q.d1 = __ldg(P.rays_d + 3 * r + 1);
        q.d2 = __ldg(P.rays_d + 3 * r + 2);
        q.a0 = __ldg(P.rays_o + 3 * r + 0);
        q.a1 = __ldg(P.rays_o + 3 * r + 1);
        q.a2 = __ldg(P.rays_o + 3 * r + 2);
    }
    return q;
}

__device__ __forceinline__ Point make_point(const umhs_points &P, const PointRaw &q) {
    Point p;
    p.dx = q.d0;
    p.dy = q.d1;
    p.dz = q.d2;
    if (P.positions != nullptr) {
        p.wx = q.a0;
        p.wy = q.a1;
        p.wz = q.a2;
        p.tmid = 0.0f;
    } else {
        // origins + directions * (starts + ends) / 2   (nerfstudio Frustums.get_positions)
        p.wx = q.a0 + (p.dx * q.tm) / 2.0f;
        p.wy = q.a1 + (p.dy * q.tm) / 2.0f;
        p.wz = q.a2 + (p.dz * q.tm) / 2.0f;
        p.tmid = q.tm / 2.0f;
    }
    float x = p.wx, y = p.wy, z = p.wz;
    if (P.contraction) {
        const float mag = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
        if (!(mag < 1.0f)) {
            const float s = 2.0f - (1.0f / mag);
            x = s * (x / mag);
            y = s * (y / mag);
            z = s * (z / mag);
        }
        x = (x + 2.0f) / 4.0f;
        y = (y + 2.0f) / 4.0f;
        z = (z + 2.0f) / 4.0f;
    } else {
        x = (x - P.aabb[0]) / (P.aabb[3] - P.aabb[0]);
        y = (y - P.aabb[1]) / (P.aabb[4] - P.aabb[1]);
        z = (z - P.aabb[2]) / (P.aabb[5] - P.aabb[2]);
    }
    const bool sel = (x > 0.0f) && (x < 1.0f) && (y > 0.0f) && (y < 1.0f) && (z > 0.0f) && (z < 1.0f);
    p.sel = sel ? 1.0f : 0.0f;
    p.x = sel ? x : 0.0f;
    p.y = sel ? y : 0.0f;
    p.z = sel ? z : 0.0f;
    return p;
}

__device__ __forceinline__ Point eval_point(const umhs_points &P, int64_t n, const PointIdx &pi) {
    return make_point(P, load_point_raw(P, n, pi));
}

__device__ __forceinline__ Point eval_point(const umhs_points &P, int64_t n) {
    return eval_point(P, n, load_point_idx(P, n));
}

// ---- "tile" layout of the hash encodings / their gradients between the hash-grid and the tensor-core MLP kernels:
// bf16, 128-sample tiles of 8 KB stored exactly as the tensor core wants its K-major operand (8 x 8 core matrices, no
// swizzle): element (sample n, column c) lives at
//     (n >> 7) * 8192 + (c >> 3) * 2048 + ((n & 127) >> 3) * 128 + (n & 7) * 16 + (c & 7) * 2
// so a tile is ONE contiguous 8 KB block that goes to shared memory without conversion (four 16-byte pieces per row, or a
// single bulk copy), and a warp's 32 consecutive rows write 512 contiguous bytes per 8-column piece.  Half the bytes of
// the fp32 row-major form, which stays the format of the module-boundary / fp32 kernels.
__device__ __forceinline__ int64_t enc_tile_piece(int64_t n, int piece) {  // byte offset of columns [8 piece, 8 piece + 8)
    return (n >> 7) * 8192 + (int64_t)piece * 2048 + ((n & 127) >> 3) * 128 + (n & 7) * 16;
}

}  // namespace umhs
