// hash_grid.cu -- multiresolution hash-grid encode, forward (gather) and backward (scatter).
//
// Replaces nerfstudio HashEncoding.pytorch_fwd reached from ref umhsnerf/umhs_field.py:320
// (self.mlp_base(positions_flat)) and its autograd backward.  Arithmetic follows
// oracle/nerf3p.py::hash_encode: ceil/floor corners, ALL levels hashed with
// (x*1) ^ (y*2654435761) ^ (z*805459861) mod T, blend in the reference's association order.
//
// Mapping: one thread per sample, loop over the L levels.  Consecutive threads are consecutive
// samples of one ray, so at coarse levels a warp reads a handful of distinct 8-byte entries
// (served from L1/L2); the whole 64 MiB table fits the 126 MB L2.  Each level issues 8
// independent float2 gathers; two levels are kept in flight (16 loads per thread).
// Output enc [n, 2L] fp32 row-major (level-major columns), written as float4 per 2 levels.
//
// x-pairs: the hash multiplies x by 1, so for an EVEN floor coordinate fx the floor and ceil corners of a cell edge
// along x are (fx ^ h) and ((fx + 1) ^ h) = (fx ^ h) ^ 1: the two rows of one 16-byte aligned pair of table rows.
// Such a sample (half of them at fine levels) gathers its 8 corners with 4 LDG.128 and scatters its gradient with
// 4 red.global.add.v4.f32 instead of 8 eight-byte accesses.
//
// Algorithmic bytes per sample: fwd 8*L*8 gathered + 12 in + 8L out;
// bwd 8L in + 8*L*8 read-modify-write.
#include "common.cuh"
#include "hash_scatter.cuh"

namespace umhs {

// the 8 corner rows of one level: 4 aligned row pairs when `pair`, 8 single rows otherwise
__device__ __forceinline__ void gather_level(const float2 *__restrict__ table, const uint32_t (&idx)[8], bool pair,
                                             float2 (&f)[8]) {
    constexpr int kFloor[4] = {3, 2, 7, 6}, kCeil[4] = {0, 1, 4, 5};
    if (pair) {
        const float4 *t4 = reinterpret_cast<const float4 *>(table);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const uint32_t fi = idx[kFloor[p]];
            const float4 v = __ldg(t4 + (fi >> 1));
            const bool odd = (fi & 1u) != 0u;  // the floor corner is the second row of the pair
            f[kFloor[p]] = odd ? make_float2(v.z, v.w) : make_float2(v.x, v.y);
            f[kCeil[p]] = odd ? make_float2(v.x, v.y) : make_float2(v.z, v.w);
        }
    } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) f[c] = __ldg(table + idx[c]);
    }
}

__device__ __forceinline__ float2 blend(const float2 (&f)[8], float ox, float oy, float oz) {
    const float ux = 1.0f - ox, uy = 1.0f - oy, uz = 1.0f - oz;
    float2 r;
    {
        const float f03 = f[0].x * ox + f[3].x * ux;
        const float f12 = f[1].x * ox + f[2].x * ux;
        const float f56 = f[5].x * ox + f[6].x * ux;
        const float f47 = f[4].x * ox + f[7].x * ux;
        const float f0312 = f03 * oy + f12 * uy;
        const float f4756 = f47 * oy + f56 * uy;
        r.x = f0312 * oz + f4756 * uz;
    }
    {
        const float f03 = f[0].y * ox + f[3].y * ux;
        const float f12 = f[1].y * ox + f[2].y * ux;
        const float f56 = f[5].y * ox + f[6].y * ux;
        const float f47 = f[4].y * ox + f[7].y * ux;
        const float f0312 = f03 * oy + f12 * uy;
        const float f4756 = f47 * oy + f56 * uy;
        r.y = f0312 * oz + f4756 * uz;
    }
    return r;
}

struct SrcPoints {
    umhs_points P;
    __device__ __forceinline__ int64_t count(int64_t n) const { return effective_n(P, n); }
    __device__ __forceinline__ void get(int64_t n, float &x, float &y, float &z) const {
        const Point p = eval_point(P, n);
        x = p.x;
        y = p.y;
        z = p.z;
    }
};

struct SrcX01 {
    const float *x01;
    __device__ __forceinline__ int64_t count(int64_t n) const { return n; }
    __device__ __forceinline__ void get(int64_t n, float &x, float &y, float &z) const {
        x = __ldg(x01 + 3 * n + 0);
        y = __ldg(x01 + 3 * n + 1);
        z = __ldg(x01 + 3 * n + 2);
    }
};

constexpr int kHashThreads = 256;

__device__ __forceinline__ uint32_t pack_bf16x2_rn(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}

// TILES: enc is the bf16 tile buffer (common.cuh: enc_tile_piece; L must be 16), written as one 16-byte piece per four
// levels; otherwise fp32 rows [n, 2L].
template <class Src, bool PAIRS, bool TILES>
__global__ void __launch_bounds__(kHashThreads)
hash_encode_fwd_kernel(Src src, int64_t n_total, const float2 *__restrict__ table, const float *__restrict__ scalings,
                       int L, int log2T, float *__restrict__ enc) {
    const int64_t n = (int64_t)blockIdx.x * kHashThreads + threadIdx.x;
    if (n >= src.count(n_total)) return;
    float x, y, z;
    src.get(n, x, y, z);
    const uint32_t tmask = (log2T >= 32) ? 0xffffffffu : ((1u << log2T) - 1u);
    float *out = enc + n * (2 * L);
    uint8_t *tiles = reinterpret_cast<uint8_t *>(enc);
    uint32_t half_piece[2] = {0u, 0u};  // TILES: columns of the first level pair of a piece
    int l = 0;
    for (; l + 1 < L; l += 2) {
        uint32_t i0[8], i1[8];
        float ox0, oy0, oz0, ox1, oy1, oz1;
        const bool p0 = level_setup(x, y, z, __ldg(scalings + l), tmask, (uint32_t)l << log2T, i0, ox0, oy0, oz0);
        const bool p1 =
            level_setup(x, y, z, __ldg(scalings + l + 1), tmask, (uint32_t)(l + 1) << log2T, i1, ox1, oy1, oz1);
        float2 f0[8], f1[8];
        gather_level(table, i0, PAIRS && p0, f0);
        gather_level(table, i1, PAIRS && p1, f1);
        const float2 a = blend(f0, ox0, oy0, oz0);
        const float2 b = blend(f1, ox1, oy1, oz1);
        if (TILES) {
            if ((l & 2) == 0) {
                half_piece[0] = pack_bf16x2_rn(a.x, a.y);
                half_piece[1] = pack_bf16x2_rn(b.x, b.y);
            } else {
                *reinterpret_cast<uint4 *>(tiles + enc_tile_piece(n, l >> 2)) =
                    make_uint4(half_piece[0], half_piece[1], pack_bf16x2_rn(a.x, a.y), pack_bf16x2_rn(b.x, b.y));
            }
        } else if ((L & 1) == 0) {
            *reinterpret_cast<float4 *>(out + 2 * l) = make_float4(a.x, a.y, b.x, b.y);
        } else {
            *reinterpret_cast<float2 *>(out + 2 * l) = a;
            *reinterpret_cast<float2 *>(out + 2 * l + 2) = b;
        }
    }
    if (l < L) {
        uint32_t i0[8];
        float ox0, oy0, oz0;
        const bool p0 = level_setup(x, y, z, __ldg(scalings + l), tmask, (uint32_t)l << log2T, i0, ox0, oy0, oz0);
        float2 f0[8];
        gather_level(table, i0, PAIRS && p0, f0);
        *reinterpret_cast<float2 *>(out + 2 * l) = blend(f0, ox0, oy0, oz0);
    }
}

// Scatter kernel: 128-thread CTAs, grid-stride over the samples.
// Measured and dropped (round 1): scattering chunk i of the samples on a second stream, on a grid capped to 3 CTAs per
// SM so that it fits beside the persistent tensor-core MLP backward of chunk i+1.  The scatter needs its occupancy
// (capped to 12 warps per SM it is ~2x slower) and every extra launch of the persistent MLP kernel costs ~0.05 ms:
// 2 chunks 1.54 -> 1.68 ms per step, 3 chunks 1.77 ms.
constexpr int kScatterThreads = 128;

template <class Src, bool PAIRS, bool TILES>
__global__ void __launch_bounds__(kScatterThreads)
hash_encode_bwd_kernel(Src src, int64_t n_total, const float *__restrict__ scalings, int L, int log2T,
                       const float *__restrict__ d_enc, float2 *__restrict__ d_table, int agg_levels, int l_begin,
                       int l_end) {
    const int lane = threadIdx.x & 31;
    const int64_t n_eff = src.count(n_total);
    // CTA-uniform trip count: the warp-aggregated levels shuffle across the whole warp
    for (int64_t base = (int64_t)blockIdx.x * kScatterThreads; base < n_eff; base += (int64_t)gridDim.x * kScatterThreads) {
        const int64_t n = base + threadIdx.x;
        const bool valid = n < n_eff;
        float x = 0.f, y = 0.f, z = 0.f;
        if (valid) src.get(n, x, y, z);
        const float *g = d_enc + (valid ? n : 0) * (2 * L);
        const uint8_t *gt = reinterpret_cast<const uint8_t *>(d_enc);  // TILES: bf16 tile buffer
        scatter_sample<PAIRS>(x, y, z, valid, [&](int l) -> float2 {
            if (TILES)  // one bf16 pair inside the row's 16-byte piece l / 4
                return unpack_tile_pair(__ldg(reinterpret_cast<const uint32_t *>(gt + enc_tile_piece(n, l >> 2)) + (l & 3)));
            return __ldg(reinterpret_cast<const float2 *>(g + 2 * l));
        }, scalings, l_begin, l_end, log2T, d_table, agg_levels, lane);
    }
}

static int check_grid(const umhs_hash_grid *g) {
    if (!g || !g->table || !g->scalings) return -1;
    if (g->num_levels < 1 || g->num_levels > UMHS_MAX_LEVELS) return -1;
    if (g->log2_table_size < 1 || g->log2_table_size > 28) return -1;
    if ((int64_t)g->num_levels << g->log2_table_size > (int64_t)1 << 32) return -1;
    return 0;
}

// number of leading (coarse) levels whose scatter is warp-aggregated: cell edge >= ~2 sample steps
static int agg_levels_for(const umhs_hash_grid *g) { return g->num_levels < 10 ? g->num_levels : 10; }

static int scatter_grid(int64_t n, int max_ctas) {
    const int64_t want = (n + kScatterThreads - 1) / kScatterThreads;
    return (int)((max_ctas > 0 && want > max_ctas) ? max_ctas : want);
}

}  // namespace umhs

using namespace umhs;

extern "C" int umhs_hash_encode_fwd(const umhs_points *pts, int64_t n, const umhs_hash_grid *grid, float *enc,
                                    void *stream) {
    if (n <= 0) return UMHS_OK;
    UMHS_CHECK_ARG(pts != nullptr && enc != nullptr, "null argument");
    UMHS_CHECK_ARG(check_grid(grid) == 0, "bad hash grid");
    UMHS_CHECK_ARG(pts->positions || (pts->rays_o && pts->rays_d && pts->ray_indices && pts->t_starts && pts->t_ends),
                   "points need positions or the ray form");
    if (n <= 0) return UMHS_OK;
    SrcPoints src{*pts};
    hash_encode_fwd_kernel<decltype(src), true, false><<<div_up(n, kHashThreads), kHashThreads, 0, as_stream(stream)>>>(
        src, n, reinterpret_cast<const float2 *>(grid->table), grid->scalings, grid->num_levels,
        grid->log2_table_size, enc);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

extern "C" int umhs_hash_encode_fwd_tiles(const umhs_points *pts, int64_t n, const umhs_hash_grid *grid, void *enc_tiles,
                                          void *stream) {
    if (n <= 0) return UMHS_OK;
    UMHS_CHECK_ARG(pts != nullptr && enc_tiles != nullptr, "null argument");
    UMHS_CHECK_ARG(check_grid(grid) == 0 && grid->num_levels == 16, "the tile layout is 16 levels x 2 features wide");
    UMHS_CHECK_ARG(pts->positions || (pts->rays_o && pts->rays_d && pts->ray_indices && pts->t_starts && pts->t_ends),
                   "points need positions or the ray form");
    UMHS_CHECK_ARG((reinterpret_cast<uintptr_t>(enc_tiles) & 15u) == 0, "enc_tiles must be 16-byte aligned");
    SrcPoints src{*pts};
    hash_encode_fwd_kernel<decltype(src), true, true><<<div_up(n, kHashThreads), kHashThreads, 0, as_stream(stream)>>>(
        src, n, reinterpret_cast<const float2 *>(grid->table), grid->scalings, grid->num_levels,
        grid->log2_table_size, reinterpret_cast<float *>(enc_tiles));
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

extern "C" int umhs_hash_encode_bwd_tiles_levels(const umhs_points *pts, int64_t n, const umhs_hash_grid *grid,
                                                 const void *d_enc_tiles, float *d_table, int32_t level_begin,
                                                 int32_t level_end, void *stream) {
    if (n <= 0) return UMHS_OK;
    UMHS_CHECK_ARG(pts != nullptr && d_enc_tiles != nullptr && d_table != nullptr, "null argument");
    UMHS_CHECK_ARG(check_grid(grid) == 0 && grid->num_levels == 16, "the tile layout is 16 levels x 2 features wide");
    UMHS_CHECK_ARG(level_begin >= 0 && level_begin <= level_end && level_end <= grid->num_levels, "bad level range");
    if (level_begin == level_end) return UMHS_OK;
    SrcPoints src{*pts};
    hash_encode_bwd_kernel<decltype(src), true, true><<<scatter_grid(n, 0), kScatterThreads, 0, as_stream(stream)>>>(
        src, n, grid->scalings, grid->num_levels, grid->log2_table_size, reinterpret_cast<const float *>(d_enc_tiles),
        reinterpret_cast<float2 *>(d_table), agg_levels_for(grid), level_begin, level_end);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

extern "C" int umhs_hash_encode_bwd_tiles(const umhs_points *pts, int64_t n, const umhs_hash_grid *grid,
                                          const void *d_enc_tiles, float *d_table, void *stream) {
    UMHS_CHECK_ARG(grid != nullptr, "null argument");
    return umhs_hash_encode_bwd_tiles_levels(pts, n, grid, d_enc_tiles, d_table, 0, grid->num_levels, stream);
}

extern "C" int umhs_hash_encode_bwd(const umhs_points *pts, int64_t n, const umhs_hash_grid *grid, const float *d_enc,
                                    float *d_table, void *stream) {
    if (n <= 0) return UMHS_OK;
    UMHS_CHECK_ARG(pts != nullptr && d_enc != nullptr && d_table != nullptr, "null argument");
    UMHS_CHECK_ARG(check_grid(grid) == 0, "bad hash grid");
    if (n <= 0) return UMHS_OK;
    SrcPoints src{*pts};
    const int max_ctas = 0;
    hash_encode_bwd_kernel<decltype(src), true, false><<<scatter_grid(n, max_ctas), kScatterThreads, 0, as_stream(stream)>>>(
        src, n, grid->scalings, grid->num_levels, grid->log2_table_size, d_enc, reinterpret_cast<float2 *>(d_table),
        agg_levels_for(grid), 0, grid->num_levels);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

extern "C" int umhs_hash_encode_x01_fwd(const float *x01, int64_t n, const umhs_hash_grid *grid, float *enc,
                                        void *stream) {
    if (n <= 0) return UMHS_OK;
    UMHS_CHECK_ARG(x01 != nullptr && enc != nullptr, "null argument");
    UMHS_CHECK_ARG(check_grid(grid) == 0, "bad hash grid");
    if (n <= 0) return UMHS_OK;
    SrcX01 src{x01};
    hash_encode_fwd_kernel<decltype(src), true, false><<<div_up(n, kHashThreads), kHashThreads, 0, as_stream(stream)>>>(
        src, n, reinterpret_cast<const float2 *>(grid->table), grid->scalings, grid->num_levels,
        grid->log2_table_size, enc);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}

extern "C" int umhs_hash_encode_x01_bwd(const float *x01, int64_t n, const umhs_hash_grid *grid, const float *d_enc,
                                        float *d_table, void *stream) {
    if (n <= 0) return UMHS_OK;
    UMHS_CHECK_ARG(x01 != nullptr && d_enc != nullptr && d_table != nullptr, "null argument");
    UMHS_CHECK_ARG(check_grid(grid) == 0, "bad hash grid");
    if (n <= 0) return UMHS_OK;
    SrcX01 src{x01};
    const int max_ctas = 0;
    hash_encode_bwd_kernel<decltype(src), true, false><<<scatter_grid(n, max_ctas), kScatterThreads, 0, as_stream(stream)>>>(
        src, n, grid->scalings, grid->num_levels, grid->log2_table_size, d_enc, reinterpret_cast<float2 *>(d_table),
        agg_levels_for(grid), 0, grid->num_levels);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}
