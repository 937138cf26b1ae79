// batch.cu -- one training batch straight from device-resident images: pixel sampling, per-pixel gather of the RGB
// and hyperspectral ground truth, camera ray generation.  One kernel.
//
// Replaces UMHSDataManager.next_train (ref umhsnerf/data/umhs_datamanager.py:95-108):
//   batch = train_pixel_sampler.sample(image_batch)    -> nerfstudio PixelSampler: floor(rand * [N, H, W]),
//                                                         {key: value[c, y, x]}, indices[:, 0] = image_idx[c]
//   ray_bundle = train_ray_generator(batch["indices"]) -> nerfstudio RayGenerator / Cameras.generate_rays
// on hyperspectral cubes loaded as in ref umhsnerf/data/utils/hs_dataloader.py:46-50 (H x W x B per image).
// Arithmetic follows oracle/datamanager_oracle.py (perspective cameras, no distortion).
//
// Mapping: one warp per 32 rays: lane = ray for the index / camera math and the per-ray vectors, then the warp copies
// the 32 pixels' B-band spectra as contiguous rows (128 bands fp32 = 512 B = one float4 per lane), the pixel index
// broadcast by shuffle.  HBM-bound: a random 4B-byte row read and a 4B-byte row written per ray; algorithmic bytes per ray = 12 (draws) + 24 (indices) + 32 (origin, direction, area,
// norm) + sizeof(hs) B + 4 B (+ image in / out).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace umhs {

constexpr int kBatchWarps = 2;  // 64 rays per CTA: a 16 k-ray batch still spreads over every SM
constexpr int kBatchThreads = kBatchWarps * 32;

template <int DT>
struct HsElem;
template <>
struct HsElem<UMHS_DTYPE_F32> {
    using type = float;
    static __device__ __forceinline__ float get(const float *p, int64_t i) { return __ldg(p + i); }
};
template <>
struct HsElem<UMHS_DTYPE_F16> {
    using type = __half;
    static __device__ __forceinline__ float get(const __half *p, int64_t i) {
        return __half2float(__ushort_as_half(__ldg(reinterpret_cast<const unsigned short *>(p) + i)));
    }
};
template <>
struct HsElem<UMHS_DTYPE_BF16> {
    using type = __nv_bfloat16;
    static __device__ __forceinline__ float get(const __nv_bfloat16 *p, int64_t i) {
        return __uint_as_float((uint32_t)__ldg(reinterpret_cast<const unsigned short *>(p) + i) << 16);
    }
};

// four consecutive elements starting at element 4*j of a row whose first element is 8- (16-bit) / 16-byte (fp32) aligned
template <int DT>
__device__ __forceinline__ float4 load4(const typename HsElem<DT>::type *row, int j) {
    if (DT == UMHS_DTYPE_F32) {
        return __ldg(reinterpret_cast<const float4 *>(row) + j);
    } else {
        const uint2 v = __ldg(reinterpret_cast<const uint2 *>(row) + j);
        float4 r;
        if (DT == UMHS_DTYPE_F16) {
            const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&v.x));
            const float2 b = __half22float2(*reinterpret_cast<const __half2 *>(&v.y));
            r = make_float4(a.x, a.y, b.x, b.y);
        } else {
            r = make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u),
                            __uint_as_float(v.y << 16), __uint_as_float(v.y & 0xffff0000u));
        }
        return r;
    }
}

__device__ __forceinline__ int draw_index(float u, int bound) {
    const int i = (int)(u * (float)bound);  // fp32 product, truncation (torch: (rand * bound).long())
    return i < bound ? (i < 0 ? 0 : i) : bound - 1;
}

// One warp per 32 rays.  Phase 1: lane l does the scalar work of ray base + l (indices, camera ray, RGB pixel) -- one
// pass of ~150 instructions per 32 rays; the first form of this kernel had every lane repeat it for ONE ray per warp
// and was bound by instruction issue (200 warp instructions per ray, 1.7 TB/s).  Phase 2: the warp copies the 32
// spectra row by row (pixel index broadcast by shuffle), four rows in flight.
template <int DT>
__global__ void __launch_bounds__(kBatchThreads, 16)
sample_batch_kernel(umhs_batch_in in, umhs_batch_out out, int vec, int rpw) {
    using T = typename HsElem<DT>::type;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int B = in.num_bands, C = in.image_channels;
    const T *hs = reinterpret_cast<const T *>(in.hs);
    // rpw rays per warp (32 for large batches; fewer for small ones, upper lanes idle in phase 1, so that a 16 k-ray
    // batch is still thousands of warps and not 512 warps copying 32 rows one after the other)
    for (int64_t base = ((int64_t)blockIdx.x * kBatchWarps + warp) * rpw; base < in.n_rays;
         base += (int64_t)gridDim.x * kBatchWarps * rpw) {
        const int64_t r = base + lane;
        const bool valid = lane < rpw && r < in.n_rays;
        int64_t pix = 0;
        if (valid) {
            const int c = draw_index(__ldg(in.u + 3 * r + 0), in.num_images);
            const int y = draw_index(__ldg(in.u + 3 * r + 1), in.height);
            const int x = draw_index(__ldg(in.u + 3 * r + 2), in.width);
            const int64_t cam = in.image_idx ? __ldg(in.image_idx + c) : (int64_t)c;
            pix = ((int64_t)c * in.height + y) * in.width + x;
            if (out.indices != nullptr) {
                out.indices[3 * r + 0] = cam;
                out.indices[3 * r + 1] = y;
                out.indices[3 * r + 2] = x;
            }
            if (in.image != nullptr && out.image != nullptr) {
                for (int ch = 0; ch < C; ++ch)
                    out.image[r * C + ch] = in.image_dtype == UMHS_DTYPE_U8
                        ? (float)__ldg(reinterpret_cast<const uint8_t *>(in.image) + pix * C + ch) / 255.0f
                        : __ldg(reinterpret_cast<const float *>(in.image) + pix * C + ch);
            }
            // ---- camera ray through the pixel centre (Cameras.generate_rays, perspective)
            if (in.c2w != nullptr) {
                const float4 k = __ldg(reinterpret_cast<const float4 *>(in.intrinsics) + cam);  // fx, fy, cx, cy
                const float *M = in.c2w + cam * 12;
                const float px = (float)x + 0.5f, py = (float)y + 0.5f;
                // three rays: the pixel's, one pixel to the right, one pixel down (for the pixel area)
                const float cxs[3] = {(px - k.z) / k.x, (px - k.z + 1.0f) / k.x, (px - k.z) / k.x};
                const float cys[3] = {-(py - k.w) / k.y, -(py - k.w) / k.y, -(py - k.w + 1.0f) / k.y};
                float m[12];
#pragma unroll
                for (int i = 0; i < 12; ++i) m[i] = __ldg(M + i);
                float d[3][3], nrm0 = 0.f;
#pragma unroll
                for (int v = 0; v < 3; ++v) {
                    float w[3], s = 0.f;
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        w[i] = (cxs[v] * m[4 * i + 0] + cys[v] * m[4 * i + 1]) + (-1.0f) * m[4 * i + 2];
                        s += w[i] * w[i];
                    }
                    const float n = fmaxf(sqrtf(s), 1e-10f);
                    if (v == 0) nrm0 = sqrtf(s);
#pragma unroll
                    for (int i = 0; i < 3; ++i) d[v][i] = w[i] / n;
                }
                float dx = 0.f, dy = 0.f;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    dx += (d[0][i] - d[1][i]) * (d[0][i] - d[1][i]);
                    dy += (d[0][i] - d[2][i]) * (d[0][i] - d[2][i]);
                }
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    if (out.origins) out.origins[3 * r + i] = m[4 * i + 3];
                    if (out.directions) out.directions[3 * r + i] = d[0][i];
                }
                if (out.pixel_area) out.pixel_area[r] = sqrtf(dx) * sqrtf(dy);
                if (out.directions_norm) out.directions_norm[r] = nrm0;
            }
        }
        // ---- spectra of the 32 pixels: one contiguous row each
        if (hs != nullptr && out.hs_image != nullptr) {
            const int64_t left = in.n_rays - base;
            const int m32 = left < rpw ? (int)left : rpw;
#pragma unroll 4
            for (int j = 0; j < m32; ++j) {
                const int64_t pj = __shfl_sync(kFull, pix, j);
                const T *src = hs + pj * B;
                float *dst = out.hs_image + (base + j) * B;
                if (vec) {
                    for (int q = lane; 4 * q < B; q += 32) reinterpret_cast<float4 *>(dst)[q] = load4<DT>(src, q);
                } else {
                    for (int b = lane; b < B; b += 32) dst[b] = HsElem<DT>::get(src, b);
                }
            }
        }
    }
}

}  // namespace umhs

using namespace umhs;

extern "C" int umhs_sample_batch(const umhs_batch_in *in, const umhs_batch_out *out, void *stream) {
    UMHS_CHECK_ARG(in != nullptr && out != nullptr, "null argument");
    if (in->n_rays <= 0) return UMHS_OK;
    UMHS_CHECK_ARG(in->u != nullptr, "the uniform draws are required");
    UMHS_CHECK_ARG(in->num_images > 0 && in->height > 0 && in->width > 0, "empty image batch");
    UMHS_CHECK_ARG(in->hs == nullptr || (in->num_bands > 0 && in->num_bands <= 4096), "bad band count");
    UMHS_CHECK_ARG(in->image == nullptr || (in->image_channels >= 1 && in->image_channels <= 4), "bad channel count");
    UMHS_CHECK_ARG(in->image == nullptr || in->image_dtype == UMHS_DTYPE_F32 || in->image_dtype == UMHS_DTYPE_U8,
                   "image must be fp32 or uint8");
    UMHS_CHECK_ARG(in->c2w == nullptr || in->intrinsics != nullptr, "cameras need intrinsics");
    // measured at 128 fp32 bands: 32 rays per warp 3.9 TB/s at 262 k rays but 30 us at 16 k rays (one ray per warp: 18 us)
    const int rpw = in->n_rays >= 131072 ? 32 : (in->n_rays >= 32768 ? 8 : 4);
    const int64_t want = (in->n_rays + kBatchWarps * rpw - 1) / (kBatchWarps * rpw);
    const int64_t cap = (int64_t)kNumSMs * 16;
    const int grid = (int)(want < cap ? want : cap);
    const uintptr_t a_in = reinterpret_cast<uintptr_t>(in->hs), a_out = reinterpret_cast<uintptr_t>(out->hs_image);
    const bool b4 = in->num_bands % 4 == 0;
    cudaStream_t st = as_stream(stream);
    switch (in->hs_dtype) {
        case UMHS_DTYPE_F32:
            sample_batch_kernel<UMHS_DTYPE_F32><<<grid, kBatchThreads, 0, st>>>(
                *in, *out, b4 && a_in % 16 == 0 && a_out % 16 == 0, rpw);
            break;
        case UMHS_DTYPE_F16:
            sample_batch_kernel<UMHS_DTYPE_F16><<<grid, kBatchThreads, 0, st>>>(
                *in, *out, b4 && a_in % 8 == 0 && a_out % 16 == 0, rpw);
            break;
        case UMHS_DTYPE_BF16:
            sample_batch_kernel<UMHS_DTYPE_BF16><<<grid, kBatchThreads, 0, st>>>(
                *in, *out, b4 && a_in % 8 == 0 && a_out % 16 == 0, rpw);
            break;
        default:
            UMHS_CHECK_ARG(false, "hs_dtype must be UMHS_DTYPE_F32 / F16 / BF16");
    }
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}
