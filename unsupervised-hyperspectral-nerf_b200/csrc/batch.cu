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
// Mapping: one warp per ray.  Every lane repeats the few scalar operations of the index / camera math (no
// divergence, no broadcast), lanes 0-2 write the per-ray vectors, and the warp copies the pixel's B-band spectrum
// as one contiguous row (128 bands fp32 = 512 B = one float4 per lane).  HBM-bound: a random 4B-byte row read and a
// 4B-byte row written per ray; algorithmic bytes per ray = 12 (draws) + 24 (indices) + 32 (origin, direction, area,
// norm) + sizeof(hs) B + 4 B (+ image in / out).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace umhs {

constexpr int kBatchWarps = 8;
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

template <int DT>
__global__ void __launch_bounds__(kBatchThreads, 6)
sample_batch_kernel(umhs_batch_in in, umhs_batch_out out, int vec) {
    using T = typename HsElem<DT>::type;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int B = in.num_bands, C = in.image_channels;
    const T *hs = reinterpret_cast<const T *>(in.hs);
    for (int64_t r = (int64_t)blockIdx.x * kBatchWarps + warp; r < in.n_rays; r += (int64_t)gridDim.x * kBatchWarps) {
        const int c = draw_index(__ldg(in.u + 3 * r + 0), in.num_images);
        const int y = draw_index(__ldg(in.u + 3 * r + 1), in.height);
        const int x = draw_index(__ldg(in.u + 3 * r + 2), in.width);
        const int64_t cam = in.image_idx ? __ldg(in.image_idx + c) : (int64_t)c;
        const int64_t pix = ((int64_t)c * in.height + y) * in.width + x;
        // ---- spectrum of the pixel: one contiguous row
        if (hs != nullptr && out.hs_image != nullptr) {
            const T *src = hs + pix * B;
            float *dst = out.hs_image + r * B;
            if (vec) {
                for (int j = lane; 4 * j < B; j += 32) reinterpret_cast<float4 *>(dst)[j] = load4<DT>(src, j);
            } else {
                for (int b = lane; b < B; b += 32) dst[b] = HsElem<DT>::get(src, b);
            }
        }
        if (in.image != nullptr && out.image != nullptr && lane < C) {
            out.image[r * C + lane] = in.image_dtype == UMHS_DTYPE_U8
                ? (float)__ldg(reinterpret_cast<const uint8_t *>(in.image) + pix * C + lane) / 255.0f
                : __ldg(reinterpret_cast<const float *>(in.image) + pix * C + lane);
        }
        if (out.indices != nullptr && lane < 3) out.indices[3 * r + lane] = lane == 0 ? cam : (lane == 1 ? y : x);
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
            if (lane < 3) {
                if (out.origins) out.origins[3 * r + lane] = lane == 0 ? m[3] : (lane == 1 ? m[7] : m[11]);
                if (out.directions) out.directions[3 * r + lane] = lane == 0 ? d[0][0] : (lane == 1 ? d[0][1] : d[0][2]);
            }
            if (lane == 0) {
                if (out.pixel_area) out.pixel_area[r] = sqrtf(dx) * sqrtf(dy);
                if (out.directions_norm) out.directions_norm[r] = nrm0;
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
    const int64_t want = (in->n_rays + kBatchWarps - 1) / kBatchWarps;
    const int64_t cap = (int64_t)kNumSMs * 6;  // 40 registers: 6 CTAs of 8 warps per SM
    const int grid = (int)(want < cap ? want : cap);
    const uintptr_t a_in = reinterpret_cast<uintptr_t>(in->hs), a_out = reinterpret_cast<uintptr_t>(out->hs_image);
    const bool b4 = in->num_bands % 4 == 0;
    cudaStream_t st = as_stream(stream);
    switch (in->hs_dtype) {
        case UMHS_DTYPE_F32:
            sample_batch_kernel<UMHS_DTYPE_F32><<<grid, kBatchThreads, 0, st>>>(
                *in, *out, b4 && a_in % 16 == 0 && a_out % 16 == 0);
            break;
        case UMHS_DTYPE_F16:
            sample_batch_kernel<UMHS_DTYPE_F16><<<grid, kBatchThreads, 0, st>>>(
                *in, *out, b4 && a_in % 8 == 0 && a_out % 16 == 0);
            break;
        case UMHS_DTYPE_BF16:
            sample_batch_kernel<UMHS_DTYPE_BF16><<<grid, kBatchThreads, 0, st>>>(
                *in, *out, b4 && a_in % 8 == 0 && a_out % 16 == 0);
            break;
        default:
            UMHS_CHECK_ARG(false, "hs_dtype must be UMHS_DTYPE_F32 / F16 / BF16");
    }
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}
