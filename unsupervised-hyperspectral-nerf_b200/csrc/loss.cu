// loss.cu -- the training losses of ref umhsnerf/umhs_model.py:329-383 and their gradients in ONE pass.
//
//   spectral_loss = w_spec * mean((spectral - hs_image)^2)                      (ref :369; w_spec = 5, quirk 4)
//   rgb_loss      = w_rgb  * mean((rgb [+ bg (1 - acc)] - blend(image, bg))^2)  (ref :358-362, :370;
//                   nerfstudio RGBRenderer.blend_background_for_loss_computation: the prediction is blended over
//                   the background only when that background is the per-step RANDOM draw (blend_pred = 1); a named
//                   colour (white / black) only blends a 4-channel ground-truth image (blend_pred = 0))
// The reference evaluates this with ~15 elementwise / reduction launches forward and backward.  The gradients of a
// scalar loss do not depend on anything but these same inputs, so the kernel writes d loss / d spectral,
// d loss / d rgb and d loss / d accumulation right away; the host feeds them to the compositing backward.
// HBM streaming: 2 B-wide rows read, 1 written per ray.
#include "common.cuh"

namespace umhs {

constexpr int kLossThreads = 256;

__global__ void __launch_bounds__(kLossThreads)
train_loss_kernel(int64_t R, int B, int image_channels, const float *__restrict__ spectral,
                  const float *__restrict__ hs_image, const float *__restrict__ rgb, const float *__restrict__ acc,
                  const float *__restrict__ image, const float *__restrict__ bg, int blend_pred, float w_spec, float w_rgb,
                  float *__restrict__ losses, float *__restrict__ d_spectral, float *__restrict__ d_rgb,
                  float *__restrict__ d_acc) {
    __shared__ float red[2][kLossThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float s_spec = 0.f, s_rgb = 0.f;
    // spectral term: flat grid-stride over R * B (float4 when B % 4 == 0)
    const int64_t n = R * B;
    const float gs = 2.0f * w_spec / (float)n;
    if ((B & 3) == 0) {
        const float4 *a4 = reinterpret_cast<const float4 *>(spectral), *b4 = reinterpret_cast<const float4 *>(hs_image);
        float4 *g4 = reinterpret_cast<float4 *>(d_spectral);
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n / 4; i += (int64_t)gridDim.x * blockDim.x) {
            const float4 a = a4[i], b = __ldg(b4 + i);
            const float4 d = make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
            s_spec += d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
            g4[i] = make_float4(gs * d.x, gs * d.y, gs * d.z, gs * d.w);
        }
    } else {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
            const float d = spectral[i] - __ldg(hs_image + i);
            s_spec += d * d;
            d_spectral[i] = gs * d;
        }
    }
    // rgb term: one thread per ray
    if (w_rgb != 0.0f && rgb != nullptr) {
        const float gr = 2.0f * w_rgb / (float)(3 * R);
        for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < R; r += (int64_t)gridDim.x * blockDim.x) {
            const float a = acc[r];
            float da = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float b = bg ? __ldg(bg + 3 * r + c) : 0.f;
                const float bp = blend_pred ? b : 0.f;  // background under the prediction: random draws only
                const float pred = rgb[3 * r + c] + bp * (1.0f - a);
                float gt = __ldg(image + image_channels * r + c);
                if (image_channels == 4) {
                    const float al = __ldg(image + 4 * r + 3);
                    gt = gt * al + b * (1.0f - al);
                }
                const float d = pred - gt;
                s_rgb += d * d;
                d_rgb[3 * r + c] = gr * d;
                da -= gr * d * bp;
            }
            d_acc[r] = da;
        }
    }
    s_spec = warp_sum(s_spec);
    s_rgb = warp_sum(s_rgb);
    if (lane == 0) {
        red[0][warp] = s_spec;
        red[1][warp] = s_rgb;
    }
    __syncthreads();
    if (warp == 0) {
        float a = lane < kLossThreads / 32 ? red[0][lane] : 0.f, b = lane < kLossThreads / 32 ? red[1][lane] : 0.f;
        a = warp_sum(a);
        b = warp_sum(b);
        if (lane == 0) {
            atomicAdd(losses + 0, a * w_spec / (float)n);
            if (w_rgb != 0.0f) atomicAdd(losses + 1, b * w_rgb / (float)(3 * R));
        }
    }
}

}  // namespace umhs

using namespace umhs;

extern "C" int umhs_train_loss(int64_t n_rays, int32_t num_bands, int32_t image_channels, const float *spectral,
                               const float *hs_image, const float *rgb, const float *accumulation,
                               const float *image, const float *background, int32_t blend_pred, float w_spectral,
                               float w_rgb, float *losses, float *d_spectral, float *d_rgb, float *d_accumulation, void *stream) {
    UMHS_CHECK_ARG(spectral && hs_image && losses && d_spectral, "null argument");
    UMHS_CHECK_ARG(w_rgb == 0.0f || (rgb && accumulation && image && d_rgb && d_accumulation), "rgb term needs its buffers");
    UMHS_CHECK_ARG(image_channels == 3 || image_channels == 4, "image must have 3 or 4 channels");
    UMHS_CHECK_ARG(num_bands >= 1, "bad band count");
    UMHS_CHECK_CUDA(cudaMemsetAsync(losses, 0, 2 * sizeof(float), as_stream(stream)));
    if (n_rays <= 0) return UMHS_OK;
    const int64_t want = (n_rays * num_bands / 4 + kLossThreads - 1) / kLossThreads;
    const int grid = (int)(want < kNumSMs * 4 ? (want < 1 ? 1 : want) : kNumSMs * 4);
    train_loss_kernel<<<grid, kLossThreads, 0, as_stream(stream)>>>(n_rays, num_bands, image_channels, spectral, hs_image,
                                                                  rgb, accumulation, image, background, blend_pred,
                                                                  w_spectral, w_rgb, losses, d_spectral, d_rgb, d_accumulation);
    UMHS_CHECK_LAUNCH();
    return UMHS_OK;
}
