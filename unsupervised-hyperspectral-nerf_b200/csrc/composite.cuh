// composite.cuh -- pieces shared by the two compositing kernel families (composite.cu: one warp per ray, used when
// the specular head is on; composite_group.cu: one lane group per ray, the no-specular training / sweep path).
#pragma once
#include "common.cuh"

namespace umhs {

// sRGB transfer function of ref umhsnerf/utils/spec_to_rgb.py:103-110 and its derivative
__device__ __forceinline__ float gamma_fwd(float x) {
    return x < 0.0031308f ? 12.92f * x : 1.055f * powf(fmaxf(x, 1e-6f), 1.0f / 2.4f) - 0.055f;
}
__device__ __forceinline__ float gamma_grad(float x) {
    return x < 0.0031308f ? 12.92f : (1.055f / 2.4f) * powf(fmaxf(x, 1e-6f), 1.0f / 2.4f - 1.0f);
}

// lane-group form (composite_group.cu); both return UMHS_OK or an error code like the entry points
int composite_group_fwd(const umhs_composite_in *in, const umhs_composite_out *out, cudaStream_t st);
int composite_group_bwd(const umhs_composite_in *in, const umhs_composite_grads_in *gin,
                        const umhs_composite_grads_out *gout, cudaStream_t st);

}  // namespace umhs
