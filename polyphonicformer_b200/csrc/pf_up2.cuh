// F.interpolate(scale_factor=2, mode='bilinear', align_corners=False) of one plane, shared by pf_upsample2x (which
// materialises the map, polyphonic/kernel_update.py:133-143) and pf_panoptic (which samples it on the fly): the same
// helper with an explicit rounding order, so that both paths produce bit-identical values.
#pragma once
#include <cuda_runtime.h>

namespace pf {

// wa * a + wb * b with ONE fixed evaluation order: round(wb * b) first, then a fused multiply-add
__device__ __forceinline__ float up2_mix(float wa, float a, float wb, float b) { return __fmaf_rn(wa, a, __fmul_rn(wb, b)); }

// output column X of the horizontally up-sampled input row `row` [W]: src = (X + 0.5) / 2 - 0.5, clamped at 0
__device__ __forceinline__ float up2_hval(const float* __restrict__ row, int X, int W) {
    const int x = X >> 1;
    if (X & 1) return up2_mix(0.75f, __ldg(row + x), 0.25f, __ldg(row + min(x + 1, W - 1)));
    return x == 0 ? __ldg(row) : up2_mix(0.25f, __ldg(row + x - 1), 0.75f, __ldg(row + x));
}

// value of the x2 up-sampled plane p [H][W] at output pixel (Y, X), 0 <= Y < 2H, 0 <= X < 2W
__device__ __forceinline__ float up2_val(const float* __restrict__ p, int H, int W, int Y, int X) {
    const int y = Y >> 1;
    if (Y & 1)
        return up2_mix(0.75f, up2_hval(p + (size_t)y * W, X, W), 0.25f, up2_hval(p + (size_t)min(y + 1, H - 1) * W, X, W));
    return y == 0 ? up2_hval(p, X, W)
                  : up2_mix(0.25f, up2_hval(p + (size_t)(y - 1) * W, X, W), 0.75f, up2_hval(p + (size_t)y * W, X, W));
}

}  // namespace pf
