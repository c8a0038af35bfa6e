// IoU threshold decisions shared by the proposal (proposals.cu) and detection (postprocess.cu) NMS kernels.
// Reference arithmetic: nms.cu:16-19 / nms_cpu.cpp:6-75 ("+1" pixel convention, fp32, division then compare).
#pragma once
#include <cuda_runtime.h>

namespace mega {

__device__ __forceinline__ float box_area_plus1(const float4 a) {
  return __fmul_rn(__fadd_rn(__fsub_rn(a.z, a.x), 1.f), __fadd_rn(__fsub_rn(a.w, a.y), 1.f));
}

// iou_plus1(a, b) > thresh with the SAME outcome for every input, but without the division unless the quotient is
// within 2^-20 (relative) of the threshold: RN(inter / u) > t is decided by inter vs t*u whenever the true quotient is
// more than a few ulps away from t (t_lo = t*(1-2^-20), t_hi = t*(1+2^-20); the two products carry 2^-24 relative
// error each, far inside the band). sa / sb: box_area_plus1 of a / b.
__device__ __forceinline__ bool iou_plus1_gt(const float4 a, const float sa, const float4 b, const float sb,
                                             const float thresh, const float t_lo, const float t_hi) {
  const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
  const float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
  const float width = fmaxf(__fadd_rn(__fsub_rn(right, left), 1.f), 0.f);
  const float height = fmaxf(__fadd_rn(__fsub_rn(bottom, top), 1.f), 0.f);
  const float inter = __fmul_rn(width, height);
  const float u = __fsub_rn(__fadd_rn(sa, sb), inter);
  if (u > 0.f) {
    if (inter > __fmul_rn(t_hi, u)) return true;
    if (inter < __fmul_rn(t_lo, u)) return false;
  }
  return __fdiv_rn(inter, u) > thresh;
}

}  // namespace mega
