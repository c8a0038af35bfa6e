// Host build of csrc/image_ops.cuh -- TEST INFRASTRUCTURE ONLY (see train_ops_host.cpp): the same per-pixel body the
// CUDA kernel runs, under the C-ABI prototype of include/mega_b200.h, with host pointers; `stream` is ignored.
// Build: g++ -O2 -fPIC -shared -std=c++17 -ffp-contract=off -I mega.pytorch_b200/csrc -I include -o libimage_ops_host.so image_ops_host.cpp
#include "mega_b200.h"
#include "image_ops.cuh"

extern "C" int mega_image_transform_u8(const unsigned char* src, int src_h, int src_w, long long src_row_stride,
                                       long long src_pix_stride, long long src_ch_stride, const int* bounds_h, const int* kk_h, int ksize_h, const int* bounds_v,
                                       const int* kk_v, int ksize_v, int out_h, int out_w, const float* mean_host,
                                       const float* std_host, int to_bgr255, float* out, void* stream) {
  (void)stream;
  mega_image::ResizeGeom g;
  g.src_h = src_h, g.src_w = src_w, g.out_h = out_h, g.out_w = out_w, g.src_row_stride = src_row_stride;
  g.src_pix_stride = src_pix_stride, g.src_ch_stride = src_ch_stride;
  g.ksize_h = ksize_h, g.ksize_v = ksize_v;
  g.bounds_h = bounds_h, g.kk_h = kk_h, g.bounds_v = bounds_v, g.kk_v = kk_v;
  for (int c = 0; c < 3; ++c) g.mean[c] = mean_host[c], g.stdv[c] = std_host[c];
  g.to_bgr255 = to_bgr255 ? 1 : 0;
  const long long total = static_cast<long long>(out_h) * out_w;
  for (long long i = 0; i < total; ++i) mega_image::image_transform_item(i, g, src, out);
  return 0;
}
