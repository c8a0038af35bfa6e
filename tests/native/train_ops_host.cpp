// Host build of the per-item bodies of csrc/train_ops.cuh -- TEST INFRASTRUCTURE ONLY.
// g++ compiles the very functions the CUDA kernels of csrc/train_ops.cu loop over (the header is plain
// __host__ __device__ code) and runs every item serially, so the CPU test-suite can check the index arithmetic and the
// gradient formulas of the training-side ops against oracle/train_ops_oracle.py without a GPU.
// The entry points carry the SAME NAMES AND PROTOTYPES as the ABI v4 functions of include/mega_b200.h (the header is
// included, so a drifting signature does not compile); pointers are host pointers and `stream` is ignored. The CPU
// tests patch these over the ctypes handles of the real library to run mega_core/_C.py end to end on CPU tensors.
// Build: g++ -O2 -fPIC -shared -std=c++17 -I mega.pytorch_b200/csrc -I include -o libtrain_ops_host.so train_ops_host.cpp
#include "mega_b200.h"   // the prototypes below must match the C ABI: the compiler checks them
#include "train_ops.cuh"

using namespace mega_train;

struct HostAdd {
  void operator()(float* p, float v) const { *p += v; }
};

extern "C" {

int mega_roi_align_backward_nchw(const float* grad, const float* rois, int num_rois, float spatial_scale, int pooled_h,
                                 int pooled_w, int batch, int channels, int height, int width, int sampling_ratio,
                                 float* grad_input, void* stream) {
  (void)batch, (void)stream;
  const long long items = roi_align_bwd_items(num_rois, channels, pooled_h, pooled_w);
  for (long long i = 0; i < items; ++i)
    roi_align_bwd_item(i, grad, rois, spatial_scale, channels, height, width, pooled_h, pooled_w, sampling_ratio,
                       grad_input, HostAdd());
  return 0;
}

int mega_roi_pool_forward(const float* input, const float* rois, int num_rois, float spatial_scale, int channels,
                          int height, int width, int pooled_h, int pooled_w, float* output, int* argmax, void* stream) {
  (void)stream;
  const long long total = static_cast<long long>(num_rois) * channels * pooled_h * pooled_w;
  for (long long i = 0; i < total; ++i)
    roi_pool_fwd_item(i, input, rois, spatial_scale, channels, height, width, pooled_h, pooled_w, output, argmax);
  return 0;
}

int mega_roi_pool_backward(const float* grad, const int* argmax, const float* rois, int num_rois, int channels,
                           int height, int width, int pooled_h, int pooled_w, float* grad_input, void* stream) {
  (void)stream;
  const long long total = static_cast<long long>(num_rois) * channels * pooled_h * pooled_w;
  for (long long i = 0; i < total; ++i)
    roi_pool_bwd_item(i, grad, argmax, rois, channels, height, width, pooled_h, pooled_w, grad_input, HostAdd());
  return 0;
}

static DcnGeom make_geom(int batch, int channels, int height, int width, int kh, int kw, int pad_h, int pad_w,
                         int stride_h, int stride_w, int dil_h, int dil_w, int dg, int ldp) {
  DcnGeom g;
  g.batch = batch, g.channels = channels, g.height = height, g.width = width, g.kh = kh, g.kw = kw;
  g.pad_h = pad_h, g.pad_w = pad_w, g.stride_h = stride_h, g.stride_w = stride_w, g.dil_h = dil_h, g.dil_w = dil_w;
  g.deformable_group = dg;
  g.ho = (height + 2 * pad_h - (dil_h * (kh - 1) + 1)) / stride_h + 1;
  g.wo = (width + 2 * pad_w - (dil_w * (kw - 1) + 1)) / stride_w + 1;
  g.ldp = ldp;
  return g;
}

int mega_deform_im2col_kq(const float* input, const float* offset, const float* mask, int batch, int channels,
                          int height, int width, int kh, int kw, int pad_h, int pad_w, int stride_h, int stride_w,
                          int dil_h, int dil_w, int deformable_group, int ldp, float* cols, void* stream) {
  (void)stream;
  const DcnGeom g = make_geom(batch, channels, height, width, kh, kw, pad_h, pad_w, stride_h, stride_w, dil_h, dil_w,
                              deformable_group, ldp);
  const long long total = static_cast<long long>(channels) * batch * g.ho * g.wo;
  for (long long i = 0; i < total; ++i) dcn_im2col_kq_item(i, g, input, offset, mask, cols);
  return 0;
}

int mega_deform_col2im_fused(const float* gcols, const float* input, const float* offset, const float* mask, int batch,
                             int channels, int height, int width, int kh, int kw, int pad_h, int pad_w, int stride_h,
                             int stride_w, int dil_h, int dil_w, int deformable_group, int ldp, float* grad_input,
                             float* grad_offset, float* grad_mask, void* stream) {
  (void)stream;
  const DcnGeom g = make_geom(batch, channels, height, width, kh, kw, pad_h, pad_w, stride_h, stride_w, dil_h, dil_w,
                              deformable_group, ldp);
  const long long total = static_cast<long long>(batch) * deformable_group * kh * kw * g.ho * g.wo;
  for (long long i = 0; i < total; ++i)
    dcn_col2im_fused_item(i, g, gcols, input, offset, mask, grad_input, grad_offset, grad_mask, HostAdd());
  return 0;
}

int mega_channel_sum_nchw(const float* x, int batch, int channels, int plane, float* out, void* stream) {
  (void)stream;
  for (int c = 0; c < channels; ++c) out[c] += dcn_channel_sum_item(c, x, batch, channels, plane);
  return 0;
}

int mega_deform_psroi_pooling_backward(const float* out_grad, const float* input, const float* rois, const float* trans,
                                       const float* top_count, int num_rois, int channels, int height, int width,
                                       int no_trans, float spatial_scale, int output_dim, int group_size,
                                       int pooled_size, int part_size, int sample_per_part, float trans_std,
                                       int num_classes, float* input_grad, float* trans_grad, void* stream) {
  (void)stream;
  PsRoiGeom g;
  g.channels = channels, g.height = height, g.width = width, g.pooled = pooled_size, g.output_dim = output_dim;
  g.group_size = group_size, g.part_size = part_size, g.sample_per_part = sample_per_part;
  g.num_classes = num_classes, g.no_trans = no_trans ? 1 : 0;
  g.channels_each_class = no_trans ? output_dim : output_dim / num_classes;
  g.spatial_scale = spatial_scale, g.trans_std = trans_std;
  const long long total = static_cast<long long>(num_rois) * output_dim * pooled_size * pooled_size;
  for (long long i = 0; i < total; ++i)
    deform_psroi_bwd_item(i, g, out_grad, top_count, input, rois, trans, input_grad, trans_grad, HostAdd());
  return 0;
}

}  // extern "C"
