// The remaining custom ops of mega_core/csrc that the north star names, forward (inference) side:
//   * sigmoid focal loss forward/backward   (csrc/cuda/SigmoidFocalLoss_cuda.cu:20-103)
//   * deformable / modulated deformable convolution: bilinear im2col
//     (csrc/cuda/deform_conv_kernel_cuda.cu:92-120, :197-250, :475-504, :578-640); the contraction that
//     follows (`addmm_` per group in deform_conv_cuda.cu:228-243, :545-567) runs on the tcgen05 GEMM,
//     so the column matrix is written K-major per output pixel: cols[b][h*Wo+w][c*kh*kw + i*kw + j];
//   * deformable position-sensitive ROI pooling forward (csrc/cuda/deform_pool_kernel_cuda.cu:31-142).
// None of them is reachable from the VID configs (STAGE_WITH_DCN all False, defaults.py:287; focal loss is
// RetinaNet-only); they complete the `_C` operator surface.
#include <float.h>
#include "common.cuh"
#include "mega_b200.h"

namespace mega {

__global__ void focal_loss_fwd_kernel(long long total, const float* __restrict__ logits, const int* __restrict__ targets,
                                      int num_classes, float gamma, float alpha, float* __restrict__ losses) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long n = i / num_classes;
    const int d = static_cast<int>(i - n * num_classes);
    const int t = targets[n];
    const float c1 = (t == (d + 1)) ? 1.f : 0.f;
    const float c2 = (t >= 0 && t != (d + 1)) ? 1.f : 0.f;
    const float x = logits[i];
    const float p = 1.f / (1.f + expf(-x));
    const float term1 = powf(1.f - p, gamma) * logf(fmaxf(p, FLT_MIN));
    const float pos = (x >= 0.f) ? 1.f : 0.f;
    const float term2 = powf(p, gamma) * (-1.f * x * pos - logf(1.f + expf(x - 2.f * x * pos)));
    float l = 0.f;
    l += -c1 * term1 * alpha;
    l += -c2 * term2 * (1.f - alpha);
    losses[i] = l;
  }
}

__global__ void focal_loss_bwd_kernel(long long total, const float* __restrict__ logits, const int* __restrict__ targets,
                                      const float* __restrict__ d_losses, int num_classes, float gamma, float alpha,
                                      float* __restrict__ d_logits) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long n = i / num_classes;
    const int d = static_cast<int>(i - n * num_classes);
    const int t = targets[n];
    const float c1 = (t == (d + 1)) ? 1.f : 0.f;
    const float c2 = (t >= 0 && t != (d + 1)) ? 1.f : 0.f;
    const float x = logits[i];
    const float p = 1.f / (1.f + expf(-x));
    const float term1 = powf(1.f - p, gamma) * (1.f - p - (p * gamma * logf(fmaxf(p, FLT_MIN))));
    const float pos = (x >= 0.f) ? 1.f : 0.f;
    const float term2 =
        powf(p, gamma) * ((-1.f * x * pos - logf(1.f + expf(x - 2.f * x * pos))) * (1.f - p) * gamma - p);
    float g = 0.f;
    g += -c1 * term1 * alpha;
    g += -c2 * term2 * (1.f - alpha);
    d_logits[i] = g * d_losses[i];
  }
}

__device__ __forceinline__ float dcn_bilinear(const float* __restrict__ plane, int height, int width, float h, float w) {
  const int h_low = static_cast<int>(floorf(h)), w_low = static_cast<int>(floorf(w));
  const int h_high = h_low + 1, w_high = w_low + 1;
  const float lh = h - h_low, lw = w - w_low, hh = 1.f - lh, hw = 1.f - lw;
  float v1 = 0.f, v2 = 0.f, v3 = 0.f, v4 = 0.f;
  if (h_low >= 0 && w_low >= 0) v1 = plane[h_low * width + w_low];
  if (h_low >= 0 && w_high <= width - 1) v2 = plane[h_low * width + w_high];
  if (h_high <= height - 1 && w_low >= 0) v3 = plane[h_high * width + w_low];
  if (h_high <= height - 1 && w_high <= width - 1) v4 = plane[h_high * width + w_high];
  const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
  return w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4;
}

// one thread per (b, c, h_col, w_col); writes the kh*kw taps of that channel into the pixel's K-major row
__global__ void deform_im2col_kernel(long long total, const float* __restrict__ im, const float* __restrict__ offset,
                                     const float* __restrict__ mask, int batch, int channels, int height, int width,
                                     int kh, int kw, int pad_h, int pad_w, int stride_h, int stride_w, int dil_h,
                                     int dil_w, int deformable_group, int ho, int wo, int kpad,
                                     float* __restrict__ cols) {
  const int cpg = channels / deformable_group;
  for (long long index = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; index < total;
       index += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(index % channels);                // channel fastest: contiguous K-major writes
    const int w_col = static_cast<int>((index / channels) % wo);
    const int h_col = static_cast<int>((index / channels / wo) % ho);
    const int b = static_cast<int>(index / channels / wo / ho);
    const int dg = c / cpg;
    const int h_in = h_col * stride_h - pad_h, w_in = w_col * stride_w - pad_w;
    const float* plane = im + (static_cast<long long>(b) * channels + c) * height * width;
    const float* off = offset + (static_cast<long long>(b) * deformable_group + dg) * 2 * kh * kw * ho * wo;
    const float* msk = mask ? mask + (static_cast<long long>(b) * deformable_group + dg) * kh * kw * ho * wo : nullptr;
    float* row = cols + ((static_cast<long long>(b) * ho + h_col) * wo + w_col) * kpad + c * kh * kw;
    for (int i = 0; i < kh; ++i) {
      for (int j = 0; j < kw; ++j) {
        const int tap = i * kw + j;
        const float offset_h = off[((2 * tap) * ho + h_col) * wo + w_col];
        const float offset_w = off[((2 * tap + 1) * ho + h_col) * wo + w_col];
        const float h_im = h_in + i * dil_h + offset_h;
        const float w_im = w_in + j * dil_w + offset_w;
        float val = 0.f;
        if (h_im > -1 && w_im > -1 && h_im < height && w_im < width) val = dcn_bilinear(plane, height, width, h_im, w_im);
        if (msk) val *= msk[(tap * ho + h_col) * wo + w_col];
        row[tap] = val;
      }
    }
  }
}

__device__ __forceinline__ float psroi_bilinear(const float* __restrict__ data, float x, float y, int width) {
  const int x1 = static_cast<int>(floorf(x)), x2 = static_cast<int>(ceilf(x));
  const int y1 = static_cast<int>(floorf(y)), y2 = static_cast<int>(ceilf(y));
  const float dist_x = x - x1, dist_y = y - y1;
  const float v11 = data[y1 * width + x1], v12 = data[y2 * width + x1];
  const float v21 = data[y1 * width + x2], v22 = data[y2 * width + x2];
  return (1 - dist_x) * (1 - dist_y) * v11 + (1 - dist_x) * dist_y * v12 + dist_x * (1 - dist_y) * v21 +
         dist_x * dist_y * v22;
}

__global__ void deform_psroi_fwd_kernel(long long count, const float* __restrict__ bottom, float spatial_scale,
                                        int channels, int height, int width, int pooled_h, int pooled_w,
                                        const float* __restrict__ rois, const float* __restrict__ trans, int no_trans,
                                        float trans_std, int sample_per_part, int output_dim, int group_size,
                                        int part_size, int num_classes, int channels_each_class,
                                        float* __restrict__ top, float* __restrict__ top_count) {
  for (long long index = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; index < count;
       index += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int pw = static_cast<int>(index % pooled_w);
    const int ph = static_cast<int>((index / pooled_w) % pooled_h);
    const int ctop = static_cast<int>((index / pooled_w / pooled_h) % output_dim);
    const long long n = index / pooled_w / pooled_h / output_dim;
    const float* r = rois + n * 5;
    const int roi_batch_ind = static_cast<int>(r[0]);
    const float roi_start_w = roundf(r[1]) * spatial_scale - 0.5f;
    const float roi_start_h = roundf(r[2]) * spatial_scale - 0.5f;
    const float roi_end_w = (roundf(r[3]) + 1.f) * spatial_scale - 0.5f;
    const float roi_end_h = (roundf(r[4]) + 1.f) * spatial_scale - 0.5f;
    const float roi_width = fmaxf(roi_end_w - roi_start_w, 0.1f);
    const float roi_height = fmaxf(roi_end_h - roi_start_h, 0.1f);
    const float bin_size_h = roi_height / static_cast<float>(pooled_h);
    const float bin_size_w = roi_width / static_cast<float>(pooled_w);
    const float sub_bin_size_h = bin_size_h / static_cast<float>(sample_per_part);
    const float sub_bin_size_w = bin_size_w / static_cast<float>(sample_per_part);
    const int part_h = static_cast<int>(floorf(static_cast<float>(ph) / pooled_h * part_size));
    const int part_w = static_cast<int>(floorf(static_cast<float>(pw) / pooled_w * part_size));
    const int class_id = ctop / channels_each_class;
    const float trans_x =
        no_trans ? 0.f : trans[(((n * num_classes + class_id) * 2) * part_size + part_h) * part_size + part_w] * trans_std;
    const float trans_y =
        no_trans ? 0.f
                 : trans[(((n * num_classes + class_id) * 2 + 1) * part_size + part_h) * part_size + part_w] * trans_std;
    float wstart = static_cast<float>(pw) * bin_size_w + roi_start_w;
    wstart += trans_x * roi_width;
    float hstart = static_cast<float>(ph) * bin_size_h + roi_start_h;
    hstart += trans_y * roi_height;
    float sum = 0.f;
    int cnt = 0;
    int gw = static_cast<int>(floorf(static_cast<float>(pw) * group_size / pooled_w));
    int gh = static_cast<int>(floorf(static_cast<float>(ph) * group_size / pooled_h));
    gw = min(max(gw, 0), group_size - 1);
    gh = min(max(gh, 0), group_size - 1);
    const float* data = bottom + (static_cast<long long>(roi_batch_ind) * channels) * height * width;
    for (int ih = 0; ih < sample_per_part; ++ih) {
      for (int iw = 0; iw < sample_per_part; ++iw) {
        float w = wstart + iw * sub_bin_size_w;
        float h = hstart + ih * sub_bin_size_h;
        if (w < -0.5f || w > width - 0.5f || h < -0.5f || h > height - 0.5f) continue;
        w = fminf(fmaxf(w, 0.f), width - 1.f);
        h = fminf(fmaxf(h, 0.f), height - 1.f);
        const int c = (ctop * group_size + gh) * group_size + gw;
        sum += psroi_bilinear(data + static_cast<long long>(c) * height * width, w, h, width);
        cnt++;
      }
    }
    top[index] = cnt == 0 ? 0.f : sum / cnt;
    top_count[index] = static_cast<float>(cnt);
  }
}

static int grid_dcn(long long total, int block) {
  long long b = (total + block - 1) / block;
  const long long cap = 148LL * 16;
  return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace mega

using namespace mega;

extern "C" int mega_sigmoid_focalloss_forward(const float* logits, const int* targets, int num_samples,
                                              int num_classes, float gamma, float alpha, float* losses,
                                              void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const long long total = static_cast<long long>(num_samples) * num_classes;
  if (total == 0) return MEGA_OK;
  focal_loss_fwd_kernel<<<grid_dcn(total, 256), 256, 0, stream>>>(total, logits, targets, num_classes, gamma, alpha,
                                                                  losses);
  MEGA_CUDA_CHECK(cudaGetLastError());
  return MEGA_OK;
}

extern "C" int mega_sigmoid_focalloss_backward(const float* logits, const int* targets, const float* d_losses,
                                               int num_samples, int num_classes, float gamma, float alpha,
                                               float* d_logits, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const long long total = static_cast<long long>(num_samples) * num_classes;
  if (total == 0) return MEGA_OK;
  focal_loss_bwd_kernel<<<grid_dcn(total, 256), 256, 0, stream>>>(total, logits, targets, d_losses, num_classes, gamma,
                                                                  alpha, d_logits);
  MEGA_CUDA_CHECK(cudaGetLastError());
  return MEGA_OK;
}

extern "C" int mega_deform_im2col(const float* input, const float* offset, const float* mask, int batch, int channels,
                                  int height, int width, int kh, int kw, int pad_h, int pad_w, int stride_h,
                                  int stride_w, int dil_h, int dil_w, int deformable_group, int kpad, float* cols,
                                  void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  MEGA_ARG_CHECK(deformable_group >= 1 && channels % deformable_group == 0,
                 "deform_im2col: channels must divide into deformable groups");
  MEGA_ARG_CHECK(kpad >= channels * kh * kw, "deform_im2col: kpad smaller than C*kh*kw");
  const int ho = (height + 2 * pad_h - (dil_h * (kh - 1) + 1)) / stride_h + 1;
  const int wo = (width + 2 * pad_w - (dil_w * (kw - 1) + 1)) / stride_w + 1;
  const long long total = static_cast<long long>(batch) * channels * ho * wo;
  if (total == 0) return MEGA_OK;
  deform_im2col_kernel<<<grid_dcn(total, 256), 256, 0, stream>>>(total, input, offset, mask, batch, channels, height,
                                                                 width, kh, kw, pad_h, pad_w, stride_h, stride_w, dil_h,
                                                                 dil_w, deformable_group, ho, wo, kpad, cols);
  MEGA_CUDA_CHECK(cudaGetLastError());
  return MEGA_OK;
}

extern "C" int mega_deform_psroi_pooling_forward(const float* input, const float* rois, const float* trans, int num_rois,
                                                 int channels, int height, int width, int no_trans, float spatial_scale,
                                                 int output_dim, int group_size, int pooled_size, int part_size,
                                                 int sample_per_part, float trans_std, int num_classes, float* out,
                                                 float* top_count, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const long long count = static_cast<long long>(num_rois) * output_dim * pooled_size * pooled_size;
  if (count == 0) return MEGA_OK;
  MEGA_ARG_CHECK(num_classes >= 1, "deform_psroi_pooling: num_classes must be >= 1");
  const int channels_each_class = no_trans ? output_dim : output_dim / num_classes;
  deform_psroi_fwd_kernel<<<grid_dcn(count, 256), 256, 0, stream>>>(
      count, input, spatial_scale, channels, height, width, pooled_size, pooled_size, rois, trans, no_trans, trans_std,
      sample_per_part, output_dim, group_size, part_size, num_classes, channels_each_class, out, top_count);
  MEGA_CUDA_CHECK(cudaGetLastError());
  return MEGA_OK;
}
