// Per-item bodies of the training-side / non-VID ops that complete the `mega_core._C` surface (SURVEY.md section 8b,
// 8f row 3): ROIAlign backward, ROIPool forward / backward, deformable-convolution backward (v1 and modulated) and
// deformable PSROI pooling backward.
//
// Every kernel of train_ops.cu is a grid-stride loop around one of these functions, which are plain
// `__host__ __device__` code: the same bodies are compiled by g++ into tests/native/libtrain_ops_host.so and compared
// with the oracle in the CPU test-suite (tests/test_train_ops_cpu.py), so index arithmetic and gradient formulas are
// checked without a GPU; the GPU tests then only have to confirm the launch plumbing.
//
// Scatter targets are updated through the `Add` functor: `red.global.add.f32` on the device, `+=` on the host.
// Summation ORDER over colliding scatters is therefore unspecified on the device, exactly as in the reference
// (atomicAdd in ROIAlign_cuda.cu:237-240, deform_conv_kernel_cuda.cu:334, deform_pool_kernel_cuda.cu:253-276).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define MEGA_HD __host__ __device__ __forceinline__
#else
#define MEGA_HD static inline
#endif

namespace mega_train {

// ----------------------------------------------------------------------------------------------- ROIAlign backward
// Reference: RoIAlignBackwardFeature + bilinear_interpolate_gradient (csrc/cuda/ROIAlign_cuda.cu:125-246).
// Layout: grad [K, C, ph, pw], rois [K, 5] = (batch, x1, y1, x2, y2), grad_in [N, C, H, W] (zero-initialised).
//
// Work item = one bin (n, ph, pw) x a strip of kRoiBwdStrip channels: the sample grid of a bin and the four bilinear
// weights of every sample do not depend on the channel, so they are computed once per sample and applied to the strip
// (the reference recomputes them for each of the C channels). Items are numbered pw-fastest, so a warp reads
// consecutive gradient values of one channel plane and scatters into one neighbourhood of one feature plane.
constexpr int kRoiBwdStrip = 8;

MEGA_HD long long roi_align_bwd_items(int num_rois, int channels, int pooled_h, int pooled_w) {
  const long long strips = (channels + kRoiBwdStrip - 1) / kRoiBwdStrip;
  return static_cast<long long>(num_rois) * strips * pooled_h * pooled_w;
}

struct BilinearGrad {
  int y_low, y_high, x_low, x_high;   // all -1: sample outside the map, contributes nothing
  float w1, w2, w3, w4;
};

MEGA_HD BilinearGrad bilinear_grad_setup(int height, int width, float y, float x) {
  BilinearGrad b;
  if (y < -1.0f || y > static_cast<float>(height) || x < -1.0f || x > static_cast<float>(width)) {
    b.y_low = b.y_high = b.x_low = b.x_high = -1;
    b.w1 = b.w2 = b.w3 = b.w4 = 0.f;
    return b;
  }
  if (y <= 0.f) y = 0.f;
  if (x <= 0.f) x = 0.f;
  b.y_low = static_cast<int>(y);
  b.x_low = static_cast<int>(x);
  if (b.y_low >= height - 1) {
    b.y_high = b.y_low = height - 1;
    y = static_cast<float>(b.y_low);
  } else {
    b.y_high = b.y_low + 1;
  }
  if (b.x_low >= width - 1) {
    b.x_high = b.x_low = width - 1;
    x = static_cast<float>(b.x_low);
  } else {
    b.x_high = b.x_low + 1;
  }
  const float ly = y - static_cast<float>(b.y_low), lx = x - static_cast<float>(b.x_low);
  const float hy = 1.f - ly, hx = 1.f - lx;
  b.w1 = hy * hx;
  b.w2 = hy * lx;
  b.w3 = ly * hx;
  b.w4 = ly * lx;
  return b;
}

template <class Add>
MEGA_HD void roi_align_bwd_item(long long item, const float* grad, const float* rois, float spatial_scale, int channels,
                                int height, int width, int pooled_h, int pooled_w, int sampling_ratio, float* grad_in,
                                Add add) {
  const int strips = (channels + kRoiBwdStrip - 1) / kRoiBwdStrip;
  const int pw = static_cast<int>(item % pooled_w);
  const int ph = static_cast<int>((item / pooled_w) % pooled_h);
  const int strip = static_cast<int>((item / pooled_w / pooled_h) % strips);
  const long long n = item / pooled_w / pooled_h / strips;
  const float* r = rois + n * 5;
  const int roi_batch = static_cast<int>(r[0]);
  const float roi_start_w = r[1] * spatial_scale, roi_start_h = r[2] * spatial_scale;
  const float roi_end_w = r[3] * spatial_scale, roi_end_h = r[4] * spatial_scale;
  const float roi_width = fmaxf(roi_end_w - roi_start_w, 1.f);
  const float roi_height = fmaxf(roi_end_h - roi_start_h, 1.f);
  const float bin_h = roi_height / static_cast<float>(pooled_h);
  const float bin_w = roi_width / static_cast<float>(pooled_w);
  const int grid_h = sampling_ratio > 0 ? sampling_ratio : static_cast<int>(ceilf(roi_height / pooled_h));
  const int grid_w = sampling_ratio > 0 ? sampling_ratio : static_cast<int>(ceilf(roi_width / pooled_w));
  const float count = static_cast<float>(grid_h * grid_w);
  const int c0 = strip * kRoiBwdStrip;
  const int c1 = c0 + kRoiBwdStrip < channels ? c0 + kRoiBwdStrip : channels;
  const long long plane = static_cast<long long>(height) * width;
  const long long bins = static_cast<long long>(pooled_h) * pooled_w;
  const float* g = grad + (n * channels + c0) * bins + ph * pooled_w + pw;
  float* out = grad_in + (static_cast<long long>(roi_batch) * channels + c0) * plane;
  for (int iy = 0; iy < grid_h; ++iy) {
    const float y = roi_start_h + ph * bin_h + (iy + .5f) * bin_h / static_cast<float>(grid_h);
    for (int ix = 0; ix < grid_w; ++ix) {
      const float x = roi_start_w + pw * bin_w + (ix + .5f) * bin_w / static_cast<float>(grid_w);
      const BilinearGrad b = bilinear_grad_setup(height, width, y, x);
      if (b.x_low < 0 || b.x_high < 0 || b.y_low < 0 || b.y_high < 0) continue;
      const int o1 = b.y_low * width + b.x_low, o2 = b.y_low * width + b.x_high;
      const int o3 = b.y_high * width + b.x_low, o4 = b.y_high * width + b.x_high;
      for (int c = 0; c < c1 - c0; ++c) {
        const float top = g[c * bins];
        float* p = out + c * plane;
        add(p + o1, top * b.w1 / count);
        add(p + o2, top * b.w2 / count);
        add(p + o3, top * b.w3 / count);
        add(p + o4, top * b.w4 / count);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------- ROIPool
// Reference: RoIPoolFForward / RoIPoolFBackward (csrc/cuda/ROIPool_cuda.cu:16-112). One item per output element
// (n, c, ph, pw), pw fastest. argmax = offset inside the (batch, c) plane, -1 for an empty bin.
MEGA_HD void roi_pool_fwd_item(long long index, const float* input, const float* rois, float spatial_scale,
                               int channels, int height, int width, int pooled_h, int pooled_w, float* out,
                               int* argmax) {
  const int pw = static_cast<int>(index % pooled_w);
  const int ph = static_cast<int>((index / pooled_w) % pooled_h);
  const int c = static_cast<int>((index / pooled_w / pooled_h) % channels);
  const long long n = index / pooled_w / pooled_h / channels;
  const float* r = rois + n * 5;
  const int roi_batch = static_cast<int>(r[0]);
  const int roi_start_w = static_cast<int>(roundf(r[1] * spatial_scale));
  const int roi_start_h = static_cast<int>(roundf(r[2] * spatial_scale));
  const int roi_end_w = static_cast<int>(roundf(r[3] * spatial_scale));
  const int roi_end_h = static_cast<int>(roundf(r[4] * spatial_scale));
  const int roi_width = roi_end_w - roi_start_w + 1 > 1 ? roi_end_w - roi_start_w + 1 : 1;
  const int roi_height = roi_end_h - roi_start_h + 1 > 1 ? roi_end_h - roi_start_h + 1 : 1;
  const float bin_h = static_cast<float>(roi_height) / static_cast<float>(pooled_h);
  const float bin_w = static_cast<float>(roi_width) / static_cast<float>(pooled_w);
  int hstart = static_cast<int>(floorf(static_cast<float>(ph) * bin_h));
  int wstart = static_cast<int>(floorf(static_cast<float>(pw) * bin_w));
  int hend = static_cast<int>(ceilf(static_cast<float>(ph + 1) * bin_h));
  int wend = static_cast<int>(ceilf(static_cast<float>(pw + 1) * bin_w));
  hstart = hstart + roi_start_h < 0 ? 0 : (hstart + roi_start_h > height ? height : hstart + roi_start_h);
  hend = hend + roi_start_h < 0 ? 0 : (hend + roi_start_h > height ? height : hend + roi_start_h);
  wstart = wstart + roi_start_w < 0 ? 0 : (wstart + roi_start_w > width ? width : wstart + roi_start_w);
  wend = wend + roi_start_w < 0 ? 0 : (wend + roi_start_w > width ? width : wend + roi_start_w);
  const bool is_empty = (hend <= hstart) || (wend <= wstart);
  float maxval = is_empty ? 0.f : -3.402823466e+38f;
  int maxidx = -1;
  const float* plane = input + (static_cast<long long>(roi_batch) * channels + c) * height * width;
  for (int h = hstart; h < hend; ++h) {
    for (int w = wstart; w < wend; ++w) {
      const float v = plane[h * width + w];
      if (v > maxval) {
        maxval = v;
        maxidx = h * width + w;
      }
    }
  }
  out[index] = maxval;
  argmax[index] = maxidx;
}

template <class Add>
MEGA_HD void roi_pool_bwd_item(long long index, const float* grad, const int* argmax, const float* rois, int channels,
                               int height, int width, int pooled_h, int pooled_w, float* grad_in, Add add) {
  const int c = static_cast<int>((index / pooled_w / pooled_h) % channels);
  const long long n = index / pooled_w / pooled_h / channels;
  const int roi_batch = static_cast<int>(rois[n * 5]);
  const int am = argmax[index];
  if (am != -1) add(grad_in + (static_cast<long long>(roi_batch) * channels + c) * height * width + am, grad[index]);
}

// -------------------------------------------------------------------------------- deformable convolution, backward
// Column matrices use the reference's layout (deform_conv_cuda.cu:300-302): cols[k][q], k = c*kh*kw + i*kw + j,
// q = b*ldp + (h_col*Wo + w_col), with ldp >= Ho*Wo a multiple of 4 (the padding columns are never read here and
// must be zero in a GEMM operand) and ldq = batch*ldp floats per row.
struct DcnGeom {
  int batch, channels, height, width;       // input [B, C, H, W]
  int kh, kw, pad_h, pad_w, stride_h, stride_w, dil_h, dil_w;
  int deformable_group, ho, wo;
  int ldp;                                  // padded Ho*Wo
};

struct DcnSample {
  bool inside;
  int h_low, w_low;
  float lh, lw;      // fractional parts
  bool v1, v2, v3, v4;   // corner (low,low) (low,high) (high,low) (high,high) lies inside the map
};

MEGA_HD DcnSample dcn_sample(const DcnGeom& g, float h, float w) {
  DcnSample s;
  s.inside = (h > -1.f && w > -1.f && h < static_cast<float>(g.height) && w < static_cast<float>(g.width));
  const float hf = floorf(h), wf = floorf(w);
  s.h_low = static_cast<int>(hf);
  s.w_low = static_cast<int>(wf);
  s.lh = h - hf;
  s.lw = w - wf;
  s.v1 = s.inside && s.h_low >= 0 && s.w_low >= 0;
  s.v2 = s.inside && s.h_low >= 0 && s.w_low + 1 <= g.width - 1;
  s.v3 = s.inside && s.h_low + 1 <= g.height - 1 && s.w_low >= 0;
  s.v4 = s.inside && s.h_low + 1 <= g.height - 1 && s.w_low + 1 <= g.width - 1;
  return s;
}

// deformable im2col into the [K, B*ldp] layout: one item per (c, b, p), p fastest (coalesced stores; neighbouring
// output pixels sample neighbouring input cells). Reference: deformable_im2col_gpu_kernel /
// modulated_deformable_im2col_gpu_kernel (deform_conv_kernel_cuda.cu:197-250, :578-640). mask == nullptr: v1.
MEGA_HD void dcn_im2col_kq_item(long long index, const DcnGeom& g, const float* im, const float* offset,
                                const float* mask, float* cols) {
  const int p_total = g.ho * g.wo;
  const int p = static_cast<int>(index % p_total);
  const int b = static_cast<int>((index / p_total) % g.batch);
  const int c = static_cast<int>(index / p_total / g.batch);
  const int h_col = p / g.wo, w_col = p - h_col * g.wo;
  const int cpg = g.channels / g.deformable_group;
  const int dg = c / cpg;
  const int taps = g.kh * g.kw;
  const long long ldq = static_cast<long long>(g.batch) * g.ldp;
  const float* plane = im + (static_cast<long long>(b) * g.channels + c) * g.height * g.width;
  const float* off = offset + (static_cast<long long>(b) * g.deformable_group + dg) * 2 * taps * p_total;
  const float* msk = mask ? mask + (static_cast<long long>(b) * g.deformable_group + dg) * taps * p_total : nullptr;
  float* col = cols + static_cast<long long>(c) * taps * ldq + static_cast<long long>(b) * g.ldp + p;
  const int h_in = h_col * g.stride_h - g.pad_h, w_in = w_col * g.stride_w - g.pad_w;
  for (int i = 0; i < g.kh; ++i) {
    for (int j = 0; j < g.kw; ++j) {
      const int tap = i * g.kw + j;
      const float h = h_in + i * g.dil_h + off[(2 * tap) * p_total + p];
      const float w = w_in + j * g.dil_w + off[(2 * tap + 1) * p_total + p];
      const DcnSample s = dcn_sample(g, h, w);
      float val = 0.f;
      if (s.inside) {
        const float hh = 1.f - s.lh, hw = 1.f - s.lw;
        const float* q = plane + s.h_low * g.width + s.w_low;
        const float a1 = s.v1 ? q[0] : 0.f, a2 = s.v2 ? q[1] : 0.f;
        const float a3 = s.v3 ? q[g.width] : 0.f, a4 = s.v4 ? q[g.width + 1] : 0.f;
        val = hh * hw * a1 + hh * s.lw * a2 + s.lh * hw * a3 + s.lh * s.lw * a4;
      }
      if (msk) val *= msk[tap * p_total + p];
      col[tap * ldq] = val;
    }
  }
}

// Fused col2im + col2im_coord (+ mask gradient): one item per (b, dg, tap, p), p fastest. `gcols` = W^T . grad_out in
// the [K, B*ldp] layout. The item walks the channels of its deformable group in order (the reference's summation
// order for grad_offset / grad_mask, deform_conv_kernel_cuda.cu:394-424, :741-776), and for each channel
//   * adds d(sample)/d(offset_h), d(sample)/d(offset_w) times the column gradient to its two grad_offset values
//     (get_coordinate_weight, :153-196),
//   * adds column gradient x bilinear sample to its grad_mask value (modulated only, :764),
//   * scatters column gradient x mask x corner weight into grad_im (col2im, :292-338 / :662-712).
// The reference runs two kernels that each re-read the columns, the offsets and (coord) the image; here they share
// one pass. grad_offset / grad_mask are ASSIGNED (as in the reference), grad_im is accumulated.
template <class Add>
MEGA_HD void dcn_col2im_fused_item(long long index, const DcnGeom& g, const float* gcols, const float* im,
                                   const float* offset, const float* mask, float* grad_im, float* grad_offset,
                                   float* grad_mask, Add add) {
  const int p_total = g.ho * g.wo;
  const int taps = g.kh * g.kw;
  const int p = static_cast<int>(index % p_total);
  const int tap = static_cast<int>((index / p_total) % taps);
  const int dg = static_cast<int>((index / p_total / taps) % g.deformable_group);
  const int b = static_cast<int>(index / p_total / taps / g.deformable_group);
  const int h_col = p / g.wo, w_col = p - h_col * g.wo;
  const int i = tap / g.kw, j = tap - i * g.kw;
  const int cpg = g.channels / g.deformable_group;
  const long long ldq = static_cast<long long>(g.batch) * g.ldp;
  const long long og = (static_cast<long long>(b) * g.deformable_group + dg);
  const float* off = offset + og * 2 * taps * p_total;
  const float h = h_col * g.stride_h - g.pad_h + i * g.dil_h + off[(2 * tap) * p_total + p];
  const float w = w_col * g.stride_w - g.pad_w + j * g.dil_w + off[(2 * tap + 1) * p_total + p];
  const float m = mask ? mask[(og * taps + tap) * p_total + p] : 1.f;
  const DcnSample s = dcn_sample(g, h, w);
  float acc_h = 0.f, acc_w = 0.f, acc_m = 0.f;
  if (s.inside) {
    const float hh = 1.f - s.lh, hw = 1.f - s.lw;
    const long long plane = static_cast<long long>(g.height) * g.width;
    const long long corner = static_cast<long long>(s.h_low) * g.width + s.w_low;
    const int c_first = dg * cpg;
    const float* gc = gcols + (static_cast<long long>(c_first) * taps + tap) * ldq + static_cast<long long>(b) * g.ldp + p;
    const float* q = im + (static_cast<long long>(b) * g.channels + c_first) * plane + corner;
    float* gq = grad_im + (static_cast<long long>(b) * g.channels + c_first) * plane + corner;
    for (int c = 0; c < cpg; ++c) {
      const float col = gc[static_cast<long long>(c) * taps * ldq];
      const float a1 = s.v1 ? q[0] : 0.f, a2 = s.v2 ? q[1] : 0.f;
      const float a3 = s.v3 ? q[g.width] : 0.f, a4 = s.v4 ? q[g.width + 1] : 0.f;
      // d/dh: -(1-lw) a1 - lw a2 + (1-lw) a3 + lw a4;   d/dw: -(1-lh) a1 + (1-lh) a2 - lh a3 + lh a4
      const float wh = -hw * a1 - s.lw * a2 + hw * a3 + s.lw * a4;
      const float ww = -hh * a1 + hh * a2 - s.lh * a3 + s.lh * a4;
      acc_h += wh * col * m;
      acc_w += ww * col * m;
      acc_m += col * (hh * hw * a1 + hh * s.lw * a2 + s.lh * hw * a3 + s.lh * s.lw * a4);
      const float top = col * m;
      if (s.v1) add(gq, hh * hw * top);
      if (s.v2) add(gq + 1, hh * s.lw * top);
      if (s.v3) add(gq + g.width, s.lh * hw * top);
      if (s.v4) add(gq + g.width + 1, s.lh * s.lw * top);
      q += plane;
      gq += plane;
    }
  }
  grad_offset[(og * 2 * taps + 2 * tap) * p_total + p] = acc_h;
  grad_offset[(og * 2 * taps + 2 * tap + 1) * p_total + p] = acc_w;
  if (grad_mask) grad_mask[(og * taps + tap) * p_total + p] = acc_m;
}

// grad_bias[c] += sum_{b,p} grad_out[b, c, p]   (the `ones` GEMM of deform_conv_cuda.cu:667-672); one item per channel
// on the host, one warp per channel on the device (train_ops.cu).
MEGA_HD float dcn_channel_sum_item(int c, const float* grad_out, int batch, int channels, int p_total) {
  float s = 0.f;
  for (int b = 0; b < batch; ++b) {
    const float* row = grad_out + (static_cast<long long>(b) * channels + c) * p_total;
    for (int p = 0; p < p_total; ++p) s += row[p];
  }
  return s;
}

// ------------------------------------------------------------------------- deformable PSROI pooling, backward
// Reference: DeformablePSROIPoolBackwardAccKernel (csrc/cuda/deform_pool_kernel_cuda.cu:144-280). One item per output
// element (n, ctop, ph, pw); scatters into grad_in [N, C, H, W] and grad_trans [K, 2*num_classes, part, part].
struct PsRoiGeom {
  int channels, height, width, pooled, output_dim, group_size, part_size, sample_per_part, num_classes,
      channels_each_class, no_trans;
  float spatial_scale, trans_std;
};

template <class Add>
MEGA_HD void deform_psroi_bwd_item(long long index, const PsRoiGeom& g, const float* top_diff, const float* top_count,
                                   const float* data, const float* rois, const float* trans, float* grad_in,
                                   float* grad_trans, Add add) {
  const int pw = static_cast<int>(index % g.pooled);
  const int ph = static_cast<int>((index / g.pooled) % g.pooled);
  const int ctop = static_cast<int>((index / g.pooled / g.pooled) % g.output_dim);
  const long long n = index / g.pooled / g.pooled / g.output_dim;
  if (top_count[index] <= 0.f) return;
  const float* r = rois + n * 5;
  const int roi_batch = static_cast<int>(r[0]);
  const float roi_start_w = roundf(r[1]) * g.spatial_scale - 0.5f;
  const float roi_start_h = roundf(r[2]) * g.spatial_scale - 0.5f;
  const float roi_end_w = (roundf(r[3]) + 1.f) * g.spatial_scale - 0.5f;
  const float roi_end_h = (roundf(r[4]) + 1.f) * g.spatial_scale - 0.5f;
  const float roi_width = fmaxf(roi_end_w - roi_start_w, 0.1f);
  const float roi_height = fmaxf(roi_end_h - roi_start_h, 0.1f);
  const float bin_h = roi_height / static_cast<float>(g.pooled), bin_w = roi_width / static_cast<float>(g.pooled);
  const float sub_h = bin_h / static_cast<float>(g.sample_per_part);
  const float sub_w = bin_w / static_cast<float>(g.sample_per_part);
  const int part_h = static_cast<int>(floorf(static_cast<float>(ph) / g.pooled * g.part_size));
  const int part_w = static_cast<int>(floorf(static_cast<float>(pw) / g.pooled * g.part_size));
  const int class_id = ctop / g.channels_each_class;
  const long long t_x = (((n * g.num_classes + class_id) * 2) * g.part_size + part_h) * g.part_size + part_w;
  const long long t_y = (((n * g.num_classes + class_id) * 2 + 1) * g.part_size + part_h) * g.part_size + part_w;
  const float trans_x = g.no_trans ? 0.f : trans[t_x] * g.trans_std;
  const float trans_y = g.no_trans ? 0.f : trans[t_y] * g.trans_std;
  float wstart = static_cast<float>(pw) * bin_w + roi_start_w;
  wstart += trans_x * roi_width;
  float hstart = static_cast<float>(ph) * bin_h + roi_start_h;
  hstart += trans_y * roi_height;
  const float diff_val = top_diff[index] / top_count[index];
  int gw = static_cast<int>(floorf(static_cast<float>(pw) * g.group_size / g.pooled));
  int gh = static_cast<int>(floorf(static_cast<float>(ph) * g.group_size / g.pooled));
  gw = gw < 0 ? 0 : (gw > g.group_size - 1 ? g.group_size - 1 : gw);
  gh = gh < 0 ? 0 : (gh > g.group_size - 1 ? g.group_size - 1 : gh);
  const int c = (ctop * g.group_size + gh) * g.group_size + gw;
  const long long base = (static_cast<long long>(roi_batch) * g.channels + c) * g.height * g.width;
  const float* plane = data + base;
  float* gplane = grad_in + base;
  for (int ih = 0; ih < g.sample_per_part; ++ih) {
    for (int iw = 0; iw < g.sample_per_part; ++iw) {
      float w = wstart + iw * sub_w;
      float h = hstart + ih * sub_h;
      if (w < -0.5f || w > g.width - 0.5f || h < -0.5f || h > g.height - 0.5f) continue;
      w = fminf(fmaxf(w, 0.f), g.width - 1.f);
      h = fminf(fmaxf(h, 0.f), g.height - 1.f);
      const int x0 = static_cast<int>(floorf(w)), x1 = static_cast<int>(ceilf(w));
      const int y0 = static_cast<int>(floorf(h)), y1 = static_cast<int>(ceilf(h));
      const float dist_x = w - x0, dist_y = h - y0;
      add(gplane + y0 * g.width + x0, (1 - dist_x) * (1 - dist_y) * diff_val);
      add(gplane + y1 * g.width + x0, (1 - dist_x) * dist_y * diff_val);
      add(gplane + y0 * g.width + x1, dist_x * (1 - dist_y) * diff_val);
      add(gplane + y1 * g.width + x1, dist_x * dist_y * diff_val);
      if (g.no_trans) continue;
      const float u00 = plane[y0 * g.width + x0], u01 = plane[y1 * g.width + x0];
      const float u10 = plane[y0 * g.width + x1], u11 = plane[y1 * g.width + x1];
      float diff_x = (u11 * dist_y + u10 * (1 - dist_y) - u01 * dist_y - u00 * (1 - dist_y)) * g.trans_std * diff_val;
      diff_x *= roi_width;
      float diff_y = (u11 * dist_x + u01 * (1 - dist_x) - u10 * dist_x - u00 * (1 - dist_x)) * g.trans_std * diff_val;
      diff_y *= roi_height;
      add(grad_trans + t_x, diff_x);
      add(grad_trans + t_y, diff_y);
    }
  }
}

}  // namespace mega_train
