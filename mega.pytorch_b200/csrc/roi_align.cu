// ROIAlign forward (mega_core/csrc/cuda/ROIAlign_cuda.cu:16-122, cpu/ROIAlign_cpu.cpp:17-219).
//
// Two layouts:
//   * mega_roi_align_forward_nchw -- the `_C.roi_align_forward` contract: NCHW fp32 in,
//     [K,C,ph,pw] out, one thread per output element (same mapping as the reference kernel);
//   * mega_roi_align_forward_nhwc -- the engine's layout: NHWC feature map in, [K, ph*pw, C] out
//     (bin-major, channel-minor). One CTA per (roi, bin); threads span channels with 128-bit
//     loads, so every bilinear corner is one fully coalesced row read and the output row feeds
//     the l_fcs[0] GEMM as its K-major A operand without a transpose.
// Arithmetic is written with explicit round-to-nearest ops (no FMA contraction) in the
// reference's association order, so results are bit-identical to the C oracle.
#include "common.cuh"
#include "mega_b200.h"

namespace mega {

struct Bilinear {
  int y_low, y_high, x_low, x_high;
  float w1, w2, w3, w4;
  bool empty;
};

__device__ __forceinline__ Bilinear bilinear_setup(int height, int width, float y, float x) {
  Bilinear b;
  b.empty = (y < -1.0f || y > static_cast<float>(height) || x < -1.0f || x > static_cast<float>(width));
  if (b.empty) {
    b.y_low = b.y_high = b.x_low = b.x_high = 0;
    b.w1 = b.w2 = b.w3 = b.w4 = 0.f;
    return b;
  }
  if (y <= 0) y = 0;
  if (x <= 0) x = 0;
  int y_low = static_cast<int>(y), x_low = static_cast<int>(x), y_high, x_high;
  if (y_low >= height - 1) {
    y_high = y_low = height - 1;
    y = static_cast<float>(y_low);
  } else {
    y_high = y_low + 1;
  }
  if (x_low >= width - 1) {
    x_high = x_low = width - 1;
    x = static_cast<float>(x_low);
  } else {
    x_high = x_low + 1;
  }
  const float ly = __fsub_rn(y, static_cast<float>(y_low)), lx = __fsub_rn(x, static_cast<float>(x_low));
  const float hy = __fsub_rn(1.f, ly), hx = __fsub_rn(1.f, lx);
  b.y_low = y_low; b.y_high = y_high; b.x_low = x_low; b.x_high = x_high;
  b.w1 = __fmul_rn(hy, hx); b.w2 = __fmul_rn(hy, lx); b.w3 = __fmul_rn(ly, hx); b.w4 = __fmul_rn(ly, lx);
  return b;
}

__device__ __forceinline__ float blend(const Bilinear& b, float v1, float v2, float v3, float v4) {
  return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(b.w1, v1), __fmul_rn(b.w2, v2)), __fmul_rn(b.w3, v3)),
                   __fmul_rn(b.w4, v4));
}

struct RoiGeom {
  float start_w, start_h, bin_w, bin_h;
  int grid_w, grid_h, batch;
};

__device__ __forceinline__ RoiGeom roi_geom(const float* roi, float scale, int ph, int pw, int sampling_ratio) {
  RoiGeom g;
  g.batch = static_cast<int>(roi[0]);
  g.start_w = __fmul_rn(roi[1], scale);
  g.start_h = __fmul_rn(roi[2], scale);
  const float end_w = __fmul_rn(roi[3], scale), end_h = __fmul_rn(roi[4], scale);
  const float roi_w = fmaxf(__fsub_rn(end_w, g.start_w), 1.f);
  const float roi_h = fmaxf(__fsub_rn(end_h, g.start_h), 1.f);
  g.bin_h = __fdiv_rn(roi_h, static_cast<float>(ph));
  g.bin_w = __fdiv_rn(roi_w, static_cast<float>(pw));
  g.grid_h = sampling_ratio > 0 ? sampling_ratio : static_cast<int>(ceilf(__fdiv_rn(roi_h, static_cast<float>(ph))));
  g.grid_w = sampling_ratio > 0 ? sampling_ratio : static_cast<int>(ceilf(__fdiv_rn(roi_w, static_cast<float>(pw))));
  return g;
}

__device__ __forceinline__ float sample_coord(float start, int p, float bin, int i, int grid) {
  // start + p*bin + (i + .5f) * bin / grid   (ROIAlign_cuda.cu:106-110)
  return __fadd_rn(__fadd_rn(start, __fmul_rn(static_cast<float>(p), bin)),
                   __fdiv_rn(__fmul_rn(static_cast<float>(i) + .5f, bin), static_cast<float>(grid)));
}

__global__ void roi_align_nchw_kernel(const float* __restrict__ in, int channels, int height, int width,
                                      const float* __restrict__ rois, long long total, float scale, int ph, int pw,
                                      int sampling_ratio, float* __restrict__ out) {
  for (long long index = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; index < total;
       index += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int pwi = static_cast<int>(index % pw);
    const int phi = static_cast<int>((index / pw) % ph);
    const int c = static_cast<int>((index / pw / ph) % channels);
    const long long n = index / pw / ph / channels;
    const RoiGeom g = roi_geom(rois + n * 5, scale, ph, pw, sampling_ratio);
    const float* plane = in + (static_cast<long long>(g.batch) * channels + c) * height * width;
    const float count = static_cast<float>(g.grid_h * g.grid_w);
    float acc = 0.f;
    for (int iy = 0; iy < g.grid_h; ++iy) {
      const float y = sample_coord(g.start_h, phi, g.bin_h, iy, g.grid_h);
      for (int ix = 0; ix < g.grid_w; ++ix) {
        const float x = sample_coord(g.start_w, pwi, g.bin_w, ix, g.grid_w);
        const Bilinear b = bilinear_setup(height, width, y, x);
        float val = 0.f;
        if (!b.empty) {
          val = blend(b, plane[b.y_low * width + b.x_low], plane[b.y_low * width + b.x_high],
                      plane[b.y_high * width + b.x_low], plane[b.y_high * width + b.x_high]);
        }
        acc = __fadd_rn(acc, val);
      }
    }
    out[index] = __fdiv_rn(acc, count);
  }
}

// grid = (ph*pw, K); block = 256 threads, each thread owns float4 channel groups.
__global__ void __launch_bounds__(256)
roi_align_nhwc_kernel(const float* __restrict__ in, int channels, int height, int width, long long in_img_stride,
                      const float* __restrict__ rois, int roi_ld, int roi_box_off, const int* __restrict__ roi_batch,
                      float scale, int ph, int pw, int sampling_ratio, float* __restrict__ out,
                      long long out_roi_stride) {
  const int bin = blockIdx.x;
  const int n = blockIdx.y;
  const int phi = bin / pw, pwi = bin - phi * pw;
  float roi5[5];
  roi5[0] = roi_batch ? static_cast<float>(roi_batch[n]) : 0.f;
  const float* rb = rois + static_cast<long long>(n) * roi_ld + roi_box_off;
  if (roi_box_off < 0) {  // packed [K,5] (batch, x1, y1, x2, y2)
    rb = rois + static_cast<long long>(n) * roi_ld;
    roi5[0] = rb[0];
    rb += 1;
  }
  roi5[1] = rb[0]; roi5[2] = rb[1]; roi5[3] = rb[2]; roi5[4] = rb[3];
  const RoiGeom g = roi_geom(roi5, scale, ph, pw, sampling_ratio);
  const float* img = in + static_cast<long long>(g.batch) * in_img_stride;
  const float count = static_cast<float>(g.grid_h * g.grid_w);
  float* orow = out + static_cast<long long>(n) * out_roi_stride + static_cast<long long>(bin) * channels;
  for (int c = threadIdx.x * 4; c < channels; c += blockDim.x * 4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int iy = 0; iy < g.grid_h; ++iy) {
      const float y = sample_coord(g.start_h, phi, g.bin_h, iy, g.grid_h);
      for (int ix = 0; ix < g.grid_w; ++ix) {
        const float x = sample_coord(g.start_w, pwi, g.bin_w, ix, g.grid_w);
        const Bilinear b = bilinear_setup(height, width, y, x);
        if (b.empty) continue;  // adds exactly 0 in the reference
        const float4 v1 = ldg_f4(img + (static_cast<long long>(b.y_low) * width + b.x_low) * channels + c);
        const float4 v2 = ldg_f4(img + (static_cast<long long>(b.y_low) * width + b.x_high) * channels + c);
        const float4 v3 = ldg_f4(img + (static_cast<long long>(b.y_high) * width + b.x_low) * channels + c);
        const float4 v4 = ldg_f4(img + (static_cast<long long>(b.y_high) * width + b.x_high) * channels + c);
        acc.x = __fadd_rn(acc.x, blend(b, v1.x, v2.x, v3.x, v4.x));
        acc.y = __fadd_rn(acc.y, blend(b, v1.y, v2.y, v3.y, v4.y));
        acc.z = __fadd_rn(acc.z, blend(b, v1.z, v2.z, v3.z, v4.z));
        acc.w = __fadd_rn(acc.w, blend(b, v1.w, v2.w, v3.w, v4.w));
      }
    }
    acc.x = __fdiv_rn(acc.x, count); acc.y = __fdiv_rn(acc.y, count);
    acc.z = __fdiv_rn(acc.z, count); acc.w = __fdiv_rn(acc.w, count);
    *reinterpret_cast<float4*>(orow + c) = acc;
  }
}

// Same arithmetic, less L2 traffic: one CTA per (roi, 64-channel slice) first copies the cells the roi can touch
// (rows r0..r1 x cols c0..c1 of the map, 256 B per cell) into shared memory, then all 49 bins sample from there;
// neighbouring samples share corners, so each cell is fetched once per slice instead of up to ~8 times. ROIs whose
// footprint exceeds the shared-memory budget (whole-image boxes) read global memory directly.
constexpr int kRoiSlice = 64;            // channels per CTA
constexpr int kRoiMaxCells = 192;        // 192 cells x 256 B = 48 KB per CTA (4 CTAs per SM)

__global__ void __launch_bounds__(256)
roi_align_nhwc_cached_kernel(const float* __restrict__ in, int channels, int height, int width, long long in_img_stride,
                             const float* __restrict__ rois, int roi_ld, int roi_box_off,
                             const int* __restrict__ roi_batch, float scale, int ph, int pw, int sampling_ratio,
                             float* __restrict__ out, long long out_roi_stride) {
  extern __shared__ float4 cell_s[];   // [cells][16] float4
  const int slice = blockIdx.x;
  const int n = blockIdx.y;
  float roi5[5];
  roi5[0] = roi_batch ? static_cast<float>(roi_batch[n]) : 0.f;
  const float* rb = rois + static_cast<long long>(n) * roi_ld + roi_box_off;
  if (roi_box_off < 0) {
    rb = rois + static_cast<long long>(n) * roi_ld;
    roi5[0] = rb[0];
    rb += 1;
  }
  roi5[1] = rb[0]; roi5[2] = rb[1]; roi5[3] = rb[2]; roi5[4] = rb[3];
  const RoiGeom g = roi_geom(roi5, scale, ph, pw, sampling_ratio);
  const float* img = in + static_cast<long long>(g.batch) * in_img_stride + slice * kRoiSlice;
  // footprint of every sample this roi can take (after the reference's clamping of y, x to the map)
  const float end_h = __fadd_rn(g.start_h, __fmul_rn(g.bin_h, static_cast<float>(ph)));
  const float end_w = __fadd_rn(g.start_w, __fmul_rn(g.bin_w, static_cast<float>(pw)));
  int r0 = static_cast<int>(floorf(fmaxf(g.start_h, 0.f))), r1 = static_cast<int>(floorf(fmaxf(end_h, 0.f))) + 1;
  int c0 = static_cast<int>(floorf(fmaxf(g.start_w, 0.f))), c1 = static_cast<int>(floorf(fmaxf(end_w, 0.f))) + 1;
  r0 = min(max(r0, 0), height - 1); r1 = min(max(r1, 0), height - 1);
  c0 = min(max(c0, 0), width - 1); c1 = min(max(c1, 0), width - 1);
  const int rh = r1 - r0 + 1, rw = c1 - c0 + 1;
  const bool cached = (rh * rw <= kRoiMaxCells);
  if (cached) {
    for (int i = threadIdx.x; i < rh * rw * 16; i += blockDim.x) {
      const int cell = i >> 4, q = i & 15;
      const int rr = cell / rw, cc = cell - rr * rw;
      cell_s[i] = ldg_f4(img + (static_cast<long long>(r0 + rr) * width + (c0 + cc)) * channels + q * 4);
    }
  }
  __syncthreads();
  const int q = threadIdx.x & 15;
  const float count = static_cast<float>(g.grid_h * g.grid_w);
  float* obase = out + static_cast<long long>(n) * out_roi_stride + slice * kRoiSlice + q * 4;
  for (int bin = threadIdx.x >> 4; bin < ph * pw; bin += blockDim.x >> 4) {
    const int phi = bin / pw, pwi = bin - phi * pw;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int iy = 0; iy < g.grid_h; ++iy) {
      const float y = sample_coord(g.start_h, phi, g.bin_h, iy, g.grid_h);
      for (int ix = 0; ix < g.grid_w; ++ix) {
        const float x = sample_coord(g.start_w, pwi, g.bin_w, ix, g.grid_w);
        const Bilinear b = bilinear_setup(height, width, y, x);
        if (b.empty) continue;
        float4 v1, v2, v3, v4;
        const bool inside = cached && b.y_low >= r0 && b.y_high <= r1 && b.x_low >= c0 && b.x_high <= c1;
        if (inside) {
          v1 = cell_s[((b.y_low - r0) * rw + (b.x_low - c0)) * 16 + q];
          v2 = cell_s[((b.y_low - r0) * rw + (b.x_high - c0)) * 16 + q];
          v3 = cell_s[((b.y_high - r0) * rw + (b.x_low - c0)) * 16 + q];
          v4 = cell_s[((b.y_high - r0) * rw + (b.x_high - c0)) * 16 + q];
        } else {
          v1 = ldg_f4(img + (static_cast<long long>(b.y_low) * width + b.x_low) * channels + q * 4);
          v2 = ldg_f4(img + (static_cast<long long>(b.y_low) * width + b.x_high) * channels + q * 4);
          v3 = ldg_f4(img + (static_cast<long long>(b.y_high) * width + b.x_low) * channels + q * 4);
          v4 = ldg_f4(img + (static_cast<long long>(b.y_high) * width + b.x_high) * channels + q * 4);
        }
        acc.x = __fadd_rn(acc.x, blend(b, v1.x, v2.x, v3.x, v4.x));
        acc.y = __fadd_rn(acc.y, blend(b, v1.y, v2.y, v3.y, v4.y));
        acc.z = __fadd_rn(acc.z, blend(b, v1.z, v2.z, v3.z, v4.z));
        acc.w = __fadd_rn(acc.w, blend(b, v1.w, v2.w, v3.w, v4.w));
      }
    }
    acc.x = __fdiv_rn(acc.x, count); acc.y = __fdiv_rn(acc.y, count);
    acc.z = __fdiv_rn(acc.z, count); acc.w = __fdiv_rn(acc.w, count);
    *reinterpret_cast<float4*>(obase + static_cast<long long>(bin) * channels) = acc;
  }
}

}  // namespace mega

using namespace mega;

extern "C" int mega_roi_align_forward_nchw(const float* input, int batch, int channels, int height, int width,
                                           const float* rois, int num_rois, float spatial_scale, int pooled_h,
                                           int pooled_w, int sampling_ratio, float* output, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  (void)batch;
  MEGA_ARG_CHECK(pooled_h > 0 && pooled_w > 0 && channels > 0, "roi_align: bad pooled size / channels");
  const long long total = static_cast<long long>(num_rois) * channels * pooled_h * pooled_w;
  if (total == 0) return MEGA_OK;  // ROIAlign_cuda.cu:278-281
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  roi_align_nchw_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(input, channels, height, width, rois, total,
                                                                      spatial_scale, pooled_h, pooled_w,
                                                                      sampling_ratio, output);
  MEGA_CUDA_CHECK(cudaGetLastError());
  return MEGA_OK;
}

extern "C" int mega_roi_align_forward_nhwc(const float* input, int channels, int height, int width,
                                           long long in_img_stride, const float* rois, int roi_ld, int roi_box_off,
                                           const int* roi_batch, int num_rois, float spatial_scale, int pooled_h,
                                           int pooled_w, int sampling_ratio, float* output, long long out_roi_stride,
                                           void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  MEGA_ARG_CHECK((channels & 3) == 0, "roi_align_nhwc: channels must be a multiple of 4");
  MEGA_ARG_CHECK((reinterpret_cast<uintptr_t>(input) & 15) == 0 && (reinterpret_cast<uintptr_t>(output) & 15) == 0 &&
                     (out_roi_stride & 3) == 0,
                 "roi_align_nhwc: 16-byte alignment required");
  if (num_rois == 0) return MEGA_OK;
  if ((channels % kRoiSlice) == 0) {
    static bool configured = false;
    if (!configured) {
      MEGA_CUDA_CHECK(cudaFuncSetAttribute(roi_align_nhwc_cached_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           kRoiMaxCells * 256));
      configured = true;
    }
    dim3 cgrid(channels / kRoiSlice, num_rois);
    roi_align_nhwc_cached_kernel<<<cgrid, 256, kRoiMaxCells * 256, stream>>>(
        input, channels, height, width, in_img_stride, rois, roi_ld, roi_box_off, roi_batch, spatial_scale, pooled_h,
        pooled_w, sampling_ratio, output, out_roi_stride);
    MEGA_CUDA_CHECK(cudaGetLastError());
    return MEGA_OK;
  }
  dim3 grid(pooled_h * pooled_w, num_rois);
  roi_align_nhwc_kernel<<<grid, 256, 0, stream>>>(input, channels, height, width, in_img_stride, rois, roi_ld,
                                                  roi_box_off, roi_batch, spatial_scale, pooled_h, pooled_w,
                                                  sampling_ratio, output, out_roi_stride);
  MEGA_CUDA_CHECK(cudaGetLastError());
  return MEGA_OK;
}
