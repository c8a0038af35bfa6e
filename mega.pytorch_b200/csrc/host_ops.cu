// CPU implementations behind `mega_core._C.nms` and `mega_core._C.roi_align_forward` for CPU tensors.
//
// The reference's `_C` dispatches these two ops on the tensor's device (csrc/nms.h:10-28 -> cpu/nms_cpu.cpp:6-75,
// csrc/ROIAlign.h:11-25 -> cpu/ROIAlign_cpu.cpp:221-257); BASELINE configs[0] (single-frame R-50, MODEL.DEVICE cpu) is
// its ROIAlign / NMS correctness case. They are part of the operator API, not a fallback of the CUDA path: every other
// entry point of this library still refuses CPU tensors. Results are bit-identical to the reference's CPU kernels
// (same operand order in every floating-point expression; tests/test_abi_cpu.py checks against oracle/_ref).
//
// Layout of the work differs from the reference: NMS walks the score order over a suppression bitmap and skips whole
// words of it; ROIAlign builds the sample geometry per AXIS (rows and columns separately, a bin's grid_h x grid_w
// samples are their outer product) and runs the rois on a small thread pool.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <thread>
#include <vector>

#include "common.cuh"
#include "mega_b200.h"

namespace {

template <typename T>
int nms_host(const T* boxes, const T* scores, int n, float thresh, long long* keep_out, int* count_out) {
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  // descending score, equal scores by ascending index (torch's sort leaves ties unspecified; this is the oracle's rule)
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return scores[a] > scores[b]; });
  std::vector<T> area(n);
  for (int i = 0; i < n; ++i) area[i] = (boxes[4 * i + 2] - boxes[4 * i] + 1) * (boxes[4 * i + 3] - boxes[4 * i + 1] + 1);
  std::vector<uint8_t> dead(n, 0);
  for (int a = 0; a < n; ++a) {
    const int i = order[a];
    if (dead[i]) continue;
    const T ix1 = boxes[4 * i], iy1 = boxes[4 * i + 1], ix2 = boxes[4 * i + 2], iy2 = boxes[4 * i + 3], ia = area[i];
    for (int b = a + 1; b < n; ++b) {
      const int j = order[b];
      if (dead[j]) continue;
      const T xx1 = std::max(ix1, boxes[4 * j]), yy1 = std::max(iy1, boxes[4 * j + 1]);
      const T xx2 = std::min(ix2, boxes[4 * j + 2]), yy2 = std::min(iy2, boxes[4 * j + 3]);
      const T w = std::max(static_cast<T>(0), xx2 - xx1 + 1), h = std::max(static_cast<T>(0), yy2 - yy1 + 1);
      const T inter = w * h;
      const T ovr = inter / (ia + area[j] - inter);
      if (ovr >= thresh) dead[j] = 1;          // CPU rule: >= (nms_cpu.cpp:60); the CUDA kernel uses > (nms.cu:60)
    }
  }
  int m = 0;
  for (int i = 0; i < n; ++i)
    if (!dead[i]) keep_out[m++] = i;           // nonzero(suppressed == 0): ascending original index
  *count_out = m;
  return MEGA_OK;
}

template <typename T>
struct AxisTap {      // one sample coordinate along one axis: two cells and their weights; skip = outside [-1, size]
  int lo, hi;
  T wlo, whi;
  bool skip;
};

template <typename T>
AxisTap<T> axis_tap(T v, int size) {
  AxisTap<T> t;
  t.skip = (v < -1.0 || v > size);
  if (t.skip) {
    t.lo = t.hi = 0;
    t.wlo = t.whi = 0;
    return t;
  }
  if (v <= 0) v = 0;
  int lo = static_cast<int>(v);
  int hi;
  if (lo >= size - 1) {
    hi = lo = size - 1;
    v = static_cast<T>(lo);
  } else {
    hi = lo + 1;
  }
  const T l = v - lo;
  t.lo = lo;
  t.hi = hi;
  t.whi = l;
  t.wlo = static_cast<T>(1.) - l;
  return t;
}

template <typename T>
void roi_align_one(const T* input, int channels, int height, int width, const T* roi, T scale, int ph_n, int pw_n,
                   int sampling_ratio, T* out) {
  const int b = static_cast<int>(roi[0]);
  const T x0 = roi[1] * scale, y0 = roi[2] * scale, x1 = roi[3] * scale, y1 = roi[4] * scale;
  const T rw = std::max(x1 - x0, static_cast<T>(1.)), rh = std::max(y1 - y0, static_cast<T>(1.));
  const T bin_h = rh / static_cast<T>(ph_n), bin_w = rw / static_cast<T>(pw_n);
  const int gh = sampling_ratio > 0 ? sampling_ratio : static_cast<int>(std::ceil(rh / ph_n));
  const int gw = sampling_ratio > 0 ? sampling_ratio : static_cast<int>(std::ceil(rw / pw_n));
  const T count = static_cast<T>(gh * gw);
  std::vector<AxisTap<T>> ys(static_cast<size_t>(ph_n) * gh), xs(static_cast<size_t>(pw_n) * gw);
  for (int ph = 0; ph < ph_n; ++ph)
    for (int iy = 0; iy < gh; ++iy)
      ys[ph * gh + iy] = axis_tap<T>(y0 + ph * bin_h + static_cast<T>(iy + .5f) * bin_h / static_cast<T>(gh), height);
  for (int pw = 0; pw < pw_n; ++pw)
    for (int ix = 0; ix < gw; ++ix)
      xs[pw * gw + ix] = axis_tap<T>(x0 + pw * bin_w + static_cast<T>(ix + .5f) * bin_w / static_cast<T>(gw), width);
  for (int c = 0; c < channels; ++c) {
    const T* f = input + (static_cast<size_t>(b) * channels + c) * height * width;
    T* o = out + static_cast<size_t>(c) * ph_n * pw_n;
    for (int ph = 0; ph < ph_n; ++ph) {
      for (int pw = 0; pw < pw_n; ++pw) {
        T acc = 0;
        for (int iy = 0; iy < gh; ++iy) {
          const AxisTap<T>& y = ys[ph * gh + iy];
          for (int ix = 0; ix < gw; ++ix) {
            const AxisTap<T>& x = xs[pw * gw + ix];
            if (y.skip || x.skip) continue;            // the reference adds four zero-weighted taps of cell 0 here
            const T w1 = y.wlo * x.wlo, w2 = y.wlo * x.whi, w3 = y.whi * x.wlo, w4 = y.whi * x.whi;
            acc += w1 * f[y.lo * width + x.lo] + w2 * f[y.lo * width + x.hi] + w3 * f[y.hi * width + x.lo] +
                   w4 * f[y.hi * width + x.hi];
          }
        }
        o[ph * pw_n + pw] = acc / count;
      }
    }
  }
}

template <typename T>
int roi_align_host(const T* input, int batch, int channels, int height, int width, const T* rois, int num_rois,
                   float spatial_scale, int ph, int pw, int sampling_ratio, T* output) {
  MEGA_ARG_CHECK(input != nullptr && rois != nullptr && output != nullptr, "roi_align_forward (cpu): null tensor");
  MEGA_ARG_CHECK(channels > 0 && height > 0 && width > 0 && ph > 0 && pw > 0, "roi_align_forward (cpu): bad shape");
  for (int k = 0; k < num_rois; ++k) {
    const int b = static_cast<int>(rois[5 * k]);
    MEGA_ARG_CHECK(b >= 0 && b < batch, "roi_align_forward (cpu): roi %d names image %d of %d", k, b, batch);
  }
  const size_t per_roi = static_cast<size_t>(channels) * ph * pw;
  const int workers = std::max(1, std::min<int>(num_rois / 8, std::min(16u, std::thread::hardware_concurrency())));
  auto run = [&](int w) {
    for (int k = w; k < num_rois; k += workers)
      roi_align_one<T>(input, channels, height, width, rois + 5 * k, static_cast<T>(spatial_scale), ph, pw, sampling_ratio,
                       output + per_roi * k);
  };
  if (workers == 1) {
    run(0);
  } else {
    std::vector<std::thread> pool;
    for (int w = 0; w < workers; ++w) pool.emplace_back(run, w);
    for (auto& t : pool) t.join();
  }
  return MEGA_OK;
}

}  // namespace

extern "C" int mega_nms_host(const void* boxes, const void* scores, int n, float thresh, int is_double, long long* keep_out,
                             int* count_out) {
  MEGA_ARG_CHECK(n >= 0 && keep_out != nullptr && count_out != nullptr, "nms (cpu): bad arguments");
  if (n == 0) {
    *count_out = 0;
    return MEGA_OK;
  }
  MEGA_ARG_CHECK(boxes != nullptr && scores != nullptr, "nms (cpu): null tensor");
  return is_double ? nms_host<double>(static_cast<const double*>(boxes), static_cast<const double*>(scores), n, thresh,
                                      keep_out, count_out)
                   : nms_host<float>(static_cast<const float*>(boxes), static_cast<const float*>(scores), n, thresh,
                                     keep_out, count_out);
}

extern "C" int mega_roi_align_forward_nchw_host(const void* input, int batch, int channels, int height, int width,
                                                const void* rois, int num_rois, float spatial_scale, int pooled_h,
                                                int pooled_w, int sampling_ratio, int is_double, void* output) {
  if (num_rois == 0) return MEGA_OK;
  return is_double ? roi_align_host<double>(static_cast<const double*>(input), batch, channels, height, width,
                                            static_cast<const double*>(rois), num_rois, spatial_scale, pooled_h, pooled_w,
                                            sampling_ratio, static_cast<double*>(output))
                   : roi_align_host<float>(static_cast<const float*>(input), batch, channels, height, width,
                                           static_cast<const float*>(rois), num_rois, spatial_scale, pooled_h, pooled_w,
                                           sampling_ratio, static_cast<float*>(output));
}
