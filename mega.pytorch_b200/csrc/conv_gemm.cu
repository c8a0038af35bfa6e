// Implicit-GEMM convolution / GEMM on the 5th-gen tensor cores (tcgen05, TF32 operands,
// FP32 accumulate in TMEM), fed by TMA with the 128-byte shared-memory swizzle.
//
// One kernel serves every dense contraction of the MEGA hot path:
//   * backbone / res5 / RPN-head convolutions (1x1, 3x3, 3x3 dilated) over NHWC maps
//     (reference: mega_core/modeling/backbone/resnet.py:324-344, rpn/rpn.py:99-106),
//     with FrozenBatchNorm scale/bias (layers/batch_norm.py:26-31), residual add and ReLU
//     folded into the epilogue;
//   * Linear layers (make_layers.py:80-92) as a 1x1 "convolution" over an H=1 image;
//   * the per-head Q.K^T and P.V' products of the relation module
//     (roi_box_feature_extractors.py:602-646) through the batch (grid.z) offsets.
//
// Tiling: the M tile is a th x tw rectangle of 128 output pixels, so the A operand of filter
// tap (r,s) is the same rectangle shifted by (r,s)*dilation - pad: a plain 4-D tiled TMA load
// with out-of-bounds zero fill supplies the padding. K is consumed in slabs of 32 floats
// (= one 128 B swizzle row) per tap. Warp roles: warp 0 TMA producer, warp 1 MMA issuer,
// warps 2-5 epilogue (TMEM -> registers -> global).
#include "common.cuh"
#include "mega_b200.h"

namespace mega {

constexpr int kBM = 128;        // UMMA M (one CTA)
constexpr int kBK = 32;         // floats per K slab = 128 bytes
constexpr int kUmmaK = 8;       // tf32
constexpr int kThreads = 192;   // 6 warps

struct ConvGemmParams {
  int tiles_w, tiles_h, tile_w, tile_h;
  int out_h, out_w, n_img;
  int taps_r, taps_s, dil, pad;
  int k_chunks;  // ceil(Cin / 32)
  int cout;
  float* out;
  long long out_ld;
  const float* scale;
  const float* bias;
  const float* residual;
  long long res_ld;
  int relu;
  int a_c_off, a_n_off, b_k_off, b_n_off;
  long long out_z_off;
  long long res_z_off;
  int bias_z_off;
  int splits;
  float* partial;  // [splits][M_total][cout] when splits > 1
};

template <int BN, int STAGES>
struct SmemLayout {
  static constexpr int kABytes = kBM * 128;
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = STAGES * kStageBytes;
  static constexpr int kTotal = kBarOffset + (2 * STAGES + 1) * 8 + 16 + 1024;  // + align slack
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const ConvGemmParams p) {
  using L = SmemLayout<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  // the 128B swizzle pattern is a function of the absolute smem address: align to 1024 B
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- tile coordinates
  const int tile = blockIdx.x;
  const int tw_i = tile % p.tiles_w;
  const int th_i = (tile / p.tiles_w) % p.tiles_h;
  const int img = tile / (p.tiles_w * p.tiles_h);
  const int h0 = th_i * p.tile_h;
  const int w0 = tw_i * p.tile_w;
  const int n0 = blockIdx.y * BN;
  const int batch = blockIdx.z / p.splits;
  const int split = blockIdx.z % p.splits;
  const int total_kb = p.taps_r * p.taps_s * p.k_chunks;
  const int kb_begin = static_cast<int>((static_cast<long long>(total_kb) * split) / p.splits);
  const int kb_end = static_cast<int>((static_cast<long long>(total_kb) * (split + 1)) / p.splits);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, BN);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        const int tap = kb / p.k_chunks;
        const int kc = kb - tap * p.k_chunks;
        const int r = tap / p.taps_s;
        const int s = tap - r * p.taps_s;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* a_dst = smem + stage * L::kStageBytes;
        uint8_t* b_dst = a_dst + L::kABytes;
        mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
        tma_load_4d(a_dst, &tmA, &full_bar[stage], kc * kBK + batch * p.a_c_off,
                    w0 + s * p.dil - p.pad, h0 + r * p.dil - p.pad, img + batch * p.a_n_off);
        tma_load_3d(b_dst, &tmB, &full_bar[stage], kc * kBK + batch * p.b_k_off,
                    n0 + batch * p.b_n_off, tap);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc<2>(kBM, BN);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + stage * L::kStageBytes);
        const uint32_t b_addr = a_addr + L::kABytes;
        const uint64_t adesc = umma_desc_sw128(a_addr);
        const uint64_t bdesc = umma_desc_sw128(b_addr);
#pragma unroll
        for (int k = 0; k < kBK / kUmmaK; ++k) {
          // advance 8 floats = 32 B inside the swizzle row: +2 in 16-byte units
          umma_tf32(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc,
                    (kb > kb_begin || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;
    const int hl = row / p.tile_w;
    const int wl = row - hl * p.tile_w;
    const int h = h0 + hl;
    const int w = w0 + wl;
    const bool row_ok = (h < p.out_h) && (w < p.out_w);
    const long long pix = (static_cast<long long>(img) * p.out_h + h) * p.out_w + w;

    if (kb_end > kb_begin) {
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
    }
    const bool have_acc = kb_end > kb_begin;
    float* out_row;
    const float* res_row = nullptr;
    if (p.splits > 1) {
      const long long m_total = static_cast<long long>(p.n_img) * p.out_h * p.out_w;
      out_row = p.partial + ((static_cast<long long>(blockIdx.z) * m_total) + pix) * p.cout;
    } else {
      out_row = p.out + batch * p.out_z_off + pix * p.out_ld;
      if (p.residual) res_row = p.residual + batch * p.res_z_off + pix * p.res_ld;
    }
    const float* scale_p = p.scale ? p.scale + batch * p.bias_z_off : nullptr;
    const float* bias_p = p.bias ? p.bias + batch * p.bias_z_off : nullptr;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(out_row) & 15) == 0) &&
                        (res_row == nullptr || (reinterpret_cast<uintptr_t>(res_row) & 15) == 0);
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t acc[32];
      __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the divergent stores
      if (have_acc) {
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, acc);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = 0u;
      }
      const int nb = n0 + c * 32;
      if (!row_ok || nb >= p.cout) continue;
      if (p.splits > 1) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          if (nb + j + 3 < p.cout && vec_ok) {
            float4 v = make_float4(__uint_as_float(acc[j]), __uint_as_float(acc[j + 1]),
                                   __uint_as_float(acc[j + 2]), __uint_as_float(acc[j + 3]));
            *reinterpret_cast<float4*>(out_row + nb + j) = v;
          } else {
            for (int t = 0; t < 4; ++t)
              if (nb + j + t < p.cout) out_row[nb + j + t] = __uint_as_float(acc[j + t]);
          }
        }
        continue;
      }
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const int n = nb + j;
        if (n + 3 < p.cout && vec_ok) {
          float4 v = make_float4(__uint_as_float(acc[j]), __uint_as_float(acc[j + 1]),
                                 __uint_as_float(acc[j + 2]), __uint_as_float(acc[j + 3]));
          if (scale_p) {
            const float4 sc = ldg_f4(scale_p + n);
            v.x *= sc.x; v.y *= sc.y; v.z *= sc.z; v.w *= sc.w;
          }
          if (bias_p) {
            const float4 bi = ldg_f4(bias_p + n);
            v.x += bi.x; v.y += bi.y; v.z += bi.z; v.w += bi.w;
          }
          if (res_row) {
            const float4 rr = ldg_f4(res_row + n);
            v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
          }
          if (p.relu) {
            v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
          }
          *reinterpret_cast<float4*>(out_row + n) = v;
        } else {
          for (int t = 0; t < 4; ++t) {
            if (n + t < p.cout) {
              float v = __uint_as_float(acc[j + t]);
              if (scale_p) v *= __ldg(scale_p + n + t);
              if (bias_p) v += __ldg(bias_p + n + t);
              if (res_row) v += __ldg(res_row + n + t);
              if (p.relu) v = fmaxf(v, 0.f);
              out_row[n + t] = v;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

// sums split-K partials and applies the epilogue
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int splits, long long m_total,
                                     int cout, float* __restrict__ out, long long out_ld,
                                     const float* __restrict__ scale, const float* __restrict__ bias,
                                     const float* __restrict__ residual, long long res_ld, int relu) {
  const long long total = m_total * cout;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long m = i / cout;
    const int n = static_cast<int>(i - m * cout);
    float v = 0.f;
    for (int s = 0; s < splits; ++s) v += partial[(static_cast<long long>(s) * m_total + m) * cout + n];
    if (scale) v *= scale[n];
    if (bias) v += bias[n];
    if (residual) v += residual[m * res_ld + n];
    if (relu) v = fmaxf(v, 0.f);
    out[m * out_ld + n] = v;
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || p == nullptr) {
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

static int g_tf32_round = 1;  // TMA converts fp32 -> tf32 (round to nearest) while loading

template <int BN, int STAGES>
static int launch_cfg(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvGemmParams& p, dim3 grid,
                      cudaStream_t stream) {
  using L = SmemLayout<BN, STAGES>;
  static bool configured = false;
  if (!configured) {
    MEGA_CUDA_CHECK(cudaFuncSetAttribute(conv_gemm_tf32_kernel<BN, STAGES>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    configured = true;
  }
  conv_gemm_tf32_kernel<BN, STAGES><<<grid, kThreads, L::kTotal, stream>>>(tmA, tmB, p);
  MEGA_CUDA_CHECK(cudaGetLastError());
  return MEGA_OK;
}

}  // namespace mega

using namespace mega;

extern "C" int mega_set_tf32_rounding(int enable) {
  int old = g_tf32_round;
  g_tf32_round = enable ? 1 : 0;
  return old;
}

extern "C" int mega_conv_gemm_tf32(const mega_conv_gemm_desc* d, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  MEGA_ARG_CHECK(d != nullptr, "conv_gemm: null descriptor");
  MEGA_ARG_CHECK(d->tile_h > 0 && d->tile_w > 0 && d->tile_h * d->tile_w == kBM,
                 "conv_gemm: tile_h*tile_w must be 128 (got %dx%d)", d->tile_h, d->tile_w);
  MEGA_ARG_CHECK(d->tile_w <= 256 && d->tile_h <= 256, "conv_gemm: tile too large for a TMA box");
  MEGA_ARG_CHECK(d->block_n == 32 || d->block_n == 64 || d->block_n == 128 || d->block_n == 256,
                 "conv_gemm: block_n must be 32/64/128/256");
  MEGA_ARG_CHECK((reinterpret_cast<uintptr_t>(d->a) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->b) & 15) == 0,
                 "conv_gemm: operand base pointers must be 16-byte aligned");
  MEGA_ARG_CHECK((d->a_stride_w % 4) == 0 && (d->a_stride_h % 4) == 0 && (d->a_stride_n % 4) == 0,
                 "conv_gemm: activation strides must be multiples of 4 floats");
  MEGA_ARG_CHECK((d->b_stride_n % 4) == 0 && (d->b_stride_tap % 4) == 0,
                 "conv_gemm: weight strides must be multiples of 4 floats");
  MEGA_ARG_CHECK(d->batch >= 1 && d->splits >= 1, "conv_gemm: batch/splits must be >= 1");
  MEGA_ARG_CHECK(d->splits == 1 || (d->batch == 1 && d->partial != nullptr),
                 "conv_gemm: split-K needs batch==1 and a partial workspace");
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    mega_set_error("conv_gemm: cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return MEGA_ERR_CUDA;
  }
  const CUtensorMapDataType dt = g_tf32_round ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;

  CUtensorMap tmA, tmB;
  {
    cuuint64_t gdim[4] = {static_cast<cuuint64_t>(d->a_c), static_cast<cuuint64_t>(d->a_w),
                          static_cast<cuuint64_t>(d->a_h), static_cast<cuuint64_t>(d->a_n)};
    cuuint64_t gstr[3] = {static_cast<cuuint64_t>(d->a_stride_w) * 4, static_cast<cuuint64_t>(d->a_stride_h) * 4,
                          static_cast<cuuint64_t>(d->a_stride_n) * 4};
    cuuint32_t box[4] = {static_cast<cuuint32_t>(kBK), static_cast<cuuint32_t>(d->tile_w),
                         static_cast<cuuint32_t>(d->tile_h), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmA, dt, 4, const_cast<float*>(d->a), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      mega_set_error("conv_gemm: encode A tensor map failed (CUresult %d) dims %d %d %d %d strides %lld %lld %lld",
                     static_cast<int>(r), d->a_c, d->a_w, d->a_h, d->a_n, d->a_stride_w, d->a_stride_h,
                     d->a_stride_n);
      return MEGA_ERR_CUDA;
    }
  }
  {
    const int taps = d->taps_r * d->taps_s;
    cuuint64_t gdim[3] = {static_cast<cuuint64_t>(d->b_k), static_cast<cuuint64_t>(d->b_n),
                          static_cast<cuuint64_t>(taps)};
    cuuint64_t gstr[2] = {static_cast<cuuint64_t>(d->b_stride_n) * 4,
                          static_cast<cuuint64_t>(taps > 1 ? d->b_stride_tap : d->b_stride_n * d->b_n) * 4};
    cuuint32_t box[3] = {static_cast<cuuint32_t>(kBK), static_cast<cuuint32_t>(d->block_n), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&tmB, dt, 3, const_cast<float*>(d->b), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      mega_set_error("conv_gemm: encode B tensor map failed (CUresult %d) dims %d %d %d strides %lld %lld",
                     static_cast<int>(r), d->b_k, d->b_n, taps, d->b_stride_n, d->b_stride_tap);
      return MEGA_ERR_CUDA;
    }
  }

  ConvGemmParams p;
  p.tile_w = d->tile_w;
  p.tile_h = d->tile_h;
  p.tiles_w = mega_ceil_div(d->out_w, d->tile_w);
  p.tiles_h = mega_ceil_div(d->out_h, d->tile_h);
  p.out_h = d->out_h;
  p.out_w = d->out_w;
  p.n_img = d->n_img;
  p.taps_r = d->taps_r;
  p.taps_s = d->taps_s;
  p.dil = d->dil;
  p.pad = d->pad;
  p.k_chunks = mega_ceil_div(d->k_per_tap, kBK);
  p.cout = d->cout;
  p.out = d->out;
  p.out_ld = d->out_ld;
  p.scale = d->scale;
  p.bias = d->bias;
  p.residual = d->residual;
  p.res_ld = d->res_ld;
  p.relu = d->relu;
  p.a_c_off = d->a_c_off;
  p.a_n_off = d->a_n_off;
  p.b_k_off = d->b_k_off;
  p.b_n_off = d->b_n_off;
  p.out_z_off = d->out_z_off;
  p.res_z_off = d->res_z_off;
  p.bias_z_off = d->bias_z_off;
  p.splits = d->splits;
  p.partial = d->partial;

  dim3 grid(p.tiles_w * p.tiles_h * p.n_img, mega_ceil_div(d->cout, d->block_n), d->batch * d->splits);
  int rc;
  switch (d->block_n) {
    case 32: rc = launch_cfg<32, 8>(tmA, tmB, p, grid, stream); break;
    case 64: rc = launch_cfg<64, 8>(tmA, tmB, p, grid, stream); break;
    case 128: rc = launch_cfg<128, 6>(tmA, tmB, p, grid, stream); break;
    default: rc = launch_cfg<256, 4>(tmA, tmB, p, grid, stream); break;
  }
  if (rc != MEGA_OK) return rc;
  if (d->splits > 1) {
    const long long m_total = static_cast<long long>(d->n_img) * d->out_h * d->out_w;
    const long long total = m_total * d->cout;
    int blocks = static_cast<int>((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    splitk_reduce_kernel<<<blocks, 256, 0, stream>>>(d->partial, d->splits, m_total, d->cout, d->out, d->out_ld,
                                                      d->scale, d->bias, d->residual, d->res_ld, d->relu);
    MEGA_CUDA_CHECK(cudaGetLastError());
  }
  return MEGA_OK;
}
