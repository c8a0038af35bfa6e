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
constexpr int kMaxCtas = 148;   // persistent grid: one CTA per SM

struct ConvGemmParams {
  int tiles_w, tiles_h, tile_w, tile_h;
  int out_h, out_w, n_img;
  int taps_r, taps_s, dil, pad;
  int k_chunks;  // ceil(Cin / 32)
  int cout;
  const float* scale;
  const float* bias;
  int has_residual;
  int relu;
  int a_c_off, a_n_off, b_k_off, b_n_off;
  int out_c_off, out_n_off;   // per-batch coordinate offsets of the output / residual tensors
  int res_c_off, res_n_off;
  int bias_z_off;
  int box_w, box_h;           // per-warp store box: 32 output pixels = box_h x box_w
  // stream-K decomposition
  int m_tiles, n_tiles;      // per batch entry
  int kb_per_tile;           // taps * k_chunks
  long long total_units;     // batch * m_tiles * n_tiles * kb_per_tile
  long long total_tiles;     // batch * m_tiles * n_tiles
  int stream_k;              // 1: k-block granular split across CTAs, 0: whole tiles round-robin
  float* part_ws;            // [grid][2][128][BN] partial accumulators
  int* counters;             // [tiles], zero between launches
};

template <int BN, int STAGES, bool SPLIT3 = false>
struct SmemLayout {
  static constexpr int kABytes = kBM * 128;
  static constexpr int kBBytes = BN * 128;
  static constexpr int kHalf = kABytes + kBBytes;                 // 3xTF32: the low-part tiles follow at +kHalf
  static constexpr int kStageBytes = SPLIT3 ? 2 * kHalf : kHalf;
  static constexpr int kEpiOffset = STAGES * kStageBytes;       // 4 warps x (2 out + 2 residual) x 4 KB
  static constexpr int kEpiBytes = 4 * 4 * 4096;
  static constexpr int kBarOffset = kEpiOffset + kEpiBytes;
  static constexpr int kTotal = kBarOffset + (3 * STAGES + 4 + 8) * 8 + 32 + 1024;  // + align slack
};

struct TileCoord {
  int img, h0, w0, n0, batch;
};

__device__ __forceinline__ TileCoord decode_tile(const ConvGemmParams& p, long long t, int bn) {
  TileCoord c;
  const int m_tile = static_cast<int>(t % p.m_tiles);
  const long long rest = t / p.m_tiles;
  const int n_tile = static_cast<int>(rest % p.n_tiles);
  c.batch = static_cast<int>(rest / p.n_tiles);
  const int tw_i = m_tile % p.tiles_w;
  const int th_i = (m_tile / p.tiles_w) % p.tiles_h;
  c.img = m_tile / (p.tiles_w * p.tiles_h);
  c.h0 = th_i * p.tile_h;
  c.w0 = tw_i * p.tile_w;
  c.n0 = n_tile * bn;
  return c;
}

__device__ __forceinline__ long long cta_first_unit(long long total, int grid, int c) {
  return (total * c) / grid;
}

// the CTA whose unit range [first(c), first(c+1)) contains unit u
__device__ __forceinline__ int unit_owner(long long total, int grid, long long u) {
  int c = static_cast<int>((u * grid) / total);
  if (c >= grid) c = grid - 1;
  while (c + 1 < grid && cta_first_unit(total, grid, c + 1) <= u) ++c;
  while (c > 0 && cta_first_unit(total, grid, c) > u) --c;
  return c;
}

// the (tile, k-block range) items of one CTA, identical for the three warp roles
struct WorkIter {
  long long u, u_end, tile, tiles;
  int KB, grid;
  bool sk;
  __device__ __forceinline__ WorkIter(const ConvGemmParams& p, int cta, int grid_)
      : tile(cta), tiles(p.total_tiles), KB(p.kb_per_tile), grid(grid_), sk(p.stream_k != 0) {
    u = cta_first_unit(p.total_units, grid_, cta);
    u_end = cta_first_unit(p.total_units, grid_, cta + 1);
  }
  __device__ __forceinline__ bool next(long long& t, int& kb0, int& kb1) {
    if (sk) {
      if (u >= u_end) return false;
      t = u / KB;
      kb0 = static_cast<int>(u - t * KB);
      kb1 = static_cast<int>(min(static_cast<long long>(KB), kb0 + (u_end - u)));
      u += kb1 - kb0;
      return true;
    }
    if (tile >= tiles) return false;
    t = tile;
    kb0 = 0;
    kb1 = KB;
    tile += grid;
    return true;
  }
};

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// Persistent stream-K kernel. The work is the list of (tile, k-block) units, tiles ordered
// (batch, n-tile, m-tile) with m fastest; CTA c owns the contiguous unit range
// [c*U/G, (c+1)*U/G). A tile whose k-blocks straddle CTAs is finished by the last CTA to
// arrive, which sums the partial accumulators (in CTA order -> deterministic) and runs the
// epilogue. Accumulators are double-buffered in TMEM so the epilogue of item i overlaps the
// MMAs of item i+1.
// SPLIT3 ("3xTF32"): operands stay full fp32 in shared memory; four extra warps split every staged tile into
// hi = fp32 truncated to TF32 and lo = x - hi (exact), and each k-step issues hi*hi + hi*lo + lo*hi into the same
// accumulator: ~2^-19 relative error instead of 2^-11, for the strict-parity mode.
template <int BN, int STAGES, bool SPLIT3>
__global__ void __launch_bounds__(kThreads + (SPLIT3 ? 128 : 0), 1)
conv_gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmRes,
                      const ConvGemmParams p) {
  using L = SmemLayout<BN, STAGES, SPLIT3>;
  // accumulators: two ping-pong buffers (+ a master accumulator in 3xTF32 mode, see kSegLen)
  constexpr uint32_t kAccBufs = SPLIT3 ? 3 : 2;
  constexpr uint32_t kTmemCols = (kAccBufs * BN <= 64) ? 64 : (kAccBufs * BN <= 128) ? 128 : (kAccBufs * BN <= 256) ? 256 : 512;
  constexpr uint32_t kAccStride = SPLIT3 ? BN : kTmemCols / 2;
  // The tensor core adds into its fp32 accumulator with truncation, a bias that grows with the length of the
  // accumulation chain (measured ~1e-3 relative after 3000 k-blocks). The strict mode therefore restarts the
  // TMEM accumulator every kSegLen k-blocks and folds the segments into a master accumulator (also in TMEM)
  // with round-to-nearest fp32 adds done by the epilogue warps.
  constexpr int kSegLen = SPLIT3 ? 32 : 0x7fffffff;
  extern __shared__ uint8_t smem_raw[];
  // the 128B swizzle pattern is a function of the absolute smem address: align to 1024 B
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* split_bar = empty_bar + STAGES;       // [STAGES] (3xTF32 only)
  uint64_t* tmem_full_bar = split_bar + STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint64_t* res_bar = tmem_empty_bar + 2;         // [4 warps][2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 8);
  int* epi_flag = reinterpret_cast<int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int grid = gridDim.x;
  const int cta = blockIdx.x;
  const long long U = p.total_units;
  const int KB = p.kb_per_tile;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmOut);
    if (p.has_residual) prefetch_tmap(&tmRes);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&split_bar[s], 4);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], 4);
    }
    for (int b = 0; b < 8; ++b) mbar_init(&res_bar[b], 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, kTmemCols);
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
      WorkIter it(p, cta, grid);
      long long t;
      int kb0, kb1;
      while (it.next(t, kb0, kb1)) {
        const TileCoord tc = decode_tile(p, t, BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          const int tap = kb / p.k_chunks;
          const int kc = kb - tap * p.k_chunks;
          const int r = tap / p.taps_s;
          const int s = tap - r * p.taps_s;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* a_dst = smem + stage * L::kStageBytes;
          uint8_t* b_dst = a_dst + L::kABytes;
          mbar_arrive_expect_tx(&full_bar[stage], L::kHalf);   // bytes delivered by the two TMA loads
          tma_load_4d(a_dst, &tmA, &full_bar[stage], kc * kBK + tc.batch * p.a_c_off,
                      tc.w0 + s * p.dil - p.pad, tc.h0 + r * p.dil - p.pad, tc.img + tc.batch * p.a_n_off);
          tma_load_3d(b_dst, &tmB, &full_bar[stage], kc * kBK + tc.batch * p.b_k_off,
                      tc.n0 + tc.batch * p.b_n_off, tap);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc<2>(kBM, BN);
      int stage = 0;
      uint32_t phase = 0;
      WorkIter it(p, cta, grid);
      long long t;
      int kb0, kb1;
      int item = 0;
      while (it.next(t, kb0, kb1)) {
        for (int s0 = kb0, s1 = 0; s0 < kb1; s0 = s1, ++item) {
          s1 = (kb1 - s0 > kSegLen) ? s0 + kSegLen : kb1;
          const int buf = item & 1;
          const uint32_t use = static_cast<uint32_t>(item >> 1);
          mbar_wait(&tmem_empty_bar[buf], (use & 1) ^ 1);   // epilogue drained this accumulator
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + buf * kAccStride;
          for (int kb = s0; kb < s1; ++kb) {
            mbar_wait(SPLIT3 ? &split_bar[stage] : &full_bar[stage], phase);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem + stage * L::kStageBytes);
            const uint32_t b_addr = a_addr + L::kABytes;
            const uint64_t adesc = umma_desc_sw128(a_addr);
            const uint64_t bdesc = umma_desc_sw128(b_addr);
#pragma unroll
            for (int k = 0; k < kBK / kUmmaK; ++k) {
              // advance 8 floats = 32 B inside the swizzle row: +2 in 16-byte units
              umma_tf32(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > s0 || k > 0) ? 1u : 0u);
              if (SPLIT3) {
                const uint64_t alo = umma_desc_sw128(a_addr + L::kHalf), blo = umma_desc_sw128(b_addr + L::kHalf);
                umma_tf32(tmem_d, adesc + 2 * k, blo + 2 * k, idesc, 1u);
                umma_tf32(tmem_d, alo + 2 * k, bdesc + 2 * k, idesc, 1u);
              }
            }
            umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
          umma_commit(&tmem_full_bar[buf]);
        }
      }
    }
  } else if (warp >= 6) {
    // ===================== operand splitter (3xTF32 only, warps 6..9) =====================
    if (SPLIT3) {
      const int stid = threadIdx.x - kThreads;
      int stage = 0;
      uint32_t phase = 0;
      WorkIter it(p, cta, grid);
      long long t;
      int kb0, kb1;
      while (it.next(t, kb0, kb1)) {
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          uint8_t* base = smem + stage * L::kStageBytes;
          constexpr int kVecs = L::kHalf / 16;
#pragma unroll 4
          for (int v = stid; v < kVecs; v += 128) {
            const float4 x = *reinterpret_cast<const float4*>(base + v * 16);
            float4 hi, lo;
            hi.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u); lo.x = x.x - hi.x;
            hi.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u); lo.y = x.y - hi.y;
            hi.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u); lo.z = x.z - hi.z;
            hi.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u); lo.w = x.w - hi.w;
            *reinterpret_cast<float4*>(base + v * 16) = hi;
            *reinterpret_cast<float4*>(base + L::kHalf + v * 16) = lo;
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&split_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;
    const int epi_tid = (warp - 2) * 32 + lane;
    uint8_t* epi_out = smem + L::kEpiOffset + (warp - 2) * 16384;   // 2 x 4 KB store staging
    uint8_t* epi_res = epi_out + 8192;                              // 2 x 4 KB residual staging
    uint64_t* rbar = res_bar + (warp - 2) * 2;
    uint32_t rphase = 0;
    WorkIter it(p, cta, grid);
    long long t;
    int kb0, kb1;
    int item = 0;   // accumulator-segment counter (ping-pong bookkeeping shared with the MMA warp)
    const uint32_t lane_bits = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t master_row = tmem_base + 2 * kAccStride + lane_bits;
    for (int tile_item = 0; it.next(t, kb0, kb1); ++tile_item) {
      const TileCoord tc = decode_tile(p, t, BN);
      // ---- 3xTF32 only: fold every segment but the last into the master accumulator (RN fp32 adds)
      bool has_master = false;
      int s0 = kb0;
      for (; SPLIT3 && s0 + kSegLen < kb1; s0 += kSegLen, ++item) {
        const int fb = item & 1;
        mbar_wait(&tmem_full_bar[fb], static_cast<uint32_t>(item >> 1) & 1);
        tc_fence_after();
        const uint32_t seg_row = tmem_base + fb * kAccStride + lane_bits;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t a[32];
          __syncwarp();
          tmem_ld_32x32(seg_row + c * 32, a);
          tmem_ld_wait();
          if (has_master) {
            uint32_t m[32];
            tmem_ld_32x32(master_row + c * 32, m);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) a[j] = __float_as_uint(__fadd_rn(__uint_as_float(a[j]), __uint_as_float(m[j])));
          }
          tmem_st_32x32(master_row + c * 32, a);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[fb]);
        has_master = true;
      }
      const int buf = item & 1;
      const uint32_t use = static_cast<uint32_t>(item >> 1);
      ++item;
      mbar_wait(&tmem_full_bar[buf], use & 1);
      tc_fence_after();
      const uint32_t tmem_row = tmem_base + buf * kAccStride + lane_bits;
      // accumulator chunk c (32 columns of this thread's row): last segment (+ master)
      auto load_acc = [&](int c, uint32_t (&acc)[32]) {
        __syncwarp();  // tcgen05.ld is .sync.aligned
        tmem_ld_32x32(tmem_row + c * 32, acc);
        tmem_ld_wait();
        if (SPLIT3 && has_master) {
          uint32_t m[32];
          tmem_ld_32x32(master_row + c * 32, m);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] = __float_as_uint(__fadd_rn(__uint_as_float(acc[j]), __uint_as_float(m[j])));
        }
      };
      const bool complete = (kb0 == 0 && kb1 == KB);
      bool finalize = complete;
      int c_first = cta, c_last = cta;
      if (!complete) {
        // ---- publish this CTA's partial accumulator, then find out whether it arrived last
        float* my_ws = p.part_ws + ((static_cast<long long>(cta) * 2 + (tile_item == 0 ? 0 : 1)) * kBM + row) * BN;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t acc[32];
          load_acc(c, acc);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 v = make_float4(__uint_as_float(acc[j]), __uint_as_float(acc[j + 1]),
                                   __uint_as_float(acc[j + 2]), __uint_as_float(acc[j + 3]));
            __stcg(reinterpret_cast<float4*>(my_ws + c * 32 + j), v);
          }
        }
        __threadfence();
        epi_bar_sync();
        c_first = unit_owner(U, grid, t * KB);
        c_last = unit_owner(U, grid, t * KB + KB - 1);
        if (epi_tid == 0) {
          const int parts = c_last - c_first + 1;
          const int old = atomicAdd(&p.counters[t], 1);
          const int last = (old == parts - 1);
          if (last) p.counters[t] = 0;   // every part has arrived: leave the counter clean for the next launch
          *epi_flag = last;
        }
        epi_bar_sync();
        finalize = (*epi_flag != 0);
        if (finalize) __threadfence();
      }
      if (finalize) {
        // Output pixels of this warp: tile rows [32q, 32q+32) = a box_h x box_w rectangle. Results go
        // registers -> 128B-swizzled smem -> one TMA store per 32-column chunk (full-line writes,
        // image-edge and channel-edge clipping by the TMA unit); the residual arrives the same way.
        const int r0 = q * 32;
        const int bh0 = r0 / p.tile_w, bw0 = r0 - bh0 * p.tile_w;
        const int st_w = tc.w0 + bw0, st_h = tc.h0 + bh0;
        const int out_n = tc.img + tc.batch * p.out_n_off;
        const int res_n = tc.img + tc.batch * p.res_n_off;
        const float* scale_p = p.scale ? p.scale + tc.batch * p.bias_z_off : nullptr;
        const float* bias_p = p.bias ? p.bias + tc.batch * p.bias_z_off : nullptr;
        const int nchunks = min(BN / 32, (p.cout - tc.n0 + 31) / 32);
        const uint32_t sw = static_cast<uint32_t>(lane & 7);
        if (p.has_residual && lane == 0 && nchunks > 0) {
          mbar_arrive_expect_tx(&rbar[0], 4096);
          tma_load_4d(epi_res, &tmRes, &rbar[0], tc.n0 + tc.batch * p.res_c_off, st_w, st_h, res_n);
        }
#pragma unroll 1
        for (int c = 0; c < nchunks; ++c) {
          uint32_t acc[32];
          load_acc(c, acc);
          const int nb = tc.n0 + c * 32;
          if (!complete) {
            // deterministic reduction: parts summed in CTA order, own part from TMEM
            float sum[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) sum[j] = 0.f;
            for (int oc = c_first; oc <= c_last; ++oc) {
              if (oc == cta) {
#pragma unroll
                for (int j = 0; j < 32; ++j) sum[j] += __uint_as_float(acc[j]);
              } else {
                const int slot = (cta_first_unit(U, grid, oc) >= t * KB) ? 0 : 1;
                const float* ws = p.part_ws + ((static_cast<long long>(oc) * 2 + slot) * kBM + row) * BN + c * 32;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  const float4 v = __ldcg(reinterpret_cast<const float4*>(ws + j));
                  sum[j] += v.x; sum[j + 1] += v.y; sum[j + 2] += v.z; sum[j + 3] += v.w;
                }
              }
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = __float_as_uint(sum[j]);
          }
          const uint8_t* rsrc = nullptr;
          if (p.has_residual) {
            const int rb = c & 1;
            if (c + 1 < nchunks && lane == 0) {   // prefetch the next residual chunk into the other buffer
              mbar_arrive_expect_tx(&rbar[rb ^ 1], 4096);
              tma_load_4d(epi_res + (rb ^ 1) * 4096, &tmRes, &rbar[rb ^ 1], nb + 32 + tc.batch * p.res_c_off, st_w,
                          st_h, res_n);
            }
            mbar_wait(&rbar[rb], (rphase >> rb) & 1u);
            rphase ^= (1u << rb);
            rsrc = epi_res + rb * 4096 + lane * 128;
          }
          // the out staging buffer (c & 1) was handed to a TMA store two chunks ago: wait until read
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
          uint8_t* dst = epi_out + (c & 1) * 4096 + lane * 128;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const int n = nb + j;
            float4 v = make_float4(__uint_as_float(acc[j]), __uint_as_float(acc[j + 1]),
                                   __uint_as_float(acc[j + 2]), __uint_as_float(acc[j + 3]));
            if (n < p.cout) {   // cout is padded to 4 by the host wrapper's buffers; tail lanes are clipped by TMA
              if (scale_p) {
                const float4 sc = ldg_f4(scale_p + n);
                v.x *= sc.x; v.y *= sc.y; v.z *= sc.z; v.w *= sc.w;
              }
              if (bias_p) {
                const float4 bi = ldg_f4(bias_p + n);
                v.x += bi.x; v.y += bi.y; v.z += bi.z; v.w += bi.w;
              }
            }
            const uint32_t chunk = (static_cast<uint32_t>(j >> 2) ^ sw) << 4;
            if (rsrc) {
              const float4 rr = *reinterpret_cast<const float4*>(rsrc + chunk);
              v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
            }
            if (p.relu) {
              v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
            }
            *reinterpret_cast<float4*>(dst + chunk) = v;
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&tmOut, epi_out + (c & 1) * 4096, nb + tc.batch * p.out_c_off, st_w, st_h, out_n);
            tma_store_commit();
          }
        }
      }
      // release the accumulator buffer to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
    }
    if (lane == 0) tma_store_wait<0>();   // global writes complete before the CTA retires
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
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

static int g_num_sms = 0;
constexpr int kCounterSlots = 65536;   // ints at the head of the workspace
constexpr int kMinUnits = 4;

template <int BN, int STAGES, bool SPLIT3 = false>
static int launch_cfg(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut,
                      const CUtensorMap& tmRes, const ConvGemmParams& p, dim3 grid, cudaStream_t stream) {
  using L = SmemLayout<BN, STAGES, SPLIT3>;
  static bool configured = false;
  if (!configured) {
    MEGA_CUDA_CHECK(cudaFuncSetAttribute(conv_gemm_tf32_kernel<BN, STAGES, SPLIT3>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    configured = true;
  }
  conv_gemm_tf32_kernel<BN, STAGES, SPLIT3><<<grid, kThreads + (SPLIT3 ? 128 : 0), L::kTotal, stream>>>(tmA, tmB, tmOut,
                                                                                                      tmRes, p);
  MEGA_CUDA_CHECK(cudaGetLastError());
  return MEGA_OK;
}

}  // namespace mega

using namespace mega;

extern "C" long long mega_conv_gemm_workspace_bytes(void) {
  return static_cast<long long>(kCounterSlots) * sizeof(int) +
         static_cast<long long>(kMaxCtas) * 2 * kBM * 256 * sizeof(float);
}

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
  MEGA_ARG_CHECK(d->block_n == 32 || d->block_n == 64 || d->block_n == 96 || d->block_n == 128 ||
                     d->block_n == 160 || d->block_n == 192 || d->block_n == 256,
                 "conv_gemm: block_n must be one of 32/64/96/128/160/192/256");
  MEGA_ARG_CHECK((reinterpret_cast<uintptr_t>(d->a) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->b) & 15) == 0,
                 "conv_gemm: operand base pointers must be 16-byte aligned");
  MEGA_ARG_CHECK((d->a_stride_w % 4) == 0 && (d->a_stride_h % 4) == 0 && (d->a_stride_n % 4) == 0,
                 "conv_gemm: activation strides must be multiples of 4 floats");
  MEGA_ARG_CHECK(d->out != nullptr && (reinterpret_cast<uintptr_t>(d->out) & 15) == 0 && (d->out_ld % 4) == 0,
                 "conv_gemm: output must be 16-byte aligned with a row pitch multiple of 4 floats");
  MEGA_ARG_CHECK(d->residual == nullptr ||
                     ((reinterpret_cast<uintptr_t>(d->residual) & 15) == 0 && (d->res_ld % 4) == 0),
                 "conv_gemm: residual must be 16-byte aligned with a row pitch multiple of 4 floats");
  MEGA_ARG_CHECK(d->out_c_off == 0 || d->cout == d->block_n,
                 "conv_gemm: channel-offset batching needs cout == block_n (got %d vs %d)", d->cout, d->block_n);
  MEGA_ARG_CHECK((d->b_stride_n % 4) == 0 && (d->b_stride_tap % 4) == 0,
                 "conv_gemm: weight strides must be multiples of 4 floats");
  MEGA_ARG_CHECK(d->batch >= 1, "conv_gemm: batch must be >= 1");
  MEGA_ARG_CHECK(d->workspace != nullptr && d->workspace_bytes >= mega_conv_gemm_workspace_bytes(),
                 "conv_gemm: workspace missing or smaller than mega_conv_gemm_workspace_bytes()");
  MEGA_ARG_CHECK((reinterpret_cast<uintptr_t>(d->workspace) & 255) == 0, "conv_gemm: workspace must be 256-byte aligned");
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    mega_set_error("conv_gemm: cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return MEGA_ERR_CUDA;
  }
  const bool strict = d->precision == 1;
  MEGA_ARG_CHECK(d->precision == 0 || d->precision == 1, "conv_gemm: precision must be 0 (tf32) or 1 (3xtf32)");
  MEGA_ARG_CHECK(!strict || d->block_n == 64 || d->block_n == 128, "conv_gemm: 3xtf32 supports block_n 64 / 128");
  const CUtensorMapDataType dt =
      (g_tf32_round && !strict) ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;

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

  // per-warp store / residual boxes: 32 channels x (box_h x box_w = 32 output pixels), 128B swizzle
  CUtensorMap tmOut, tmRes;
  {
    const int box_w = d->tile_w < 32 ? d->tile_w : 32;
    const int box_h = 32 / box_w;
    for (int which = 0; which < 2; ++which) {
      const float* base = which == 0 ? d->out : d->residual;
      CUtensorMap* tm = which == 0 ? &tmOut : &tmRes;
      if (base == nullptr) {
        *tm = tmOut;
        continue;
      }
      const long long ld = which == 0 ? d->out_ld : d->res_ld;
      const int c_off = which == 0 ? d->out_c_off : d->res_c_off;
      const int n_off = which == 0 ? d->out_n_off : d->res_n_off;
      cuuint64_t gdim[4] = {static_cast<cuuint64_t>(d->cout + (d->batch - 1) * c_off), static_cast<cuuint64_t>(d->out_w),
                            static_cast<cuuint64_t>(d->out_h),
                            static_cast<cuuint64_t>(d->n_img + (d->batch - 1) * n_off)};
      cuuint64_t gstr[3] = {static_cast<cuuint64_t>(ld) * 4, static_cast<cuuint64_t>(ld) * d->out_w * 4,
                            static_cast<cuuint64_t>(ld) * d->out_w * d->out_h * 4};
      cuuint32_t box[4] = {32, static_cast<cuuint32_t>(box_w), static_cast<cuuint32_t>(box_h), 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), gdim, gstr, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        mega_set_error("conv_gemm: encode %s tensor map failed (CUresult %d): C %d W %d H %d N %d ld %lld",
                       which == 0 ? "output" : "residual", static_cast<int>(r), static_cast<int>(gdim[0]), d->out_w,
                       d->out_h, static_cast<int>(gdim[3]), ld);
        return MEGA_ERR_CUDA;
      }
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
  p.scale = d->scale;
  p.bias = d->bias;
  p.has_residual = d->residual != nullptr;
  p.relu = d->relu;
  p.a_c_off = d->a_c_off;
  p.a_n_off = d->a_n_off;
  p.b_k_off = d->b_k_off;
  p.b_n_off = d->b_n_off;
  p.out_c_off = d->out_c_off;
  p.out_n_off = d->out_n_off;
  p.res_c_off = d->res_c_off;
  p.res_n_off = d->res_n_off;
  p.bias_z_off = d->bias_z_off;
  p.box_w = d->tile_w < 32 ? d->tile_w : 32;
  p.box_h = 32 / p.box_w;
  p.m_tiles = p.tiles_w * p.tiles_h * p.n_img;
  p.n_tiles = mega_ceil_div(d->cout, d->block_n);
  p.kb_per_tile = d->taps_r * d->taps_s * p.k_chunks;
  const long long tiles = static_cast<long long>(d->batch) * p.m_tiles * p.n_tiles;
  p.total_units = tiles * p.kb_per_tile;
  p.total_tiles = tiles;
  p.stream_k = d->stream_k ? 1 : 0;
  MEGA_ARG_CHECK(tiles <= kCounterSlots, "conv_gemm: %lld output tiles exceed the %d counter slots", tiles, kCounterSlots);
  MEGA_ARG_CHECK(p.total_units > 0, "conv_gemm: empty problem");
  p.counters = static_cast<int*>(d->workspace);
  p.part_ws = reinterpret_cast<float*>(static_cast<char*>(d->workspace) + kCounterSlots * sizeof(int));
  if (g_num_sms == 0) {
    int dev = 0;
    MEGA_CUDA_CHECK(cudaGetDevice(&dev));
    MEGA_CUDA_CHECK(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    if (g_num_sms > kMaxCtas) g_num_sms = kMaxCtas;
  }
  // persistent grid: one CTA per SM, but never fewer than kMinUnits k-blocks of work per CTA
  long long ctas = p.stream_k ? p.total_units / kMinUnits : tiles;
  if (ctas < 1) ctas = 1;
  if (ctas > g_num_sms) ctas = g_num_sms;
  if (d->max_ctas > 0 && ctas > d->max_ctas) ctas = d->max_ctas;
  dim3 grid(static_cast<unsigned>(ctas), 1, 1);
  int rc;
  if (strict) {
    rc = d->block_n == 64 ? launch_cfg<64, 3, true>(tmA, tmB, tmOut, tmRes, p, grid, stream)
                          : launch_cfg<128, 2, true>(tmA, tmB, tmOut, tmRes, p, grid, stream);
    return rc;
  }
  switch (d->block_n) {
    case 32: rc = launch_cfg<32, 6>(tmA, tmB, tmOut, tmRes, p, grid, stream); break;
    case 64: rc = launch_cfg<64, 6>(tmA, tmB, tmOut, tmRes, p, grid, stream); break;
    case 96: rc = launch_cfg<96, 5>(tmA, tmB, tmOut, tmRes, p, grid, stream); break;
    case 128: rc = launch_cfg<128, 4>(tmA, tmB, tmOut, tmRes, p, grid, stream); break;
    case 160: rc = launch_cfg<160, 4>(tmA, tmB, tmOut, tmRes, p, grid, stream); break;
    case 192: rc = launch_cfg<192, 3>(tmA, tmB, tmOut, tmRes, p, grid, stream); break;
    default: rc = launch_cfg<256, 3>(tmA, tmB, tmOut, tmRes, p, grid, stream); break;
  }
  return rc;
}

