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
// warps 2-5 epilogue (TMEM -> registers -> global); the strict modes add warps 6-9: operand splitters (3xTF32) or a second
// set of epilogue warps (3xFP16 on split-fp16 tensors, the mode the strict engine runs; see kModeF16x3 below).
#include "common.cuh"
#pragma once
#include "mega_b200.h"

namespace mega {

constexpr int kBM = 128;        // UMMA M (one CTA)
// operand arithmetic of a launch
constexpr int kModeTf32 = 0;    // fp32 operands in HBM, rounded to TF32 by the TMA load; K slab = 32 floats
constexpr int kModeSplit3 = 1;  // "3xTF32": fp32 operands split hi/lo in shared memory
constexpr int kModeF16 = 2;     // fp16 operands in HBM (10-bit mantissa like TF32, half the bytes, twice the
                                // tensor-pipe rate); K slab = 64 halves
constexpr int kModeF16x3 = 3;   // "3xFP16": operands stored SPLIT in HBM -- every group of 32 K-values is one 128-byte row
                                // [32 hi halves | 32 lo halves] with hi = fp16(x), lo = fp16(x - hi): the same 4 bytes per
                                // value as fp32 and 22 mantissa bits like 3xTF32, but NO split work in the kernel (the
                                // producing epilogue / the weight packer did it) and kind::f16 MMAs; K slab = 32 values
// every mode stages K slabs of 128 bytes per row (one swizzle row) and issues 4 MMAs of 32 bytes of K each
// (3xFP16: 2 k-steps x 3 products over the hi / lo halves of the row)
__host__ __device__ constexpr int mode_bk(int mode) { return mode == kModeF16 ? 64 : 32; }
constexpr int kThreads = 192;   // 6 warps
constexpr int kMaxCtas = 148;   // persistent grid: one CTA per SM

struct ConvGemmParams {
  int tiles_w, tiles_h, tile_w, tile_h;
  int out_h, out_w, n_img;
  int taps_r, taps_s, dil, pad;
  int pad_w;                  // left padding (pad applies to the rows)
  int stride_h, stride_w;     // convolution stride: output (h, w) reads input (h*stride_h + r*dil - pad, ...)
  int k_chunks;  // ceil(Cin / BK)
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
  int seg_len;               // 3xTF32 only: k-blocks accumulated in TMEM before the RN fold into the master accumulator
  int b_lo_tap_off;          // 3xTF32 only: > 0: B's low parts are stored as taps [b_lo_tap_off, 2 * b_lo_tap_off) of the B tensor
  int a_tmem;                // 3xFP16 only: 1 = the MMA warp copies every staged A tile into tensor memory (tcgen05.cp) and issues
                             // the MMAs in the TS form (A from TMEM, only B fetched from shared memory)
  int res_split;             // 3xFP16 only: the residual tensor is in the split-fp16 format (else fp32)
  float acc_scale;           // 3xFP16 only: the accumulator is multiplied by this power of two first (weights are stored
                             // scaled by its inverse so that their low halves stay normal fp16 numbers)
};

template <int BN, int STAGES, int MODE = kModeTf32>
struct SmemLayout {
  static constexpr bool SPLIT3 = MODE == kModeSplit3;
  static constexpr int kABytes = kBM * 128;
  static constexpr int kBBytes = BN * 128;
  static constexpr int kHalf = kABytes + kBBytes;                 // bytes the two TMA loads of a k-block deliver
  // 3xTF32: [A raw | B raw (= the hi operand: the MMA ignores the low 13 bits) | B lo]; A's hi / lo parts live in TENSOR memory
  static constexpr int kStageBytes = SPLIT3 ? kHalf + kBBytes : kHalf;
  static constexpr int kEpiOffset = STAGES * kStageBytes;       // 4 warps x (2 out + 2 residual) x 4 KB
  static constexpr int kEpiBytes = 4 * 4 * 4096;
  static constexpr int kBarOffset = kEpiOffset + kEpiBytes;
  static constexpr int kSbCols = BN <= 128 ? 128 : 256;
  static constexpr int kSbOffset = kBarOffset + 512;              // [scale | bias][kSbCols] floats of the tile being finished
  static constexpr int kTotal = kSbOffset + 8 * kSbCols + 1024;   // + align slack
};

struct TileCoord {
  int img, h0, w0, n0, batch;
};

// Work-list indices are 32-bit (the host checks total_units * grid < 2^31): 64-bit divisions cost hundreds of cycles
// on the single-thread critical paths of the producer / MMA roles.
__device__ __forceinline__ TileCoord decode_tile(const ConvGemmParams& p, int t, int bn) {
  TileCoord c;
  const int m_tile = t % p.m_tiles;
  const int rest = t / p.m_tiles;
  const int n_tile = rest % p.n_tiles;
  c.batch = rest / p.n_tiles;
  const int tw_i = m_tile % p.tiles_w;
  const int th_i = (m_tile / p.tiles_w) % p.tiles_h;
  c.img = m_tile / (p.tiles_w * p.tiles_h);
  c.h0 = th_i * p.tile_h;
  c.w0 = tw_i * p.tile_w;
  c.n0 = n_tile * bn;
  return c;
}

__device__ __forceinline__ int cta_first_unit(int total, int grid, int c) {
  return static_cast<int>((static_cast<unsigned>(total) * static_cast<unsigned>(c)) / static_cast<unsigned>(grid));
}

// the CTA whose unit range [first(c), first(c+1)) contains unit u
__device__ __forceinline__ int unit_owner(int total, int grid, int u) {
  int c = static_cast<int>((static_cast<unsigned>(u) * static_cast<unsigned>(grid)) / static_cast<unsigned>(total));
  if (c >= grid) c = grid - 1;
  while (c + 1 < grid && cta_first_unit(total, grid, c + 1) <= u) ++c;
  while (c > 0 && cta_first_unit(total, grid, c) > u) --c;
  return c;
}

// the (tile, k-block range) items of one CTA, identical for the three warp roles
struct WorkIter {
  int u, u_end, tile, tiles;
  int KB, grid;
  bool sk;
  __device__ __forceinline__ WorkIter(const ConvGemmParams& p, int cta, int grid_)
      : tile(cta), tiles(static_cast<int>(p.total_tiles)), KB(p.kb_per_tile), grid(grid_), sk(p.stream_k != 0) {
    u = cta_first_unit(static_cast<int>(p.total_units), grid_, cta);
    u_end = cta_first_unit(static_cast<int>(p.total_units), grid_, cta + 1);
  }
  __device__ __forceinline__ bool next(int& t, int& kb0, int& kb1) {
    if (sk) {
      if (u >= u_end) return false;
      t = u / KB;
      kb0 = u - t * KB;
      kb1 = min(KB, kb0 + (u_end - u));
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
__device__ __forceinline__ void epi_bar_sync8() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // 8 epilogue warps

// Persistent stream-K kernel. The work is the list of (tile, k-block) units, tiles ordered
// (batch, n-tile, m-tile) with m fastest; CTA c owns the contiguous unit range
// [c*U/G, (c+1)*U/G). A tile whose k-blocks straddle CTAs is finished by the last CTA to
// arrive, which sums the partial accumulators (in CTA order -> deterministic) and runs the
// epilogue. Accumulators are double-buffered in TMEM so the epilogue of item i overlaps the
// MMAs of item i+1.
// SPLIT3 ("3xTF32"): operands arrive as full fp32; four extra warps split every staged tile into hi = fp32 truncated to
// TF32 and lo = x - hi (exact), and each k-step issues hi*hi + hi*lo + lo*hi into the same accumulator: ~2^-19 relative
// error instead of 2^-11, for the strict-parity mode. Round 1 kept all four split tiles in shared memory (224 KB of smem
// traffic per 128 x 128 x 32 k-block: TMA 32 + split 96 + twelve MMAs reading both operands 96), which bound the mode at
// ~100 TFLOP/s. Now the A operand is a TENSOR-MEMORY operand (tcgen05.mma "TS" form): a splitter thread reads its row of
// the raw A tile once (8 x LDS.128) and writes hi / lo straight into two 32-column TMEM slabs with tcgen05.st; the MMAs
// read A from TMEM and only B (the raw tile as hi, lo beside it) from shared memory: 128 KB per k-block, and the stage shrinks
// from 64 to 48 KB (3 stages instead of 2 at block_n 128).
// OUT16: output (and residual) tensors are fp16; the epilogue then works in chunks of 64 columns (= one 128-byte
// swizzle row of halves) instead of 32.
// Programmatic dependent launch: the prologue (barrier init, TMEM allocation, descriptor prefetch) runs before
// griddepcontrol.wait, i.e. overlapped with the tail of the previous kernel on the stream; nothing before the wait
// touches global memory.
template <int BN, int STAGES, int MODE, bool OUT16>
__global__ void __launch_bounds__(kThreads + ((MODE == kModeSplit3 || MODE == kModeF16x3) ? 128 : 0), 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmRes,
                      const ConvGemmParams p) {
  constexpr bool SPLIT3 = MODE == kModeSplit3;
  constexpr bool PK = MODE == kModeF16x3;       // split-fp16 operands; OUT16 then means "split-fp16 output" (else fp32)
  constexpr bool SEG = SPLIT3 || PK;            // segmented accumulation with a round-to-nearest master accumulator
  constexpr bool OUTH = OUT16 && !PK;           // plain fp16 output
  constexpr int kBK = mode_bk(MODE);
  constexpr int CW = OUTH ? 64 : 32;   // epilogue chunk: columns per 128-byte output row segment
  static_assert(!OUTH || BN % 64 == 0, "fp16 output needs block_n % 64 == 0");
  static_assert(SmemLayout<BN, STAGES, MODE>::kTotal <= 227 * 1024, "pipeline + staging exceed the 227 KB of a CTA");
  using L = SmemLayout<BN, STAGES, MODE>;
  // accumulators: two ping-pong buffers (+ a master accumulator in 3xTF32 mode, see kSegLen)
  // 3xFP16: eight epilogue warps, the master accumulator in their REGISTERS (64 columns per thread), see the epilogue below
  constexpr bool REGM = PK;
  constexpr uint32_t kAccBufs = (SEG && !REGM) ? 3 : 2;
  // 3xTF32: three accumulators + two A slabs of 64 columns (hi: 32 columns of K, lo: the next 32)
  constexpr uint32_t kNeedCols = kAccBufs * BN + (SPLIT3 ? 128 : 0) + (PK ? 64 : 0);   // 3xFP16: two 32-column A slabs
  static_assert(kNeedCols <= 512, "accumulators + A slabs exceed the 512 TMEM columns");
  constexpr uint32_t kTmemCols = (kNeedCols <= 64) ? 64 : (kNeedCols <= 128) ? 128 : (kNeedCols <= 256) ? 256 : 512;
  constexpr uint32_t kAccStride = (SEG && !REGM) ? BN : kTmemCols / 2;
  constexpr uint32_t kASlab = 3 * BN;        // first column of A slab 0 (3xTF32); slab s at + 64 s
  constexpr uint32_t kASlabPk = kTmemCols - 64;   // 3xFP16 (a_tmem): slab s = 32 columns at + 32 s: [hi k0 | hi k1 | lo k0 | lo k1] x 8
  // The tensor core adds into its fp32 accumulator with truncation, a bias that grows with the length of the
  // accumulation chain (measured ~1e-3 relative after 3000 k-blocks; ~2e-5 after 32, which the chaotic position
  // embedding of the relation module amplifies to 5e-3 on the final logits). The strict mode therefore restarts the
  // TMEM accumulator every kSegLen k-blocks (4 k-blocks = 48 truncating adds, <= 3e-6 relative) and folds the segments into a master accumulator (also in TMEM)
  // with round-to-nearest fp32 adds done by the epilogue warps.
  const int kSegLen = SEG ? p.seg_len : 0x7fffffff;   // k-blocks per accumulator segment (mega_set_split3_seg_len, default 4)
  extern __shared__ uint8_t smem_raw[];
  // the 128B swizzle pattern is a function of the absolute smem address: align to 1024 B
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* split_bar = empty_bar + STAGES;       // [STAGES] (3xTF32 only)
  uint64_t* tmem_full_bar = split_bar + STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint64_t* res_bar = tmem_empty_bar + 2;         // [4 warps][2] (3xFP16: [8 warps][2])
  uint64_t* aslab_empty_bar = res_bar + 16;        // [2] (3xTF32: the MMAs that read A slab s have completed)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aslab_empty_bar + 2);
  int* epi_flag = reinterpret_cast<int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int grid = gridDim.x;
  const int cta = blockIdx.x;
  const int U = static_cast<int>(p.total_units);
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
      mbar_init(&tmem_empty_bar[b], PK ? 8 : 4);
    }
    for (int b = 0; b < 16; ++b) mbar_init(&res_bar[b], 1);
    for (int b = 0; b < 2; ++b) mbar_init(&aslab_empty_bar[b], 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();               // the previous kernel's results (and its reads of our outputs) are complete
  griddep_launch_dependents();  // let the next kernel's prologue overlap this kernel

  if (warp == 0) {
    // ===================== TMA producer =====================
    // lane 0 loads the A (activation) tile and posts the expected byte count, lane 1 the B (weight) tile, so the two
    // descriptor-based copies of a k-block are issued in parallel; tap / chunk indices advance incrementally
    if (lane < 2) {
      int stage = 0;
      uint32_t phase = 0;
      WorkIter it(p, cta, grid);
      int t;
      int kb0, kb1;
      const int k_chunks = p.k_chunks, taps_s = p.taps_s;
      while (it.next(t, kb0, kb1)) {
        const TileCoord tc = decode_tile(p, t, BN);
        int tap = kb0 / k_chunks;
        int kc = kb0 - tap * k_chunks;
        int r = tap / taps_s;
        int sx = tap - r * taps_s;
        const int a_c0 = tc.batch * p.a_c_off, a_n = tc.img + tc.batch * p.a_n_off;
        const int b_k0 = tc.batch * p.b_k_off, b_n = tc.n0 + tc.batch * p.b_n_off;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* a_dst = smem + stage * L::kStageBytes;
          const bool b_lo_ready = SPLIT3 && p.b_lo_tap_off > 0;    // pre-split weights: the low parts come by TMA too
          if (lane == 0) {
            mbar_arrive_expect_tx(&full_bar[stage], L::kHalf + (b_lo_ready ? L::kBBytes : 0));   // bytes the TMA loads deliver
            tma_load_4d(a_dst, &tmA, &full_bar[stage], kc * kBK + a_c0, tc.w0 * p.stride_w + sx * p.dil - p.pad_w,
                        tc.h0 * p.stride_h + r * p.dil - p.pad, a_n);
          } else {
            tma_load_3d(a_dst + L::kABytes, &tmB, &full_bar[stage], kc * kBK + b_k0, b_n, tap);
            if (b_lo_ready)
              tma_load_3d(a_dst + L::kABytes + L::kBBytes, &tmB, &full_bar[stage], kc * kBK + b_k0, b_n, tap + p.b_lo_tap_off);
          }
          if (++kc == k_chunks) {
            kc = 0;
            ++tap;
            if (++sx == taps_s) {
              sx = 0;
              ++r;
            }
          }
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
      const uint32_t idesc = (MODE == kModeF16 || PK) ? umma_idesc<0>(kBM, BN) : umma_idesc<2>(kBM, BN);
      int stage = 0;
      uint32_t phase = 0;
      WorkIter it(p, cta, grid);
      int t;
      int kb0, kb1;
      int item = 0;
      uint32_t kbn = 0;            // k-blocks issued by this CTA (3xTF32: A slab = kbn & 1)
      while (it.next(t, kb0, kb1)) {
        for (int s0 = kb0, s1 = 0; s0 < kb1; s0 = s1, ++item) {
          s1 = (kb1 - s0 > kSegLen) ? s0 + kSegLen : kb1;
          const int buf = item & 1;
          const uint32_t use = static_cast<uint32_t>(item >> 1);
          mbar_wait(&tmem_empty_bar[buf], (use & 1) ^ 1);   // epilogue drained this accumulator
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + buf * kAccStride;
          for (int kb = s0; kb < s1; ++kb, ++kbn) {
            mbar_wait(SPLIT3 ? &split_bar[stage] : &full_bar[stage], phase);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem + stage * L::kStageBytes);
            const uint32_t b_addr = a_addr + L::kABytes;
            const uint64_t adesc = umma_desc_sw128(a_addr);
            const uint64_t bdesc = umma_desc_sw128(b_addr);
            if (SPLIT3) {
              // A from tensor memory: slab (kbn & 1), hi in its columns [0, 32), lo in [32, 64); 8 columns of K per MMA
              const uint32_t a_hi = tmem_base + kASlab + (kbn & 1u) * 64u, a_lo = a_hi + 32u;
              const uint64_t blo = umma_desc_sw128(b_addr + L::kBBytes);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_tf32_ts(tmem_d, a_hi + 8 * k, bdesc + 2 * k, idesc, (kb > s0 || k > 0) ? 1u : 0u);
                umma_tf32_ts(tmem_d, a_hi + 8 * k, blo + 2 * k, idesc, 1u);
                umma_tf32_ts(tmem_d, a_lo + 8 * k, bdesc + 2 * k, idesc, 1u);
              }
              umma_commit(&aslab_empty_bar[kbn & 1u]);   // the splitter may overwrite this A slab
            } else if (PK && p.a_tmem) {
              // TS form: the four 128 x 32-byte slices of the A tile (hi / lo of the two k-steps) go to a tensor-memory slab by
              // tcgen05.cp (same descriptors the SS-form MMAs would read them through; cp and mma issued by one thread
              // execute in order), then the MMAs fetch only B from shared memory: 24 KB of operand reads per k-block
              // instead of 48 KB (+ 16 KB read once by the copies)
              const uint32_t a_tm = tmem_base + kASlabPk + (kbn & 1u) * 32u;
#pragma unroll
              for (int c = 0; c < 4; ++c) tmem_cp_128x256b(a_tm + 8 * c, adesc + 2 * c);
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                umma_f16_ts(tmem_d, a_tm + 8 * k, bdesc + 2 * k, idesc, (kb > s0 || k > 0) ? 1u : 0u);
                umma_f16_ts(tmem_d, a_tm + 8 * k, bdesc + 4 + 2 * k, idesc, 1u);
                umma_f16_ts(tmem_d, a_tm + 16 + 8 * k, bdesc + 2 * k, idesc, 1u);
              }
            } else if (PK) {
              // a staged row = [32 hi halves | 32 lo halves] of 32 K-values: k-step j (16 values) reads hi at byte 32 j and
              // lo at byte 64 + 32 j of the swizzle row (descriptor units of 16 bytes); hi.hi + hi.lo + lo.hi
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                umma_f16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > s0 || k > 0) ? 1u : 0u);
                umma_f16(tmem_d, adesc + 2 * k, bdesc + 4 + 2 * k, idesc, 1u);
                umma_f16(tmem_d, adesc + 4 + 2 * k, bdesc + 2 * k, idesc, 1u);
              }
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                // advance 32 B (8 floats / 16 halves) inside the swizzle row: +2 in 16-byte units
                if (MODE == kModeF16) {
                  umma_f16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > s0 || k > 0) ? 1u : 0u);
                } else {
                  umma_tf32(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > s0 || k > 0) ? 1u : 0u);
                }
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
  } else if (warp >= 6 && !PK) {
    // ===================== operand splitter (3xTF32 only, warps 6..9) =====================
    if (SPLIT3) {
      const int stid = threadIdx.x - kThreads;            // 0..127
      const int arow = (warp & 3) * 32 + lane;            // the A-tile row = TMEM lane this thread may write
      const uint32_t lane_bits = static_cast<uint32_t>((warp & 3) * 32) << 16;
      int stage = 0;
      uint32_t phase = 0;
      uint32_t kbn = 0;
      WorkIter it(p, cta, grid);
      int t;
      int kb0, kb1;
      while (it.next(t, kb0, kb1)) {
        for (int kb = kb0; kb < kb1; ++kb, ++kbn) {
          mbar_wait(&full_bar[stage], phase);
          uint8_t* base = smem + stage * L::kStageBytes;
          // ---- A: this thread's row (32 floats = 128 bytes, 16-byte chunks swizzled by row & 7) -> hi / lo -> TMEM
          {
            const uint8_t* rowp = base + arow * 128;
            uint32_t hi[32], lo[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 x = *reinterpret_cast<const float4*>(rowp + ((j ^ (arow & 7)) << 4));
              const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const uint32_t h = __float_as_uint(xs[e]) & 0xffffe000u;
                hi[4 * j + e] = h;
                lo[4 * j + e] = __float_as_uint(xs[e] - __uint_as_float(h));
              }
            }
            const uint32_t slab = kbn & 1u;
            // the MMAs of k-block kbn - 2 (the previous user of this slab) have completed
            if (kbn >= 2) mbar_wait(&aslab_empty_bar[slab], ((kbn >> 1) - 1) & 1u);
            tc_fence_after();
            const uint32_t ta = tmem_base + kASlab + slab * 64u + lane_bits;
            __syncwarp();
            tmem_st_32x32(ta, hi);
            tmem_st_32x32(ta + 32u, lo);
          }
          // ---- B: lo = x - trunc_tf32(x) into the region behind the raw tile. The raw tile itself serves as the hi
          //      operand: kind::tf32 reads the upper 19 bits of a 32-bit container and IGNORES the low 13 mantissa bits
          //      (truncation, not rounding -- measured with tools/tf32_trunc_probe.py: (1 + 0.75 * 2^-10) * 1 = 1.0 on
          //      both operand sides), so writing the masked copy back would only cost shared-memory bandwidth
          if (p.b_lo_tap_off == 0) {      // (pre-split weights: the producer fetched lo by TMA, nothing to do for B)
            uint8_t* bb = base + L::kABytes;
            constexpr int kVecs = L::kBBytes / 16;
#pragma unroll 4
            for (int v = stid; v < kVecs; v += 128) {
              const float4 x = *reinterpret_cast<const float4*>(bb + v * 16);
              float4 l;
              l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
              l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
              l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
              l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
              *reinterpret_cast<float4*>(bb + L::kBBytes + v * 16) = l;
            }
          }
          tmem_st_wait();
          tc_fence_before();
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
  } else if constexpr (PK) {
    // ===================== epilogue, 3xFP16 (warps 2..9) =====================
    // Two warps share a TMEM lane quarter (32 tile rows); each owns HALF of the tile's columns: kCPW chunks of 32. What
    // the 4-warp epilogue above measured on this mode (10-13 us per 128 x 128 tile at K = 256 against 2.7 us of MMAs) came
    // from (i) the RN folds -- segment + master read from tensor memory (64 B/cycle per SM) and written back, 0.85-1.2 us
    // each: here the master lives in REGISTERS (64 columns per thread) and a fold reads the segment once; (ii) two
    // CTA-wide barriers around L1-missing scale / bias loads per tile: here the BN scale is folded into the packed
    // weights, the bias slice of the NEXT tile is fetched during the current one (double-buffered, one barrier per tile);
    // (iii) separate residual staging: here the residual lands in the store staging buffer itself (every thread reads its
    // 128-byte row into registers before it writes the same row), issued for both chunks at the start of the tile.
    constexpr int kCPW = BN / 64;          // 32-column chunks per warp
    const int q = warp & 3;                // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;      // which half of the columns
    const int row = q * 32 + lane;
    const int epi_tid = (warp - 2) * 32 + lane;                              // 0 .. 255
    uint8_t* stage_buf = smem + L::kEpiOffset + (warp - 2) * (kCPW * 4096);  // kCPW x 4 KB: residual in, result out
    uint64_t* rbar = res_bar + (warp - 2) * 2;
    float* bias_s = reinterpret_cast<float*>(smem + L::kSbOffset);           // [2][BN]
    uint32_t rphase = 0;
    WorkIter it(p, cta, grid);
    int t;
    int kb0, kb1;
    int item = 0;
    const uint32_t lane_bits = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t col_base = static_cast<uint32_t>(half * kCPW * 32);
    const uint32_t sw = static_cast<uint32_t>(lane & 7);
    const bool res_split = p.res_split != 0;
    const float slope = p.relu == 2 ? 0.1f : 0.f;
    float master[kCPW * 32];
    for (int tile_item = 0; it.next(t, kb0, kb1); ++tile_item) {
      const TileCoord tc = decode_tile(p, t, BN);
      // ---- fold every segment but the last into the register master (round-to-nearest fp32 adds)
      bool has_master = false;
      int s0 = kb0;
      for (; s0 + kSegLen < kb1; s0 += kSegLen, ++item) {
        const int fb = item & 1;
        mbar_wait(&tmem_full_bar[fb], static_cast<uint32_t>(item >> 1) & 1);
        tc_fence_after();
        const uint32_t seg_row = tmem_base + fb * kAccStride + lane_bits + col_base;
#pragma unroll
        for (int j = 0; j < kCPW; ++j) {
          uint32_t a[32];
          __syncwarp();
          tmem_ld_32x32(seg_row + j * 32, a);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float v = __uint_as_float(a[i]);
            master[j * 32 + i] = has_master ? __fadd_rn(v, master[j * 32 + i]) : v;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[fb]);
        has_master = true;
      }
      const int buf = item & 1;
      const uint32_t use = static_cast<uint32_t>(item >> 1);
      ++item;
      const bool complete = (kb0 == 0 && kb1 == KB);
      const int r0 = q * 32;
      const int bh0 = r0 / p.tile_w, bw0 = r0 - bh0 * p.tile_w;
      const int st_w = tc.w0 + bw0, st_h = tc.h0 + bh0;
      const int res_n = tc.img + tc.batch * p.res_n_off;
      const int nchunks = min(BN / 32, (p.cout - tc.n0 + 31) / 32);
      const int bsel = tile_item & 1;
      // ---- bias slices: [bsel] holds this tile's (written at the end of the previous tile, or right here for the first)
      epi_bar_sync8();     // every warp is done with the tile before: its bias buffer may be refilled, this one's is visible
      if (tile_item == 0) {
        if (epi_tid < BN) {
          const int n = tc.n0 + epi_tid;
          bias_s[epi_tid] = (p.bias && n < p.cout) ? __ldg(p.bias + tc.batch * p.bias_z_off + n) : 0.f;
        }
        epi_bar_sync8();
      }
      float next_bias = 0.f;
      bool has_next = false;
      {
        WorkIter peek = it;
        int t2, k0, k1;
        has_next = peek.next(t2, k0, k1);
        if (has_next && epi_tid < BN) {
          const TileCoord tc2 = decode_tile(p, t2, BN);
          const int n = tc2.n0 + epi_tid;
          next_bias = (p.bias && n < p.cout) ? __ldg(p.bias + tc2.batch * p.bias_z_off + n) : 0.f;
        }
      }
      // ---- the staging buffers are free once the previous tile's stores have read them; then the residual may land there
      if (lane == 0) tma_store_wait_read<0>();
      __syncwarp();
      auto issue_residual = [&]() {
        if (p.has_residual && lane == 0) {
#pragma unroll
          for (int j = 0; j < kCPW; ++j) {
            const int cj = half * kCPW + j;
            if (cj < nchunks) {
              mbar_arrive_expect_tx(&rbar[j], 4096);
              tma_load_4d(stage_buf + j * 4096, &tmRes, &rbar[j], tc.n0 + cj * 32 + tc.batch * p.res_c_off, st_w, st_h, res_n);
            }
          }
        }
      };
      if (complete) issue_residual();
      mbar_wait(&tmem_full_bar[buf], use & 1);
      tc_fence_after();
      const uint32_t tmem_row = tmem_base + buf * kAccStride + lane_bits + col_base;
      // chunk j of my columns: last segment (+ master). One chunk is live at a time (168 registers per thread at 320 threads)
      auto load_chunk = [&](const int j, float (&acc)[32]) {
        uint32_t raw[32];
        __syncwarp();
        tmem_ld_32x32(tmem_row + j * 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float v = __uint_as_float(raw[i]);
          acc[i] = has_master ? __fadd_rn(v, master[j * 32 + i]) : v;
        }
      };
      const int my_chunks = max(0, min(kCPW, nchunks - half * kCPW));
      auto release_acc = [&]() {      // hand the accumulator buffer back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
      };
      bool finalize = complete;
      int c_first = cta, c_last = cta;
      if (!complete) {
        // ---- publish this CTA's partial accumulator (my columns), then find out whether it arrived last
        float* my_ws = p.part_ws + ((static_cast<long long>(cta) * 2 + (tile_item == 0 ? 0 : 1)) * kBM + row) * BN + col_base;
#pragma unroll
        for (int j = 0; j < kCPW; ++j) {
          float acc[32];
          load_chunk(j, acc);
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            __stcg(reinterpret_cast<float4*>(my_ws + j * 32 + i), make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]));
        }
        __threadfence();
        epi_bar_sync8();
        c_first = unit_owner(U, grid, t * KB);
        c_last = unit_owner(U, grid, t * KB + KB - 1);
        if (epi_tid == 0) {
          const int parts = c_last - c_first + 1;
          const int old = atomicAdd(&p.counters[t], 1);
          const int last = (old == parts - 1);
          if (last) p.counters[t] = 0;   // every part has arrived: leave the counter clean for the next launch
          *epi_flag = last;
        }
        epi_bar_sync8();
        finalize = (*epi_flag != 0);
        if (finalize) {
          __threadfence();
          issue_residual();
        }
      }
      if (!finalize || my_chunks == 0) release_acc();
      if (finalize) {
        const int out_n = tc.img + tc.batch * p.out_n_off;
#pragma unroll
        for (int j = 0; j < kCPW; ++j) {
          if (j < my_chunks) {
            const int cj = half * kCPW + j;
            float acc[32];
            load_chunk(j, acc);
            if (j + 1 == my_chunks) release_acc();
            if (!complete) {
              // deterministic reduction: parts summed in CTA order, own part from tensor memory
              float sum[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) sum[i] = 0.f;
              for (int oc = c_first; oc <= c_last; ++oc) {
                if (oc == cta) {
#pragma unroll
                  for (int i = 0; i < 32; ++i) sum[i] += acc[i];
                } else {
                  const int slot = (cta_first_unit(U, grid, oc) >= t * KB) ? 0 : 1;
                  const float* ws = p.part_ws + ((static_cast<long long>(oc) * 2 + slot) * kBM + row) * BN + col_base + j * 32;
#pragma unroll
                  for (int i = 0; i < 32; i += 4) {
                    const float4 v = __ldcg(reinterpret_cast<const float4*>(ws + i));
                    sum[i] += v.x; sum[i + 1] += v.y; sum[i + 2] += v.z; sum[i + 3] += v.w;
                  }
                }
              }
#pragma unroll
              for (int i = 0; i < 32; ++i) acc[i] = sum[i];
            }
            uint8_t* rowp = stage_buf + j * 4096 + lane * 128;
            const float4* biv = reinterpret_cast<const float4*>(bias_s + bsel * BN + cj * 32);
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 bi = biv[i >> 2];
              acc[i] = fmaf(acc[i], p.acc_scale, bi.x); acc[i + 1] = fmaf(acc[i + 1], p.acc_scale, bi.y);
              acc[i + 2] = fmaf(acc[i + 2], p.acc_scale, bi.z); acc[i + 3] = fmaf(acc[i + 3], p.acc_scale, bi.w);
            }
            if (p.has_residual) {
              mbar_wait(&rbar[j], (rphase >> j) & 1u);
              rphase ^= (1u << j);
              if (res_split) {    // 16-byte chunks 0..3: hi halves of values 8c .. 8c+7, chunks 4..7: their lo halves
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  const uint4 rh = *reinterpret_cast<const uint4*>(rowp + ((static_cast<uint32_t>(c) ^ sw) << 4));
                  const uint4 rl = *reinterpret_cast<const uint4*>(rowp + ((static_cast<uint32_t>(4 + c) ^ sw) << 4));
                  const uint32_t hs[4] = {rh.x, rh.y, rh.z, rh.w}, ls[4] = {rl.x, rl.y, rl.z, rl.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float2 a2 = h2_to_f2(hs[e]), b2 = h2_to_f2(ls[e]);
                    acc[8 * c + 2 * e] += a2.x + b2.x;
                    acc[8 * c + 2 * e + 1] += a2.y + b2.y;
                  }
                }
              } else {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                  const float4 rr = *reinterpret_cast<const float4*>(rowp + ((static_cast<uint32_t>(c) ^ sw) << 4));
                  acc[4 * c] += rr.x; acc[4 * c + 1] += rr.y; acc[4 * c + 2] += rr.z; acc[4 * c + 3] += rr.w;
                }
              }
            }
            // (every read of the row precedes the first write below: the result goes back to the same bytes)
            if (p.relu) {
#pragma unroll
              for (int i = 0; i < 32; ++i) acc[i] = fmaxf(acc[i], slope * acc[i]);
            }
            if (OUT16) {      // split-fp16 result
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                uint32_t hh[4], ll[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  hh[e] = f2_to_h2_sat(acc[8 * c + 2 * e], acc[8 * c + 2 * e + 1]);
                  const float2 back = h2_to_f2(hh[e]);
                  ll[e] = f2_to_h2_sat(acc[8 * c + 2 * e] - back.x, acc[8 * c + 2 * e + 1] - back.y);
                }
                *reinterpret_cast<uint4*>(rowp + ((static_cast<uint32_t>(c) ^ sw) << 4)) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
                *reinterpret_cast<uint4*>(rowp + ((static_cast<uint32_t>(4 + c) ^ sw) << 4)) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
              }
            } else {
#pragma unroll
              for (int c = 0; c < 8; ++c)
                *reinterpret_cast<float4*>(rowp + ((static_cast<uint32_t>(c) ^ sw) << 4)) =
                    make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_4d(&tmOut, stage_buf + j * 4096, tc.n0 + cj * 32 + tc.batch * p.out_c_off, st_w, st_h, out_n);
              tma_store_commit();
            }
          }
        }
      }
      if (has_next && epi_tid < BN) bias_s[(bsel ^ 1) * BN + epi_tid] = next_bias;
    }
    if (lane == 0) tma_store_wait<0>();   // global writes complete before the CTA retires
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;
    const int epi_tid = (warp - 2) * 32 + lane;
    uint8_t* epi_out = smem + L::kEpiOffset + (warp - 2) * 16384;   // 2 x 4 KB store staging
    uint8_t* epi_res = epi_out + 8192;                              // 2 x 4 KB residual staging
    uint64_t* rbar = res_bar + (warp - 2) * 2;
    float* sb_s = reinterpret_cast<float*>(smem + L::kSbOffset);
    uint32_t rphase = 0;
    WorkIter it(p, cta, grid);
    int t;
    int kb0, kb1;
    int item = 0;   // accumulator-segment counter (ping-pong bookkeeping shared with the MMA warp)
    const uint32_t lane_bits = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t master_row = tmem_base + 2 * kAccStride + lane_bits;
    for (int tile_item = 0; it.next(t, kb0, kb1); ++tile_item) {
      const TileCoord tc = decode_tile(p, t, BN);
      // ---- 3xTF32 only: fold every segment but the last into the master accumulator (RN fp32 adds)
      bool has_master = false;
      int s0 = kb0;
      for (; SEG && s0 + kSegLen < kb1; s0 += kSegLen, ++item) {
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
      const bool complete = (kb0 == 0 && kb1 == KB);
      // ---- while the MMAs of this tile run: stage its scale / bias slice in shared memory (the chunk loop then reads
      //      them with broadcast LDS instead of L1-missing global loads) and start the first residual load
      const int r0 = q * 32;
      const int bh0 = r0 / p.tile_w, bw0 = r0 - bh0 * p.tile_w;
      const int st_w = tc.w0 + bw0, st_h = tc.h0 + bh0;
      const int res_n = tc.img + tc.batch * p.res_n_off;
      const int nchunks = min(BN / CW, (p.cout - tc.n0 + CW - 1) / CW);
      epi_bar_sync();   // every warp is done with the previous tile's scale / bias
      for (int i = epi_tid; i < BN; i += 128) {
        const int n = tc.n0 + i;
        const int zoff = tc.batch * p.bias_z_off;
        sb_s[i] = (p.scale && n < p.cout) ? __ldg(p.scale + zoff + n) : 1.f;
        sb_s[L::kSbCols + i] = (p.bias && n < p.cout) ? __ldg(p.bias + zoff + n) : 0.f;
      }
      if (complete && p.has_residual && lane == 0 && nchunks > 0) {
        mbar_arrive_expect_tx(&rbar[0], 4096);
        tma_load_4d(epi_res, &tmRes, &rbar[0], tc.n0 + tc.batch * p.res_c_off, st_w, st_h, res_n);
      }
      epi_bar_sync();
      mbar_wait(&tmem_full_bar[buf], use & 1);
      tc_fence_after();
      const uint32_t tmem_row = tmem_base + buf * kAccStride + lane_bits;
      // accumulator chunk c (32 columns of this thread's row): last segment (+ master)
      auto load_acc = [&](int c, uint32_t (&acc)[32]) {
        __syncwarp();  // tcgen05.ld is .sync.aligned
        tmem_ld_32x32(tmem_row + c * 32, acc);
        tmem_ld_wait();
        if (SEG && has_master) {
          uint32_t m[32];
          tmem_ld_32x32(master_row + c * 32, m);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] = __float_as_uint(__fadd_rn(__uint_as_float(acc[j]), __uint_as_float(m[j])));
        }
      };
      bool finalize = complete;
      int c_first = cta, c_last = cta;
      if (!complete) {
        // ---- publish this CTA's partial accumulator, then find out whether it arrived last
        float* my_ws = p.part_ws + ((static_cast<long long>(cta) * 2 + (tile_item == 0 ? 0 : 1)) * kBM + row) * BN;
        auto publish = [&](const int c) {
          uint32_t acc[32];
          load_acc(c, acc);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 v = make_float4(__uint_as_float(acc[j]), __uint_as_float(acc[j + 1]),
                                   __uint_as_float(acc[j + 2]), __uint_as_float(acc[j + 3]));
            __stcg(reinterpret_cast<float4*>(my_ws + c * 32 + j), v);
          }
        };
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) publish(c);
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
        const int out_n = tc.img + tc.batch * p.out_n_off;
        const bool has_sb = (p.scale != nullptr) || (p.bias != nullptr);
        const uint32_t sw = static_cast<uint32_t>(lane & 7);
        if (!complete && p.has_residual && lane == 0 && nchunks > 0) {   // (whole tiles started this load earlier)
          mbar_arrive_expect_tx(&rbar[0], 4096);
          tma_load_4d(epi_res, &tmRes, &rbar[0], tc.n0 + tc.batch * p.res_c_off, st_w, st_h, res_n);
        }
        auto finish_chunk = [&](const int c) {
          const int nb = tc.n0 + c * CW;
          const uint8_t* rsrc = nullptr;
          if (p.has_residual) {
            const int rb = c & 1;
            if (c + 1 < nchunks && lane == 0) {   // prefetch the next residual chunk into the other buffer
              mbar_arrive_expect_tx(&rbar[rb ^ 1], 4096);
              tma_load_4d(epi_res + (rb ^ 1) * 4096, &tmRes, &rbar[rb ^ 1], nb + CW + tc.batch * p.res_c_off, st_w, st_h,
                          res_n);
            }
            mbar_wait(&rbar[rb], (rphase >> rb) & 1u);
            rphase ^= (1u << rb);
            rsrc = epi_res + rb * 4096 + lane * 128;
          }
          // the out staging buffer (c & 1) was handed to a TMA store two chunks ago: wait until read
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
          uint8_t* dst = epi_out + (c & 1) * 4096 + lane * 128;
          // 32 accumulator columns at a time (keeps the live registers at 32 + a handful: the 64-wide form spilled)
        #pragma unroll
          for (int h = 0; h < CW / 32; ++h) {
            uint32_t raw[32];
            load_acc(c * (CW / 32) + h, raw);
            float acc[32];
        #pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(raw[j]);
            const int col0 = c * CW + h * 32;     // first column of this half inside the tile
            if (!complete) {
              // deterministic reduction: parts summed in CTA order, own part from TMEM
              float sum[32];
        #pragma unroll
              for (int j = 0; j < 32; ++j) sum[j] = 0.f;
              for (int oc = c_first; oc <= c_last; ++oc) {
                if (oc == cta) {
        #pragma unroll
                  for (int j = 0; j < 32; ++j) sum[j] += acc[j];
                } else {
                  const int slot = (cta_first_unit(U, grid, oc) >= t * KB) ? 0 : 1;
                  const float* ws = p.part_ws + ((static_cast<long long>(oc) * 2 + slot) * kBM + row) * BN + col0;
        #pragma unroll
                  for (int j = 0; j < 32; j += 4) {
                    const float4 v = __ldcg(reinterpret_cast<const float4*>(ws + j));
                    sum[j] += v.x; sum[j + 1] += v.y; sum[j + 2] += v.z; sum[j + 3] += v.w;
                  }
                }
              }
        #pragma unroll
              for (int j = 0; j < 32; ++j) acc[j] = sum[j];
            }
            if (has_sb) {
              const float4* scv = reinterpret_cast<const float4*>(sb_s + col0);
              const float4* biv = reinterpret_cast<const float4*>(sb_s + L::kSbCols + col0);
        #pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 sc = scv[j >> 2], bi = biv[j >> 2];
                acc[j] = fmaf(acc[j], sc.x, bi.x); acc[j + 1] = fmaf(acc[j + 1], sc.y, bi.y);
                acc[j + 2] = fmaf(acc[j + 2], sc.z, bi.z); acc[j + 3] = fmaf(acc[j + 3], sc.w, bi.w);
              }
            }
            const float slope = p.relu == 2 ? 0.1f : 0.f;
            if (OUTH) {
              // 64 halves per staging row: 16-byte groups of 8 halves, swizzled like the TMA box
        #pragma unroll
              for (int j = 0; j < 32; j += 8) {
                const uint32_t chunk = (static_cast<uint32_t>((h * 32 + j) >> 3) ^ sw) << 4;
                float v[8];
        #pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = acc[j + e];
                if (rsrc) {
                  const uint4 rr = *reinterpret_cast<const uint4*>(rsrc + chunk);
                  const float2 r0v = h2_to_f2(rr.x), r1v = h2_to_f2(rr.y), r2v = h2_to_f2(rr.z), r3v = h2_to_f2(rr.w);
                  v[0] += r0v.x; v[1] += r0v.y; v[2] += r1v.x; v[3] += r1v.y;
                  v[4] += r2v.x; v[5] += r2v.y; v[6] += r3v.x; v[7] += r3v.y;
                }
                if (p.relu) {
        #pragma unroll
                  for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], slope * v[e]);
                }
                uint4 o;
                o.x = f2_to_h2(v[0], v[1]); o.y = f2_to_h2(v[2], v[3]);
                o.z = f2_to_h2(v[4], v[5]); o.w = f2_to_h2(v[6], v[7]);
                *reinterpret_cast<uint4*>(dst + chunk) = o;
              }
            } else {
        #pragma unroll
              for (int j = 0; j < 32; j += 4) {
                float4 v = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
                const uint32_t chunk = (static_cast<uint32_t>(j >> 2) ^ sw) << 4;
                if (rsrc) {
                  const float4 rr = *reinterpret_cast<const float4*>(rsrc + chunk);
                  v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
                }
                if (p.relu) {
                  v.x = fmaxf(v.x, slope * v.x); v.y = fmaxf(v.y, slope * v.y);
                  v.z = fmaxf(v.z, slope * v.z); v.w = fmaxf(v.w, slope * v.w);
                }
                *reinterpret_cast<float4*>(dst + chunk) = v;
              }
            }
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&tmOut, epi_out + (c & 1) * 4096, nb + tc.batch * p.out_c_off, st_w, st_h, out_n);
            tma_store_commit();
          }
        };
#pragma unroll 1
        for (int c = 0; c < nchunks; ++c) finish_chunk(c);
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

// ------------------------------------------------------------------ launch
// pdl != 0: launched with programmatic stream serialization, i.e. this kernel's prologue may start while the
// previous kernel on the stream drains (the kernel itself orders its memory accesses with griddepcontrol.wait).
template <int BN, int STAGES, int MODE, bool OUT16>
static int launch_cfg(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut,
                      const CUtensorMap& tmRes, const ConvGemmParams& p, dim3 grid, cudaStream_t stream, int pdl) {
  using L = SmemLayout<BN, STAGES, MODE>;
  static bool configured = false;
  if (!configured) {
    MEGA_CUDA_CHECK(cudaFuncSetAttribute(conv_gemm_kernel<BN, STAGES, MODE, OUT16>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kThreads + ((MODE == kModeSplit3 || MODE == kModeF16x3) ? 128 : 0), 1, 1);
  cfg.dynamicSmemBytes = L::kTotal;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  MEGA_CUDA_CHECK(cudaLaunchKernelEx(&cfg, conv_gemm_kernel<BN, STAGES, MODE, OUT16>, tmA, tmB, tmOut, tmRes, p));
  return MEGA_OK;
}

// fp16-operand instantiations live in their own translation unit (conv_gemm_f16.cu)
int launch_conv_gemm_f16(int block_n, int out_f16, const CUtensorMap& tmA, const CUtensorMap& tmB,
                         const CUtensorMap& tmOut, const CUtensorMap& tmRes, const ConvGemmParams& p, dim3 grid,
                         cudaStream_t stream, int pdl);

// split-fp16 ("3xFP16") instantiations: conv_gemm_f16x3.cu
int launch_conv_gemm_f16x3(int block_n, int out_split, const CUtensorMap& tmA, const CUtensorMap& tmB,
                           const CUtensorMap& tmOut, const CUtensorMap& tmRes, const ConvGemmParams& p, dim3 grid,
                           cudaStream_t stream, int pdl);

}  // namespace mega
