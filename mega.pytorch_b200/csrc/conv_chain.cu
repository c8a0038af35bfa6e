// A whole chain of dependent convolutions / GEMMs in ONE persistent kernel.
//
// The backbone of the MEGA hot path is a strictly sequential chain of ~100 small implicit GEMMs per frame pair
// (ResNet-101 res2..res5, modeling/backbone/resnet.py:324-344; RPN head, rpn/rpn.py:99-106). Each of them is 2-10 us
// of tensor-core work at M = 4788 output pixels, so one-kernel-per-layer execution is dominated by what surrounds
// the math: launch, barrier/TMEM set-up, descriptor fetch, pipeline fill and drain (measured 17-30 us per layer,
// profiles/r01_ncu_full_conv_gemm_f16_res4_raw.csv: tensor pipe active 4-10 % of the kernel's duration).
//
// Here the per-layer kernel body (TMA producer warp / tcgen05 MMA warp / 8 epilogue warps, double-buffered TMEM
// accumulators, persistent stream-K work list -- see conv_gemm_kernel.cuh) is wrapped in a loop over a device-side
// table of layers. One CTA per SM stays resident for the whole chain; mbarriers, the TMEM allocation and the smem
// ring are set up once; layers are separated by a grid-wide barrier (one atomic counter, release/acquire) instead of
// a kernel boundary. Tensor maps live in the layer table in global memory.
//
// Barrier depth. With depth 1 layer l starts when every CTA has finished layer l-1: the tensor pipe idles through every
// layer's tail (last epilogue ~2.2 us, store drain + gpu-scope release ~1 us, barrier ~0.8 us, first operand fetch
// ~0.8 us: 5-6 us against 3-10 us of MMAs per ResNet-101 layer at two 600x1000 frames -- tools/trace_chain.py). With
// depth 2 layer l only waits for layer l-2 (one arrival counter per layer parity), so a table that INTERLEAVES two
// independent chains A0 B0 A1 B1 ... (the per-frame branch of two halves of an image batch) keeps the TMA / MMA warps of
// every CTA streaming chain B's layer while chain A's tail drains, and vice versa. Stream-K partial sums and tile
// counters of odd layers live in the second half of the workspace (a CTA may already publish partials of layer l+1
// while a slower CTA still reduces layer l).
//
// Restrictions of a chain: fp16 operands (kind::f16), block_n <= 128 (one 32 KB smem stage holds A 128 x 64 and B
// block_n x 64 halves), output fp16 or fp32 per layer.
#include "conv_gemm_kernel.cuh"

namespace mega {

constexpr int kChainStages = 5;
constexpr int kChainStageBytes = 32768;     // A tile 16 KB + B tile (<= 128 rows) 16 KB
constexpr int kChainABytes = 16384;
constexpr int kChainEpiBytes = 4 * 4 * 4096;
constexpr int kChainBarOffset = kChainStages * kChainStageBytes + kChainEpiBytes;
constexpr int kChainSbOffset = kChainBarOffset + 256;      // [scale | bias][128] floats of the tile being finished
constexpr int kChainSmem = kChainSbOffset + 1024 + 1024;
constexpr uint32_t kChainTmemCols = 256;    // two accumulators of up to 128 fp32 columns
constexpr uint32_t kChainAccStride = 128;

struct alignas(128) ChainLayer {
  CUtensorMap tmA, tmB, tmOut, tmRes;
  ConvGemmParams p;
  int block_n;
  int out16;
  int active_ctas;   // CTAs that take part in this layer's work list (<= grid); the others only pass the barrier
  int cta_rot;       // physical CTA (cta_rot + i) % grid plays logical CTA i of this layer's work list (depth-2 chains:
                     // the work lists of consecutive layers start where the previous one ended, so a layer whose tile
                     // count is not a multiple of the grid does not leave the same SMs idle every time)
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// generic <-> async proxy ordering (TMA loads of data other CTAs wrote with TMA stores / generic stores)
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// all CTAs of the grid have finished `target / grid` layers
__device__ __forceinline__ void grid_wait(const unsigned* ctr, unsigned target) {
  if (ld_acquire_u32(ctr) >= target) {
    fence_proxy_async_all();
    return;
  }
  const uint64_t t0 = globaltimer_ns();
  uint32_t spins = 0;
  while (ld_acquire_u32(ctr) < target) {
    __nanosleep(40);
    if ((++spins & 0x3ff) == 0 && globaltimer_ns() - t0 > 2000000000ull) {
      printf("mega: chain grid barrier timeout block %d thread %d target %u have %u\n", blockIdx.x, threadIdx.x, target,
             ld_acquire_u32(ctr));
      __trap();
    }
  }
  fence_proxy_async_all();
}

struct PipeState {
  int stage;
  uint32_t phase;
};

// optional in-kernel event trace of ONE CTA (diagnostics, tools/trace_chain.py): (tag, SM clock) pairs per role
struct ChainTrace {
  unsigned long long* buf;   // [3 roles][kTraceCap][2] or NULL
  int cta;
  int level;                 // 0: every event; 1: per-layer events only (a 100-layer chain fits the buffer);
                             // 2 + l: every event of layer l only
};
constexpr int kTraceCap = 4096;
struct TraceCursor {
  unsigned long long* p;
  int n;
  int level;
  __device__ __forceinline__ void put(unsigned long long tag) {
    if (p != nullptr && n < kTraceCap) {
      const unsigned code = static_cast<unsigned>(tag & 0xff);
      if (level == 1 && code >= 3 && code <= 7) return;      // per-k-block / per-tile events
      if (level >= 2 && static_cast<int>(tag >> 32) != level - 2) return;
      p[2 * n] = tag;
      p[2 * n + 1] = static_cast<unsigned long long>(clock64());
      ++n;
    }
  }
};
__device__ __forceinline__ TraceCursor trace_cursor(const ChainTrace& tr, int role, int cta) {
  TraceCursor c;
  c.p = (tr.buf != nullptr && cta == tr.cta) ? tr.buf + static_cast<long long>(role) * kTraceCap * 2 : nullptr;
  c.n = 0;
  c.level = tr.level;
  return c;
}
#define TR_TAG(layer, idx, code) ((static_cast<unsigned long long>(layer) << 32) | (static_cast<unsigned long long>(idx) << 8) | (code))

// ------------------------------------------------------------------ epilogue of one layer (8 warps)
// Eight epilogue warps: warp w reads TMEM lane quarter (w & 3) (the hardware restriction: a warp touches lanes
// 32*(warp_id % 4) .. +31), and of the tile's 64-column (fp16 out) / 32-column (fp32 out) chunks it takes those with
// chunk % 2 == (w - 2) / 4. Round 1 ran 4 warps over all chunks with ~300 executed instructions per 32 columns (generic
// LD / ST to the staging buffers, per-element branches on the layer's flags, FMUL + FMNMX for every ReLU): 4.1 us per
// 128 x 128 tile with a residual against 1.2 us of MMAs for K = 256 (tools/trace_backbone.py), which made every
// 1x1-expand layer of the backbone epilogue-bound. Here the flags are template parameters of the inner loop, staging
// goes through explicit ld/st.shared with precomputed swizzled offsets, and two warps share a lane quarter.
// (Tried on top and reverted: walking the work list one tile ahead to prefetch the next tile's scale / bias into registers
// and its residual chunk by TMA. The residual still arrived 0.5 us after the accumulator -- it queues behind the main
// loop's operand loads in the SM's TMA FIFO -- and the second decode_tile per tile cost more than the barriers it saved:
// backbone chain 1.61 -> 1.70 ms at 2 images, 3.75 -> 4.04 ms at 8.)
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kChainThreads = 64 + kEpiThreads;

__device__ __forceinline__ void epi_bar_sync_all() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// 32 accumulator columns (one thread = one output pixel) -> scale / bias -> (+ residual) -> activation -> staging row.
// sb: shared address of scale[col0 ..] (bias at +512 bytes); dst / rsrc: shared addresses of this lane's 128-byte staging
// rows; off[g]: byte offset of 16-byte group g inside the swizzled row. H: which 32-column half of a 64-half row (OUT16).
template <bool OUT16, bool RES, int RELU>
__device__ __forceinline__ void epi_half(const uint32_t (&raw)[32], uint32_t sb, uint32_t dst, uint32_t rsrc,
                                         const uint32_t (&off)[8], const int H) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const uint4 sc = lds128(sb + j * 4), bi = lds128(sb + 512 + j * 4);
    v[j] = fmaf(__uint_as_float(raw[j]), __uint_as_float(sc.x), __uint_as_float(bi.x));
    v[j + 1] = fmaf(__uint_as_float(raw[j + 1]), __uint_as_float(sc.y), __uint_as_float(bi.y));
    v[j + 2] = fmaf(__uint_as_float(raw[j + 2]), __uint_as_float(sc.z), __uint_as_float(bi.z));
    v[j + 3] = fmaf(__uint_as_float(raw[j + 3]), __uint_as_float(sc.w), __uint_as_float(bi.w));
  }
  if (OUT16) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {          // 8 halves = 16 bytes per group
      const int j = g * 8;
      if (RES) {
        const uint4 rr = lds128(rsrc + off[H * 4 + g]);
        const float2 r0 = h2_to_f2(rr.x), r1 = h2_to_f2(rr.y), r2 = h2_to_f2(rr.z), r3 = h2_to_f2(rr.w);
        v[j] += r0.x; v[j + 1] += r0.y; v[j + 2] += r1.x; v[j + 3] += r1.y;
        v[j + 4] += r2.x; v[j + 5] += r2.y; v[j + 6] += r3.x; v[j + 7] += r3.y;
      }
      if (RELU == 1) {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[j + e] = fmaxf(v[j + e], 0.f);
      } else if (RELU == 2) {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[j + e] = fmaxf(v[j + e], 0.1f * v[j + e]);
      }
      uint4 o;
      o.x = f2_to_h2(v[j], v[j + 1]); o.y = f2_to_h2(v[j + 2], v[j + 3]);
      o.z = f2_to_h2(v[j + 4], v[j + 5]); o.w = f2_to_h2(v[j + 6], v[j + 7]);
      sts128(dst + off[H * 4 + g], o);
    }
  } else {
#pragma unroll
    for (int g = 0; g < 8; ++g) {          // 4 floats = 16 bytes per group
      const int j = g * 4;
      if (RES) {
        const uint4 rr = lds128(rsrc + off[g]);
        v[j] += __uint_as_float(rr.x); v[j + 1] += __uint_as_float(rr.y);
        v[j + 2] += __uint_as_float(rr.z); v[j + 3] += __uint_as_float(rr.w);
      }
      if (RELU == 1) {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[j + e] = fmaxf(v[j + e], 0.f);
      } else if (RELU == 2) {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[j + e] = fmaxf(v[j + e], 0.1f * v[j + e]);
      }
      uint4 o;
      o.x = __float_as_uint(v[j]); o.y = __float_as_uint(v[j + 1]); o.z = __float_as_uint(v[j + 2]); o.w = __float_as_uint(v[j + 3]);
      sts128(dst + off[g], o);
    }
  }
}

template <bool OUT16>
__device__ __forceinline__ void epi_half_dispatch(const uint32_t (&raw)[32], uint32_t sb, uint32_t dst, uint32_t rsrc,
                                                  const uint32_t (&off)[8], const int H, const bool res, const int relu) {
  // the flags are uniform over the layer: one branch per 32 columns instead of one per element
  if (relu == 1) {
    if (res) epi_half<OUT16, true, 1>(raw, sb, dst, rsrc, off, H);
    else epi_half<OUT16, false, 1>(raw, sb, dst, rsrc, off, H);
  } else if (relu == 0) {
    if (res) epi_half<OUT16, true, 0>(raw, sb, dst, rsrc, off, H);
    else epi_half<OUT16, false, 0>(raw, sb, dst, rsrc, off, H);
  } else {
    if (res) epi_half<OUT16, true, 2>(raw, sb, dst, rsrc, off, H);
    else epi_half<OUT16, false, 2>(raw, sb, dst, rsrc, off, H);
  }
}

template <bool OUT16>
__device__ __noinline__ void chain_epilogue_layer(const ChainLayer* L, const ConvGemmParams& p, const int BN, uint8_t* smem,
                                                  uint64_t* tmem_full_bar, uint64_t* tmem_empty_bar, uint64_t* res_bar,
                                                  int* epi_flag, uint32_t tmem_base, int warp, int lane, int cta,
                                                  int grid, int& item, uint32_t& rphase, TraceCursor& tr, int layer) {
  constexpr int CW = OUT16 ? 64 : 32;          // columns per chunk (= one 128-byte staging row)
  constexpr int HPC = CW / 32;                 // 32-column halves per chunk
  const int ew = warp - 2;                     // epilogue warp 0..7
  const int q = warp & 3;                      // TMEM lane quarter
  const int half = ew >> 2;                    // parity of the chunks this warp takes
  const int row = q * 32 + lane;
  const int epi_tid = ew * 32 + lane;
  const uint32_t stage_base = smem_u32(smem + kChainStages * kChainStageBytes + ew * 8192);
  const uint32_t out_row = stage_base + lane * 128;          // 4 KB store staging | 4 KB residual staging per warp
  const uint32_t res_row = out_row + 4096;
  uint64_t* rbar = res_bar + ew;
  const uint32_t sb_u32 = smem_u32(smem + kChainSbOffset);
  float* sb_s = reinterpret_cast<float*>(smem + kChainSbOffset);
  const int U = static_cast<int>(p.total_units);
  const int KB = p.kb_per_tile;
  const uint32_t lane_bits = static_cast<uint32_t>(q * 32) << 16;
  const CUtensorMap* tmOut = &L->tmOut;
  const CUtensorMap* tmRes = &L->tmRes;
  const bool has_res = p.has_residual != 0;
  const int relu = p.relu;
  uint32_t off[8];
#pragma unroll
  for (int g = 0; g < 8; ++g) off[g] = (static_cast<uint32_t>(g) ^ static_cast<uint32_t>(lane & 7)) << 4;
  WorkIter it(p, cta, grid);
  int t;
  int kb0, kb1;
  for (int tile_item = 0; it.next(t, kb0, kb1); ++tile_item) {
    const TileCoord tc = decode_tile(p, t, BN);
    const int buf = item & 1;
    const uint32_t use = static_cast<uint32_t>(item >> 1);
    ++item;
    const bool complete = (kb0 == 0 && kb1 == KB);
    const int r0 = q * 32;
    const int bh0 = r0 / p.tile_w, bw0 = r0 - bh0 * p.tile_w;
    const int st_w = tc.w0 + bw0, st_h = tc.h0 + bh0;
    const int res_n = tc.img + tc.batch * p.res_n_off;
    const int nchunks = min(BN / CW, (p.cout - tc.n0 + CW - 1) / CW);
    // ---- while the MMAs of this tile run: stage its scale / bias slice in shared memory and start this warp's first
    //      residual load
    epi_bar_sync_all();   // every warp is done with the previous tile's scale / bias
    if (epi_tid < BN) {
      const int n = tc.n0 + epi_tid;
      const int zoff = tc.batch * p.bias_z_off;
      sb_s[epi_tid] = (p.scale && n < p.cout) ? __ldg(p.scale + zoff + n) : 1.f;
      sb_s[128 + epi_tid] = (p.bias && n < p.cout) ? __ldg(p.bias + zoff + n) : 0.f;
    }
    if (complete && has_res && lane == 0 && half < nchunks) {
      mbar_arrive_expect_tx(rbar, 4096);
      tma_load_4d(reinterpret_cast<void*>(smem + kChainStages * kChainStageBytes + ew * 8192 + 4096), tmRes, rbar,
                  tc.n0 + half * CW + tc.batch * p.res_c_off, st_w, st_h, res_n);
    }
    epi_bar_sync_all();
    mbar_wait(&tmem_full_bar[buf], use & 1);
    tc_fence_after();
    tr.put(TR_TAG(layer, tile_item, 4));
    const uint32_t tmem_row = tmem_base + buf * kChainAccStride + lane_bits;
    bool finalize = complete;
    int c_first = cta, c_last = cta;
    if (!complete) {
      // ---- publish this CTA's partial accumulator (each warp its own columns), then find out whether it arrived last
      float* my_ws = p.part_ws + ((static_cast<long long>(cta) * 2 + (tile_item == 0 ? 0 : 1)) * kBM + row) * BN;
      for (int c = half; c < BN / CW; c += 2) {
#pragma unroll 1
        for (int h = 0; h < HPC; ++h) {
          const int g32 = c * HPC + h;
          uint32_t acc[32];
          __syncwarp();
          tmem_ld_32x32(tmem_row + g32 * 32, acc);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 v = make_float4(__uint_as_float(acc[j]), __uint_as_float(acc[j + 1]), __uint_as_float(acc[j + 2]),
                                   __uint_as_float(acc[j + 3]));
            __stcg(reinterpret_cast<float4*>(my_ws + g32 * 32 + j), v);
          }
        }
      }
      __threadfence();
      epi_bar_sync_all();
      c_first = unit_owner(U, grid, t * KB);
      c_last = unit_owner(U, grid, t * KB + KB - 1);
      if (epi_tid == 0) {
        const int parts = c_last - c_first + 1;
        const int old = atomicAdd(&p.counters[t], 1);
        const int last = (old == parts - 1);
        if (last) p.counters[t] = 0;
        *epi_flag = last;
      }
      epi_bar_sync_all();
      finalize = (*epi_flag != 0);
      if (finalize) __threadfence();
    }
    if (finalize) {
      const int out_n = tc.img + tc.batch * p.out_n_off;
#pragma unroll 1
      for (int c = half; c < nchunks; c += 2) {
        const int nb = tc.n0 + c * CW;
        if (has_res) {
          if ((!complete || c != half) && lane == 0) {   // (whole tiles started their first load before the MMAs)
            mbar_arrive_expect_tx(rbar, 4096);
            tma_load_4d(reinterpret_cast<void*>(smem + kChainStages * kChainStageBytes + ew * 8192 + 4096), tmRes, rbar,
                        nb + tc.batch * p.res_c_off, st_w, st_h, res_n);
          }
          mbar_wait(rbar, (rphase >> ew) & 1u);
          rphase ^= (1u << ew);
          tr.put(TR_TAG(layer, tile_item, 5));
        }
        // this warp's store staging was handed to a TMA store one chunk (usually one tile) ago: wait until it was read
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
#pragma unroll
        for (int h = 0; h < HPC; ++h) {
          const int col0 = c * CW + h * 32;     // first column of this half inside the tile
          uint32_t raw[32];
          __syncwarp();
          tmem_ld_32x32(tmem_row + col0, raw);
          tmem_ld_wait();
          if (!complete) {
            // deterministic reduction: parts summed in CTA order, own part from TMEM
            float sum[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) sum[j] = 0.f;
            for (int oc = c_first; oc <= c_last; ++oc) {
              if (oc == cta) {
#pragma unroll
                for (int j = 0; j < 32; ++j) sum[j] += __uint_as_float(raw[j]);
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
            for (int j = 0; j < 32; ++j) raw[j] = __float_as_uint(sum[j]);
          }
          epi_half_dispatch<OUT16>(raw, sb_u32 + col0 * 4, out_row, res_row, off, h, has_res, relu);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(tmOut, smem + kChainStages * kChainStageBytes + ew * 8192, nb + tc.batch * p.out_c_off, st_w, st_h,
                       out_n);
          tma_store_commit();
        }
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
    tr.put(TR_TAG(layer, tile_item, 6));
  }
}

__global__ void __launch_bounds__(kChainThreads, 1)
conv_chain_kernel(const ChainLayer* __restrict__ layers, const int n_layers, unsigned* sync, const ChainTrace trace,
                  const int depth) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment (the 128B swizzle pattern is a function of the absolute address) as an OFFSET into the shared
  // array, so that the compiler keeps the shared address space (ld/st.shared instead of generic accesses)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kChainBarOffset);
  uint64_t* empty_bar = full_bar + kChainStages;
  uint64_t* tmem_full_bar = empty_bar + kChainStages;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;         // [2]
  uint64_t* res_bar = tmem_empty_bar + 2;               // [8 epilogue warps]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 8);
  int* epi_flag = reinterpret_cast<int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int grid = gridDim.x;
  const int cta = blockIdx.x;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kChainStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], kEpiWarps);
    }
    for (int b = 0; b < kEpiWarps; ++b) mbar_init(&res_bar[b], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, kChainTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();
  griddep_launch_dependents();

  if (warp == 0) {
    // ===================== TMA producer =====================
    // lane 0 loads the A (activation) tile and posts the expected byte count, lane 1 loads the B (weight) tile: the two
    // descriptor-based copies of a k-block are issued in parallel (a single thread needs ~650 cycles per k-block for
    // wait + expect + 2 TMA issues, measured with tools/trace_chain.py; the MMAs of a k-block take ~260)
    if (lane < 2) {
      PipeState ps = {0, 0};
      TraceCursor tr = trace_cursor(trace, 0, cta);
      if (lane != 0) tr.p = nullptr;
      for (int l = 0; l < n_layers; ++l) {
        const ChainLayer* L = layers + l;
        if (lane == 0) prefetch_tmap(&L->tmA); else prefetch_tmap(&L->tmB);
        // everything the loop needs from the layer table is fetched BEFORE the grid barrier
        const ConvGemmParams p = L->p;
        const int BN = L->block_n;
        const int act = L->active_ctas;
        const int k_chunks = p.k_chunks, taps_s = p.taps_s, dil = p.dil, pad = p.pad, pad_w = p.pad_w;
        const int stride_h = p.stride_h, stride_w = p.stride_w;
        const int a_c_off = p.a_c_off, a_n_off = p.a_n_off, b_k_off = p.b_k_off, b_n_off = p.b_n_off;
        const CUtensorMap* tmA = &L->tmA;
        const CUtensorMap* tmB = &L->tmB;
        const uint32_t tx_bytes = static_cast<uint32_t>((kBM + BN) * 128);
        tr.put(TR_TAG(l, 0, 1));
        // layer l may read what layers <= l - depth wrote: all CTAs have arrived l / depth times at counter l % depth
        if (l >= depth) grid_wait(sync + (l % depth), static_cast<unsigned>(l / depth) * grid);
        tr.put(TR_TAG(l, 0, 2));
        int lc = cta - L->cta_rot;
        if (lc < 0) lc += grid;
        if (lc >= act) continue;
        WorkIter it(p, lc, act);
        int t;
        int kb0, kb1;
        while (it.next(t, kb0, kb1)) {
          const TileCoord tc = decode_tile(p, t, BN);
          int tap = kb0 / k_chunks;
          int kc = kb0 - tap * k_chunks;
          int r = tap / taps_s;
          int sx = tap - r * taps_s;
          const int a_c0 = tc.batch * a_c_off, a_n = tc.img + tc.batch * a_n_off;
          const int b_k0 = tc.batch * b_k_off, b_n = tc.n0 + tc.batch * b_n_off;
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&empty_bar[ps.stage], ps.phase ^ 1);
            tr.put(TR_TAG(l, kb, 3));
            uint8_t* a_dst = smem + ps.stage * kChainStageBytes;
            if (lane == 0) {
              mbar_arrive_expect_tx(&full_bar[ps.stage], tx_bytes);
              tma_load_4d(a_dst, tmA, &full_bar[ps.stage], kc * 64 + a_c0, tc.w0 * stride_w + sx * dil - pad_w,
                          tc.h0 * stride_h + r * dil - pad, a_n);
            } else {
              tma_load_3d(a_dst + kChainABytes, tmB, &full_bar[ps.stage], kc * 64 + b_k0, b_n, tap);
            }
            if (++kc == k_chunks) {
              kc = 0;
              ++tap;
              if (++sx == taps_s) {
                sx = 0;
                ++r;
              }
            }
            if (++ps.stage == kChainStages) {
              ps.stage = 0;
              ps.phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      PipeState ps = {0, 0};
      int item = 0;
      TraceCursor tr = trace_cursor(trace, 1, cta);
      for (int l = 0; l < n_layers; ++l) {
        const ChainLayer* L = layers + l;
        const ConvGemmParams p = L->p;
        const int BN = L->block_n;
        const int act = L->active_ctas;
        int lc = cta - L->cta_rot;
        if (lc < 0) lc += grid;
        if (lc >= act) continue;
        const uint32_t idesc = umma_idesc<0>(kBM, BN);
        WorkIter it(p, lc, act);
        int t;
        int kb0, kb1;
        while (it.next(t, kb0, kb1)) {
          const int buf = item & 1;
          const uint32_t use = static_cast<uint32_t>(item >> 1);
          ++item;
          mbar_wait(&tmem_empty_bar[buf], (use & 1) ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + buf * kChainAccStride;
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&full_bar[ps.stage], ps.phase);
            tc_fence_after();
            tr.put(TR_TAG(l, kb, 7));
            const uint32_t a_addr = smem_u32(smem + ps.stage * kChainStageBytes);
            const uint64_t adesc = umma_desc_sw128(a_addr);
            const uint64_t bdesc = umma_desc_sw128(a_addr + kChainABytes);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            umma_commit(&empty_bar[ps.stage]);
            if (++ps.stage == kChainStages) {
              ps.stage = 0;
              ps.phase ^= 1;
            }
          }
          umma_commit(&tmem_full_bar[buf]);
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int epi_tid = (warp - 2) * 32 + lane;
    int item = 0;
    uint32_t rphase = 0;
    TraceCursor tr = trace_cursor(trace, 2, cta);
    if (epi_tid != 0) tr.p = nullptr;
    for (int l = 0; l < n_layers; ++l) {
      const ChainLayer* L = layers + l;
      if (lane == 0) {
        prefetch_tmap(&L->tmOut);
        prefetch_tmap(&L->tmRes);
      }
      // residual / partial-sum reads of this layer must see what the other CTAs wrote in earlier layers
      // (one poller per CTA; the named barrier passes the acquired state on to the other epilogue threads)
      if (l >= depth) {
        if (epi_tid == 0) grid_wait(sync + (l % depth), static_cast<unsigned>(l / depth) * grid);
        epi_bar_sync_all();
        fence_proxy_async_all();
      }
      const ConvGemmParams p = L->p;
      const int act = L->active_ctas;
      int lc = cta - L->cta_rot;
      if (lc < 0) lc += grid;
      if (lc < act) {
        if (L->out16) {
          chain_epilogue_layer<true>(L, p, L->block_n, smem, tmem_full_bar, tmem_empty_bar, res_bar, epi_flag, tmem_base,
                                     warp, lane, lc, act, item, rphase, tr, l);
        } else {
          chain_epilogue_layer<false>(L, p, L->block_n, smem, tmem_full_bar, tmem_empty_bar, res_bar, epi_flag, tmem_base,
                                      warp, lane, lc, act, item, rphase, tr, l);
        }
      }
      tr.put(TR_TAG(l, 0, 8));
      // this CTA's part of layer l is complete and visible: arrive at the grid barrier
      // (stream-K partial sums were fenced by their writers; the TMA stores of every warp are complete after its
      //  wait_group; the named barrier orders all of that before thread 0's single gpu-scope release)
      if (lane == 0) tma_store_wait<0>();
      tr.put(TR_TAG(l, 0, 9));
      epi_bar_sync_all();
      if (epi_tid == 0) {
        fence_proxy_async_all();
        __threadfence();
        atomicAdd(sync + (l % depth), 1u);
      }
      tr.put(TR_TAG(l, 0, 10));
    }
    // last CTA out resets the barrier words for the next launch (every CTA has passed every barrier by then)
    if (epi_tid == 0) {
      const unsigned old = atomicAdd(sync + depth, 1u);
      if (old == static_cast<unsigned>(grid) - 1) {
        for (int i = 0; i <= depth; ++i) sync[i] = 0;
        __threadfence();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kChainTmemCols);
  }
}

// defined in conv_gemm.cu: validates a descriptor and encodes its tensor maps / kernel parameters
int encode_conv_gemm_problem(const mega_conv_gemm_desc* d, CUtensorMap* tmA, CUtensorMap* tmB, CUtensorMap* tmOut,
                             CUtensorMap* tmRes, ConvGemmParams* p, int* ctas);

}  // namespace mega

using namespace mega;

extern "C" long long mega_conv_chain_plan_bytes(int n_layers) {
  return static_cast<long long>(n_layers) * static_cast<long long>(sizeof(ChainLayer));
}

extern "C" int mega_conv_chain_encode2(const mega_conv_gemm_desc* descs, int n_layers, void* plan_host,
                                       long long plan_bytes, int* grid_out, int depth) {
  MEGA_ARG_CHECK(descs != nullptr && plan_host != nullptr && n_layers > 0, "conv_chain: bad arguments");
  MEGA_ARG_CHECK(depth == 1 || depth == 2, "conv_chain: barrier depth must be 1 or 2 (got %d)", depth);
  MEGA_ARG_CHECK(plan_bytes >= mega_conv_chain_plan_bytes(n_layers), "conv_chain: plan buffer too small");
  MEGA_ARG_CHECK((reinterpret_cast<uintptr_t>(plan_host) & 127) == 0, "conv_chain: plan buffer must be 128-byte aligned");
  ChainLayer* out = static_cast<ChainLayer*>(plan_host);
  int grid = 1;
  for (int l = 0; l < n_layers; ++l) {
    const mega_conv_gemm_desc* d = descs + l;
    MEGA_ARG_CHECK(d->precision == kModeF16, "conv_chain: layer %d: chains run fp16 operands only", l);
    MEGA_ARG_CHECK(d->block_n <= 128, "conv_chain: layer %d: block_n %d > 128", l, d->block_n);
    MEGA_ARG_CHECK(d->workspace == descs[0].workspace, "conv_chain: every layer must name the same workspace");
    ChainLayer* L = out + l;
    int ctas = 0;
    const int rc = encode_conv_gemm_problem(d, &L->tmA, &L->tmB, &L->tmOut, &L->tmRes, &L->p, &ctas);
    if (rc != MEGA_OK) return rc;
    L->block_n = d->block_n;
    L->out16 = d->out_f16 ? 1 : 0;
    L->active_ctas = ctas;
    L->cta_rot = 0;
    if (ctas > grid) grid = ctas;
    if (depth == 2) {
      // two layers are in flight: odd layers take the second half of the tile counters and of the partial-sum area
      // (block_n <= 128 in a chain, so one half holds kMaxCtas x 2 x 128 x 128 floats)
      MEGA_ARG_CHECK(L->p.total_tiles <= 32768, "conv_chain: layer %d: %lld tiles exceed the 32768 counters of a depth-2 chain", l,
                     L->p.total_tiles);
      if (l & 1) {
        L->p.counters += 32768;
        L->p.part_ws += static_cast<long long>(kMaxCtas) * 2 * kBM * 128;
      }
    }
  }
  if (depth == 2) {
    // rolling work lists: layer l starts at the physical CTA where layer l-1's list ended
    long long start = 0;
    for (int l = 0; l < n_layers; ++l) {
      out[l].cta_rot = static_cast<int>(start % grid);
      // whole-tile layers occupy min(tiles, ctas) CTAs for (about) one tile time each; advance by the tiles of the last,
      // partial wave so that the next layer begins on the CTAs this one leaves idle
      const long long tiles = out[l].p.total_tiles;
      const int act = out[l].active_ctas;
      start += out[l].p.stream_k ? act : (tiles % act == 0 ? act : tiles % act);
    }
  }
  if (grid_out) *grid_out = grid;
  return MEGA_OK;
}

extern "C" int mega_conv_chain_encode(const mega_conv_gemm_desc* descs, int n_layers, void* plan_host,
                                      long long plan_bytes, int* grid_out) {
  return mega_conv_chain_encode2(descs, n_layers, plan_host, plan_bytes, grid_out, 1);
}

static ChainTrace g_chain_trace = {nullptr, 0, 0};

/* diagnostics: the next launches record an in-kernel event trace of CTA `cta` into trace_dev
 * (3 * 4096 * 2 uint64, zero it first); trace_dev == NULL switches tracing off */
extern "C" int mega_conv_chain_set_trace(void* trace_dev, int cta) {
  g_chain_trace.buf = static_cast<unsigned long long*>(trace_dev);
  g_chain_trace.cta = cta;
  g_chain_trace.level = 0;
  return MEGA_OK;
}

/* level 1: only the per-layer events (layer begin / barrier passed / tiles done / stores drained / arrived) */
extern "C" int mega_conv_chain_set_trace2(void* trace_dev, int cta, int level) {
  g_chain_trace.buf = static_cast<unsigned long long*>(trace_dev);
  g_chain_trace.cta = cta;
  g_chain_trace.level = level;
  return MEGA_OK;
}

extern "C" int mega_conv_chain_launch2(const void* plan_device, int n_layers, int grid, void* sync_words, void* stream_v,
                                       int pdl, int depth) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  MEGA_ARG_CHECK(plan_device != nullptr && sync_words != nullptr && n_layers > 0 && grid > 0 && grid <= kMaxCtas,
                 "conv_chain_launch: bad arguments (grid %d)", grid);
  MEGA_ARG_CHECK(depth == 1 || depth == 2, "conv_chain_launch: barrier depth must be 1 or 2 (got %d)", depth);
  MEGA_ARG_CHECK((reinterpret_cast<uintptr_t>(plan_device) & 127) == 0, "conv_chain_launch: plan must be 128-byte aligned");
  static bool configured = false;
  if (!configured) {
    MEGA_CUDA_CHECK(cudaFuncSetAttribute(conv_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kChainSmem));
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(grid), 1, 1);
  cfg.blockDim = dim3(kChainThreads, 1, 1);
  cfg.dynamicSmemBytes = kChainSmem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  MEGA_CUDA_CHECK(cudaLaunchKernelEx(&cfg, conv_chain_kernel, static_cast<const ChainLayer*>(plan_device), n_layers,
                                     static_cast<unsigned*>(sync_words), g_chain_trace, depth));
  return MEGA_OK;
}

extern "C" int mega_conv_chain_launch(const void* plan_device, int n_layers, int grid, void* sync_words, void* stream_v,
                                      int pdl) {
  return mega_conv_chain_launch2(plan_device, n_layers, grid, sync_words, stream_v, pdl, 1);
}
