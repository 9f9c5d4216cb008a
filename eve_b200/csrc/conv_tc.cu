// tcgen05 implicit-GEMM convolution for sm_100a: TMA-staged NHWC tiles -> shared memory
// (128B swizzle) -> tcgen05.mma (accumulators in TMEM) -> tcgen05.ld epilogue.
//
// Replaces the cuDNN/ATen convolution calls behind every nn.Conv2d on the hot path whose
// GEMM is dense enough for the tensor cores (torchvision ResNet layers via eye_net.py:48-50,
// RefineNet blocks refine_net.py:45-62, ConvRNN gates common.py:338-398).
//
// GEMM view (stride 1, "same" padding):  D[m=(n,h,w)][co] = sum_{tap=(r,q)} sum_ci
//        X[n, h+r-pad, w+q-pad, ci] * Wk[co][tap*Cin+ci]
//  * A operand: for one filter tap the 128 rows of a tile are a (bw x bh x bn) box of
//    pixels shifted by the tap offset, 64 channels deep -- one 4-D TMA box load from the NHWC
//    tensor; out-of-bounds pixels are zero-filled by TMA, which IS the zero padding.
//  * B operand: weights pre-arranged K-major [Cout][taps*Cin], one 2-D TMA box per K block.
//  * Precision: fp32 activations/weights are split into bf16 hi + lo planes; with npass = 3
//    the kernel issues hi*hi + hi*lo + lo*hi per K step (products exact to ~2^-16, fp32
//    accumulation in TMEM), npass = 1 is plain bf16.
//  * Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc), warps 2..5 =
//    epilogue (TMEM -> registers -> +bias/+addend -> fp32 NHWC global).
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace eve {
namespace {

// ------------------------------------------------------------------ PTX wrappers --
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Arrival that publishes nothing but "this thread's tcgen05.ld results are in registers": the
// default .release form makes the thread wait until its earlier global stores are acknowledged
// (measured on the halo-row kernel: ~500 cycles per output row between the last STG and the next
// instruction), which the accumulator hand-back does not need.
__device__ __forceinline__ void mbar_arrive_relaxed(uint64_t* bar) {
  asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#ifdef EVE_TC_DEBUG_SPIN
  // debug builds: a protocol mistake reports the barrier instead of hanging the GPU
  for (long long spins = 0; !mbar_try_wait(bar, parity); ++spins) {
    if (spins > (1ll << 24)) {
      printf("mbar_wait timeout: block %d warp %d bar+%u parity %u\n", (int)blockIdx.x, (int)(threadIdx.x >> 5),
             smem_u32(bar) & 0xFFFu, parity);
      __trap();
    }
  }
#else
  while (!mbar_try_wait(bar, parity)) {
  }
#endif
}
// explicit shared-space vector accesses (a pointer derived from the dynamic shared-memory base
// compiles to generic LD / ST, which queue behind the global stores of the epilogue)
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(addr)
               : "memory");
  return v;
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// the same 2-D tile delivered to the same CTA-relative offset of every CTA in `mask` (and counted on
// the mbarrier at the same offset in each of them)
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* map, uint64_t* bar,
                                               int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmap_prefetch(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols)
               : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; issued by one thread on behalf of the CTA
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once every previously issued tcgen05.mma has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// ... on the mbarrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
      " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

// Shared-memory matrix descriptors.  `cw` = channels per row of the staged box (64, 32 or 16
// bf16 = 128, 64 or 32 bytes), which fixes the TMA / UMMA swizzle mode:
//   64 -> SWIZZLE_128B (layout type 2), 32 -> SWIZZLE_64B (4), 16 -> SWIZZLE_32B (6);
// 8 consecutive rows form one swizzle atom of 8 * row_bytes bytes (= the stride byte offset).
__device__ __forceinline__ uint64_t desc_layout_bits(int cw) {
  return cw == 64 ? (2ull << 61) : (cw == 32 ? (4ull << 61) : (6ull << 61));
}
// K-major (forward / dgrad operands: K = channels runs along the row)
__device__ __forceinline__ uint64_t kmajor_desc(uint32_t saddr, int cw) {
  const uint64_t sbo = (uint64_t)((cw * 2 * 8) >> 4);
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (sbo << 32) | (1ull << 46) |
         desc_layout_bits(cw);
}
// MN-major (wgrad operands: M/N = channels runs along the row, K = pixel rows); lbo = byte
// distance between consecutive cw-channel blocks
__device__ __forceinline__ uint64_t mnmajor_desc(uint32_t saddr, uint32_t lbo_bytes, int cw) {
  const uint64_t sbo = (uint64_t)((cw * 2 * 8) >> 4);
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         (sbo << 32) | (1ull << 46) | desc_layout_bits(cw);
}

struct TcParams {
  int N, OH, OW, Cin, Cout, stride;   // OH x OW: the pixel grid the M tiles walk over
  // filter taps as explicit tables: input offset (in input pixels, before the traversal
  // stride is applied to the tile origin) and the K offset of the tap's weights in B
  int ntaps;
  int tap_dh[9], tap_dw[9], tap_koff[9];
  // where grid pixel (h, w) lands in the output tensor: (h*out_mul + out_ah, w*out_mul + out_aw)
  // of an out_H x out_W image (the parity classes of a stride-2 data gradient use mul = 2)
  int out_mul, out_ah, out_aw, out_H, out_W;
  int bw, bh, bn;           // grid-pixel box of one M tile (bw == OW)
  int tiles_h;              // ceil(OH / bh)
  int tiles_m;              // M tiles: tiles_h * ceil(N / bn)
  int tiles_co;             // N tiles: Cout / BN
  int kc;                   // channels per K chunk: 64, 32 or 16
  int kchunks;              // Cin / kc
  int fmt;                  // operand format: 0 = fp16 planes, 1 = bf16 planes
  // shared-memory ring geometry (sized for the actual chunk width, so that thin-channel layers
  // keep enough bytes in flight): slot = [A hi][A lo][B hi][B lo]
  int a_bytes, b_bytes, stage_bytes, stages;
  float out_scale;          // accumulator scale undoing the weight pre-scale (fp16 planes)
  const float* bias;        // [Cout] or null
  const float* addend;      // [N,OH,OW,Cout] or null
  float* out;               // [N,OH,OW,Cout]
};

constexpr int kTileM = 128;
constexpr int kThreads = 192;

constexpr int kMaxStages = 24;
constexpr int kSmemBudget = 224 * 1024;
constexpr int kBarrierBytes = 512;     // (2 * kMaxStages + 4) mbarriers + the TMEM slot
// box-kernel epilogue: per epilogue warp a 32 pixel x 32 channel fp32 transpose tile (4 KB) and the
// 32 output row offsets of its pixels
constexpr int kEpiTileBytes = 4 * 32 * 128;
constexpr int kEpiBytes = kEpiTileBytes + 128 * 8;

// Stacked-B issue (split operands, BN <= 64).  Measured on B200 (tools/probe_mma.py): with both
// operands in shared memory one tcgen05.mma M=128 x N x K=16 costs max(N / 2, ~(4096 + 32 N) / 120)
// cycles -- 43 / 44 / 51 / 68 / 132 for N = 16 / 32 / 64 / 128 / 256: below N = 128 the 4 KB A-tile
// read, not the tensor pipe, sets the pace.  The three products of a split K step read A_hi twice;
// since the lo weight plane sits directly behind the hi plane at the same row pitch, ONE MMA with
// N = 2 BN over [B_hi ; B_lo] yields A_hi B_hi in accumulator columns [0, BN) and A_hi B_lo in
// [BN, 2 BN), a second MMA adds A_lo B_hi to [0, BN), and the epilogue sums the two halves in fp32:
// 51 + 68 instead of 3 x 51 cycles at BN = 64, 43 + 44 instead of 3 x 43 at BN = 16.
// STACK is a property of the LAYER (Cout <= 64 or not), not of the tile width a launch ends up
// with: few-tile layers narrow BN to spread over more SMs, and a frame's result must not depend on
// the batch it is processed in (the two issue orders round differently).
template <int BN, int NPASS, bool STACK = false>
struct TcCfg {
  static constexpr int kPlanes = NPASS == 3 ? 2 : 1;
  static constexpr bool kStack = NPASS == 3 && BN <= 64 && STACK;
  static constexpr int kAccCols = kStack ? 2 * BN : BN;       // accumulator columns of one tile
  static constexpr uint32_t kTmemCols = kAccCols < 32 ? 32 : kAccCols;   // per accumulator buffer (x2)
};

// PAIR: launched as clusters of two CTAs that walk PAIRS of pixel tiles with the same output-channel
// tile.  The weight tile of a stage is the same for both, so each CTA loads HALF of it (BN / 2 rows)
// and TMA-multicasts that half into both CTAs' rings: the L2 -> shared-memory stream, which bounds
// this kernel (one A box per tap: ~46 B/clk/SM measured against ~13 B/clk of DRAM), carries the
// weights once per pair.  A stage may be refilled only when BOTH CTAs have consumed it, so every
// MMA-done commit arrives on both CTAs' `empty` barriers (count 2).  Arithmetic per tile is
// unchanged (bit-identical results).  MEASURED (tools/conv_table.py, EVE_B200_TC_PAIR=1): 2-8 %
// slower on every 64- and 128-channel layer -- each SM still ingests the whole stage (its own A box,
// its half of B and the peer's half), so what the multicast halves is L2 read traffic, not the bytes
// through the SM's input port that actually pace the kernel, and the two rings advance in lockstep.
// Kept behind `tc_pair` (default off) as the evidence.  cta_group::2 MMAs add nothing either:
// measured (tools/probe_mma.py), an M = 256 pair instruction costs each SM the same 67 cycles at
// N = 128 (46 vs 51 at N = 64): at N >= 128 the instruction already runs at the tensor pipe's rate.
// DUAL: TWO issuing warps (warp 1: even ring stages of a tile, warp 6: odd ones), each accumulating
// into its own TMEM accumulator; the epilogue adds the two.  The idea: the issuing thread does not
// run ahead of the tensor pipe, so one issuer's per-stage bookkeeping and barrier round trip would
// overlap the other's MMAs (two issuers interleave at the pipe's rate: tools/probe_mma.py).
// MEASURED after the issue loop had been slimmed down: 4-11 % SLOWER on the 64-channel layers
// (32.7 vs 32.3 ms per step) -- what is left between the MMAs is waiting for operands to enter the
// SM, which a second issuer cannot hide, plus a second accumulator to drain.  Off by default
// (`tc_dual`), kept with its test as the evidence.
template <int BN, int NPASS, bool STACK, bool PAIR = false, bool DUAL = false>
__global__ void __launch_bounds__(DUAL ? kThreads + 32 : kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
               const TcParams p) {
  static_assert(!(PAIR && DUAL), "the pair and the dual-issuer variants are not combined");
  constexpr uint32_t kAccSets = DUAL ? 2u : 1u;      // partial accumulators per TMEM buffer
  // Persistent: each CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...  The accumulator
  // is double-buffered in TMEM so the epilogue of tile i overlaps the K loop of tile i+1; the
  // shared-memory ring simply keeps rolling across tile boundaries.
  using Cfg = TcCfg<BN, NPASS, STACK>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  const int kStages = p.stages;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * p.stage_bytes);
  uint64_t* empty = full + kStages;
  uint64_t* tmem_full = empty + kStages;          // [2]
  uint64_t* tmem_empty = tmem_full + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint8_t* epi = smem + kStages * p.stage_bytes + kBarrierBytes;   // transpose tiles + row offsets

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int iters = p.ntaps * p.kchunks;
  // DUAL: a ring slot must always belong to the same issuer (an issuer that skipped a phase of a
  // slot's barrier could not tell phase n from phase n - 2 by parity), so the ring depth is even and
  // every tile occupies an EVEN number of ring positions: an odd stage count is padded with one
  // empty position that the producer merely arrives on and its owner merely releases.  Then ring
  // position parity == stage parity within the tile == issuer, for every tile, whatever the batch.
  const int ring_iters = DUAL ? (iters + 1) & ~1 : iters;
  // work items: tiles (tm, tco), or for PAIR pair-rows (2 tm-tiles, tco) of which this CTA takes tile
  // 2 * row + rank; all roles of both CTAs walk the same item sequence
  const uint32_t rank = PAIR ? cluster_rank() : 0u;
  const int item0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int item_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int total_tiles = (PAIR ? (p.tiles_m + 1) / 2 : p.tiles_m) * p.tiles_co;

  if (warp == 0 && lane == 0) {
    tmap_prefetch(&tmA_hi);
    tmap_prefetch(&tmB_hi);
    if (NPASS == 3) {
      tmap_prefetch(&tmA_lo);
      tmap_prefetch(&tmB_lo);
    }
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], PAIR ? 2 : 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], DUAL ? 2 : 1);   // one commit per issuing warp
      mbar_init(&tmem_empty[b], 128);   // every epilogue thread arrives
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * kAccSets * Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();      // the peer's barriers exist before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      const uint32_t rows = p.bw * p.bh * p.bn;
      const uint32_t tx_a = rows * (uint32_t)(p.kc * 2) * Cfg::kPlanes;
      const uint32_t tx_b = (uint32_t)BN * (uint32_t)(p.kc * 2) * Cfg::kPlanes;
      const uint32_t half_bytes = (uint32_t)(BN / 2) * (uint32_t)(p.kc * 2);
      uint32_t s = 0, ph = 0;   // ring position, kept incrementally
      for (int tile = item0; tile < total_tiles; tile += item_step) {
        const int tco = tile % p.tiles_co;
        const int tm = PAIR ? 2 * (tile / p.tiles_co) + (int)rank : tile / p.tiles_co;
        const bool active = tm < p.tiles_m;
        const int h0 = (tm % p.tiles_h) * p.bh;
        const int n0 = (tm / p.tiles_h) * p.bn;
        const int co0 = tco * BN;
        int tap = 0, kc = 0;
        for (int it = 0; it < ring_iters; ++it) {
          mbar_wait(&empty[s], ph ^ 1);
          if (DUAL && it >= iters) {           // the pad position: no data, just complete the phase
            mbar_arrive(&full[s]);
            if (++s == (uint32_t)kStages) {
              s = 0;
              ph ^= 1u;
            }
            continue;
          }
          uint8_t* st = smem + s * p.stage_bytes;
          const int wi = p.tap_dw[tap], hi = h0 * p.stride + p.tap_dh[tap];
          const int kb = p.tap_koff[tap] + kc * p.kc;
          mbar_expect_tx(&full[s], (active ? tx_a : 0u) + tx_b);
          if (active) {
            tma_load_4d(st, &tmA_hi, &full[s], kc * p.kc, wi, hi, n0);
            if (NPASS == 3) tma_load_4d(st + p.a_bytes, &tmA_lo, &full[s], kc * p.kc, wi, hi, n0);
          }
          if (PAIR) {
            // this CTA's half of the weight tile, into both CTAs' stage s
            const int cb = co0 + (int)rank * (BN / 2);
            tma_load_2d_mc(st + p.a_bytes * Cfg::kPlanes + rank * half_bytes, &tmB_hi, &full[s], kb, cb, 3);
            if (NPASS == 3)
              tma_load_2d_mc(st + p.a_bytes * 2 + p.b_bytes + rank * half_bytes, &tmB_lo, &full[s], kb, cb, 3);
          } else {
            tma_load_2d(st + p.a_bytes * Cfg::kPlanes, &tmB_hi, &full[s], kb, co0);
            if (NPASS == 3) tma_load_2d(st + p.a_bytes * 2 + p.b_bytes, &tmB_lo, &full[s], kb, co0);
          }
          if (++kc == p.kchunks) {
            kc = 0;
            ++tap;
          }
          if (++s == (uint32_t)kStages) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1 || (DUAL && warp == 6)) {
    // ===================== MMA issuer =====================
    const int me = warp == 1 ? 0 : 1;     // DUAL: which half of a tile's stages this warp issues
    // instruction descriptor: D fp32, A/B fp16 or bf16 (format field 0 / 1), both K-major,
    // M = 128, N = BN
    const uint32_t idesc = (1u << 4) | ((uint32_t)p.fmt << 7) | ((uint32_t)p.fmt << 10) |
                           ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | ((uint32_t)p.fmt << 7) | ((uint32_t)p.fmt << 10) |
                            ((uint32_t)((2 * BN) >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
    const int ksteps = p.kc >> 4;
    // ONE elected lane runs the whole issue loop, waits included.  Measured (EVE_B200_TC_KC=32: twice
    // the stages per tile made the 64-channel layers 1.44x slower): a ring stage cost ~370 cycles of
    // issuer overhead next to ~480 cycles of MMAs -- an elect / reconverge pair, two runtime
    // divisions, four descriptor encodings and the barrier round trip per stage, none of which
    // overlaps the MMAs because the issuing thread does not run ahead of the tensor pipe.  Now: stage
    // index and phase kept incrementally, descriptors = one hoisted constant + (address >> 4).
    if (elect_one()) {
      const uint64_t dconst = kmajor_desc(0u, p.kc);
      const uint32_t smem16 = smem_u32(smem) >> 4;
      const uint32_t stage16 = (uint32_t)p.stage_bytes >> 4;
      const uint32_t a16 = (uint32_t)p.a_bytes >> 4, b16 = (uint32_t)p.b_bytes >> 4;
      uint32_t s = 0, ph = 0, local = 0;
      for (int tile = item0; tile < total_tiles; tile += item_step) {
        if (PAIR && 2 * (tile / p.tiles_co) + (int)rank >= p.tiles_m) {
          // no tile of its own in this pair-row: the CTA still receives (and supplies half of) the
          // weight stages; release each one for both CTAs as soon as it has landed
          for (int it = 0; it < iters; ++it) {
            mbar_wait(&full[s], ph);
            umma_commit_mc(&empty[s], 3);
            if (++s == (uint32_t)kStages) {
              s = 0;
              ph ^= 1u;
            }
          }
          continue;
        }
        const uint32_t buf = local & 1;
        const uint32_t use = local >> 1;
        ++local;
        mbar_wait(&tmem_empty[buf], (use & 1) ^ 1);     // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (buf * kAccSets + (uint32_t)me) * Cfg::kTmemCols;
        for (int it = 0; it < ring_iters; ++it) {
          if (DUAL && (it & 1) != me) {      // the other issuer's ring position
            if (++s == (uint32_t)kStages) {
              s = 0;
              ph ^= 1u;
            }
            continue;
          }
          const int first = DUAL ? me : 0;   // this issuer's first stage of the tile starts its accumulator
          mbar_wait(&full[s], ph);
          tc_fence_after();
          if (DUAL && it >= iters) {         // the pad position: release it
            umma_commit(&empty[s]);
            if (++s == (uint32_t)kStages) {
              s = 0;
              ph ^= 1u;
            }
            continue;
          }
          const uint64_t da_hi = dconst + (uint64_t)(smem16 + s * stage16);
          const uint64_t da_lo = da_hi + a16;
          const uint64_t db_hi = da_hi + a16 * Cfg::kPlanes;
          const uint64_t db_lo = db_hi + b16;
#pragma unroll
          for (int k = 0; k < 4; ++k) {              // kc <= 64: at most four K = 16 steps per stage
            if (k >= ksteps) break;
            const uint64_t adv = (uint64_t)(k * 2);  // 16 elements = 32 bytes >> 4
            if (Cfg::kStack) {
              // [A_hi B_hi | A_hi B_lo] from one N = 2 BN instruction, then A_lo B_hi
              umma_bf16(tmem_d, da_hi + adv, db_hi + adv, idesc2, ((it - first) | k) != 0);
              umma_bf16(tmem_d, da_lo + adv, db_hi + adv, idesc, 1);
            } else if (NPASS == 3) {
              // small terms first so they are not absorbed by a large partial sum
              umma_bf16(tmem_d, da_lo + adv, db_hi + adv, idesc, ((it - first) | k) != 0);
              umma_bf16(tmem_d, da_hi + adv, db_lo + adv, idesc, 1);
              umma_bf16(tmem_d, da_hi + adv, db_hi + adv, idesc, 1);
            } else {
              umma_bf16(tmem_d, da_hi + adv, db_hi + adv, idesc, ((it - first) | k) != 0);
            }
          }
          if (PAIR) umma_commit_mc(&empty[s], 3);
          else umma_commit(&empty[s]);
          if (!DUAL && it == iters - 1) umma_commit(&tmem_full[buf]);
          if (++s == (uint32_t)kStages) {
            s = 0;
            ph ^= 1u;
          }
        }
        // DUAL: each issuer reports the end of ITS stages (an issuer without a stage -- a single-stage
        // tile -- arrives at once and its accumulator is not read)
        if (DUAL) umma_commit(&tmem_full[buf]);
      }
    }
  } else {
    // ===================== epilogue =====================
    const int quad = warp & 3;              // TMEM lane quadrant this warp may read
    const int m = quad * 32 + lane;         // tile row = TMEM lane
    const int iw = m % p.bw;
    const int ih = (m / p.bw) % p.bh;
    const int in = m / (p.bw * p.bh);
    constexpr int CW = BN < 32 ? BN : 32;   // columns handled per TMEM load
    // A TMEM lane is one output pixel: stored straight from the accumulator registers, every
    // STG.128 of a warp touches 32 different 128-byte lines and half-fills a sector in each (ncu on
    // the 64-channel layers: the epilogue warps were busy 77 % of the time, 42 % of that waiting for
    // the store queue to read its registers, and handed accumulators back late).  Each warp now
    // transposes a 32 pixel x CW channel chunk through a private XOR-swizzled shared-memory tile and
    // does scale / bias / residual and the store with lane-contiguous 16-byte pieces (whole 64- or
    // 128-byte runs per pixel, 512 bytes per instruction); the accumulator goes back to the MMA warp
    // as soon as its last chunk is in registers.
    constexpr int PP = CW / 4;              // 16-byte pieces per pixel and chunk
    constexpr int PXP = 32 / PP;            // pixels per 512-byte pass
    const uint32_t tile_s = smem_u32(epi) + (uint32_t)quad * (32 * 128);
    unsigned long long* s_row = reinterpret_cast<unsigned long long*>(epi + kEpiTileBytes) + quad * 32;
    auto swz = [](int px) { return PP == 4 ? ((px >> 1) & 3) : (px & 7); };
    const int piece = lane % PP, prow = lane / PP;
    const float scale = p.out_scale;        // a power of two: fma(v, scale, bias) rounds as v * scale + bias
    uint32_t local = 0;
    for (int tile = item0; tile < total_tiles; tile += item_step) {
      const int tco = tile % p.tiles_co;
      const int tm = PAIR ? 2 * (tile / p.tiles_co) + (int)rank : tile / p.tiles_co;
      if (PAIR && tm >= p.tiles_m) continue;
      const uint32_t buf = local & 1;
      const uint32_t use = local >> 1;
      ++local;
      const int h = (tm % p.tiles_h) * p.bh + ih;
      const int n = (tm / p.tiles_h) * p.bn + in;
      const int co0 = tco * BN;
      const bool valid = in < p.bn && h < p.OH && n < p.N;
      const size_t row = ((size_t)(n * p.out_H + h * p.out_mul + p.out_ah) * p.out_W +
                          iw * p.out_mul + p.out_aw) * p.Cout + co0;
      __syncwarp();                          // the previous tile's readers are done with s_row
      s_row[lane] = valid ? (unsigned long long)row : ~0ull;
      __syncwarp();
      const uint32_t tmem_d = tmem_base + buf * kAccSets * Cfg::kTmemCols + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        // everything the coalesced side needs from global memory is requested before the accumulator
        // chunk is loaded (the shared-memory accesses below are ordering points for the compiler)
        unsigned long long ro[PP];
        float4 add4[PP];
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias) {
          const float* bp = p.bias + co0 + c0 + piece * 4;
          b4 = make_float4(__ldg(bp), __ldg(bp + 1), __ldg(bp + 2), __ldg(bp + 3));
        }
#pragma unroll
        for (int i = 0; i < PP; ++i) {
          ro[i] = s_row[i * PXP + prow];
          add4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.addend && ro[i] != ~0ull)
            add4[i] = __ldg(reinterpret_cast<const float4*>(p.addend + (size_t)ro[i] + (size_t)(c0 + piece * 4)));
        }
        if (c0 == 0) {
          mbar_wait(&tmem_full[buf], use & 1);
          tc_fence_after();
        }
        float v[32];
        tmem_ld32(tmem_d + (uint32_t)c0, v);
        if (Cfg::kStack) {
          if (BN < 32) {
#pragma unroll
            for (int j = 0; j < BN; ++j) v[j] += v[BN + j];      // both halves came with one load
          } else {
            float u[32];
            tmem_ld32(tmem_d + (uint32_t)(BN + c0), u);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += u[j];
          }
        }
        if (DUAL && iters > 1) {             // the second issuer's partial sums
          float u[32];
          tmem_ld32(tmem_d + Cfg::kTmemCols + (uint32_t)c0, u);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += u[j];
          if (Cfg::kStack) {
            tmem_ld32(tmem_d + Cfg::kTmemCols + (uint32_t)(BN + c0), u);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += u[j];
          }
        }
        if (c0 + 32 >= BN) {
          // all of this thread's TMEM reads have completed (tcgen05.wait::ld): release the buffer
          tc_fence_before();
          mbar_arrive_relaxed(&tmem_empty[buf]);
        }
        __syncwarp();                        // the previous chunk has been read out of the tile
#pragma unroll
        for (int j = 0; j < PP; ++j)
          sts128(tile_s + (uint32_t)(lane * (CW * 4) + ((j ^ swz(lane)) << 4)),
                 make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
        __syncwarp();
#pragma unroll
        for (int i = 0; i < PP; ++i) {
          const int px = i * PXP + prow;
          float4 a = lds128(tile_s + (uint32_t)(px * (CW * 4) + ((piece ^ swz(px)) << 4)));
          if (ro[i] != ~0ull) {
            const size_t o = (size_t)ro[i] + (size_t)(c0 + piece * 4);
            a.x = fmaf(a.x, scale, b4.x);
            a.y = fmaf(a.y, scale, b4.y);
            a.z = fmaf(a.z, scale, b4.z);
            a.w = fmaf(a.w, scale, b4.w);
            if (p.addend) {
              a.x += add4[i].x; a.y += add4[i].y; a.z += add4[i].z; a.w += add4[i].w;
            }
            *reinterpret_cast<float4*>(p.out + o) = a;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();      // nobody leaves while the peer may still write its ring / barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * kAccSets * Cfg::kTmemCols);
  }
}

// ============================================================= halo-row kernel (W == 128) ==
// 3x3 stride-1 "same" convolutions over 128-pixel-wide maps (RefineNet level 0, refine_net.py:
// 45-62,213-222) have so few channels (16..64) that the generic kernel above is bound by the
// L2 -> shared-memory operand stream (every input pixel is re-read once per filter tap, 9x) and
// by the single MMA-issuing thread (a few MMAs per ring stage).  Here an M tile is one output
// image row; input rows are staged ONCE in a ring of row slots (each 130 pixels wide: w = -1..128,
// zero-filled by TMA outside the image) and reused by the three output rows that touch them; the
// three horizontal taps are the SAME staged row read through shared-memory descriptors whose
// start address is shifted by q pixels (measured on B200: the 32/64/128-byte swizzle is a
// function of the absolute shared-memory address, so a row-shifted start needs no base offset).
// All nine [Cout][Cin] weight tiles stay resident in shared memory; channel counts are template
// parameters so that the issuing thread only adds immediates to two descriptor bases per MMA.
// A CTA owns strips of consecutive output rows of one image.
struct TcRowParams {
  int N, H, Cout;
  int strips, rows_per_strip, items;   // work item = (image, strip); items = N * strips
  int fmt;
  int slots;                           // ring depth (>= 4)
  float out_scale;
  const float* bias;
  const float* addend;
  float* out;
};

constexpr int kRowBox = 130;           // 128 pixels + one halo pixel on each side

template <int BN, int NPASS, int KC, int KS>
struct RowCfg {
  static constexpr int kPlanes = NPASS == 3 ? 2 : 1;
  static constexpr int kRowBytes = (kRowBox * KC * 2 + 1023) / 1024 * 1024;   // one plane of one row
  static constexpr int kSlotBytes = kRowBytes * kPlanes;
  static constexpr bool kStack = NPASS == 3 && BN <= 64;       // see TcCfg
  // one tap, one plane; with the stacked-B issue the lo plane must follow the hi plane directly
  static constexpr int kTapBytes = (kStack || BN * KC * 2 >= 1024) ? BN * KC * 2 : 1024;
  static constexpr int kTaps = KS * KS;                 // 3x3 or 1x1
  static constexpr int kHalo = KS / 2;
  static constexpr int kWeightBytes = (kTaps * kPlanes * kTapBytes + 1023) / 1024 * 1024;
  static constexpr int kStageBytes = 4 * 32 * BN * 4;   // epilogue transpose tiles: 4 warps x 32 pixels x BN floats
  static constexpr int kFixedBytes = kWeightBytes + kStageBytes + 1024 + kBarrierBytes;
  static constexpr int kSlotsRaw = (227 * 1024 - kFixedBytes) / kSlotBytes;
  static constexpr int kSlots = kSlotsRaw > kMaxStages ? kMaxStages : kSlotsRaw;
  static constexpr int kAccCols = kStack ? 2 * BN : BN;
  static constexpr uint32_t kTmemCols = kAccCols < 32 ? 32 : kAccCols;
};

template <int BN, int NPASS, int KC, int KS>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_row_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                   const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                   const TcRowParams p) {
  using Cfg = RowCfg<BN, NPASS, KC, KS>;
  constexpr int kPlanes = Cfg::kPlanes;
  constexpr int kHalo = Cfg::kHalo;
  constexpr uint32_t kTmemCols = Cfg::kTmemCols;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  uint8_t* wsm = smem;                                        // [taps][planes][kTapBytes]
  uint8_t* stage = smem + Cfg::kWeightBytes;                  // [4 epilogue warps][32 pixels][BN] fp32
  uint8_t* ring = stage + Cfg::kStageBytes;                   // [slots][planes][kRowBytes]
  const int slots = p.slots;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)slots * Cfg::kSlotBytes);
  uint64_t* empty = full + slots;
  uint64_t* wfull = empty + slots;
  uint64_t* tmem_full = wfull + 1;      // [2]
  uint64_t* tmem_empty = tmem_full + 2; // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tmap_prefetch(&tmA_hi);
    tmap_prefetch(&tmB_hi);
    if (NPASS == 3) {
      tmap_prefetch(&tmA_lo);
      tmap_prefetch(&tmB_lo);
    }
    for (int s = 0; s < slots; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(wfull, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_expect_tx(wfull, (uint32_t)Cfg::kTaps * kPlanes * (uint32_t)(BN * KC * 2));
#pragma unroll 1
      for (int t = 0; t < Cfg::kTaps; ++t) {
        tma_load_2d(wsm + (size_t)(t * kPlanes) * Cfg::kTapBytes, &tmB_hi, wfull, t * KC, 0);
        if (NPASS == 3)
          tma_load_2d(wsm + (size_t)(t * kPlanes + 1) * Cfg::kTapBytes, &tmB_lo, wfull, t * KC, 0);
      }
      constexpr uint32_t tx = (uint32_t)(kRowBox * KC * 2) * kPlanes;
      uint32_t g = 0;   // staged-row counter: slot = g % slots
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const int n = item / p.strips;
        const int h0 = (item - n * p.strips) * p.rows_per_strip;
        const int h1 = min(p.H, h0 + p.rows_per_strip);
        for (int hr = h0 - kHalo; hr < h1 + kHalo; ++hr, ++g) {
          const int s = g % slots;
          const uint32_t ph = (g / slots) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* dst = ring + (size_t)s * Cfg::kSlotBytes;
          mbar_expect_tx(&full[s], tx);
          tma_load_4d(dst, &tmA_hi, &full[s], 0, -1, hr, n);
          if (NPASS == 3) tma_load_4d(dst + Cfg::kRowBytes, &tmA_lo, &full[s], 0, -1, hr, n);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = (1u << 4) | ((uint32_t)p.fmt << 7) | ((uint32_t)p.fmt << 10) |
                           ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | ((uint32_t)p.fmt << 7) | ((uint32_t)p.fmt << 10) |
                            ((uint32_t)((2 * BN) >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
    constexpr int kSteps = KC / 16;
    constexpr uint32_t kPix16 = (uint32_t)(KC * 2) >> 4;        // one pixel, in 16-byte units
    constexpr uint32_t kRow16 = (uint32_t)Cfg::kRowBytes >> 4;
    constexpr uint32_t kTap16 = (uint32_t)Cfg::kTapBytes >> 4;
    mbar_wait(wfull, 0);
    tc_fence_after();
    // descriptor = constant high part + (shared address >> 4); all later offsets are immediates
    const uint64_t desc0 = kmajor_desc(0u, KC);
    const uint64_t b_base = desc0 + (uint64_t)(smem_u32(wsm) >> 4);
    const uint32_t ring16 = smem_u32(ring) >> 4;
    // ring position of entry g kept incrementally (slot, phase): no runtime division in the issue
    // loop, and a staged row is waited for once, not by each of the three output rows that read it
    // One elected lane runs the whole loop (waits included): a per-row elect / reconverge pair
    // around each batch of MMAs costs more issue slots than the MMAs themselves at 16 channels.
    if (elect_one()) {
      uint32_t local = 0, sg = 0, pg = 0;
      int confirmed = 0;              // entries g .. g + confirmed - 1 are known to have landed
      auto advance = [&](uint32_t& sl, uint32_t& ph) {
        if (++sl == (uint32_t)slots) {
          sl = 0;
          ph ^= 1u;
        }
      };
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const int n = item / p.strips;
        const int h0 = (item - n * p.strips) * p.rows_per_strip;
        const int h1 = min(p.H, h0 + p.rows_per_strip);
        for (int h = h0; h < h1; ++h, ++local) {
          const uint32_t buf = local & 1;
          const uint32_t use = local >> 1;
          mbar_wait(&tmem_empty[buf], (use & 1) ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + buf * kTmemCols;
          uint32_t se = sg, pe = pg;
#pragma unroll
          for (int r = 0; r < KS; ++r) {
            const uint32_t s = se;                       // staged row h + r - halo
            if (r >= confirmed) {
              mbar_wait(&full[s], pe);
              tc_fence_after();
            }
            const uint64_t a_base = desc0 + (uint64_t)(ring16 + s * ((uint32_t)Cfg::kSlotBytes >> 4));
#pragma unroll
            for (int q = 0; q < KS; ++q) {
#pragma unroll
              for (int k = 0; k < kSteps; ++k) {
                constexpr uint32_t two = 2;
                // the staged box starts at pixel -1: tap q reads it shifted by q + 1 - halo pixels
                const uint64_t da_hi = a_base + (uint64_t)((q + 1 - kHalo) * kPix16 + k * two);
                const uint64_t da_lo = da_hi + kRow16;
                const uint64_t db_hi = b_base + (uint64_t)((r * KS + q) * kPlanes * kTap16 + k * two);
                const uint64_t db_lo = db_hi + kTap16;
                if (Cfg::kStack) {
                  umma_bf16(tmem_d, da_hi, db_hi, idesc2, (r | q | k) != 0);
                  umma_bf16(tmem_d, da_lo, db_hi, idesc, 1);
                } else if (NPASS == 3) {
                  umma_bf16(tmem_d, da_lo, db_hi, idesc, (r | q | k) != 0);
                  umma_bf16(tmem_d, da_hi, db_lo, idesc, 1);
                  umma_bf16(tmem_d, da_hi, db_hi, idesc, 1);
                } else {
                  umma_bf16(tmem_d, da_hi, db_hi, idesc, (r | q | k) != 0);
                }
              }
            }
            if (r == KS - 1) {
              umma_commit(&empty[sg]);                  // row h - halo is not needed any more
              umma_commit(&tmem_full[buf]);
            }
            advance(se, pe);
          }
          advance(sg, pg);
          confirmed = KS - 1;
        }
        // end of the strip: the trailing halo rows (h1 - 1, h1) are released as well
        if (KS == 3) {
          umma_commit(&empty[sg]);
          advance(sg, pg);
          umma_commit(&empty[sg]);
          advance(sg, pg);
        }
        confirmed = 0;
      }
    }
  } else {
    // ===================== epilogue =====================
    const int quad = warp & 3;
    const int m = quad * 32 + lane;         // output pixel w = TMEM lane
    // A TMEM lane is one output pixel, so the accumulator registers of a thread are BN consecutive
    // floats of ITS pixel: stored directly, every STG.128 of a warp touches 32 different sectors and
    // half-fills each (measured: the L1 -> crossbar request path was the busiest unit of the kernel and
    // the next row's instructions waited ~500 cycles for the store queue to read its registers).  The
    // warp therefore transposes through a private shared-memory tile -- raw accumulators in, XOR-
    // swizzled by pixel so that both sides are bank-conflict free -- and does bias / scale / residual
    // and the store with lane-contiguous 16-byte pieces: 512 contiguous bytes per instruction.  The
    // accumulator is handed back to the MMA warp as soon as it is in registers.
    constexpr int RB = BN * 4;                    // bytes of one output pixel
    constexpr int NCH = BN / 4;                   // 16-byte pieces per pixel == 512-byte passes per warp tile
    const uint32_t tile = smem_u32(stage) + (uint32_t)quad * (32 * RB);
    auto swz = [](int px) { return BN == 16 ? ((px >> 1) & 3) : (px & 7); };
    // coalesced side: pass i, lane l holds piece c_l of pixel px0 + i * ppi
    constexpr int ppi = 512 / RB;                 // pixels per pass
    const int c_l = lane % NCH;
    const int px0 = lane / NCH;
    float bias4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) bias4[j] = p.bias ? __ldg(p.bias + c_l * 4 + j) : 0.f;
    const float scale = p.out_scale;              // a power of two: fma(v, scale, bias) rounds once, as v * scale + bias did
    uint32_t local = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const int n = item / p.strips;
      const int h0 = (item - n * p.strips) * p.rows_per_strip;
      const int h1 = min(p.H, h0 + p.rows_per_strip);
      for (int h = h0; h < h1; ++h, ++local) {
        const uint32_t buf = local & 1;
        const uint32_t use = local >> 1;
        // this warp's 32 pixels x BN floats are contiguous in NHWC
        const size_t base = ((size_t)(n * p.H + h) * kTileM + quad * 32) * BN + (size_t)lane * 4;
        // the residual addend does not depend on the accumulator: fetch it before waiting
        float4 add4[NCH];
        if (p.addend) {
          const float4* a4 = reinterpret_cast<const float4*>(p.addend + base);
#pragma unroll
          for (int i = 0; i < NCH; ++i) add4[i] = __ldg(a4 + i * 32);
        }
        mbar_wait(&tmem_full[buf], use & 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * kTmemCols + ((uint32_t)(quad * 32) << 16);
#pragma unroll
        for (int c0 = 0; c0 < BN; c0 += 32) {
          constexpr int CW = BN < 32 ? BN : 32;
          float v[32];
          tmem_ld32(tmem_d + (uint32_t)c0, v);
          if (Cfg::kStack) {
            if (BN < 32) {
#pragma unroll
              for (int j = 0; j < BN; ++j) v[j] += v[BN + j];
            } else {
              float u[32];
              tmem_ld32(tmem_d + (uint32_t)(BN + c0), u);
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] += u[j];
            }
          }
          if (c0 + 32 >= BN) {
            tc_fence_before();
            mbar_arrive_relaxed(&tmem_empty[buf]);
          }
#pragma unroll
          for (int j = 0; j < CW / 4; ++j) {
            const int c = c0 / 4 + j;
            sts128(tile + (uint32_t)(lane * RB + ((c ^ swz(lane)) << 4)),
                   make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
          }
        }
        __syncwarp();
        float4* o4 = reinterpret_cast<float4*>(p.out + base);
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
          const int px = px0 + i * ppi;
          float4 a = lds128(tile + (uint32_t)(px * RB + ((c_l ^ swz(px)) << 4)));
          a.x = fmaf(a.x, scale, bias4[0]);
          a.y = fmaf(a.y, scale, bias4[1]);
          a.z = fmaf(a.z, scale, bias4[2]);
          a.w = fmaf(a.w, scale, bias4[3]);
          if (p.addend) {
            a.x += add4[i].x; a.y += add4[i].y; a.z += add4[i].z; a.w += add4[i].w;
          }
          o4[i * 32] = a;
        }
        __syncwarp();          // the tile is rewritten by the next row
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * kTmemCols);
  }
}

// ================================================================ padded-strip kernel ==
// 3x3 stride-1 "same" convolutions whose maps are narrower than 128 pixels (EyeNet layer1,
// RefineNet levels 1-3: eye_net.py:48-50, refine_net.py:45-62).  The generic kernel above loads one
// A box per filter tap, i.e. every input pixel crosses L2 -> shared memory nine times; at 64
// channels that stream (125 B/clk/SM needed, ~42 available) and not the tensor pipe sets the pace.
// Here a work item stages its input ONCE per K chunk as a zero-padded strip: a TMA box
// [bn images][R + 2 rows][W + 2 pixels][KC channels] starting at (w, h) = (-1, h0 - 1) lands in
// shared memory as consecutive pixel rows of KC channels, row-major in a space padded to W + 2
// columns.  An M tile is 128 CONSECUTIVE positions of that padded space, and filter tap (r, q) of
// all of them is the same staged strip read through a descriptor whose start address is shifted by
// r * (W + 2) + q pixels (the swizzle is a function of the absolute shared-memory address, DESIGN
// 3b).  The two pad positions per row (and the halo rows between images when bn > 1) produce junk
// accumulator rows that the epilogue skips.  Each weight tile [BN][KC] of a (tap, chunk) is
// streamed once per item through a small ring and used by all T tiles of the item (T accumulators
// in TMEM), so operand traffic per item is  A * (R + 2) / R  +  9 taps * B  instead of 9 * (A + B)
// per tile.  Five roles: A producer, B producer, MMA issuer, four epilogue warps.
struct TcStripParams {
  int N, H, W, Cout, Wp;
  int R, bn;                 // output rows per strip; images per item (bn > 1 only when R == H)
  int strips, groups;        // ceil(H / R), ceil(N / bn)
  int S;                     // (R + 2) * Wp: staged pixel rows per image
  int tiles_co, items;       // Cout / BN; groups * strips * tiles_co
  int kchunks;               // Cin / KC
  int Tmax, nsets;           // accumulators per set, sets in TMEM (2 = epilogue overlaps the next item)
  int a_plane_bytes;         // one plane of one A buffer
  int b_stages;
  int fmt;
  float out_scale;
  const float* bias;
  const float* addend;
  float* out;
};

constexpr int kStripThreads = 224;      // 7 warps

template <int BN, int KC>
__global__ void __launch_bounds__(kStripThreads, 1)
conv_tc_strip_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                     const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                     const TcStripParams p) {
  constexpr int kBPlane = BN * KC * 2 < 1024 ? 1024 : BN * KC * 2;    // one weight plane of a stage
  constexpr bool kStack = BN <= 64;                                   // stacked-B issue, see TcCfg
  static_assert(!kStack || kBPlane == BN * KC * 2, "stacked B needs the lo plane right behind hi");
  constexpr int kAcc = kStack ? 2 * BN : BN;                          // accumulator columns per tile
  constexpr uint32_t kPix16 = (uint32_t)(KC * 2) >> 4;                // one pixel row, 16-byte units
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  uint8_t* a_buf = smem;                                         // [2 buffers][2 planes][a_plane_bytes]
  uint8_t* b_ring = smem + 4 * (size_t)p.a_plane_bytes;          // [b_stages][2 planes][kBPlane]
  uint64_t* a_full = reinterpret_cast<uint64_t*>(b_ring + (size_t)p.b_stages * 2 * kBPlane);
  uint64_t* a_empty = a_full + 2;
  uint64_t* b_full = a_empty + 2;
  uint64_t* b_empty = b_full + p.b_stages;
  uint64_t* tmem_full = b_empty + p.b_stages;    // [2]
  uint64_t* tmem_empty = tmem_full + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint8_t* epi = reinterpret_cast<uint8_t*>(a_full) + kBarrierBytes;   // transpose tiles + row offsets

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(p.nsets * p.Tmax * kAcc)) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) {
    tmap_prefetch(&tmA_hi);
    tmap_prefetch(&tmA_lo);
    tmap_prefetch(&tmB_hi);
    tmap_prefetch(&tmB_lo);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 128);
    }
    for (int i = 0; i < p.b_stages; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== A producer: one padded strip per (item, K chunk) =====================
    if (elect_one()) {
      const uint32_t tx = (uint32_t)(p.bn * p.S) * (uint32_t)(KC * 2) * 2u;
      uint32_t ga = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const int si = item / p.tiles_co;
        const int h0 = (si % p.strips) * p.R;
        const int n0 = (si / p.strips) * p.bn;
        for (int c = 0; c < p.kchunks; ++c, ++ga) {
          const uint32_t buf = ga & 1;
          mbar_wait(&a_empty[buf], ((ga >> 1) & 1) ^ 1);
          uint8_t* dst = a_buf + (size_t)buf * 2 * p.a_plane_bytes;
          mbar_expect_tx(&a_full[buf], tx);
          tma_load_4d(dst, &tmA_hi, &a_full[buf], c * KC, -1, h0 - 1, n0);
          tma_load_4d(dst + p.a_plane_bytes, &tmA_lo, &a_full[buf], c * KC, -1, h0 - 1, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== B producer: one weight tile per (item, K chunk, tap) =====================
    if (elect_one()) {
      constexpr uint32_t tx = (uint32_t)(BN * KC * 2) * 2u;
      const int Cin = p.kchunks * KC;
      uint32_t gb = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const int co0 = (item % p.tiles_co) * BN;
        for (int c = 0; c < p.kchunks; ++c) {
          for (int tap = 0; tap < 9; ++tap, ++gb) {
            const uint32_t s = gb % (uint32_t)p.b_stages;
            mbar_wait(&b_empty[s], ((gb / (uint32_t)p.b_stages) & 1) ^ 1);
            uint8_t* dst = b_ring + (size_t)s * 2 * kBPlane;
            mbar_expect_tx(&b_full[s], tx);
            tma_load_2d(dst, &tmB_hi, &b_full[s], tap * Cin + c * KC, co0);
            tma_load_2d(dst + kBPlane, &tmB_lo, &b_full[s], tap * Cin + c * KC, co0);
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = (1u << 4) | ((uint32_t)p.fmt << 7) | ((uint32_t)p.fmt << 10) |
                           ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | ((uint32_t)p.fmt << 7) | ((uint32_t)p.fmt << 10) |
                            ((uint32_t)((2 * BN) >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
    constexpr int kSteps = KC / 16;
    const uint64_t desc0 = kmajor_desc(0u, KC);
    const uint32_t a16 = smem_u32(a_buf) >> 4, b16 = smem_u32(b_ring) >> 4;
    const uint32_t aplane16 = (uint32_t)p.a_plane_bytes >> 4;
    constexpr uint32_t bplane16 = (uint32_t)kBPlane >> 4;
    // one elected lane runs the whole loop, waits included; ring position kept incrementally (see
    // conv_tc_kernel: per-stage issuer overhead does not overlap the MMAs)
    if (elect_one()) {
      uint32_t ga = 0, sb = 0, pb = 0, local = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++local) {
        const int si = item / p.tiles_co;
        const int h0 = (si % p.strips) * p.R;
        const int n0 = (si / p.strips) * p.bn;
        const int rows = min(p.R, p.H - h0), imgs = min(p.bn, p.N - n0);
        const int T = ((imgs - 1) * p.S + rows * p.Wp + kTileM - 1) / kTileM;
        const uint32_t set = p.nsets == 2 ? (local & 1) : 0u;
        const uint32_t use = p.nsets == 2 ? (local >> 1) : local;
        mbar_wait(&tmem_empty[set], (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_set = tmem_base + set * (uint32_t)(p.Tmax * kAcc);
        for (int c = 0; c < p.kchunks; ++c, ++ga) {
          const uint32_t buf = ga & 1;
          mbar_wait(&a_full[buf], (ga >> 1) & 1);
          const uint64_t a_desc = desc0 + (uint64_t)(a16 + buf * 2 * aplane16);
          uint32_t tap_off = 0;                     // (r * Wp + q) pixels
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&b_full[sb], pb);
            tc_fence_after();
            const uint64_t a_tap = a_desc + (uint64_t)(tap_off * kPix16);
            const uint64_t b_tap = desc0 + (uint64_t)(b16 + sb * 2 * bplane16);
            for (int t = 0; t < T; ++t) {
              const uint64_t a_t = a_tap + (uint64_t)((uint32_t)t * (uint32_t)kTileM * kPix16);
              const uint32_t tmem_d = tmem_set + (uint32_t)(t * kAcc);
#pragma unroll
              for (int k = 0; k < kSteps; ++k) {
                const uint64_t da_hi = a_t + (uint64_t)(k * 2);
                const uint64_t da_lo = da_hi + aplane16;
                const uint64_t db_hi = b_tap + (uint64_t)(k * 2);
                const uint64_t db_lo = db_hi + bplane16;
                if (kStack) {
                  umma_bf16(tmem_d, da_hi, db_hi, idesc2, (c | tap | k) != 0);
                  umma_bf16(tmem_d, da_lo, db_hi, idesc, 1);
                } else {
                  // small terms first so they are not absorbed by a large partial sum
                  umma_bf16(tmem_d, da_lo, db_hi, idesc, (c | tap | k) != 0);
                  umma_bf16(tmem_d, da_hi, db_lo, idesc, 1);
                  umma_bf16(tmem_d, da_hi, db_hi, idesc, 1);
                }
              }
            }
            umma_commit(&b_empty[sb]);
            if (tap == 8) {
              umma_commit(&a_empty[buf]);
              if (c == p.kchunks - 1) umma_commit(&tmem_full[set]);
            }
            if (++sb == (uint32_t)p.b_stages) {
              sb = 0;
              pb ^= 1u;
            }
            // next tap: one pixel to the right, or to the start of the next padded row
            tap_off += (tap % 3 == 2) ? (uint32_t)(p.Wp - 2) : 1u;
          }
        }
      }
    }
  } else {
    // ===================== epilogue (warps 3..6: TMEM lane quadrant = warp & 3) =====================
    // as in conv_tc_kernel: 32-channel chunks transposed through a per-warp swizzled shared-memory
    // tile, scale / bias / residual and the stores on lane-contiguous 16-byte pieces
    const int quad = warp & 3;
    const int m = quad * 32 + lane;
    const uint32_t tile_s = smem_u32(epi) + (uint32_t)quad * (32 * 128);
    unsigned long long* s_row = reinterpret_cast<unsigned long long*>(epi + kEpiTileBytes) + quad * 32;
    const int piece = lane & 7, prow = lane >> 3;
    const float scale = p.out_scale;
    uint32_t local = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++local) {
      const int tco = item % p.tiles_co;
      const int si = item / p.tiles_co;
      const int h0 = (si % p.strips) * p.R;
      const int n0 = (si / p.strips) * p.bn;
      const int rows = min(p.R, p.H - h0), imgs = min(p.bn, p.N - n0);
      const int T = ((imgs - 1) * p.S + rows * p.Wp + kTileM - 1) / kTileM;
      const int co0 = tco * BN;
      const uint32_t set = p.nsets == 2 ? (local & 1) : 0u;
      const uint32_t use = p.nsets == 2 ? (local >> 1) : local;
      const uint32_t tmem_set = tmem_base + set * (uint32_t)(p.Tmax * kAcc) + ((uint32_t)(quad * 32) << 16);
      for (int t = 0; t < T; ++t) {
        const int j = t * kTileM + m;            // position in the padded space of the item
        const int i = j / p.S;
        const int rem = j - i * p.S;
        const int hh = rem / p.Wp;
        const int ww = rem - hh * p.Wp;
        const bool valid = i < imgs && hh < rows && ww < p.W;
        const size_t row = (((size_t)(n0 + i) * p.H + (h0 + hh)) * p.W + ww) * p.Cout + co0;
        __syncwarp();                            // the previous tile's readers are done with s_row
        s_row[lane] = valid ? (unsigned long long)row : ~0ull;
        __syncwarp();
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          unsigned long long ro[8];
          float4 add4[8];
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.bias) {
            const float* bp = p.bias + co0 + c0 + piece * 4;
            b4 = make_float4(__ldg(bp), __ldg(bp + 1), __ldg(bp + 2), __ldg(bp + 3));
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            ro[q] = s_row[q * 4 + prow];
            add4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.addend && ro[q] != ~0ull)
              add4[q] = __ldg(reinterpret_cast<const float4*>(p.addend + (size_t)ro[q] + (size_t)(c0 + piece * 4)));
          }
          if (t == 0 && c0 == 0) {
            mbar_wait(&tmem_full[set], use & 1);
            tc_fence_after();
          }
          float v[32];
          tmem_ld32(tmem_set + (uint32_t)(t * kAcc + c0), v);
          if (kStack) {
            float u[32];
            tmem_ld32(tmem_set + (uint32_t)(t * kAcc + BN + c0), u);
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) v[jj] += u[jj];
          }
          if (t == T - 1 && c0 + 32 >= BN) {
            tc_fence_before();
            mbar_arrive_relaxed(&tmem_empty[set]);   // the item's accumulators are all in registers
          }
          __syncwarp();                          // the previous chunk has been read out of the tile
#pragma unroll
          for (int q = 0; q < 8; ++q)
            sts128(tile_s + (uint32_t)(lane * 128 + ((q ^ (lane & 7)) << 4)),
                   make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
          __syncwarp();
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int px = q * 4 + prow;
            float4 a = lds128(tile_s + (uint32_t)(px * 128 + ((piece ^ (px & 7)) << 4)));
            if (ro[q] != ~0ull) {
              const size_t o = (size_t)ro[q] + (size_t)(c0 + piece * 4);
              a.x = fmaf(a.x, scale, b4.x);
              a.y = fmaf(a.y, scale, b4.y);
              a.z = fmaf(a.z, scale, b4.z);
              a.w = fmaf(a.w, scale, b4.w);
              if (p.addend) {
                a.x += add4[q].x; a.y += add4[q].y; a.z += add4[q].z; a.w += add4[q].w;
              }
              *reinterpret_cast<float4*>(p.out + o) = a;
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
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ================================================================= persistent ConvGRU ==
// CGRUCell (common.py:388-415) over a whole sequence in ONE kernel, for the 64-feature 5x8
// bottleneck of GazeRefineNet (refine_net.py:151-176):
//     r, z = sigmoid(conv3x3(cat[x, h]; W1) + b1)        n = tanh(conv3x3(cat[r * h, x]; W2) + b2)
//     h'   = (1 - z) n + z h
// The x halves of both convolutions do not depend on the recurrence: the caller evaluates them for
// all T steps at once (gx1 = conv(x; W1[:, :nf]) + b1, gx2 = conv(x; W2[:, nf:]) + b2, two batched
// launches).  What is left per step -- conv(h; W1[:, nf:]) and conv(r * h; W2[:, :nf]) on a 5x8 map
// -- is 0.3 GFLOP but strictly sequential; as separate launches it cost ~40 us per step.  Here one
// CTA owns one clip and walks its T steps:
//   * the hidden state lives in shared memory as the fp16 hi/lo operand planes of a zero-padded
//     7x10 strip (so the nine taps are descriptor shifts, as in the strip kernel) and in registers
//     as fp32;
//   * D^T[out channel][position] = W[out channel][K] . act[position][K]: the weights are the M = 128
//     operand (streamed tap by tap from L2 through a TMA ring, they never change), the 48 padded
//     positions the N operand, so every TMEM lane is an output channel and the gate math runs on
//     all 128 epilogue threads with coalesced channel-major global accesses;
//   * the epilogue threads write r * h and h' straight back into the operand planes (manual
//     128-byte swizzle) and hand over to the MMA warp through mbarriers: no global round trip and
//     no launch inside the time loop.
// The per-step tape the BPTT kernels read (r, z, n, h and the h-halves of xh / cat2) is written
// from the same registers.
struct CgruSeqParams {
  int B, T;
  const float* gx1;     // [T][B][40][128]  x-half of the gate convolution + bias
  const float* gx2;     // [T][B][40][64]   x-half of the candidate convolution + bias
  const float* h0;      // [B][40][64]
  float *r, *z, *n, *h; // [T][B][40][64]
  float* xh;            // [T][B][40][128]: this kernel fills channels [64, 128) with h_{t-1}
  float* cat2;          // [T][B][40][128]: this kernel fills channels [0, 64) with r * h_{t-1}
  float out_scale;      // undoes the 2^6 pre-scale of the fp16 weight planes
};

constexpr int kCgNf = 64, kCgH = 5, kCgW = 8, kCgWp = kCgW + 2, kCgPix = kCgH * kCgW;
constexpr int kCgPos = 48;                       // padded positions covered (last valid one is 47)
constexpr int kCgActRows = 88;                   // staged rows one operand plane may be read at
constexpr int kCgActPlane = kCgActRows * 128;    // bytes (1024-aligned)
constexpr int kCgStageBytes = 2 * 128 * kCgNf * 2;   // [hi][lo] x 128 out channels x 64 in channels
constexpr int kCgStages = 5;
constexpr int kCgSmem = kCgStages * kCgStageBytes + 4 * kCgActPlane + kCgPix * kCgNf * 4 + 1024 + 512;

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// value v of channel c at padded row p of an operand plane pair (K-major, SWIZZLE_128B)
__device__ __forceinline__ void cg_store_act(uint8_t* plane_hi, int p, int c, float v) {
  const __half hh = __float2half_rn(v);
  const __half ll = __float2half_rn(v - __half2float(hh));
  const uint32_t off = (uint32_t)p * 128u + ((((uint32_t)c >> 3) ^ ((uint32_t)p & 7u)) << 4) +
                       ((uint32_t)c & 7u) * 2u;
  *reinterpret_cast<__half*>(plane_hi + off) = hh;
  *reinterpret_cast<__half*>(plane_hi + kCgActPlane + off) = ll;
}

__global__ void __launch_bounds__(kThreads, 1)
cgru_seq_fwd_kernel(const __grid_constant__ CUtensorMap tmW1_hi, const __grid_constant__ CUtensorMap tmW1_lo,
                    const __grid_constant__ CUtensorMap tmW2_hi, const __grid_constant__ CUtensorMap tmW2_lo,
                    const CgruSeqParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  uint8_t* ring = smem;                                        // [stages][hi | lo][128 x 64 fp16]
  uint8_t* act1 = ring + kCgStages * kCgStageBytes;            // h planes      [hi | lo][88 rows]
  uint8_t* act2 = act1 + 2 * kCgActPlane;                      // r * h planes
  float* zbuf = reinterpret_cast<float*>(act2 + 2 * kCgActPlane);   // [40][64]
  uint64_t* full = reinterpret_cast<uint64_t*>(zbuf + kCgPix * kCgNf);
  uint64_t* empty = full + kCgStages;
  uint64_t* d_full = empty + kCgStages;      // [2]: gate / candidate accumulators complete
  uint64_t* act_ready = d_full + 2;          // [2]: h planes / r*h planes written
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(act_ready + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x;
  const int B = p.B, T = p.T;

  if (warp == 0 && lane == 0) {
    tmap_prefetch(&tmW1_hi);
    tmap_prefetch(&tmW1_lo);
    tmap_prefetch(&tmW2_hi);
    tmap_prefetch(&tmW2_lo);
    for (int i = 0; i < kCgStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&d_full[i], 1);
      mbar_init(&act_ready[i], 128);
    }
    fence_barrier_init();
  }
  // pad rows / columns of the operand planes stay zero for the whole kernel
  for (int i = threadIdx.x; i < 4 * kCgActPlane / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(act1)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (warp == 1) tmem_alloc(tmem_slot, 128);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== weight producer: 18 tap tiles per step, forever the same =====================
    if (elect_one()) {
      uint32_t g = 0;
      for (int t = 0; t < T; ++t) {
        for (int conv = 0; conv < 2; ++conv) {
          const CUtensorMap* mh = conv == 0 ? &tmW1_hi : &tmW2_hi;
          const CUtensorMap* ml = conv == 0 ? &tmW1_lo : &tmW2_lo;
          const uint32_t tx = conv == 0 ? 2u * 128u * kCgNf * 2u : 2u * 64u * kCgNf * 2u;
          for (int tap = 0; tap < 9; ++tap, ++g) {
            const uint32_t s = g % kCgStages;
            mbar_wait(&empty[s], ((g / kCgStages) & 1) ^ 1);
            uint8_t* dst = ring + (size_t)s * kCgStageBytes;
            mbar_expect_tx(&full[s], tx);
            tma_load_2d(dst, mh, &full[s], tap * kCgNf, 0);
            tma_load_2d(dst + kCgStageBytes / 2, ml, &full[s], tap * kCgNf, 0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // D fp32, A = weights (fp16, K-major, M = 128), B = activations (fp16, K-major, N = 48)
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(kCgPos >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
    const uint64_t desc0 = kmajor_desc(0u, 64);
    const uint32_t ring16 = smem_u32(ring) >> 4;
    const uint32_t act16[2] = {smem_u32(act1) >> 4, smem_u32(act2) >> 4};
    constexpr uint32_t kLo16W = (uint32_t)(kCgStageBytes / 2) >> 4;
    constexpr uint32_t kLo16A = (uint32_t)kCgActPlane >> 4;
    if (elect_one()) {               // one lane runs the whole loop, waits included
      uint32_t s = 0, ph = 0;
      for (int t = 0; t < T; ++t) {
        for (int conv = 0; conv < 2; ++conv) {
          mbar_wait(&act_ready[conv], (uint32_t)t & 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + (uint32_t)(conv * 64);
          uint32_t tap_off = 0;        // (r * Wp + q) padded positions
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&full[s], ph);
            tc_fence_after();
            const uint64_t w_hi = desc0 + (uint64_t)(ring16 + s * ((uint32_t)kCgStageBytes >> 4));
            const uint64_t a_hi = desc0 + (uint64_t)(act16[conv] + tap_off * 8u);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t dw_hi = w_hi + (uint64_t)(k * 2), dw_lo = dw_hi + kLo16W;
              const uint64_t da_hi = a_hi + (uint64_t)(k * 2), da_lo = da_hi + kLo16A;
              umma_bf16(tmem_d, dw_lo, da_hi, idesc, (tap | k) != 0);
              umma_bf16(tmem_d, dw_hi, da_lo, idesc, 1);
              umma_bf16(tmem_d, dw_hi, da_hi, idesc, 1);
            }
            umma_commit(&empty[s]);
            if (tap == 8) umma_commit(&d_full[conv]);
            if (++s == (uint32_t)kCgStages) {
              s = 0;
              ph ^= 1u;
            }
            tap_off += (tap % 3 == 2) ? (uint32_t)(kCgWp - 2) : 1u;
          }
        }
      }
    }
  } else {
    // ===================== gates (128 threads: TMEM lane = output channel) =====================
    const int quad = warp & 3;
    const int co = quad * 32 + lane;           // gate convolution: r for co < 64, z for co >= 64
    const int c = co & 63;
    const bool is_r = co < 64;
    const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
    float h[kCgPix];
    // initial state -> registers, operand planes and the h-half of xh[0]
    if (is_r) {
#pragma unroll
      for (int pix = 0; pix < kCgPix; ++pix) {
        const float v = __ldg(p.h0 + ((size_t)b * kCgPix + pix) * kCgNf + c);
        h[pix] = v;
        cg_store_act(act1, (pix / kCgW + 1) * kCgWp + (pix % kCgW) + 1, c, v);
        p.xh[((size_t)b * kCgPix + pix) * (2 * kCgNf) + kCgNf + c] = v;
      }
    }
    fence_proxy_async();
    mbar_arrive(&act_ready[0]);
    for (int t = 0; t < T; ++t) {
      const size_t frame = (size_t)t * B + b;
      // ---- r, z
      float gx[kCgPix];
#pragma unroll
      for (int pix = 0; pix < kCgPix; ++pix)
        gx[pix] = __ldg(p.gx1 + (frame * kCgPix + pix) * (2 * kCgNf) + co);
      mbar_wait(&d_full[0], (uint32_t)t & 1);
      tc_fence_after();
      float v[kCgPos];
      tmem_ld32(tlane, v);
      tmem_ld16(tlane + 32u, v + 32);
      // critical path first: the values the next convolution needs go to shared memory and the
      // MMA warp is released; the tape stores to global memory follow while it is already issuing
#pragma unroll
      for (int j = 0; j < kCgPos; ++j) {
        if (j % kCgWp < kCgW) {
          const int pix = (j / kCgWp) * kCgW + j % kCgWp;
          const float s = 1.f / (1.f + expf(-(v[j] * p.out_scale + gx[pix])));
          v[j] = s;
          if (is_r) cg_store_act(act2, j + kCgWp + 1, c, s * h[pix]);
          else zbuf[pix * kCgNf + c] = s;
        }
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&act_ready[1]);
      asm volatile("bar.sync 1, 128;" ::: "memory");      // z (threads 64..127) -> zbuf -> threads 0..63
#pragma unroll
      for (int j = 0; j < kCgPos; ++j) {
        if (j % kCgWp < kCgW) {
          const int pix = (j / kCgWp) * kCgW + j % kCgWp;
          const size_t o = (frame * kCgPix + pix) * kCgNf + c;
          if (is_r) {
            p.r[o] = v[j];
            p.cat2[(frame * kCgPix + pix) * (2 * kCgNf) + c] = v[j] * h[pix];
          } else {
            p.z[o] = v[j];
          }
        }
      }
      // ---- candidate and the new state (threads 0..63)
      if (is_r) {
#pragma unroll
        for (int pix = 0; pix < kCgPix; ++pix)
          gx[pix] = __ldg(p.gx2 + (frame * kCgPix + pix) * kCgNf + c);
      }
      mbar_wait(&d_full[1], (uint32_t)t & 1);
      tc_fence_after();
      if (is_r) {
        tmem_ld32(tlane + 64u, v);
        tmem_ld16(tlane + 96u, v + 32);
#pragma unroll
        for (int j = 0; j < kCgPos; ++j) {
          if (j % kCgWp < kCgW) {
            const int pix = (j / kCgWp) * kCgW + j % kCgWp;
            const float nv = tanhf(v[j] * p.out_scale + gx[pix]);
            const float zv = zbuf[pix * kCgNf + c];
            const float hn = (1.f - zv) * nv + zv * h[pix];
            v[j] = nv;
            h[pix] = hn;
            cg_store_act(act1, j + kCgWp + 1, c, hn);
          }
        }
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&act_ready[0]);
      if (is_r) {
#pragma unroll
        for (int j = 0; j < kCgPos; ++j) {
          if (j % kCgWp < kCgW) {
            const int pix = (j / kCgWp) * kCgW + j % kCgWp;
            const size_t o = (frame * kCgPix + pix) * kCgNf + c;
            p.n[o] = v[j];
            p.h[o] = h[pix];
            if (t + 1 < T)
              p.xh[(((size_t)(t + 1) * B + b) * kCgPix + pix) * (2 * kCgNf) + kCgNf + c] = h[pix];
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

// ---- backward through time of the same cell (BPTT), again one CTA per clip.  Per step t (T-1..0):
//     d    = dL/dh_t (from the decoder) + carry
//     dg2  = d (1 - z)(1 - n^2)        dz' = d (h_{t-1} - n) z (1 - z)        carry' = d z
//     drh  = dgrad(conv2 | r*h half)(dg2)      dr' = drh h_{t-1} r (1 - r)     carry' += drh r
//     carry' += dgrad(conv1 | h half)([dr', dz'])
// Only the h halves of the two data gradients sit in the recurrence; the x halves (the gradient
// handed to the encoder) and both weight gradients are evaluated afterwards, batched over all T
// steps, from the dg1 / dg2 sequences this kernel stores.  D^T[in channel][position] =
// Wd[in channel][(tap, out channel)] . dg[position + tap][out channel]: the flipped, transposed
// filters are the M operand (64 real rows), the gate gradients -- written by the epilogue threads
// as bf16 hi/lo operand planes of a zero-padded 7x10 strip -- the N operand.
struct CgruSeqBwdParams {
  int B, T;
  const float* dout;                   // [T][B][40][64]
  float* dcarry;                       // [B][40][64]: in dL/dh_T, out dL/dh_0
  const float *r, *z, *n, *h, *h0;     // forward tape
  float* dg1;                          // [T][B][40][128] = [dr', dz']
  float* dg2;                          // [T][B][40][64]
};

constexpr int kCgBStage = 2 * 64 * kCgNf * 2;          // [hi | lo] x 64 in channels x 64 (tap, co) columns
constexpr int kCgBStages = 8;
// one extra half stage behind the ring: the M = 128 read of the last stage's lo plane runs 8 KB past it
constexpr int kCgBSmem = kCgBStages * kCgBStage + kCgBStage + 6 * kCgActPlane + 1024 + 512;

__device__ __forceinline__ void cg_store_grad(uint8_t* plane_hi, int p, int c, float v) {
  const __nv_bfloat16 hh = __float2bfloat16_rn(v);
  const __nv_bfloat16 ll = __float2bfloat16_rn(v - __bfloat162float(hh));
  const uint32_t off = (uint32_t)p * 128u + ((((uint32_t)c >> 3) ^ ((uint32_t)p & 7u)) << 4) +
                       ((uint32_t)c & 7u) * 2u;
  *reinterpret_cast<__nv_bfloat16*>(plane_hi + off) = hh;
  *reinterpret_cast<__nv_bfloat16*>(plane_hi + kCgActPlane + off) = ll;
}

__global__ void __launch_bounds__(kThreads, 1)
cgru_seq_bwd_kernel(const __grid_constant__ CUtensorMap tmW2_hi, const __grid_constant__ CUtensorMap tmW2_lo,
                    const __grid_constant__ CUtensorMap tmW1_hi, const __grid_constant__ CUtensorMap tmW1_lo,
                    const CgruSeqBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  uint8_t* ring = smem;                                          // [stages][hi 8 KB | lo 8 KB] (+ pad)
  uint8_t* g2p = ring + (kCgBStages + 1) * kCgBStage;            // dg2 planes  [hi | lo][88 rows]
  uint8_t* g1p = g2p + 2 * kCgActPlane;                          // dg1 planes  [chunk][hi | lo][88 rows]
  uint64_t* full = reinterpret_cast<uint64_t*>(g1p + 4 * kCgActPlane);
  uint64_t* empty = full + kCgBStages;
  uint64_t* d_full = empty + kCgBStages;     // [2]
  uint64_t* act_ready = d_full + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(act_ready + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x;
  const int B = p.B, T = p.T;

  if (warp == 0 && lane == 0) {
    tmap_prefetch(&tmW1_hi);
    tmap_prefetch(&tmW1_lo);
    tmap_prefetch(&tmW2_hi);
    tmap_prefetch(&tmW2_lo);
    for (int i = 0; i < kCgBStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&d_full[i], 1);
      mbar_init(&act_ready[i], 128);
    }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 6 * kCgActPlane / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(g2p)[i] = make_uint4(0u, 0u, 0u, 0u);
  // the half stage behind the ring is only ever read into junk accumulator rows: keep it finite
  for (int i = threadIdx.x; i < kCgBStage / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(ring + kCgBStages * kCgBStage)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (warp == 1) tmem_alloc(tmem_slot, 128);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== filter producer: 9 + 18 tiles per step =====================
    if (elect_one()) {
      constexpr uint32_t tx = 2u * 64u * kCgNf * 2u;
      uint32_t g = 0;
      for (int t = 0; t < T; ++t) {
        for (int conv = 0; conv < 2; ++conv) {
          const CUtensorMap* mh = conv == 0 ? &tmW2_hi : &tmW1_hi;
          const CUtensorMap* ml = conv == 0 ? &tmW2_lo : &tmW1_lo;
          const int tiles = conv == 0 ? 9 : 18;          // (tap, 64-wide out-channel chunk)
          for (int i = 0; i < tiles; ++i, ++g) {
            const uint32_t s = g % kCgBStages;
            mbar_wait(&empty[s], ((g / kCgBStages) & 1) ^ 1);
            uint8_t* dst = ring + (size_t)s * kCgBStage;
            mbar_expect_tx(&full[s], tx);
            tma_load_2d(dst, mh, &full[s], i * kCgNf, 0);
            tma_load_2d(dst + kCgBStage / 2, ml, &full[s], i * kCgNf, 0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // D fp32, A = flipped filters (bf16, K-major, M = 128 with 64 real rows), B = gate gradients (bf16)
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kCgPos >> 3) << 17) |
                               ((uint32_t)(kTileM >> 4) << 24);
    const uint64_t desc0 = kmajor_desc(0u, 64);
    const uint32_t ring16 = smem_u32(ring) >> 4;
    const uint32_t g2_16 = smem_u32(g2p) >> 4, g1_16 = smem_u32(g1p) >> 4;
    constexpr uint32_t kLo16W = (uint32_t)(kCgBStage / 2) >> 4;
    constexpr uint32_t kLo16A = (uint32_t)kCgActPlane >> 4;
    if (elect_one()) {               // one lane runs the whole loop, waits included
      uint32_t s = 0, ph = 0;
      for (int t = 0; t < T; ++t) {
        for (int conv = 0; conv < 2; ++conv) {
          mbar_wait(&act_ready[conv], (uint32_t)t & 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + (uint32_t)(conv * 64);
          const int tiles = conv == 0 ? 9 : 18;
#pragma unroll 1
          for (int i = 0; i < tiles; ++i) {
            mbar_wait(&full[s], ph);
            tc_fence_after();
            const int tap = conv == 0 ? i : (i >> 1);
            const int chunk = conv == 0 ? 0 : (i & 1);
            const int rr = tap / 3, qq = tap - 3 * rr;
            const uint64_t w_hi = desc0 + (uint64_t)(ring16 + s * ((uint32_t)kCgBStage >> 4));
            const uint32_t planes16 = conv == 0 ? g2_16 : g1_16 + (uint32_t)chunk * 2u * kLo16A;
            const uint64_t a_hi = desc0 + (uint64_t)(planes16 + (uint32_t)(rr * kCgWp + qq) * 8u);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t dw_hi = w_hi + (uint64_t)(k * 2), dw_lo = dw_hi + kLo16W;
              const uint64_t da_hi = a_hi + (uint64_t)(k * 2), da_lo = da_hi + kLo16A;
              umma_bf16(tmem_d, dw_lo, da_hi, idesc, (i | k) != 0);
              umma_bf16(tmem_d, dw_hi, da_lo, idesc, 1);
              umma_bf16(tmem_d, dw_hi, da_hi, idesc, 1);
            }
            umma_commit(&empty[s]);
            if (i == tiles - 1) umma_commit(&d_full[conv]);
            if (++s == (uint32_t)kCgBStages) {
              s = 0;
              ph ^= 1u;
            }
          }
        }
      }
    }
  } else {
    // ===================== gate gradients (TMEM lane = input channel; lanes 0..63 are real) =====================
    const int quad = warp & 3;
    const int c = quad * 32 + lane;
    const bool live = c < kCgNf;
    const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
    float dc[kCgPix];
    if (live) {
#pragma unroll
      for (int pix = 0; pix < kCgPix; ++pix) dc[pix] = p.dcarry[((size_t)b * kCgPix + pix) * kCgNf + c];
    }
    for (int step = 0; step < T; ++step) {
      const int t = T - 1 - step;
      const size_t frame = (size_t)t * B + b;
      const float* hprev = t == 0 ? p.h0 + (size_t)b * kCgPix * kCgNf
                                  : p.h + ((size_t)(t - 1) * B + b) * kCgPix * kCgNf;
      float v[kCgPos];
      // ---- dg2, dz', direct carry
      if (live) {
#pragma unroll
        for (int j = 0; j < kCgPos; ++j) {
          if (j % kCgWp < kCgW) {
            const int pix = (j / kCgWp) * kCgW + j % kCgWp;
            const size_t o = (frame * kCgPix + pix) * kCgNf + c;
            const float d = __ldg(p.dout + o) + dc[pix];
            const float zv = __ldg(p.z + o), nv = __ldg(p.n + o);
            const float hp = __ldg(hprev + (size_t)pix * kCgNf + c);
            const float g2 = d * (1.f - zv) * (1.f - nv * nv);
            const float dz = d * (hp - nv) * zv * (1.f - zv);
            dc[pix] = d * zv;
            v[j] = dz;
            cg_store_grad(g2p, j + kCgWp + 1, c, g2);
            cg_store_grad(g1p + 2 * kCgActPlane, j + kCgWp + 1, c, dz);       // chunk 1 = dz'
            p.dg2[o] = g2;
          }
        }
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&act_ready[0]);
      if (live) {
#pragma unroll
        for (int j = 0; j < kCgPos; ++j) {
          if (j % kCgWp < kCgW) {
            const int pix = (j / kCgWp) * kCgW + j % kCgWp;
            p.dg1[(frame * kCgPix + pix) * (2 * kCgNf) + kCgNf + c] = v[j];
          }
        }
      }
      // ---- d(r*h) -> dr', carry += d(r*h) r
      mbar_wait(&d_full[0], (uint32_t)step & 1);
      tc_fence_after();
      if (live) {
        tmem_ld32(tlane, v);
        tmem_ld16(tlane + 32u, v + 32);
#pragma unroll
        for (int j = 0; j < kCgPos; ++j) {
          if (j % kCgWp < kCgW) {
            const int pix = (j / kCgWp) * kCgW + j % kCgWp;
            const size_t o = (frame * kCgPix + pix) * kCgNf + c;
            const float rv = __ldg(p.r + o);
            const float hp = __ldg(hprev + (size_t)pix * kCgNf + c);
            const float dr = v[j] * hp * rv * (1.f - rv);
            dc[pix] += v[j] * rv;
            v[j] = dr;
            cg_store_grad(g1p, j + kCgWp + 1, c, dr);                           // chunk 0 = dr'
          }
        }
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&act_ready[1]);
      if (live) {
#pragma unroll
        for (int j = 0; j < kCgPos; ++j) {
          if (j % kCgWp < kCgW) {
            const int pix = (j / kCgWp) * kCgW + j % kCgWp;
            p.dg1[(frame * kCgPix + pix) * (2 * kCgNf) + c] = v[j];
          }
        }
      }
      // ---- carry += dgrad(conv1 | h half)
      mbar_wait(&d_full[1], (uint32_t)step & 1);
      tc_fence_after();
      if (live) {
        tmem_ld32(tlane + 64u, v);
        tmem_ld16(tlane + 96u, v + 32);
#pragma unroll
        for (int j = 0; j < kCgPos; ++j) {
          if (j % kCgWp < kCgW) dc[(j / kCgWp) * kCgW + j % kCgWp] += v[j];
        }
      }
      tc_fence_before();
    }
    if (live) {
#pragma unroll
      for (int pix = 0; pix < kCgPix; ++pix) p.dcarry[((size_t)b * kCgPix + pix) * kCgNf + c] = dc[pix];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

// =============================================================== weight gradient ====
// dW[co][(r,q)][ci] = sum over pixels p=(n,h,w) of dy[n, h-r+pad, w-q+pad, co] * x[n,h,w,ci]
// as a GEMM whose K dimension is the pixel index.  Both operands come straight from the NHWC
// bf16 planes with the same 4-D TMA boxes as the forward pass, i.e. they sit in shared memory
// as [pixel rows][channels] -- "MN-major" for the tensor core (instruction-descriptor major
// bits set):
//   A (M side): dy boxes shifted by the filter tap (TMA zero-fill = the padding); the 128 rows
//               of an M block are 128/uw "units", each one (tap, uw-wide co block) pair,
//               uw = min(64, Cout)
//   B (N side): the unshifted x box, nblk = min(Cin, 256) input channels in xw-wide blocks
// Each CTA owns one (M block, N block, pixel split) and accumulates its split in TMEM; a
// deterministic second kernel sums the splits into the OIHW gradient.
struct TcWgradParams {
  int N, H, W, Cin, Cout, KH, KW, pad;
  int bw, bh, bn, rows;     // pixel box of one K tile (rows = bw*bh*bn <= 64, multiple of 16)
  int tiles_w, tiles_h, tiles_total;
  int tiles_per_split;
  int uw, upb;              // unit width (channels) and units per M block (128 / uw)
  int units;                // taps * (Cout / uw)
  int cout_blocks;          // Cout / uw
  int xw;                   // channels per N-side block: min(64, C_n)
  int nblk;                 // N-side channels per CTA (multiple of xw, <= 256)
  int stages;
  // swap == 0: M side = dy (shifted by pad - tap), N side = x            (stride 1)
  // swap == 1: M side = x (shifted by tap - pad, traversal stride 2), N side = dy (stride 2);
  //            then uw / units / cout_blocks describe the *input* channels and the epilogue
  //            writes the transposed tile
  int swap, mstride;
  int fmt_m, fmt_n;         // operand formats of the M-side / N-side planes (0 = fp16, 1 = bf16)
  float* part;              // [splits][Cout][taps*Cin]
};

constexpr int kWgRows = 64;   // pixel rows reserved per staged block
constexpr int kWgradChainPixels = 4096;   // target length of one fp32 accumulation chain

template <int NPASS>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmD_hi,
                     const __grid_constant__ CUtensorMap tmD_lo,
                     const __grid_constant__ CUtensorMap tmX_hi,
                     const __grid_constant__ CUtensorMap tmX_lo, const TcWgradParams p) {
  constexpr int kPlanes = NPASS == 3 ? 2 : 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  const int nbB = p.nblk / p.xw;
  const int a_block = kWgRows * p.uw * 2;          // bytes reserved per dy unit
  const int b_block = kWgRows * p.xw * 2;          // bytes reserved per x block
  const int a_bytes = kTileM * kWgRows * 2;        // upb * a_block == 128 * 64 * 2 always
  const int bplane_bytes = nbB * b_block;          // one plane of the N-side blocks
  // stage = [M hi][M lo][N hi blocks][N lo blocks]: the lo N blocks follow the hi ones at the
  // block pitch, so ONE MN-major descriptor with N = 2 nblk reads [hi ; lo] (stacked-B issue)
  const int stage_bytes = (((a_bytes + bplane_bytes) * kPlanes) + 1023) & ~1023;
  const bool stack = NPASS == 3 && p.nblk <= 128;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* empty = full + p.stages;
  uint64_t* tmem_full = empty + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int mb = blockIdx.x;          // M block: units upb*mb .. upb*mb + upb - 1
  const int nb = blockIdx.y;          // N block
  const int sp = blockIdx.z;          // pixel split
  const int t_begin = sp * p.tiles_per_split;
  const int t_end = min(p.tiles_total, t_begin + p.tiles_per_split);
  const int iters = max(t_end - t_begin, 0);
  uint32_t tmem_cols = 32;                 // allocation granularity: powers of two >= 32
  while (tmem_cols < (uint32_t)(stack ? 2 * p.nblk : p.nblk)) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) {
    tmap_prefetch(&tmD_hi);
    tmap_prefetch(&tmX_hi);
    if (NPASS == 3) {
      tmap_prefetch(&tmD_lo);
      tmap_prefetch(&tmX_lo);
    }
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      const uint32_t tx = (uint32_t)p.rows * (uint32_t)(p.upb * p.uw + nbB * p.xw) * 2u * kPlanes;
      for (int it = 0; it < iters; ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (it / p.stages) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* st = smem + s * stage_bytes;
        int tile = t_begin + it;
        const int tw = tile % p.tiles_w;
        tile /= p.tiles_w;
        const int th = tile % p.tiles_h;
        const int tn = tile / p.tiles_h;
        const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;
        mbar_expect_tx(&full[s], tx);
#pragma unroll
        for (int pl = 0; pl < kPlanes; ++pl) {
          uint8_t* abase = st + pl * a_bytes;
          uint8_t* bbase = st + kPlanes * a_bytes + pl * bplane_bytes;
          const CUtensorMap* md = pl == 0 ? &tmD_hi : &tmD_lo;
          const CUtensorMap* mx = pl == 0 ? &tmX_hi : &tmX_lo;
          for (int b = 0; b < p.upb; ++b) {
            const int u = min(p.upb * mb + b, p.units - 1);
            const int tap = u / p.cout_blocks;
            const int cb = u - tap * p.cout_blocks;
            const int r = tap / p.KW, q = tap - r * p.KW;
            const int dw = p.swap ? q - p.pad : p.pad - q;
            const int dh = p.swap ? r - p.pad : p.pad - r;
            tma_load_4d(abase + b * a_block, md, &full[s], cb * p.uw, w0 * p.mstride + dw,
                        h0 * p.mstride + dh, n0);
          }
          for (int b = 0; b < nbB; ++b)
            tma_load_4d(bbase + b * b_block, mx, &full[s], nb * p.nblk + b * p.xw, w0, h0, n0);
        }
      }
    }
  } else if (warp == 1) {
    // D fp32, A/B fp16 or bf16 (independently), both MN-major, M = 128, N = nblk
    const uint32_t idesc = (1u << 4) | ((uint32_t)p.fmt_m << 7) | ((uint32_t)p.fmt_n << 10) |
                           (1u << 15) | (1u << 16) |
                           ((uint32_t)(p.nblk >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
    const uint32_t idesc2 = (idesc & ~(0x3Fu << 17)) | ((uint32_t)((2 * p.nblk) >> 3) << 17);
    const int ksteps = p.rows >> 4;
    // one elected lane, ring position kept incrementally, descriptor constants hoisted (the issuer's
    // per-stage and per-K-step overhead does not overlap the MMAs: see conv_tc_kernel)
    if (elect_one()) {
      const uint64_t dac = mnmajor_desc(0u, a_block, p.uw);
      const uint64_t dbc = mnmajor_desc(0u, b_block, p.xw);
      const uint32_t smem16 = smem_u32(smem) >> 4, stage16 = (uint32_t)stage_bytes >> 4;
      const uint32_t a16 = (uint32_t)a_bytes >> 4, bp16 = (uint32_t)bplane_bytes >> 4;
      const uint32_t ka16 = 2u * (uint32_t)p.uw, kb16 = 2u * (uint32_t)p.xw;   // 16 pixel rows, in 16-byte units
      uint32_t s = 0, ph = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_hi16 = smem16 + s * stage16;
        const uint32_t b_hi16 = a_hi16 + kPlanes * a16;
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t dah = dac + (uint64_t)(a_hi16 + (uint32_t)k * ka16);
          const uint64_t dbh = dbc + (uint64_t)(b_hi16 + (uint32_t)k * kb16);
          if (stack) {
            // [M_hi N_hi | M_hi N_lo] from one N = 2 nblk instruction, then M_lo N_hi
            umma_bf16(tmem_base, dah, dbh, idesc2, (it | k) != 0);
            umma_bf16(tmem_base, dah + a16, dbh, idesc, 1);
          } else if (NPASS == 3) {
            umma_bf16(tmem_base, dah + a16, dbh, idesc, (it | k) != 0);
            umma_bf16(tmem_base, dah, dbh + bp16, idesc, 1);
            umma_bf16(tmem_base, dah, dbh, idesc, 1);
          } else {
            umma_bf16(tmem_base, dah, dbh, idesc, (it | k) != 0);
          }
        }
        umma_commit(&empty[s]);
        if (it == iters - 1) umma_commit(tmem_full);
        if (++s == (uint32_t)p.stages) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
  } else {
    const int quad = warp & 3;
    const int m = quad * 32 + lane;
    const int u = p.upb * mb + m / p.uw;
    const bool valid = u < p.units;
    const int tap = valid ? u / p.cout_blocks : 0;
    const int cb = valid ? u - tap * p.cout_blocks : 0;
    const int mc = cb * p.uw + (m % p.uw);      // channel on the M side (co, or ci when swapped)
    const size_t KK = (size_t)p.KH * p.KW * p.Cin;
    float* dst = p.swap
        ? p.part + ((size_t)sp * p.Cout + (size_t)nb * p.nblk) * KK + (size_t)tap * p.Cin + mc
        : p.part + ((size_t)sp * p.Cout + mc) * KK + (size_t)tap * p.Cin + (size_t)nb * p.nblk;
    if (iters > 0) {
      mbar_wait(tmem_full, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int c0 = 0; c0 < p.nblk; c0 += 32) {
      float v[32];
      if (iters > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, v);
        if (stack) {
          if (p.nblk < 32) {
            for (int j = 0; j < p.nblk; ++j) v[j] += v[p.nblk + j];
          } else {
            float u[32];
            tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(p.nblk + c0), u);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += u[j];
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      if (valid && !p.swap) {
        float4* o4 = reinterpret_cast<float4*>(dst + c0);
        const int nq = min(32, p.nblk - c0) >> 2;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j < nq) o4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      } else if (valid) {
        const int nv = min(32, p.nblk - c0);      // column n = output channel: rows KK apart
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nv) dst[(size_t)(c0 + j) * KK] = v[j];
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ===================================================== halo-row weight gradient (W == 128) ==
// Same staging idea as conv_tc_row_kernel for dW of 3x3 stride-1 convolutions over 128-pixel-wide
// maps: ring entry rho holds the dy row rho (130 pixels, halo zero-filled by TMA) and the x row
// rho.  For the x row h and filter row r the GEMM  D_r[(q, co)][ci] += dy[h-r+1, w-q+1, co] *
// x[h, w, ci]  runs over K = the 128 pixels of the row in 16-pixel steps; the three horizontal
// taps are three M blocks of ONE MN-major descriptor whose leading-dimension offset is one pixel
// (M block b reads the dy row shifted by b pixels, i.e. q = 2 - b; the remaining M blocks read
// further-shifted data into TMEM lanes nobody looks at).  Each CTA keeps three accumulators
// (one per filter row) in TMEM over all the rows it owns and writes one partial gradient; the
// deterministic second stage (wgrad_reduce) sums the per-CTA partials.
struct TcWgRowParams {
  int N, H;
  int strips, rows_per_strip, items;
  int slots;
  float* part;              // [gridDim.x][Cout][KS * KS * Cin]
};

template <int NPASS, int CIN, int COUT, int KS>
struct WgRowCfg {
  static constexpr int kPlanes = NPASS == 3 ? 2 : 1;
  static constexpr int kDRowBytes = (kRowBox * COUT * 2 + 1023) / 1024 * 1024;   // one dy plane
  static constexpr int kXRowBytes = kTileM * CIN * 2;                            // one x plane
  static constexpr int kSlotBytes = kPlanes * (kDRowBytes + kXRowBytes);
  static constexpr int kSlotsRaw = (227 * 1024 - 1024 - kBarrierBytes) / kSlotBytes;
  static constexpr int kSlots = kSlotsRaw > kMaxStages ? kMaxStages : kSlotsRaw;
  // two accumulator sets (three filter rows x CIN columns each), flushed alternately
  static constexpr int kHalo = KS / 2;
  // stacked-B issue (see TcCfg): the x lo plane sits exactly one leading-dimension offset behind
  // the hi plane, so N = 2 CIN reads [x_hi ; x_lo]; 2 sets x 3 filter rows x 2 CIN columns must fit
  static constexpr bool kStack = NPASS == 3 && CIN <= 32;
  static constexpr uint32_t kAcc = kStack ? 2 * CIN : CIN;       // accumulator columns per filter row
  static constexpr uint32_t kSetCols = KS * kAcc;
  static constexpr uint32_t kTmemCols = 2 * kSetCols <= 128 ? 128u : (2 * kSetCols <= 256 ? 256u : 512u);
};
constexpr int kWgFlushRows = kWgradChainPixels / kTileM;   // image rows per accumulation chain

template <int NPASS, int CIN, int COUT, int KS>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_wgrad_row_kernel(const __grid_constant__ CUtensorMap tmD_hi,
                         const __grid_constant__ CUtensorMap tmD_lo,
                         const __grid_constant__ CUtensorMap tmX_hi,
                         const __grid_constant__ CUtensorMap tmX_lo, const TcWgRowParams p) {
  using Cfg = WgRowCfg<NPASS, CIN, COUT, KS>;
  constexpr int kPlanes = Cfg::kPlanes;
  constexpr int kHalo = Cfg::kHalo;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  const int slots = p.slots;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)slots * Cfg::kSlotBytes);
  uint64_t* empty = full + slots;
  uint64_t* tmem_full = empty + slots;     // [2]
  uint64_t* tmem_empty = tmem_full + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tmap_prefetch(&tmD_hi);
    tmap_prefetch(&tmX_hi);
    if (NPASS == 3) {
      tmap_prefetch(&tmD_lo);
      tmap_prefetch(&tmX_lo);
    }
    for (int s = 0; s < slots; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // rows this CTA owns (all roles walk the same items in the same order)
  int total_rows = 0;
  for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
    const int n = item / p.strips;
    const int h0 = (item - n * p.strips) * p.rows_per_strip;
    total_rows += min(p.H, h0 + p.rows_per_strip) - h0;
  }

  if (warp == 0) {
    if (elect_one()) {
      constexpr uint32_t tx_d = (uint32_t)(kRowBox * COUT * 2) * kPlanes;
      constexpr uint32_t tx_x = (uint32_t)(kTileM * CIN * 2) * kPlanes;
      uint32_t g = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const int n = item / p.strips;
        const int h0 = (item - n * p.strips) * p.rows_per_strip;
        const int h1 = min(p.H, h0 + p.rows_per_strip);
        for (int hr = h0 - kHalo; hr < h1 + kHalo; ++hr, ++g) {
          const int s = g % slots;
          mbar_wait(&empty[s], ((g / slots) & 1) ^ 1);
          uint8_t* dst = ring + (size_t)s * Cfg::kSlotBytes;
          const bool with_x = hr >= h0 && hr < h1;
          mbar_expect_tx(&full[s], tx_d + (with_x ? tx_x : 0u));
          tma_load_4d(dst, &tmD_hi, &full[s], 0, -1, hr, n);
          if (NPASS == 3) tma_load_4d(dst + Cfg::kDRowBytes, &tmD_lo, &full[s], 0, -1, hr, n);
          if (with_x) {
            uint8_t* xd = dst + kPlanes * Cfg::kDRowBytes;
            tma_load_4d(xd, &tmX_hi, &full[s], 0, 0, hr, n);
            if (NPASS == 3) tma_load_4d(xd + Cfg::kXRowBytes, &tmX_lo, &full[s], 0, 0, hr, n);
          }
        }
      }
    }
  } else if (warp == 1) {
    // D fp32, A/B bf16, both MN-major, M = 128, N = CIN
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                               ((uint32_t)(CIN >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
    constexpr uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                ((uint32_t)((2 * CIN) >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
    // per 16-pixel K step the operands advance by 16 pixel rows; everything in 16-byte units
    constexpr uint32_t kDStep16 = (uint32_t)(16 * COUT * 2) >> 4;
    constexpr uint32_t kXStep16 = (uint32_t)(16 * CIN * 2) >> 4;
    constexpr uint32_t kDRow16 = (uint32_t)Cfg::kDRowBytes >> 4;
    constexpr uint32_t kXRow16 = (uint32_t)Cfg::kXRowBytes >> 4;
    constexpr uint32_t kSlot16 = (uint32_t)Cfg::kSlotBytes >> 4;
    // M side: leading-dimension offset = one pixel -> M block b is the row shifted by b pixels
    const uint64_t d_desc0 = mnmajor_desc(0u, (uint32_t)(COUT * 2), COUT);
    const uint64_t x_desc0 = mnmajor_desc(0u, (uint32_t)Cfg::kXRowBytes, CIN);
    const uint32_t ring16 = smem_u32(ring) >> 4;
    // ring position of entry g kept incrementally (no runtime division in the issue loop); a staged
    // row is waited for once, not by each of the three x rows that read it
    // one elected lane runs the whole loop, waits included (see conv_tc_row_kernel)
    if (elect_one()) {
      uint32_t sg = 0, pg = 0;
      int confirmed = 0;
      auto advance = [&](uint32_t& sl_, uint32_t& ph_) {
        if (++sl_ == (uint32_t)slots) {
          sl_ = 0;
          ph_ ^= 1u;
        }
      };
      int row = 0;                 // rows accumulated so far by this CTA
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const int n = item / p.strips;
        const int h0 = (item - n * p.strips) * p.rows_per_strip;
        const int h1 = min(p.H, h0 + p.rows_per_strip);
        for (int h = h0; h < h1; ++h, ++row) {
          // accumulation chains of kWgFlushRows rows alternate between the two accumulator sets
          const uint32_t chain = (uint32_t)(row / kWgFlushRows);
          const uint32_t set = chain & 1;
          const bool chain_start = row % kWgFlushRows == 0;
          const bool chain_end = row % kWgFlushRows == kWgFlushRows - 1 || row == total_rows - 1;
          if (chain_start) {
            mbar_wait(&tmem_empty[set], ((chain >> 1) & 1) ^ 1);   // the epilogue drained this set
            tc_fence_after();
          }
          uint32_t sl[KS];
          {
            uint32_t se = sg, pe = pg;
#pragma unroll
            for (int j = 0; j < KS; ++j) {
              sl[j] = se;
              if (j >= confirmed) mbar_wait(&full[se], pe);
              advance(se, pe);
            }
          }
          tc_fence_after();
          const uint32_t acc0 = chain_start ? 0u : 1u;
          const uint64_t xb = x_desc0 + (uint64_t)(ring16 + sl[kHalo] * kSlot16 + kPlanes * kDRow16);
#pragma unroll
          for (int r = 0; r < KS; ++r) {
            // dy row h - r + halo is ring entry g + 2 * halo - r (entry g holds row h - halo); the
            // staged box starts at pixel -1, so a 1x1 filter reads it one pixel in
            const uint64_t db = d_desc0 + (uint64_t)(ring16 + sl[2 * kHalo - r] * kSlot16 +
                                                     (1 - kHalo) * ((uint32_t)(COUT * 2) >> 4));
            const uint32_t tmem_d = tmem_base + set * Cfg::kSetCols + (uint32_t)r * Cfg::kAcc;
#pragma unroll
            for (int ks = 0; ks < kTileM / 16; ++ks) {
              const uint64_t dah = db + (uint64_t)(ks * kDStep16);
              const uint64_t dbh = xb + (uint64_t)(ks * kXStep16);
              if (Cfg::kStack) {
                const uint64_t dal = dah + kDRow16;
                umma_bf16(tmem_d, dah, dbh, idesc2, ks == 0 ? acc0 : 1u);
                umma_bf16(tmem_d, dal, dbh, idesc, 1);
              } else if (NPASS == 3) {
                const uint64_t dal = dah + kDRow16;
                const uint64_t dbl = dbh + kXRow16;
                umma_bf16(tmem_d, dal, dbh, idesc, ks == 0 ? acc0 : 1u);
                umma_bf16(tmem_d, dah, dbl, idesc, 1);
                umma_bf16(tmem_d, dah, dbh, idesc, 1);
              } else {
                umma_bf16(tmem_d, dah, dbh, idesc, ks == 0 ? acc0 : 1u);
              }
            }
          }
          umma_commit(&empty[sl[0]]);
          if (chain_end) umma_commit(&tmem_full[set]);
          advance(sg, pg);
          confirmed = KS - 1;
        }
        if (KS == 3) {
          umma_commit(&empty[sg]);
          advance(sg, pg);
          umma_commit(&empty[sg]);
          advance(sg, pg);
        }
        confirmed = 0;
      }
    }
  } else {
    // epilogue: drain each finished chain into this CTA's partial gradient (same thread, same
    // address, fixed order: deterministic fp32 additions with round-to-nearest)
    const int quad = warp & 3;
    const int m = quad * 32 + lane;
    const int b = m / COUT;                   // M block = horizontal shift
    const int co = m - b * COUT;
    const bool valid = b < KS;
    const int q = KS - 1 - b;
    constexpr size_t KK = (size_t)KS * KS * CIN;
    const int chains = (total_rows + kWgFlushRows - 1) / kWgFlushRows;
    for (int chain = 0; chain < chains; ++chain) {
      const uint32_t set = (uint32_t)chain & 1;
      mbar_wait(&tmem_full[set], ((uint32_t)chain >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int r = 0; r < KS; ++r) {
        float* dst = p.part + ((size_t)blockIdx.x * COUT + co) * KK + (size_t)(r * KS + q) * CIN;
#pragma unroll 1
        for (int c0 = 0; c0 < CIN; c0 += 32) {
          float v[32];
          tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + set * Cfg::kSetCols +
                        (uint32_t)r * Cfg::kAcc + (uint32_t)c0, v);
          if (Cfg::kStack) {
            if (CIN < 32) {
#pragma unroll
              for (int j = 0; j < CIN; ++j) v[j] += v[CIN + j];
            } else {
              float u[32];
              tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + set * Cfg::kSetCols +
                            (uint32_t)r * Cfg::kAcc + (uint32_t)(CIN + c0), u);
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] += u[j];
            }
          }
          if (valid) {
            float4* o4 = reinterpret_cast<float4*>(dst + c0);
            constexpr int nq = (CIN < 32 ? CIN : 32) >> 2;
#pragma unroll
            for (int j = 0; j < nq; ++j) {
              float4 o = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
              if (chain > 0) {
                const float4 a = o4[j];
                o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
              }
              o4[j] = o;
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive_relaxed(&tmem_empty[set]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---- 3x3 variant with the three FILTER ROWS stacked along N.  Above, the x row h meets the dy
// rows h-1, h, h+1 in three separate instruction groups of N = Cin (2 Cin stacked) columns: at 16 / 32
// input channels every tcgen05.mma costs the ~43 cycles of its 4 KB dy-tile read whatever N is
// (DESIGN 3b), so the kernel is paced by its instruction count.  Turned around -- the dy row g meets
// the x rows g-1, g, g+1 -- the three products share the dy operand and become ONE instruction with
// N = 3 Cin: B = three consecutive x slots (the x planes live in their own rings at a pitch of one row,
// which is the descriptor's leading-dimension offset), D = [D_0 | D_1 | D_2].  The two halo entries of
// an item carry a zero x row (TMA box outside the image), so every interior instruction is uniform;
// the dy halo rows meet their single x row in two N = Cin instructions per item.  Where the three
// slots wrap around the ring the instruction is issued in two parts.  3 instead of 6 instructions per
// 16-pixel K step with split operands.
template <int NPASS, int CIN, int COUT>
struct WgRow3Cfg {
  static constexpr int kPlanes = NPASS == 3 ? 2 : 1;
  static constexpr int kDRowBytes = (kRowBox * COUT * 2 + 1023) / 1024 * 1024;   // one dy plane
  static constexpr int kXRowBytes = kTileM * CIN * 2;                            // one x plane
  static constexpr int kSlotBytes = kPlanes * (kDRowBytes + kXRowBytes);
  static constexpr int kSlotsRaw = (227 * 1024 - 1024 - kBarrierBytes) / kSlotBytes;
  static constexpr int kSlots = kSlotsRaw > kMaxStages ? kMaxStages : kSlotsRaw;
  // hi x hi and lo x hi share columns [0, 3 CIN); with stacked accumulators hi x lo goes to
  // [3 CIN, 6 CIN) and the epilogue adds the halves
  static constexpr bool kStack = NPASS == 3 && CIN <= 32;
  static constexpr uint32_t kSetCols = (kStack ? 6u : 3u) * CIN;
  // 16 output channels: the three pixel-shifted M blocks fit an M = 64 instruction (four blocks, half
  // the dy-tile read of M = 128, which is what an instruction costs at N <= 96).  Its accumulator row m
  // sits in TMEM lane 32 (m / 16) + m % 16: M block b is lanes 0..15 of epilogue warp b.
  static constexpr int kM = COUT == 16 ? 64 : kTileM;
  static constexpr uint32_t kTmemCols = 2 * kSetCols <= 128 ? 128u : (2 * kSetCols <= 256 ? 256u : 512u);
};

template <int NPASS, int CIN, int COUT>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_wgrad_row3_kernel(const __grid_constant__ CUtensorMap tmD_hi,
                          const __grid_constant__ CUtensorMap tmD_lo,
                          const __grid_constant__ CUtensorMap tmX_hi,
                          const __grid_constant__ CUtensorMap tmX_lo, const TcWgRowParams p) {
  using Cfg = WgRow3Cfg<NPASS, CIN, COUT>;
  constexpr int kPlanes = Cfg::kPlanes;
  constexpr int KS = 3;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  const int slots = p.slots;
  // [dy: slots x planes x kDRowBytes][x hi: slots x kXRowBytes][x lo: slots x kXRowBytes]
  uint8_t* xring = ring + (size_t)slots * kPlanes * Cfg::kDRowBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)slots * Cfg::kSlotBytes);
  uint64_t* empty = full + slots;
  uint64_t* tmem_full = empty + slots;     // [2]
  uint64_t* tmem_empty = tmem_full + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tmap_prefetch(&tmD_hi);
    tmap_prefetch(&tmX_hi);
    if (NPASS == 3) {
      tmap_prefetch(&tmD_lo);
      tmap_prefetch(&tmX_lo);
    }
    for (int s = 0; s < slots; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  int total_rows = 0;
  for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
    const int n = item / p.strips;
    const int h0 = (item - n * p.strips) * p.rows_per_strip;
    total_rows += min(p.H, h0 + p.rows_per_strip) - h0;
  }

  if (warp == 0) {
    if (elect_one()) {
      constexpr uint32_t tx = (uint32_t)(kRowBox * COUT * 2 + kTileM * CIN * 2) * kPlanes;
      uint32_t g = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const int n = item / p.strips;
        const int h0 = (item - n * p.strips) * p.rows_per_strip;
        const int h1 = min(p.H, h0 + p.rows_per_strip);
        for (int hr = h0 - 1; hr < h1 + 1; ++hr, ++g) {
          const int s = g % slots;
          mbar_wait(&empty[s], ((g / slots) & 1) ^ 1);
          uint8_t* dd = ring + (size_t)s * kPlanes * Cfg::kDRowBytes;
          uint8_t* xd = xring + (size_t)s * Cfg::kXRowBytes;
          // the x row of a halo entry belongs to the neighbouring strip: a box outside the image
          // (row H) is all TMA zero fill
          const int xr = (hr >= h0 && hr < h1) ? hr : p.H;
          mbar_expect_tx(&full[s], tx);
          tma_load_4d(dd, &tmD_hi, &full[s], 0, -1, hr, n);
          tma_load_4d(xd, &tmX_hi, &full[s], 0, 0, xr, n);
          if (NPASS == 3) {
            tma_load_4d(dd + Cfg::kDRowBytes, &tmD_lo, &full[s], 0, -1, hr, n);
            tma_load_4d(xd + (size_t)slots * Cfg::kXRowBytes, &tmX_lo, &full[s], 0, 0, xr, n);
          }
        }
      }
    }
  } else if (warp == 1) {
    // D fp32, A/B bf16, both MN-major, M = 128; N = nb * CIN
    constexpr uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                ((uint32_t)(Cfg::kM >> 4) << 24);
    constexpr uint32_t idescN[4] = {0u, idesc0 | ((uint32_t)(CIN >> 3) << 17),
                                    idesc0 | ((uint32_t)((2 * CIN) >> 3) << 17),
                                    idesc0 | ((uint32_t)((3 * CIN) >> 3) << 17)};
    constexpr uint32_t kDStep16 = (uint32_t)(16 * COUT * 2) >> 4;
    constexpr uint32_t kXStep16 = (uint32_t)(16 * CIN * 2) >> 4;
    constexpr uint32_t kDRow16 = (uint32_t)Cfg::kDRowBytes >> 4;
    constexpr uint32_t kXRow16 = (uint32_t)Cfg::kXRowBytes >> 4;
    const uint64_t d_desc0 = mnmajor_desc(0u, (uint32_t)(COUT * 2), COUT);
    const uint64_t x_desc0 = mnmajor_desc(0u, (uint32_t)Cfg::kXRowBytes, CIN);
    const uint32_t ring16 = smem_u32(ring) >> 4;
    const uint32_t xring16 = smem_u32(xring) >> 4;
    const uint32_t xlo16 = (uint32_t)slots * kXRow16;          // x lo ring behind the x hi ring
    if (elect_one()) {
      uint32_t sg = 0, pg = 0;
      int confirmed = 0;
      auto advance = [&](uint32_t& sl_, uint32_t& ph_) {
        if (++sl_ == (uint32_t)slots) {
          sl_ = 0;
          ph_ ^= 1u;
        }
      };
      // the dy row in slot sd against `nb` consecutive x rows from slot sx, into accumulator
      // column block cb (of CIN columns), over the 128 pixels of the row
      auto issue = [&](uint32_t tmem_set, uint32_t sd, uint32_t sx, int nb, uint32_t cb, uint32_t acc0) {
        const uint64_t db = d_desc0 + (uint64_t)(ring16 + sd * (kPlanes * kDRow16));
        const uint64_t xb = x_desc0 + (uint64_t)(xring16 + sx * kXRow16);
        const uint32_t td = tmem_set + cb * CIN;
        const uint32_t id = idescN[nb];
#pragma unroll
        for (int ks = 0; ks < kTileM / 16; ++ks) {
          const uint64_t dah = db + (uint64_t)(ks * kDStep16);
          const uint64_t dbh = xb + (uint64_t)(ks * kXStep16);
          const uint32_t acc = ks == 0 ? acc0 : 1u;
          if (Cfg::kStack) {
            umma_bf16(td, dah, dbh, id, acc);
            umma_bf16(td + 3 * CIN, dah, dbh + xlo16, id, acc);
            umma_bf16(td, dah + kDRow16, dbh, id, 1);
          } else if (NPASS == 3) {
            umma_bf16(td, dah + kDRow16, dbh, id, acc);
            umma_bf16(td, dah, dbh + xlo16, id, 1);
            umma_bf16(td, dah, dbh, id, 1);
          } else {
            umma_bf16(td, dah, dbh, id, acc);
          }
        }
      };
      int row = 0;                 // rows accumulated so far by this CTA
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const int n = item / p.strips;
        const int h0 = (item - n * p.strips) * p.rows_per_strip;
        const int h1 = min(p.H, h0 + p.rows_per_strip);
        for (int h = h0; h < h1; ++h, ++row) {
          const uint32_t chain = (uint32_t)(row / kWgFlushRows);
          const uint32_t set = chain & 1;
          const bool chain_start = row % kWgFlushRows == 0;
          const bool chain_end = row % kWgFlushRows == kWgFlushRows - 1 || row == total_rows - 1;
          if (chain_start) {
            mbar_wait(&tmem_empty[set], ((chain >> 1) & 1) ^ 1);   // the epilogue drained this set
            tc_fence_after();
          }
          // ring entries of rows h-1, h, h+1
          uint32_t sl[KS];
          {
            uint32_t se = sg, pe = pg;
#pragma unroll
            for (int j = 0; j < KS; ++j) {
              sl[j] = se;
              if (j >= confirmed) mbar_wait(&full[se], pe);
              advance(se, pe);
            }
          }
          tc_fence_after();
          const uint32_t acc0 = chain_start ? 0u : 1u;
          const uint32_t tmem_set = tmem_base + set * Cfg::kSetCols;
          // dy row h x (x rows h-1, h, h+1) -> (D_0, D_1, D_2): D_r += dy[h'-r+1] x[h'] with h' = h+r-1
          if (sl[2] == sl[0] + 2) {
            issue(tmem_set, sl[1], sl[0], 3, 0, acc0);
          } else if (sl[1] == sl[0] + 1) {
            issue(tmem_set, sl[1], sl[0], 2, 0, acc0);
            issue(tmem_set, sl[1], sl[2], 1, 2, acc0);
          } else {
            issue(tmem_set, sl[1], sl[0], 1, 0, acc0);
            issue(tmem_set, sl[1], sl[1], 2, 1, acc0);
          }
          // the strip's dy halo rows: row h0-1 meets x row h0 in D_2, row h1 meets x row h1-1 in D_0
          if (h == h0) issue(tmem_set, sl[0], sl[1], 1, 2, 1u);
          if (h == h1 - 1) issue(tmem_set, sl[2], sl[1], 1, 0, 1u);
          umma_commit(&empty[sl[0]]);
          if (chain_end) umma_commit(&tmem_full[set]);
          advance(sg, pg);
          confirmed = KS - 1;
        }
        umma_commit(&empty[sg]);
        advance(sg, pg);
        umma_commit(&empty[sg]);
        advance(sg, pg);
        confirmed = 0;
      }
    }
  } else {
    // epilogue: as in conv_tc_wgrad_row_kernel, columns [r * CIN, (r+1) * CIN) (+ 3 CIN stacked)
    const int quad = warp & 3;
    const int m = quad * 32 + lane;
    // M = 128: lane = accumulator row; M = 64 (16 output channels): block b = warp b, lanes 0..15
    const int b = Cfg::kM == 64 ? quad : m / COUT;      // M block = horizontal shift
    const int co = Cfg::kM == 64 ? (lane & 15) : m - b * COUT;
    const bool valid = b < KS && (Cfg::kM != 64 || lane < 16);
    const int q = KS - 1 - b;
    constexpr size_t KK = (size_t)KS * KS * CIN;
    const int chains = (total_rows + kWgFlushRows - 1) / kWgFlushRows;
    for (int chain = 0; chain < chains; ++chain) {
      const uint32_t set = (uint32_t)chain & 1;
      mbar_wait(&tmem_full[set], ((uint32_t)chain >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int r = 0; r < KS; ++r) {
        float* dst = p.part + ((size_t)blockIdx.x * COUT + co) * KK + (size_t)(r * KS + q) * CIN;
#pragma unroll 1
        for (int c0 = 0; c0 < CIN; c0 += 32) {
          constexpr int nc = CIN < 32 ? CIN : 32;
          float v[32];
          const uint32_t t0 = tmem_base + ((uint32_t)(quad * 32) << 16) + set * Cfg::kSetCols +
                              (uint32_t)(r * CIN + c0);
          tmem_ld32(t0, v);                  // CIN = 16: the upper half is the next block, unused
          if (Cfg::kStack) {
            float u[32];
            tmem_ld32(t0 + 3 * CIN, u);
#pragma unroll
            for (int j = 0; j < nc; ++j) v[j] += u[j];
          }
          if (valid) {
            float4* o4 = reinterpret_cast<float4*>(dst + c0);
#pragma unroll
            for (int j = 0; j < nc / 4; ++j) {
              float4 o = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
              if (chain > 0) {
                const float4 a = o4[j];
                o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
              }
              o4[j] = o;
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive_relaxed(&tmem_empty[set]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ====================================================== padded-strip weight gradient ==
// dW of 3x3 stride-1 convolutions on maps narrower than 128 pixels (EyeNet layers 1-2, RefineNet
// levels 1-2: the split-K kernel above re-loads the dy box once per filter tap and the x box once per
// M block -- 48..64 KB of operands per 8..12 MMAs, about twice what L2 -> shared memory delivers).
// As in conv_tc_wgrad_row_kernel the K dimension runs over pixels, but of a zero-padded STRIP: a work
// item stages R rows of x and R + 2 rows of dy ONCE, both as boxes that start at w = -1 and are
// W + 2 pixels wide, so that pixel p of the x strip meets pixel p + (2 - r)(W + 2) + (1 - q) of the dy
// strip for filter tap (r, q): nine descriptor start addresses into one staged strip (the single
// negative offset, tap (2, 2), is taken on the x side).  The pad pixels of x are TMA zero fill and
// contribute nothing.  One M = 128 MN-major instruction covers 128 / SUBW taps of SUBW output
// channels whose strip offsets differ by the descriptor's leading-dimension offset:
//   SUBW = 64: tap pairs, five M blocks cover the nine taps of a 64-channel output block;
//   SUBW = 32: the three horizontal taps of a filter row (+ one junk sub-block), three M blocks.
// A JOB = (output-channel block, group of M blocks) owns as many accumulators as TMEM holds
// (128-channel inputs: the taps of a block split into two jobs); the CTAs are dealt round-robin to
// the jobs and each walks its job's strips, accumulating in TMEM and flushing into ONE partial
// gradient (deterministic second stage: wgrad_reduce) at least every kWgradChainPixels pixels.
struct WgStripJob {
  int co0;                 // first output channel of the job's block
  int nblk;                // M blocks (accumulators)
  int boff[5], blbo[5];    // dy-strip offset of the first tap of a block, pixel distance between its taps
  int xoff[5];             // x-strip offset (1 for the block that holds tap (2, 2))
  int tap[5][4];           // r * 3 + q of every M sub-block, -1 = junk
};
constexpr int kWgMaxJobs = 4;
struct TcWgStripParams {
  int N, H, W, Wp, Cin, Cout;
  int R, strips, items;           // rows per strip, strips per image, N * strips (per job)
  int ksteps;                     // ceil(R * Wp / 16)
  int dy_plane, x_chunk;          // bytes of one dy plane / of one 64-channel x chunk plane (1024-aligned)
  int flush_items;                // items per accumulation chain
  int njobs;
  int period, slotjob[12];        // CTA c works on job slotjob[c % period]
  WgStripJob job[kWgMaxJobs];
  float* part;                    // [gridDim.x][Cout][9 * Cin]  (zero-filled by the host when njobs > 1)
};

template <int CIN, int SUBW>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_wgrad_strip_kernel(const __grid_constant__ CUtensorMap tmD_hi,
                           const __grid_constant__ CUtensorMap tmD_lo,
                           const __grid_constant__ CUtensorMap tmX_hi,
                           const __grid_constant__ CUtensorMap tmX_lo, const TcWgStripParams p) {
  constexpr int XW = CIN > 64 ? 64 : CIN;                  // channels of one x chunk (swizzle span)
  constexpr int XCH = CIN / XW;                            // x chunks: N = CIN spans XCH descriptor blocks
  constexpr bool kStack = CIN <= 32;                       // see TcCfg: N = 2 CIN over [x_hi ; x_lo]
  constexpr uint32_t kAcc = kStack ? 2 * CIN : CIN;        // accumulator columns of one M block
  constexpr int SUBS = kTileM / SUBW;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  const int x_plane = XCH * p.x_chunk;
  const int stage_bytes = 2 * (p.dy_plane + x_plane);      // [dy hi][dy lo][x hi chunks][x lo chunks]
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + 2 * (size_t)stage_bytes);
  uint64_t* empty = full + 2;
  uint64_t* tmem_full = empty + 2;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // this CTA's job (dealt by a short repeating pattern, jobs with more M blocks get more CTAs) and
  // its share of the job's items
  const int jb = p.slotjob[blockIdx.x % p.period];
  int cta_in_job = 0, ctas_in_job = 0;
  for (int c = 0; c < (int)gridDim.x; ++c) {
    if (p.slotjob[c % p.period] == jb) {
      if (c < (int)blockIdx.x) ++cta_in_job;
      ++ctas_in_job;
    }
  }
  const WgStripJob& job = p.job[jb];
  uint32_t tmem_cols = 32;
  {
    int mx = 0;
    for (int j = 0; j < p.njobs; ++j) mx = max(mx, p.job[j].nblk);
    while (tmem_cols < (uint32_t)mx * kAcc) tmem_cols <<= 1;
  }
  if (warp == 0 && lane == 0) {
    tmap_prefetch(&tmD_hi);
    tmap_prefetch(&tmD_lo);
    tmap_prefetch(&tmX_hi);
    tmap_prefetch(&tmX_lo);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 128);
    fence_barrier_init();
  }
  // everything a descriptor may touch outside the TMA boxes (the tails behind both strips) must be
  // finite: x is zero there, and 0 * NaN would poison dW
  for (int i = threadIdx.x; i < 2 * stage_bytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(ring)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  int my_items = 0;
  for (int item = cta_in_job; item < p.items; item += ctas_in_job) ++my_items;
  const int chains = (my_items + p.flush_items - 1) / p.flush_items;

  if (warp == 0) {
    if (elect_one()) {
      const uint32_t tx = (uint32_t)((p.R + 2) * p.Wp * SUBW * 2 + p.R * p.Wp * CIN * 2) * 2u;
      uint32_t g = 0;
      for (int item = cta_in_job; item < p.items; item += ctas_in_job, ++g) {
        const int n = item / p.strips;
        const int h0 = (item - n * p.strips) * p.R;
        const uint32_t s = g & 1;
        mbar_wait(&empty[s], ((g >> 1) & 1) ^ 1);
        uint8_t* st = ring + (size_t)s * stage_bytes;
        mbar_expect_tx(&full[s], tx);
        tma_load_4d(st, &tmD_hi, &full[s], job.co0, -1, h0 - 1, n);
        tma_load_4d(st + p.dy_plane, &tmD_lo, &full[s], job.co0, -1, h0 - 1, n);
#pragma unroll
        for (int c = 0; c < XCH; ++c) {
          tma_load_4d(st + 2 * p.dy_plane + c * p.x_chunk, &tmX_hi, &full[s], c * XW, -1, h0, n);
          tma_load_4d(st + 2 * p.dy_plane + x_plane + c * p.x_chunk, &tmX_lo, &full[s], c * XW, -1, h0, n);
        }
      }
    }
  } else if (warp == 1) {
    // D fp32, A/B bf16, both MN-major, M = 128 (SUBS taps x SUBW output channels), N = CIN (or 2 CIN)
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                               ((uint32_t)(CIN >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
    constexpr uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                ((uint32_t)((2 * CIN) >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
    constexpr uint32_t kDPix16 = (uint32_t)(SUBW * 2) >> 4;         // one pixel of the dy strip
    constexpr uint32_t kXPix16 = (uint32_t)(XW * 2) >> 4;
    constexpr uint32_t kDStep16 = 16u * kDPix16, kXStep16 = 16u * kXPix16;
    const uint32_t ring16 = smem_u32(ring) >> 4;
    const uint32_t dylo16 = (uint32_t)p.dy_plane >> 4, xlo16 = (uint32_t)x_plane >> 4;
    uint64_t ddesc[5], xdesc[5];
    // x: N spans XCH 64-channel chunks (leading-dimension offset = one chunk); with the stacked
    // issue the second block is the lo plane instead
    const uint32_t x_lbo = kStack ? (uint32_t)x_plane : (uint32_t)p.x_chunk;
#pragma unroll
    for (int b = 0; b < 5; ++b) {
      ddesc[b] = mnmajor_desc(0u, (uint32_t)job.blbo[b] * (uint32_t)(SUBW * 2), SUBW) +
                 (uint64_t)((uint32_t)job.boff[b] * kDPix16);
      xdesc[b] = mnmajor_desc(0u, x_lbo, XW) + (uint64_t)((uint32_t)job.xoff[b] * kXPix16);
    }
    const int nblk = job.nblk;
    if (elect_one()) {                 // one lane runs the whole loop, waits included
      uint32_t g = 0;
      int in_chain = 0, chain = 0, done = 0;
      for (int item = cta_in_job; item < p.items; item += ctas_in_job, ++g) {
        const uint32_t s = g & 1;
        if (in_chain == 0) {
          mbar_wait(tmem_empty, ((uint32_t)chain & 1) ^ 1);       // the epilogue drained the accumulators
          tc_fence_after();
        }
        mbar_wait(&full[s], (g >> 1) & 1);
        tc_fence_after();
        const uint32_t st16 = ring16 + s * ((uint32_t)stage_bytes >> 4);
        const uint32_t x16 = st16 + 2 * dylo16;
        for (int ks = 0; ks < p.ksteps; ++ks) {
          const uint32_t acc = (in_chain | ks) != 0 ? 1u : 0u;
#pragma unroll
          for (int b = 0; b < 5; ++b) {
            if (b < nblk) {
              const uint64_t dbh = xdesc[b] + (uint64_t)(x16 + (uint32_t)ks * kXStep16);
              const uint64_t dah = ddesc[b] + (uint64_t)(st16 + (uint32_t)ks * kDStep16);
              const uint64_t dal = dah + dylo16;
              const uint32_t tmem_d = tmem_base + (uint32_t)b * kAcc;
              if (kStack) {
                umma_bf16(tmem_d, dah, dbh, idesc2, acc);
                umma_bf16(tmem_d, dal, dbh, idesc, 1);
              } else {
                const uint64_t dbl = dbh + xlo16;
                umma_bf16(tmem_d, dal, dbh, idesc, acc);
                umma_bf16(tmem_d, dah, dbl, idesc, 1);
                umma_bf16(tmem_d, dah, dbh, idesc, 1);
              }
            }
          }
        }
        umma_commit(&empty[s]);
        ++in_chain;
        ++done;
        if (in_chain == p.flush_items || done == my_items) {
          umma_commit(tmem_full);
          in_chain = 0;
          ++chain;
        }
      }
    }
  } else {
    // epilogue: drain every finished chain into this CTA's partial gradient (same thread, same
    // address, fixed order: deterministic round-to-nearest additions)
    const int quad = warp & 3;
    const int m = quad * 32 + lane;
    const int sub = m / SUBW, co = job.co0 + m % SUBW;
    const size_t KK = (size_t)9 * CIN;
    for (int chain = 0; chain < chains; ++chain) {
      mbar_wait(tmem_full, (uint32_t)chain & 1);
      tc_fence_after();
#pragma unroll 1
      for (int b = 0; b < job.nblk; ++b) {
        const int tap = sub < SUBS ? job.tap[b][sub] : -1;
        float* dst = p.part + ((size_t)blockIdx.x * p.Cout + co) * KK + (size_t)(tap < 0 ? 0 : tap) * CIN;
#pragma unroll 1
        for (int c0 = 0; c0 < CIN; c0 += 32) {
          float v[32];
          const uint32_t ta = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)b * kAcc + (uint32_t)c0;
          tmem_ld32(ta, v);
          if (kStack) {
            float u[32];
            tmem_ld32(ta + (uint32_t)CIN, u);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += u[j];
          }
          if (tap >= 0) {
            float4* o4 = reinterpret_cast<float4*>(dst + c0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 o = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
              if (chain > 0) {
                const float4 a = o4[j];
                o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
              }
              o4[j] = o;
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive_relaxed(tmem_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// multi-job launches: each CTA's partial only holds its job's (output block, taps) region.  Sum
// every element over the CTAs of ITS job, in CTA order (deterministic), into one compact partial.
struct WgCompactArgs {
  int subw, Cin;
  int tapjob[4][9];                     // job that owns (output-channel block, tap)
  int count[kWgMaxJobs];                // CTAs of each job that had items
  unsigned char cta[kWgMaxJobs][160];   // their indices, ascending (the summation order)
};
__global__ void __launch_bounds__(256)
wgrad_strip_compact_kernel(const float* __restrict__ parts, float* __restrict__ out, long long full,
                           const WgCompactArgs a) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= full) return;
  const int co = (int)(e / (9 * a.Cin));
  const int tap = (int)((e / a.Cin) % 9);
  const int job = a.tapjob[co / a.subw][tap];
  const int n = a.count[job];
  float acc = 0.f;
  int i = 0;
  for (; i + 4 <= n; i += 4) {           // four independent loads in flight, fixed addition order
    const float v0 = __ldg(parts + (size_t)a.cta[job][i] * full + e);
    const float v1 = __ldg(parts + (size_t)a.cta[job][i + 1] * full + e);
    const float v2 = __ldg(parts + (size_t)a.cta[job][i + 2] * full + e);
    const float v3 = __ldg(parts + (size_t)a.cta[job][i + 3] * full + e);
    acc = (((acc + v0) + v1) + v2) + v3;
  }
  for (; i < n; ++i) acc += __ldg(parts + (size_t)a.cta[job][i] * full + e);
  out[e] = acc;
}

// ------------------------------------------------------------ operand preparation --
__global__ void __launch_bounds__(256)
split_bf16_kernel(const float* __restrict__ x, long long n4, __nv_bfloat16* __restrict__ hi,
                  __nv_bfloat16* __restrict__ lo) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
  float f[4] = {v.x, v.y, v.z, v.w};
  __nv_bfloat16 h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    h[j] = __float2bfloat16_rn(f[j]);
    l[j] = __float2bfloat16_rn(f[j] - __bfloat162float(h[j]));
  }
  reinterpret_cast<uint2*>(hi)[i] = *reinterpret_cast<uint2*>(h);
  if (lo) reinterpret_cast<uint2*>(lo)[i] = *reinterpret_cast<uint2*>(l);
}

// fp16 variant for forward activations (bounded by the normalisations: |x| << 65504):
// hi + lo carries 22 mantissa bits instead of bf16's 16.
__global__ void __launch_bounds__(256)
split_f16_kernel(const float* __restrict__ x, long long n4, __half* __restrict__ hi,
                 __half* __restrict__ lo) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
  float f[4] = {v.x, v.y, v.z, v.w};
  __half h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    h[j] = __float2half_rn(f[j]);
    l[j] = __float2half_rn(f[j] - __half2float(h[j]));
  }
  reinterpret_cast<uint2*>(hi)[i] = *reinterpret_cast<uint2*>(h);
  if (lo) reinterpret_cast<uint2*>(lo)[i] = *reinterpret_cast<uint2*>(l);
}

// OIHW fp32 -> K-major bf16 hi/lo.  flip == 0: forward  Wk[co][(r*KW+q)*Cin + ci]
//                                  flip == 1: dgrad    Wk[ci][((KH-1-r)*KW + (KW-1-q))*Cout + co]
__global__ void __launch_bounds__(256)
prep_weights_tc_kernel(const float* __restrict__ w, int Cout, int Cin, int KH, int KW, int flip,
                       int fmt, float scale, uint16_t* __restrict__ hi,
                       uint16_t* __restrict__ lo) {
  size_t total = (size_t)Cout * Cin * KH * KW;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int q = (int)(i % KW);
  size_t t = i / KW;
  int r = (int)(t % KH);
  t /= KH;
  int ci = (int)(t % Cin);
  int co = (int)(t / Cin);
  float v = w[i];
  size_t o;
  if (!flip)
    o = (size_t)co * ((size_t)KH * KW * Cin) + (size_t)(r * KW + q) * Cin + ci;
  else
    o = (size_t)ci * ((size_t)KH * KW * Cout) + (size_t)((KH - 1 - r) * KW + (KW - 1 - q)) * Cout + co;
  if (fmt == 1) {
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[o] = __bfloat16_as_ushort(h);
    if (lo) lo[o] = __bfloat16_as_ushort(__float2bfloat16_rn(v - __bfloat162float(h)));
  } else {
    v *= scale;   // power of two: keeps hi and lo in fp16's normal range, undone in the epilogue
    __half h = __float2half_rn(v);
    hi[o] = __half_as_ushort(h);
    if (lo) lo[o] = __half_as_ushort(__float2half_rn(v - __half2float(h)));
  }
}

// ------------------------------------------------------------------ host helpers --
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

CUtensorMapSwizzle swizzle_for(int cw) {
  return cw == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                  : (cw == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// NHWC bf16 tensor, box = [bn][bh][bw][cw]; `stride` > 1 traverses W and H with that step
// (the box then spans stride*b input pixels and delivers b of them).
int make_map_nhwc(CUtensorMap* m, const void* base, int N, int H, int W, int C, int cw, int bw,
                  int bh, int bn, int stride = 1, int fmt = 1) {
  EncodeTiledFn enc = encode_fn();
  EVE_REQUIRE(enc, EVE_ERR_CUDA, "conv_tc: cuTensorMapEncodeTiled is not available");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)cw, (cuuint32_t)(bw * stride), (cuuint32_t)(bh * stride),
                       (cuuint32_t)bn};
  cuuint32_t es[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = enc(m, fmt == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                   4, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(cw),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EVE_REQUIRE(r == CUDA_SUCCESS, EVE_ERR_CUDA,
              "conv_tc: cuTensorMapEncodeTiled(NHWC %dx%dx%dx%d box %dx%dx%dx%d s%d) failed: %d", N,
              H, W, C, bn, bh, bw, cw, stride, (int)r);
  return EVE_OK;
}

// Stem windows (conv_tc_stem_run): a rank-4 view [N][Hp][OW][32] whose rows of 32 elements start
// every `win_bytes` bytes along OW (64: one materialised window per output column; 16: windows that
// OVERLAP in a zero-padded 4-channel image, 2 pixels apart) and every `pitch_bytes` along Hp; the box
// takes every second row (the convolution's vertical stride).
int make_map_windows(CUtensorMap* m, const void* base, int N, int Hp, int OW, int win_bytes,
                     int pitch_bytes, int bw, int bh, int bn, int fmt) {
  EncodeTiledFn enc = encode_fn();
  EVE_REQUIRE(enc, EVE_ERR_CUDA, "conv_tc: cuTensorMapEncodeTiled is not available");
  cuuint64_t dims[4] = {32, (cuuint64_t)OW, (cuuint64_t)Hp, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)win_bytes, (cuuint64_t)pitch_bytes,
                           (cuuint64_t)Hp * (cuuint64_t)pitch_bytes};
  cuuint32_t box[4] = {32, (cuuint32_t)bw, (cuuint32_t)(bh * 2), (cuuint32_t)bn};
  cuuint32_t es[4] = {1, 1, 2, 1};
  CUresult r = enc(m, fmt == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                   4, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(32),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EVE_REQUIRE(r == CUDA_SUCCESS, EVE_ERR_CUDA,
              "conv_tc: cuTensorMapEncodeTiled(stem windows %dx%dx%d, %d / %d bytes) failed: %d", N, Hp,
              OW, win_bytes, pitch_bytes, (int)r);
  return EVE_OK;
}

int make_map_2d(CUtensorMap* m, const void* base, int rows, int cols, int cw, int box_rows,
                int fmt = 1) {
  EncodeTiledFn enc = encode_fn();
  EVE_REQUIRE(enc, EVE_ERR_CUDA, "conv_tc: cuTensorMapEncodeTiled is not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)cw, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, fmt == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                   2, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(cw),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EVE_REQUIRE(r == CUDA_SUCCESS, EVE_ERR_CUDA, "conv_tc: cuTensorMapEncodeTiled(B) failed: %d",
              (int)r);
  return EVE_OK;
}

// pixel box (bw, bh, bn) with bw == W, bw*bh*bn <= 128, maximising useful rows per MMA
void pick_box(int N, int H, int W, int& bw, int& bh, int& bn) {
  bw = W;
  double best = -1.0;
  bh = 1;
  bn = 1;
  for (int h = 1; h <= H && W * h <= kTileM; ++h) {
    int nmax = kTileM / (W * h);
    for (int n = 1; n <= nmax && n <= (h == H ? N : 1); ++n) {
      // spanning several images only makes sense when the box covers whole images
      double eff = ((double)H / (double)(cdiv(H, h) * h)) * ((double)(W * h * n) / kTileM) *
                   ((double)N / (double)(cdiv(N, n) * n));
      if (eff > best + 1e-9) {
        best = eff;
        bh = h;
        bn = n;
      }
    }
  }
}

template <int BN, int NPASS, bool STACK, bool PAIR, bool DUAL = false>
int launch_tc_kernel(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi,
                     const CUtensorMap& b_lo, const TcParams& p, int grid, int smem_bytes,
                     cudaStream_t s) {
  EVE_TRY(ensure_dynamic_smem((const void*)conv_tc_kernel<BN, NPASS, STACK, PAIR, DUAL>, 227 * 1024));
  constexpr int threads = DUAL ? kThreads + 32 : kThreads;
  if (!PAIR) {
    conv_tc_kernel<BN, NPASS, STACK, false, DUAL><<<grid, threads, smem_bytes, s>>>(a_hi, a_lo, b_hi, b_lo, p);
    EVE_LAUNCH_CHECK();
    return EVE_OK;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3(threads, 1, 1);
  cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  EVE_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<BN, NPASS, STACK, PAIR, false>, a_hi, a_lo, b_hi, b_lo, p));
  count_launch();
  return EVE_OK;
}

// pair: clusters of two CTAs sharing each weight stage (see conv_tc_kernel); the B maps then have
// BN / 2-row boxes
template <int BN, int NPASS, bool STACK>
int launch_tc(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi,
              const CUtensorMap& b_lo, const TcParams& p0, int tiles_m, int tiles_co,
              cudaStream_t s, bool pair = false) {
  using Cfg = TcCfg<BN, NPASS, STACK>;
  TcParams p = p0;
  p.tiles_m = tiles_m;
  p.tiles_co = tiles_co;
  // ring slots sized for this layer's chunk width (swizzle atoms need 1024-byte alignment)
  p.a_bytes = kTileM * p.kc * 2;
  // stacked-B issue reads [B_hi ; B_lo] through one descriptor: the lo plane follows the hi plane
  // without padding (BN * kc * 2 is a multiple of every swizzle atom; a stage stays 1024-aligned)
  p.b_bytes = (Cfg::kStack || BN * p.kc * 2 >= 1024) ? BN * p.kc * 2 : 1024;
  p.stage_bytes = (p.a_bytes + p.b_bytes) * Cfg::kPlanes;
  // experiment switch: two persistent CTAs per SM, each with half the ring (two MMA issuers per SM)
  static const int ctas = [] { const char* e = getenv("EVE_B200_TC_CTAS"); return e && atoi(e) == 2 ? 2 : 1; }();
  const int per_sm = (ctas == 2 && !pair && 2 * 2 * (int)Cfg::kTmemCols <= 512 &&
                      ((kSmemBudget / 2 - 1024 - kBarrierBytes - kEpiBytes) / p.stage_bytes) >= 3) ? 2 : 1;
  p.stages = (kSmemBudget / per_sm - 1024 - kBarrierBytes - kEpiBytes) / p.stage_bytes;
  const int cap = tc_stage_cap();
  if (p.stages > cap) p.stages = cap;
  const int smem_bytes = p.stages * p.stage_bytes + 1024 /*align*/ + kBarrierBytes + kEpiBytes;
  if (per_sm == 2) {
    const long long total2 = (long long)tiles_m * tiles_co;
    const int grid2 = (int)(total2 < 2 * kNumSMs ? total2 : 2 * kNumSMs);
    return launch_tc_kernel<BN, NPASS, STACK, false>(a_hi, a_lo, b_hi, b_lo, p, grid2, smem_bytes, s);
  }
  if (pair) {
    constexpr bool kHasPair = NPASS == 3 && ((BN == 128 && !STACK) || (BN == 64 && STACK));
    if (kHasPair) {
      const long long rows = (long long)((tiles_m + 1) / 2) * tiles_co;
      const int pairs = (int)(rows < kNumSMs / 2 ? rows : kNumSMs / 2);
      return launch_tc_kernel<BN, NPASS, STACK, kHasPair>(a_hi, a_lo, b_hi, b_lo, p, 2 * pairs, smem_bytes, s);
    }
  }
  const long long total = (long long)tiles_m * tiles_co;
  const int grid = (int)(total < kNumSMs ? total : kNumSMs);   // one persistent CTA per SM
  // two issuing warps with one partial accumulator each (TMEM: 2 buffers x 2 x 128 columns); a property
  // of the layer's tile shape only, never of the batch (the two partial sums round differently)
  // (64 stacked output channels: the ring is deep enough to be made even; the 128-channel tiles have
  // three 64 KB stages)
  constexpr bool kHasDual = NPASS == 3 && BN == 64 && STACK;
  if (kHasDual && get_option(OPT_TC_DUAL) != 0 && p.ntaps * p.kchunks >= 2 && p.stages >= 4) {
    p.stages &= ~1;
    return launch_tc_kernel<BN, NPASS, STACK, false, kHasDual>(a_hi, a_lo, b_hi, b_lo, p, grid, smem_bytes, s);
  }
  return launch_tc_kernel<BN, NPASS, STACK, false>(a_hi, a_lo, b_hi, b_lo, p, grid, smem_bytes, s);
}

template <int NPASS>
int launch_tc_bn(int BN, bool stack, const CUtensorMap& a_hi, const CUtensorMap& a_lo,
                 const CUtensorMap& b_hi, const CUtensorMap& b_lo, const TcParams& p, int gx, int gy,
                 cudaStream_t s, bool pair = false) {
  if (NPASS == 3 && stack) {
    switch (BN) {
      case 64: return launch_tc<64, NPASS, true>(a_hi, a_lo, b_hi, b_lo, p, gx, gy, s, pair);
      case 32: return launch_tc<32, NPASS, true>(a_hi, a_lo, b_hi, b_lo, p, gx, gy, s);
      default: return launch_tc<16, NPASS, true>(a_hi, a_lo, b_hi, b_lo, p, gx, gy, s);
    }
  }
  switch (BN) {
    case 128: return launch_tc<128, NPASS, false>(a_hi, a_lo, b_hi, b_lo, p, gx, gy, s, pair);
    case 64: return launch_tc<64, NPASS, false>(a_hi, a_lo, b_hi, b_lo, p, gx, gy, s);
    case 32: return launch_tc<32, NPASS, false>(a_hi, a_lo, b_hi, b_lo, p, gx, gy, s);
    default: return launch_tc<16, NPASS, false>(a_hi, a_lo, b_hi, b_lo, p, gx, gy, s);
  }
}

// channels per K chunk (= TMA box width and swizzle span): multiples of 64, then multiples of 32
// (32 itself and the stem's 160-wide patch matrix), then 16
inline int chunk_for(int C) { return C % 64 == 0 ? 64 : (C % 32 == 0 ? 32 : (C == 16 ? 16 : 0)); }
inline int bn_for(int Cout) {
  return Cout % 128 == 0 ? 128 : (Cout % 64 == 0 ? 64 : (Cout == 32 ? 32 : (Cout == 16 ? 16 : 0)));
}

}  // namespace

// --------------------------------------------------------------------- public API --
// Forward-style geometries the tensor-core kernel takes: 1x1 / 3x3, stride 1 with "same"
// padding or stride 2 (pad = k/2), channel counts 16, 32 or multiples of 64.
bool conv_tc_supported(const ConvGeom& g) {
  if (g.KH != g.KW || (g.KH != 1 && g.KH != 3) || g.N < 1) return false;
  if (g.pad != g.KH / 2) return false;
  if (g.stride == 1) {
    if (g.OH != g.H || g.OW != g.W) return false;
  } else if (g.stride == 2) {
    if (g.H % 2 != 0 || g.W % 2 != 0) return false;
  } else {
    return false;
  }
  if (chunk_for(g.Cin) == 0 || bn_for(g.Cout) == 0) return false;
  if (g.OW > kTileM || g.OW < 1) return false;
  return true;
}

int split_planes(const float* x, long long n, void* hi, void* lo, int fmt, cudaStream_t s) {
  EVE_REQUIRE(n % 4 == 0, EVE_ERR_SHAPE, "split_planes: element count must be a multiple of 4");
  long long n4 = n / 4;
  if (n4 == 0) return EVE_OK;
  if (fmt == 1)
    split_bf16_kernel<<<cdiv(n4, 256), 256, 0, s>>>(x, n4, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo);
  else
    split_f16_kernel<<<cdiv(n4, 256), 256, 0, s>>>(x, n4, (__half*)hi, (__half*)lo);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

int conv_tc_prep_weights(const ConvGeom& g, const float* w_oihw, bool dgrad, void* hi, void* lo,
                         int fmt, float scale, cudaStream_t s) {
  size_t total = (size_t)g.Cout * g.Cin * g.KH * g.KW;
  prep_weights_tc_kernel<<<cdiv(total, 256), 256, 0, s>>>(w_oihw, g.Cout, g.Cin, g.KH, g.KW,
                                                          dgrad ? 1 : 0, fmt, scale, (uint16_t*)hi,
                                                          (uint16_t*)lo);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

static int tc_launch(const TcParams& p0, const void* x_hi, const void* x_lo, int inN, int inH,
                     int inW, const void* w_hi, const void* w_lo, int wrows, int wcols, int npass,
                     int fmt, cudaStream_t s, int win_bytes = 0, int pitch_bytes = 0) {
  TcParams p = p0;
  pick_box(p.N, p.OH, p.OW, p.bw, p.bh, p.bn);
  p.tiles_h = cdiv(p.OH, p.bh);
  p.kc = chunk_for(p.Cin);
  {
    // experiment switch: narrower K chunks (same sequence of K = 16 steps, smaller ring stages)
    static const int env_kc = [] { const char* e = getenv("EVE_B200_TC_KC"); return e ? atoi(e) : 0; }();
    if ((env_kc == 32 || env_kc == 16) && env_kc < p.kc && p.Cin % env_kc == 0) p.kc = env_kc;
  }
  p.kchunks = p.Cin / p.kc;
  p.fmt = fmt;
  const int tiles_n = cdiv(p.N, p.bn);
  int BN = bn_for(p.Cout);
  const bool stack = BN <= 64;          // decided by the layer, before any narrowing (see TcCfg)
  // Few pixel tiles (the per-time-step ConvRNN gate convolutions have 3): narrower output-channel
  // tiles put more SMs on the layer; each CTA's serial chain of MMAs shrinks by the same factor.
  while (BN > 16 && (long long)p.tiles_h * tiles_n * (p.Cout / BN) < kNumSMs / 2) BN >>= 1;
  // CTA pairs sharing the weight stages: the split-operand kernels with 64 (stacked) or 128 output
  // channels per tile, when every pair has at least one pair-row of tiles (a launch-shape decision
  // only: the arithmetic per tile does not change, so it may depend on the batch)
  const int popt = get_option(OPT_TC_PAIR);      // 2 = whenever there are two tiles (tests)
  const long long pair_rows = (long long)((p.tiles_h * tiles_n + 1) / 2) * (p.Cout / BN);
  const bool pair = popt != 0 && npass == 3 && ((BN == 128 && !stack) || (BN == 64 && stack)) &&
                    p.tiles_h * tiles_n >= 2 && (popt == 2 || pair_rows >= kNumSMs / 2);
  const int b_box = pair ? BN / 2 : BN;
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
  if (win_bytes)
    EVE_TRY(make_map_windows(&a_hi, x_hi, inN, inH, inW, win_bytes, pitch_bytes, p.bw, p.bh, p.bn, fmt));
  else
    EVE_TRY(make_map_nhwc(&a_hi, x_hi, inN, inH, inW, p.Cin, p.kc, p.bw, p.bh, p.bn, p.stride, fmt));
  EVE_TRY(make_map_2d(&b_hi, w_hi, wrows, wcols, p.kc, b_box, fmt));
  if (npass == 3) {
    if (win_bytes)
      EVE_TRY(make_map_windows(&a_lo, x_lo, inN, inH, inW, win_bytes, pitch_bytes, p.bw, p.bh, p.bn,
                               fmt));
    else
      EVE_TRY(make_map_nhwc(&a_lo, x_lo, inN, inH, inW, p.Cin, p.kc, p.bw, p.bh, p.bn, p.stride,
                            fmt));
    EVE_TRY(make_map_2d(&b_lo, w_lo, wrows, wcols, p.kc, b_box, fmt));
  } else {
    a_lo = a_hi;
    b_lo = b_hi;
  }
  const int gx = p.tiles_h * tiles_n, gy = p.Cout / BN;
  return npass == 3 ? launch_tc_bn<3>(BN, stack, a_hi, a_lo, b_hi, b_lo, p, gx, gy, s, pair)
                    : launch_tc_bn<1>(BN, false, a_hi, a_lo, b_hi, b_lo, p, gx, gy, s);
}

// ---- halo-row kernels: planning and launch
static bool row_geometry_ok(const ConvGeom& g) {
  if (g.KH != g.KW || (g.KH != 3 && g.KH != 1) || g.stride != 1 || g.pad != g.KH / 2) return false;
  if (g.W != kTileM || g.OW != kTileM || g.OH != g.H || g.N < 1) return false;
  if (g.Cin != 16 && g.Cin != 32 && g.Cin != 64) return false;
  if (g.Cout != 16 && g.Cout != 32 && g.Cout != 64) return false;
  if (g.KH == 3 && g.Cin == 64 && g.Cout == 64) return false;   // nine resident 64x64 taps leave < 4 row slots
  return true;
}

// strips of consecutive rows per work item: balance over the SMs vs halo rows staged twice
static void pick_strips(int N, int H, int& strips, int& rows_per_strip) {
  const int forced = get_option(OPT_TC_ROW_STRIPS);
  int best_s = 1;
  double best = -1.0;
  for (int st = 1; st <= H; ++st) {
    const int rps = cdiv(H, st);
    if (cdiv(H, rps) != st) continue;             // skip strip counts that leave empty strips
    const long long items = (long long)N * st;
    const double balance = (double)items / (double)(cdiv(items, kNumSMs) * kNumSMs);
    const double eff = balance * (double)rps / (double)(rps + 1);
    if (forced ? st == forced : eff > best + 1e-9) {
      best = eff;
      best_s = st;
    }
  }
  strips = best_s;
  rows_per_strip = cdiv(H, best_s);
}

bool conv_tc_row_supported(const ConvGeom& g) {
  return get_option(OPT_TC_ROW_KERNEL) != 0 && row_geometry_ok(g);
}

template <int BN, int NPASS, int KC, int KS>
static int launch_row(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi,
                      const CUtensorMap& b_lo, TcRowParams p, cudaStream_t s) {
  using Cfg = RowCfg<BN, NPASS, KC, KS>;
  if (Cfg::kSlots < 4) {
    EVE_REQUIRE(false, EVE_ERR_SHAPE, "conv_tc_row: %d -> %d channels do not fit", KC, BN);
    return EVE_ERR_SHAPE;
  }
  EVE_TRY(ensure_dynamic_smem((const void*)conv_tc_row_kernel<BN, NPASS, KC, KS>, 227 * 1024));
  p.slots = Cfg::kSlots;
  const int smem_bytes = Cfg::kFixedBytes + Cfg::kSlots * Cfg::kSlotBytes;
  const int grid = p.items < kNumSMs ? p.items : kNumSMs;
  conv_tc_row_kernel<BN, NPASS, KC, KS><<<grid, kThreads, smem_bytes, s>>>(a_hi, a_lo, b_hi, b_lo, p);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

template <int BN, int KC>
static int launch_row_np(int ks, int npass, const CUtensorMap& a_hi, const CUtensorMap& a_lo,
                         const CUtensorMap& b_hi, const CUtensorMap& b_lo, const TcRowParams& p,
                         cudaStream_t s) {
  if (ks == 3)
    return npass == 3 ? launch_row<BN, 3, KC, 3>(a_hi, a_lo, b_hi, b_lo, p, s)
                      : launch_row<BN, 1, KC, 3>(a_hi, a_lo, b_hi, b_lo, p, s);
  return npass == 3 ? launch_row<BN, 3, KC, 1>(a_hi, a_lo, b_hi, b_lo, p, s)
                    : launch_row<BN, 1, KC, 1>(a_hi, a_lo, b_hi, b_lo, p, s);
}

template <int BN>
static int launch_row_kc(int kc, int ks, int npass, const CUtensorMap& a_hi, const CUtensorMap& a_lo,
                         const CUtensorMap& b_hi, const CUtensorMap& b_lo, const TcRowParams& p,
                         cudaStream_t s) {
  switch (kc) {
    case 16: return launch_row_np<BN, 16>(ks, npass, a_hi, a_lo, b_hi, b_lo, p, s);
    case 32: return launch_row_np<BN, 32>(ks, npass, a_hi, a_lo, b_hi, b_lo, p, s);
    default: return launch_row_np<BN, 64>(ks, npass, a_hi, a_lo, b_hi, b_lo, p, s);
  }
}

static int conv_tc_row_run(const ConvGeom& g, const void* x_hi, const void* x_lo, const void* w_hi,
                           const void* w_lo, const float* bias, const float* addend, float* y,
                           int npass, int fmt, float out_scale, cudaStream_t s) {
  EVE_REQUIRE(row_geometry_ok(g), EVE_ERR_SHAPE, "conv_tc_row: unsupported geometry");
  TcRowParams p;
  p.N = g.N; p.H = g.H; p.Cout = g.Cout;
  pick_strips(g.N, g.H, p.strips, p.rows_per_strip);
  p.items = g.N * p.strips;
  p.fmt = fmt;
  p.slots = 0;
  p.out_scale = out_scale;
  p.bias = bias; p.addend = addend; p.out = y;
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
  EVE_TRY(make_map_nhwc(&a_hi, x_hi, g.N, g.H, g.W, g.Cin, g.Cin, kRowBox, 1, 1, 1, fmt));
  EVE_TRY(make_map_2d(&b_hi, w_hi, g.Cout, g.KH * g.KW * g.Cin, g.Cin, g.Cout, fmt));
  if (npass == 3) {
    EVE_TRY(make_map_nhwc(&a_lo, x_lo, g.N, g.H, g.W, g.Cin, g.Cin, kRowBox, 1, 1, 1, fmt));
    EVE_TRY(make_map_2d(&b_lo, w_lo, g.Cout, g.KH * g.KW * g.Cin, g.Cin, g.Cout, fmt));
  } else {
    a_lo = a_hi;
    b_lo = b_hi;
  }
  switch (g.Cout) {
    case 16: return launch_row_kc<16>(g.Cin, g.KH, npass, a_hi, a_lo, b_hi, b_lo, p, s);
    case 32: return launch_row_kc<32>(g.Cin, g.KH, npass, a_hi, a_lo, b_hi, b_lo, p, s);
    default: return launch_row_kc<64>(g.Cin, g.KH, npass, a_hi, a_lo, b_hi, b_lo, p, s);
  }
}

// ---- padded-strip kernel: planning and launch
struct StripPlan {
  int BN, KC, R, bn, Tmax, nsets, a_plane_bytes, b_stages;
  double score;      // useful fraction of the issued MMA rows x wave balance
};

static int strip_bn_for(int Cout) { return Cout % 128 == 0 ? 128 : (Cout % 64 == 0 ? 64 : (Cout == 32 ? 32 : 0)); }

// efficiency of the generic kernel's pixel boxes on the same geometry (for the auto switch)
static double generic_box_efficiency(const ConvGeom& g) {
  int bw, bh, bn;
  pick_box(g.N, g.OH, g.OW, bw, bh, bn);
  return ((double)g.OH / (double)(cdiv(g.OH, bh) * bh)) * ((double)(bw * bh * bn) / kTileM) *
         ((double)g.N / (double)(cdiv(g.N, bn) * bn));
}

static bool strip_plan(const ConvGeom& g, StripPlan& best, int force_kc = 0) {
  if (g.KH != 3 || g.KW != 3 || g.stride != 1 || g.pad != 1 || g.OH != g.H || g.OW != g.W) return false;
  if (g.W + 2 > 256 || g.W >= kTileM || g.N < 1) return false;
  const int BN = strip_bn_for(g.Cout);
  if (BN == 0) return false;
  const int Wp = g.W + 2;
  const int usable = 227 * 1024 - 1024 - kBarrierBytes - kEpiBytes;
  best.score = -1.0;
  // experiment switches (tools/probe_strip.py): force the strip height / chunk width
  static const int env_r = [] { const char* e = getenv("EVE_B200_STRIP_R"); return e ? atoi(e) : 0; }();
  static const int env_kc = [] { const char* e = getenv("EVE_B200_STRIP_KC"); return e ? atoi(e) : 0; }();
  if (env_kc && !force_kc && g.Cin % env_kc == 0) force_kc = env_kc;
  const int kcs[3] = {64, 32, 16};
  for (int ki = 0; ki < 3; ++ki) {
    const int KC = kcs[ki];
    if (g.Cin % KC != 0 || (force_kc && KC != force_kc)) continue;
    const int bsz = 2 * (BN * KC * 2 < 1024 ? 1024 : BN * KC * 2);
    for (int bn = 1; bn <= 16 && bn <= g.N; ++bn) {
      for (int R = (bn == 1 ? 1 : g.H); R <= g.H; ++R) {
        if (env_r && bn == 1 && R != env_r && env_r <= g.H) continue;
        const int S = (R + 2) * Wp;
        const int Tmax = cdiv((long long)(bn - 1) * S + (long long)R * Wp, kTileM);
        const int acc = BN <= 64 ? 2 * BN : BN;        // stacked-B accumulators are twice as wide
        if (Tmax * acc > 512) break;                   // grows with R
        const int arows = Tmax * kTileM + 2 * Wp + 2;
        const int a_plane = (int)align_up((size_t)arows * KC * 2, 1024);
        int stages = (usable - 4 * a_plane) / bsz;
        if (stages < 3) break;
        if (stages > 12) stages = 12;
        const int strips = cdiv(g.H, R), groups = cdiv(g.N, bn);
        const int last_rows = g.H - (strips - 1) * R;
        // tiles over all items of one output-channel tile (the tail image group counts as full)
        const long long t_full = cdiv((long long)(bn - 1) * S + (long long)R * Wp, kTileM);
        const long long t_last = cdiv((long long)(bn - 1) * S + (long long)last_rows * Wp, kTileM);
        const long long tiles = (long long)groups * ((strips - 1) * t_full + t_last);
        const double eff = (double)g.N * g.H * g.W / ((double)tiles * kTileM);
        const long long items = (long long)groups * strips * (g.Cout / BN);
        const double bal = (double)items / (double)((long long)cdiv(items, kNumSMs) * kNumSMs);
        const int nsets = 2 * Tmax * acc <= 512 ? 2 : 1;
        const double score = eff * bal * (KC == 64 ? 1.0 : (KC == 32 ? 0.97 : 0.9)) *
                             (nsets == 2 ? 1.0 : 0.95) * ((double)R / (R + 2) * 0.1 + 0.9);
        if (score > best.score + 1e-9) {
          best = StripPlan{BN, KC, R, bn, Tmax, nsets, a_plane, stages, score};
        }
      }
    }
  }
  return best.score > 0.0;
}

// The strip and the box kernel sum a pixel's products in different orders (as do two K chunk
// widths), and a frame's result must not depend on the batch it is processed in
// (test_inference_stream_900_frames_in_chunks): the kernel choice and the chunk width are therefore
// decided for a canonical batch (one full wave of images per SM), i.e. from (H, W, Cin, Cout) alone;
// only the strip height / images per item -- which do not change any output bit -- follow the real N.
static ConvGeom canonical_batch(const ConvGeom& g) {
  ConvGeom c = g;
  c.N = 64 * kNumSMs;
  return c;
}

static bool strip_decision(const ConvGeom& g, StripPlan& canon) {
  const int opt = get_option(OPT_TC_STRIP);
  if (opt == 0) return false;
  const ConvGeom c = canonical_batch(g);
  if (!strip_plan(c, canon)) return false;
  if (opt == 2) return true;
  // auto (calibrated with tools/probe_strip.py at the bench geometries, profiles/r2_probe_strip.txt):
  // the strip kernel issues junk rows (pad columns, halo rows, partial tiles) but moves 3-7x fewer
  // operand bytes and amortises each weight tile over all tiles of an item.  It wins where the box
  // kernel's own pixel boxes are inefficient (9x16 maps: 56 % useful rows, 1.07-1.42x) and on
  // 32-channel outputs (1.02-1.36x); at 64 channels it needs both accumulator sets (18x32: 1.07-
  // 1.16x, while 32x32 / 36x64 plans only fit one set and lose 15-30 %); on the well-filled
  // 128-channel maps the box kernel already runs at 70-78 % of the split-operand peak.
  const double ge = generic_box_efficiency(c);
  if (canon.BN == 32) return true;
  if (canon.BN == 64) return canon.nsets == 2 && canon.score >= 0.78 * ge;
  return canon.score >= 0.9 * ge;
}

bool conv_tc_strip_supported(const ConvGeom& g) {
  StripPlan canon, pl;
  return strip_decision(g, canon) && strip_plan(g, pl, canon.KC);
}

template <int BN, int KC>
static int launch_strip(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi,
                        const CUtensorMap& b_lo, const TcStripParams& p, cudaStream_t s) {
  EVE_TRY(ensure_dynamic_smem((const void*)conv_tc_strip_kernel<BN, KC>, 227 * 1024));
  constexpr int kBPlane = BN * KC * 2 < 1024 ? 1024 : BN * KC * 2;
  const int smem_bytes = 4 * p.a_plane_bytes + p.b_stages * 2 * kBPlane + 1024 + kBarrierBytes + kEpiBytes;
  const int grid = p.items < kNumSMs ? p.items : kNumSMs;
  conv_tc_strip_kernel<BN, KC><<<grid, kStripThreads, smem_bytes, s>>>(a_hi, a_lo, b_hi, b_lo, p);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

template <int BN>
static int launch_strip_kc(int kc, const CUtensorMap& a_hi, const CUtensorMap& a_lo,
                           const CUtensorMap& b_hi, const CUtensorMap& b_lo, const TcStripParams& p,
                           cudaStream_t s) {
  switch (kc) {
    case 64: return launch_strip<BN, 64>(a_hi, a_lo, b_hi, b_lo, p, s);
    case 32: return launch_strip<BN, 32>(a_hi, a_lo, b_hi, b_lo, p, s);
    default: return launch_strip<BN, 16>(a_hi, a_lo, b_hi, b_lo, p, s);
  }
}

static int conv_tc_strip_run(const ConvGeom& g, const void* x_hi, const void* x_lo, const void* w_hi,
                             const void* w_lo, const float* bias, const float* addend, float* y,
                             int fmt, float out_scale, cudaStream_t s) {
  StripPlan canon, pl;
  EVE_REQUIRE(strip_decision(g, canon) && strip_plan(g, pl, canon.KC), EVE_ERR_SHAPE,
              "conv_tc_strip: unsupported geometry");
  TcStripParams p;
  p.N = g.N; p.H = g.H; p.W = g.W; p.Cout = g.Cout; p.Wp = g.W + 2;
  p.R = pl.R; p.bn = pl.bn;
  p.strips = cdiv(g.H, pl.R);
  p.groups = cdiv(g.N, pl.bn);
  p.S = (pl.R + 2) * p.Wp;
  p.tiles_co = g.Cout / pl.BN;
  p.items = p.groups * p.strips * p.tiles_co;
  p.kchunks = g.Cin / pl.KC;
  p.Tmax = pl.Tmax; p.nsets = pl.nsets;
  p.a_plane_bytes = pl.a_plane_bytes;
  p.b_stages = pl.b_stages;
  p.fmt = fmt;
  p.out_scale = out_scale;
  p.bias = bias; p.addend = addend; p.out = y;
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
  EVE_TRY(make_map_nhwc(&a_hi, x_hi, g.N, g.H, g.W, g.Cin, pl.KC, p.Wp, pl.R + 2, pl.bn, 1, fmt));
  EVE_TRY(make_map_nhwc(&a_lo, x_lo, g.N, g.H, g.W, g.Cin, pl.KC, p.Wp, pl.R + 2, pl.bn, 1, fmt));
  EVE_TRY(make_map_2d(&b_hi, w_hi, g.Cout, 9 * g.Cin, pl.KC, pl.BN, fmt));
  EVE_TRY(make_map_2d(&b_lo, w_lo, g.Cout, 9 * g.Cin, pl.KC, pl.BN, fmt));
  switch (pl.BN) {
    case 128: return launch_strip_kc<128>(pl.KC, a_hi, a_lo, b_hi, b_lo, p, s);
    case 64: return launch_strip_kc<64>(pl.KC, a_hi, a_lo, b_hi, b_lo, p, s);
    default: return launch_strip_kc<32>(pl.KC, a_hi, a_lo, b_hi, b_lo, p, s);
  }
}

// which forward-style kernel conv_tc_run() picks for `g`, and with what plan (debugging / docs)
int conv_tc_describe(const ConvGeom& g, char* buf, size_t cap) {
  if (!conv_tc_supported(g)) return snprintf(buf, cap, "not on the tensor-core path");
  if (conv_tc_row_supported(g)) return snprintf(buf, cap, "halo-row kernel");
  StripPlan canon, pl;
  const bool take = strip_decision(g, canon);
  const bool have = take ? strip_plan(g, pl, canon.KC) : strip_plan(g, pl);
  const double ge = generic_box_efficiency(g);
  if (have && take)
    return snprintf(buf, cap, "strip kernel BN=%d KC=%d R=%d bn=%d T=%d sets=%d stages=%d score=%.3f (box kernel %.3f)",
                    pl.BN, pl.KC, pl.R, pl.bn, pl.Tmax, pl.nsets, pl.b_stages, pl.score, ge);
  if (have)
    return snprintf(buf, cap, "box kernel eff=%.3f (strip plan BN=%d KC=%d R=%d bn=%d T=%d sets=%d score=%.3f)", ge,
                    pl.BN, pl.KC, pl.R, pl.bn, pl.Tmax, pl.nsets, pl.score);
  return snprintf(buf, cap, "box kernel eff=%.3f", ge);
}

// ---- persistent ConvGRU sequence kernel
bool cgru_seq_supported(int nf, int H, int W) {
  return get_option(OPT_CGRU_PERSISTENT) != 0 && conv_mode() == 1 && nf == kCgNf && H == kCgH && W == kCgW;
}

// w1h_*: fp16 hi/lo K-major planes [128][9 * 64] of gates_1.weight[:, nf:2nf] (pre-scaled by
// 1 / out_scale), w2h_*: [64][9 * 64] of gate_2.weight[:, 0:nf]
int cgru_seq_fwd(int B, int T, const void* w1h_hi, const void* w1h_lo, const void* w2h_hi,
                 const void* w2h_lo, const float* gx1, const float* gx2, const float* h0, float* r,
                 float* z, float* n, float* h, float* xh, float* cat2, float out_scale,
                 cudaStream_t s) {
  EVE_REQUIRE(B > 0 && T > 0, EVE_ERR_SHAPE, "cgru_seq_fwd: B=%d T=%d", B, T);
  CgruSeqParams p;
  p.B = B; p.T = T;
  p.gx1 = gx1; p.gx2 = gx2; p.h0 = h0;
  p.r = r; p.z = z; p.n = n; p.h = h; p.xh = xh; p.cat2 = cat2;
  p.out_scale = out_scale;
  CUtensorMap m1h, m1l, m2h, m2l;
  EVE_TRY(make_map_2d(&m1h, w1h_hi, 2 * kCgNf, 9 * kCgNf, 64, 2 * kCgNf, TC_F16));
  EVE_TRY(make_map_2d(&m1l, w1h_lo, 2 * kCgNf, 9 * kCgNf, 64, 2 * kCgNf, TC_F16));
  EVE_TRY(make_map_2d(&m2h, w2h_hi, kCgNf, 9 * kCgNf, 64, kCgNf, TC_F16));
  EVE_TRY(make_map_2d(&m2l, w2h_lo, kCgNf, 9 * kCgNf, 64, kCgNf, TC_F16));
  EVE_TRY(ensure_dynamic_smem((const void*)cgru_seq_fwd_kernel, kCgSmem));
  cgru_seq_fwd_kernel<<<B, kThreads, kCgSmem, s>>>(m1h, m1l, m2h, m2l, p);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

// wd2_*: bf16 hi/lo planes [64][9 * 64] of the data-gradient layout of gate_2.weight[:, 0:nf],
// wd1_*: [64][9 * 128] of gates_1.weight[:, nf:2nf] (conv_tc_prep_weights(dgrad = true))
int cgru_seq_bwd(int B, int T, const void* wd2_hi, const void* wd2_lo, const void* wd1_hi,
                 const void* wd1_lo, const float* dout, float* dcarry, const float* r, const float* z,
                 const float* n, const float* h, const float* h0, float* dg1, float* dg2,
                 cudaStream_t s) {
  EVE_REQUIRE(B > 0 && T > 0, EVE_ERR_SHAPE, "cgru_seq_bwd: B=%d T=%d", B, T);
  CgruSeqBwdParams p;
  p.B = B; p.T = T;
  p.dout = dout; p.dcarry = dcarry;
  p.r = r; p.z = z; p.n = n; p.h = h; p.h0 = h0;
  p.dg1 = dg1; p.dg2 = dg2;
  CUtensorMap m2h, m2l, m1h, m1l;
  EVE_TRY(make_map_2d(&m2h, wd2_hi, kCgNf, 9 * kCgNf, 64, kCgNf, TC_BF16));
  EVE_TRY(make_map_2d(&m2l, wd2_lo, kCgNf, 9 * kCgNf, 64, kCgNf, TC_BF16));
  EVE_TRY(make_map_2d(&m1h, wd1_hi, kCgNf, 9 * 2 * kCgNf, 64, kCgNf, TC_BF16));
  EVE_TRY(make_map_2d(&m1l, wd1_lo, kCgNf, 9 * 2 * kCgNf, 64, kCgNf, TC_BF16));
  EVE_TRY(ensure_dynamic_smem((const void*)cgru_seq_bwd_kernel, kCgBSmem));
  cgru_seq_bwd_kernel<<<B, kThreads, kCgBSmem, s>>>(m2h, m2l, m1h, m1l, p);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

// y[N,OH,OW,Cout] = conv(x) (+bias) (+addend).  x_hi/x_lo: 16-bit NHWC planes of the input;
// w_hi/w_lo: K-major weights [Cout][KH*KW*Cin].  npass: 3 (split operands) or 1 (plain bf16).
int conv_tc_run(const ConvGeom& g, const void* x_hi, const void* x_lo, const void* w_hi,
                const void* w_lo, const float* bias, const float* addend, float* y, int npass,
                int fmt, float out_scale, cudaStream_t s) {
  EVE_REQUIRE(conv_tc_supported(g), EVE_ERR_SHAPE, "conv_tc: unsupported geometry");
  EVE_REQUIRE(npass == 1 || npass == 3, EVE_ERR_CONFIG, "conv_tc: npass must be 1 or 3");
  if (conv_tc_row_supported(g))
    return conv_tc_row_run(g, x_hi, x_lo, w_hi, w_lo, bias, addend, y, npass, fmt, out_scale, s);
  if (npass == 3 && conv_tc_strip_supported(g))
    return conv_tc_strip_run(g, x_hi, x_lo, w_hi, w_lo, bias, addend, y, fmt, out_scale, s);
  TcParams p;
  p.N = g.N; p.OH = g.OH; p.OW = g.OW; p.Cin = g.Cin; p.Cout = g.Cout; p.stride = g.stride;
  p.ntaps = g.KH * g.KW;
  for (int r = 0; r < g.KH; ++r)
    for (int q = 0; q < g.KW; ++q) {
      int t = r * g.KW + q;
      p.tap_dh[t] = r - g.pad;
      p.tap_dw[t] = q - g.pad;
      p.tap_koff[t] = t * g.Cin;
    }
  p.out_mul = 1; p.out_ah = 0; p.out_aw = 0; p.out_H = g.OH; p.out_W = g.OW;
  p.bias = bias; p.addend = addend; p.out = y; p.out_scale = out_scale;
  return tc_launch(p, x_hi, x_lo, g.N, g.H, g.W, w_hi, w_lo, g.Cout, g.KH * g.KW * g.Cin, npass,
                   fmt, s);
}

// The 7x7 stride-2 stem (3 input channels, eye_net.py:48 via torchvision's ResNet-18 conv1) without
// a materialised im2col matrix: a filter ROW of one output pixel is 7 pixels x 3 channels = 8 pixels
// x 4 channels (zero padded) = 32 consecutive 16-bit values, so the convolution is a 7-tap "7x1"
// convolution over 32-channel windows: tap r reads window row 2 oh + r of the zero-padded planes
// (top pad 3: no negative coordinates), K = 7 x 32.  x_hi / x_lo: [N][H + 6][...] planes whose
// window of output column ow starts at ow * win_bytes (see make_map_windows); w: [64][7 * 32].
int conv_tc_stem_run(int N, int H, int W, const void* x_hi, const void* x_lo, int win_bytes,
                     int pitch_bytes, const void* w_hi, const void* w_lo, const float* bias, float* y,
                     int npass, int fmt, float out_scale, cudaStream_t s) {
  EVE_REQUIRE(H % 2 == 0 && W % 2 == 0 && W / 2 <= kTileM && N >= 1, EVE_ERR_SHAPE,
              "conv_tc_stem: unsupported geometry");
  TcParams p;
  p.N = N; p.OH = H / 2; p.OW = W / 2; p.Cin = 32; p.Cout = 64; p.stride = 2;
  p.ntaps = 7;
  for (int r = 0; r < 7; ++r) {
    p.tap_dh[r] = r;
    p.tap_dw[r] = 0;
    p.tap_koff[r] = r * 32;
  }
  p.out_mul = 1; p.out_ah = 0; p.out_aw = 0; p.out_H = p.OH; p.out_W = p.OW;
  p.bias = bias; p.addend = nullptr; p.out = y; p.out_scale = out_scale;
  return tc_launch(p, x_hi, x_lo, N, H + 6, W / 2, w_hi, w_lo, 64, 7 * 32, npass, fmt, s, win_bytes,
                   pitch_bytes);
}

// Data gradient of a stride-2 convolution (3x3 pad 1 or 1x1 pad 0) as four stride-1
// tensor-core passes, one per output-pixel parity class (a, b): dx[2i+a, 2j+b] only receives
// the taps r with (a + pad - r) even, read from dy[i + (a+pad-r)/2, ...].  `w_hi/w_lo` hold
// the dgrad weight layout of conv_tc_prep_weights(dgrad = true): row ci, column
// ((KH-1-r)*KW + (KW-1-q))*Cout + co.  dx must be zero-filled by the caller when KH == 1
// (odd pixels receive nothing); addend (if any) has dx's layout.
bool conv_tc_dgrad_s2_supported(const ConvGeom& g) {
  if (g.stride != 2 || g.KH != g.KW || (g.KH != 1 && g.KH != 3) || g.pad != g.KH / 2) return false;
  if (g.H % 2 != 0 || g.W % 2 != 0 || g.OH != g.H / 2 || g.OW != g.W / 2) return false;
  if (chunk_for(g.Cout) == 0 || bn_for(g.Cin) == 0) return false;
  if (g.OW > kTileM || g.N < 1) return false;
  return true;
}

int conv_tc_dgrad_s2_run(const ConvGeom& g, const void* d_hi, const void* d_lo, const void* w_hi,
                         const void* w_lo, const float* addend, float* dx, int npass,
                         cudaStream_t s) {
  EVE_REQUIRE(conv_tc_dgrad_s2_supported(g), EVE_ERR_SHAPE, "conv_tc_dgrad_s2: unsupported geometry");
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      TcParams p;
      p.N = g.N; p.OH = g.OH; p.OW = g.OW;       // class grid == dy grid
      p.Cin = g.Cout; p.Cout = g.Cin; p.stride = 1;
      p.ntaps = 0;
      for (int r = 0; r < g.KH; ++r) {
        if ((a + g.pad - r) % 2 != 0) continue;
        for (int q = 0; q < g.KW; ++q) {
          if ((b + g.pad - q) % 2 != 0) continue;
          int t = p.ntaps++;
          p.tap_dh[t] = (a + g.pad - r) / 2;
          p.tap_dw[t] = (b + g.pad - q) / 2;
          p.tap_koff[t] = ((g.KH - 1 - r) * g.KW + (g.KW - 1 - q)) * g.Cout;
        }
      }
      if (p.ntaps == 0) continue;                // 1x1: only the even/even class gets data
      p.out_mul = 2; p.out_ah = a; p.out_aw = b; p.out_H = g.H; p.out_W = g.W;
      p.bias = nullptr; p.addend = addend; p.out = dx; p.out_scale = 1.f;
      EVE_TRY(tc_launch(p, d_hi, d_lo, g.N, g.OH, g.OW, w_hi, w_lo, g.Cin, g.KH * g.KW * g.Cout,
                        npass, TC_BF16, s));
    }
  return EVE_OK;
}

// pixel box for the wgrad K loop: rows = bw*bh*bn <= 64 and a multiple of 16
static bool pick_wgrad_box(int N, int H, int W, int& bw, int& bh, int& bn) {
  double best = -1.0;
  bw = W <= kWgRows ? W : kWgRows;
  if (W % bw != 0) return false;
  bh = bn = 0;
  for (int h = 1; h <= H && bw * h <= kWgRows; ++h) {
    if (bw < W && h > 1) break;                 // partial rows: one image row per box
    int nmax = kWgRows / (bw * h);
    for (int n = 1; n <= nmax && n <= ((h == H && bw == W) ? N : 1); ++n) {
      int rows = bw * h * n;
      if (rows % 16 != 0) continue;
      double eff = ((double)H / (double)(cdiv(H, h) * h)) * ((double)N / (double)(cdiv(N, n) * n)) *
                   (0.5 + 0.5 * rows / (double)kWgRows);   // prefer full boxes
      if (eff > best + 1e-9) {
        best = eff;
        bh = h;
        bn = n;
      }
    }
  }
  return bh > 0;
}

bool conv_tc_wgrad_supported(const ConvGeom& g) {
  if (g.KH != g.KW || (g.KH != 1 && g.KH != 3) || g.pad != g.KH / 2) return false;
  if (g.stride == 1) {
    if (g.OH != g.H || g.OW != g.W) return false;
  } else if (g.stride == 2) {
    if (g.H % 2 != 0 || g.W % 2 != 0) return false;
  } else {
    return false;
  }
  if (chunk_for(g.Cin) == 0 || chunk_for(g.Cout) == 0) return false;
  if (g.N < 1) return false;
  int bw, bh, bn;
  return pick_wgrad_box(g.N, g.OH, g.OW, bw, bh, bn);
}

static int wgrad_stage_bytes(const TcWgradParams& p, int npass) {
  const int planes = npass == 3 ? 2 : 1;
  const int plane = kTileM * kWgRows * 2 + (p.nblk / p.xw) * kWgRows * p.xw * 2;
  return ((plane * planes) + 1023) & ~1023;      // [M hi][M lo][N hi blocks][N lo blocks]
}

static void wgrad_plan(const ConvGeom& g, TcWgradParams& p, int& mblocks, int& nblocks,
                       int& splits, int npass, int waves = 0) {
  p.N = g.N; p.H = g.OH; p.W = g.OW; p.Cin = g.Cin; p.Cout = g.Cout;   // tiles walk the dy grid
  p.KH = g.KH; p.KW = g.KW; p.pad = g.pad;
  p.swap = g.stride == 2 ? 1 : 0;
  p.mstride = g.stride;
  pick_wgrad_box(g.N, g.OH, g.OW, p.bw, p.bh, p.bn);
  p.rows = p.bw * p.bh * p.bn;
  p.tiles_w = g.OW / p.bw;
  p.tiles_h = cdiv(g.OH, p.bh);
  p.tiles_total = p.tiles_w * p.tiles_h * cdiv(g.N, p.bn);
  const int Cm = p.swap ? g.Cin : g.Cout;      // channels on the M side / N side
  const int Cn = p.swap ? g.Cout : g.Cin;
  p.uw = chunk_for(Cm);
  p.upb = kTileM / p.uw;
  p.cout_blocks = Cm / p.uw;
  p.units = g.KH * g.KW * p.cout_blocks;
  p.xw = chunk_for(Cn);
  p.nblk = Cn % 256 == 0 ? 256 : (Cn % 128 == 0 ? 128 : (Cn % 64 == 0 ? 64 : Cn));
  mblocks = cdiv(p.units, p.upb);
  nblocks = Cn / p.nblk;
  p.stages = (220 * 1024) / wgrad_stage_bytes(p, npass);
  if (p.stages > 6) p.stages = 6;
  // one CTA per SM (the ring takes the whole shared memory): the grid must not spill one CTA into
  // an extra wave, so the split count is rounded DOWN to fill `waves` full waves
  // Fewer splits are faster (less partial traffic), more splits keep each fp32 accumulation chain
  // in TMEM short (the tensor core's accumulate rounds toward zero: the error of a chain grows
  // linearly with its length).  Take the smallest number of full waves whose chains stay below
  // kWgradChainPixels, at most "tc_wgrad_waves".
  const bool sizing = waves > 0;
  if (!sizing) waves = get_option(OPT_TC_WGRAD_WAVES);
  int want = 1;
  for (int wv = sizing ? waves : 1; wv <= waves; ++wv) {
    want = (wv * kNumSMs) / (mblocks * nblocks);
    if (want < 1) want = 1;
    if ((long long)p.tiles_total * p.rows <= (long long)want * kWgradChainPixels) break;
  }
  int max_splits = cdiv(p.tiles_total, 4);          // at least 4 pixel tiles per CTA
  splits = want < max_splits ? want : max_splits;
  if (splits < 1) splits = 1;
  p.tiles_per_split = cdiv(p.tiles_total, splits);
  splits = cdiv(p.tiles_total, p.tiles_per_split);
}

// ---- halo-row weight gradient: W == 128, 3x3 stride 1, Cout in {16, 32} (three pixel-shifted M
// blocks of Cout channels must fit the 128-row MMA), Cin in {16, 32, 64}
bool conv_tc_wgrad_row_supported(const ConvGeom& g) {
  if (!get_option(OPT_TC_ROW_WGRAD) || !row_geometry_ok(g)) return false;
  return g.Cout == 16 || g.Cout == 32;
}

template <int NPASS, int CIN, int COUT, int KS>
static int launch_wgrad_row(const CUtensorMap& d_hi, const CUtensorMap& d_lo, const CUtensorMap& x_hi,
                            const CUtensorMap& x_lo, TcWgRowParams p, int grid, cudaStream_t s) {
  using Cfg = WgRowCfg<NPASS, CIN, COUT, KS>;
  static_assert(Cfg::kSlots >= 4, "halo-row wgrad: ring too shallow");
  EVE_TRY(ensure_dynamic_smem((const void*)conv_tc_wgrad_row_kernel<NPASS, CIN, COUT, KS>,
                              227 * 1024));
  p.slots = Cfg::kSlots;
  const int smem_bytes = Cfg::kSlots * Cfg::kSlotBytes + 1024 + kBarrierBytes;
  conv_tc_wgrad_row_kernel<NPASS, CIN, COUT, KS><<<grid, kThreads, smem_bytes, s>>>(d_hi, d_lo, x_hi, x_lo, p);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

template <int NPASS, int CIN, int COUT>
static int launch_wgrad_row3(const CUtensorMap& d_hi, const CUtensorMap& d_lo, const CUtensorMap& x_hi,
                             const CUtensorMap& x_lo, TcWgRowParams p, int grid, cudaStream_t s) {
  using Cfg = WgRow3Cfg<NPASS, CIN, COUT>;
  static_assert(Cfg::kSlots >= 4, "halo-row wgrad: ring too shallow");
  EVE_TRY(ensure_dynamic_smem((const void*)conv_tc_wgrad_row3_kernel<NPASS, CIN, COUT>, 227 * 1024));
  p.slots = Cfg::kSlots;
  const int smem_bytes = Cfg::kSlots * Cfg::kSlotBytes + 1024 + kBarrierBytes;
  conv_tc_wgrad_row3_kernel<NPASS, CIN, COUT><<<grid, kThreads, smem_bytes, s>>>(d_hi, d_lo, x_hi, x_lo, p);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

template <int CIN, int COUT>
static int launch_wgrad_row_np(int ks, int npass, const CUtensorMap& d_hi, const CUtensorMap& d_lo,
                               const CUtensorMap& x_hi, const CUtensorMap& x_lo,
                               const TcWgRowParams& p, int grid, cudaStream_t s) {
  // 3x3: filter rows stacked along N (tc_row_wgrad = 2, default); 1 = one instruction group per row
  if (ks == 3 && get_option(OPT_TC_ROW_WGRAD) == 2)
    return npass == 3 ? launch_wgrad_row3<3, CIN, COUT>(d_hi, d_lo, x_hi, x_lo, p, grid, s)
                      : launch_wgrad_row3<1, CIN, COUT>(d_hi, d_lo, x_hi, x_lo, p, grid, s);
  if (ks == 3)
    return npass == 3 ? launch_wgrad_row<3, CIN, COUT, 3>(d_hi, d_lo, x_hi, x_lo, p, grid, s)
                      : launch_wgrad_row<1, CIN, COUT, 3>(d_hi, d_lo, x_hi, x_lo, p, grid, s);
  return npass == 3 ? launch_wgrad_row<3, CIN, COUT, 1>(d_hi, d_lo, x_hi, x_lo, p, grid, s)
                    : launch_wgrad_row<1, CIN, COUT, 1>(d_hi, d_lo, x_hi, x_lo, p, grid, s);
}

static int conv_tc_wgrad_row_run(const ConvGeom& g, const void* d_hi, const void* d_lo,
                                 const void* x_hi, const void* x_lo, float* part, int npass,
                                 int* splits_out, cudaStream_t s) {
  TcWgRowParams p;
  p.N = g.N; p.H = g.H;
  pick_strips(g.N, g.H, p.strips, p.rows_per_strip);
  p.items = g.N * p.strips;
  p.slots = 0;
  p.part = part;
  const int grid = p.items < kNumSMs ? p.items : kNumSMs;
  CUtensorMap md_hi, md_lo, mx_hi, mx_lo;
  EVE_TRY(make_map_nhwc(&md_hi, d_hi, g.N, g.H, g.W, g.Cout, g.Cout, kRowBox, 1, 1));
  EVE_TRY(make_map_nhwc(&mx_hi, x_hi, g.N, g.H, g.W, g.Cin, g.Cin, kTileM, 1, 1));
  if (npass == 3) {
    EVE_TRY(make_map_nhwc(&md_lo, d_lo, g.N, g.H, g.W, g.Cout, g.Cout, kRowBox, 1, 1));
    EVE_TRY(make_map_nhwc(&mx_lo, x_lo, g.N, g.H, g.W, g.Cin, g.Cin, kTileM, 1, 1));
  } else {
    md_lo = md_hi;
    mx_lo = mx_hi;
  }
  *splits_out = grid;
  const int key = g.Cin * 100 + g.Cout;
  switch (key) {
    case 1616: return launch_wgrad_row_np<16, 16>(g.KH, npass, md_hi, md_lo, mx_hi, mx_lo, p, grid, s);
    case 1632: return launch_wgrad_row_np<16, 32>(g.KH, npass, md_hi, md_lo, mx_hi, mx_lo, p, grid, s);
    case 3216: return launch_wgrad_row_np<32, 16>(g.KH, npass, md_hi, md_lo, mx_hi, mx_lo, p, grid, s);
    case 3232: return launch_wgrad_row_np<32, 32>(g.KH, npass, md_hi, md_lo, mx_hi, mx_lo, p, grid, s);
    case 6416: return launch_wgrad_row_np<64, 16>(g.KH, npass, md_hi, md_lo, mx_hi, mx_lo, p, grid, s);
    case 6432: return launch_wgrad_row_np<64, 32>(g.KH, npass, md_hi, md_lo, mx_hi, mx_lo, p, grid, s);
  }
  EVE_REQUIRE(false, EVE_ERR_SHAPE, "conv_tc_wgrad_row: %d -> %d channels", g.Cin, g.Cout);
  return EVE_ERR_SHAPE;
}

// ---- padded-strip weight gradient: 3x3 stride 1; (Cout, Cin) in {64} x {32, 64, 128},
// {128} x {64, 128}, {32} x {32, 128}
static bool wgrad_strip_geometry_ok(const ConvGeom& g) {
  if (g.KH != 3 || g.KW != 3 || g.stride != 1 || g.pad != 1 || g.OH != g.H || g.OW != g.W) return false;
  if (g.N < 1 || g.W + 2 > 256 || g.W >= kTileM) return false;      // (128-pixel rows: the halo-row kernels)
  if (g.Cout == 64) return g.Cin == 32 || g.Cin == 64 || g.Cin == 128;
  if (g.Cout == 128) return g.Cin == 64 || g.Cin == 128;
  if (g.Cout == 32) return g.Cin == 32 || g.Cin == 128;
  return false;
}

struct WgStripPlan {
  int R, strips, ksteps, dy_plane, x_chunk, smem;
};
static bool wgrad_strip_plan(const ConvGeom& g, WgStripPlan& pl) {
  const int Wp = g.W + 2;
  const int subw = g.Cout == 32 ? 32 : 64;
  const int xw = g.Cin > 64 ? 64 : g.Cin, xch = g.Cin / xw;
  auto fit = [&](int R, WgStripPlan& q) {
    q.R = R;
    q.ksteps = cdiv((long long)R * Wp, 16);
    const int kpad = q.ksteps * 16;
    q.dy_plane = (int)align_up((size_t)(kpad + 2 * Wp + 8) * subw * 2, 1024);
    // + 1 pixel: the block that holds tap (2, 2) reads the x strip one pixel in
    q.x_chunk = (int)align_up((size_t)(kpad + 1) * xw * 2, 1024);
    q.smem = 2 * 2 * (q.dy_plane + xch * q.x_chunk) + 1024 + 256;
    return q.smem <= 227 * 1024 && R + 2 <= 256;
  };
  int rmax = 0;
  WgStripPlan q;
  for (int R = 1; R <= g.H; ++R)
    if (fit(R, q)) rmax = R; else break;
  if (rmax == 0) return false;
  pl.strips = cdiv(g.H, rmax);
  fit(cdiv(g.H, pl.strips), pl);
  pl.strips = cdiv(g.H, pl.R);
  return true;
}

bool conv_tc_wgrad_strip_supported(const ConvGeom& g) {
  WgStripPlan pl;
  return get_option(OPT_TC_WGRAD_STRIP) != 0 && wgrad_strip_geometry_ok(g) && wgrad_strip_plan(g, pl);
}

// M blocks of one output-channel block: dy-strip offset of the first tap, pixel distance between the
// taps of the block, x-strip offset, (r * 3 + q) per sub-block.  Tap (r, q) pairs x pixel p with dy
// pixel p + (2 - r) Wp + (1 - q).
static void wgrad_strip_blocks(int Wp, int subw, int boff[5], int blbo[5], int xoff[5], int tap[5][4],
                               int& nblk) {
  for (int b = 0; b < 5; ++b) {
    xoff[b] = 0;
    for (int j = 0; j < 4; ++j) tap[b][j] = -1;
  }
  if (subw == 64) {       // tap pairs in order of increasing strip offset
    nblk = 5;
    const int o[5] = {0, 1, Wp, 2 * Wp - 1, 2 * Wp + 1};
    const int l[5] = {1, Wp - 2, 1, 1, 1};
    const int t[5][2] = {{8, 7}, {6, 5}, {4, 3}, {2, 1}, {0, -1}};
    for (int b = 0; b < 5; ++b) {
      boff[b] = o[b]; blbo[b] = l[b];
      tap[b][0] = t[b][0]; tap[b][1] = t[b][1];
    }
    xoff[0] = 1;          // tap (2, 2): offset -1, taken on the x side
  } else {                // one filter row per block: q = 2, 1, 0 one pixel apart (+ a junk sub-block)
    nblk = 3;
    for (int b = 0; b < 3; ++b) {
      const int r = 2 - b;
      boff[b] = (2 - r) * Wp - 1 + (b == 0 ? 1 : 0);
      blbo[b] = 1;
      tap[b][0] = r * 3 + 2; tap[b][1] = r * 3 + 1; tap[b][2] = r * 3 + 0;
    }
    xoff[0] = 1;
    boff[3] = boff[4] = 0; blbo[3] = blbo[4] = 1;
  }
}

template <int CIN, int SUBW>
static int launch_wgrad_strip(const CUtensorMap& d_hi, const CUtensorMap& d_lo, const CUtensorMap& x_hi,
                              const CUtensorMap& x_lo, const TcWgStripParams& p, int grid, int smem,
                              cudaStream_t s) {
  EVE_TRY(ensure_dynamic_smem((const void*)conv_tc_wgrad_strip_kernel<CIN, SUBW>, 227 * 1024));
  conv_tc_wgrad_strip_kernel<CIN, SUBW><<<grid, kThreads, smem, s>>>(d_hi, d_lo, x_hi, x_lo, p);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

static int conv_tc_wgrad_strip_run(const ConvGeom& g, const void* d_hi, const void* d_lo,
                                   const void* x_hi, const void* x_lo, float* part, int* splits_out,
                                   cudaStream_t s) {
  WgStripPlan pl;
  EVE_REQUIRE(wgrad_strip_geometry_ok(g) && wgrad_strip_plan(g, pl), EVE_ERR_SHAPE,
              "conv_tc_wgrad_strip: unsupported geometry");
  const int subw = g.Cout == 32 ? 32 : 64;
  const int xw = g.Cin > 64 ? 64 : g.Cin;
  TcWgStripParams p;
  p.N = g.N; p.H = g.H; p.W = g.W; p.Wp = g.W + 2; p.Cin = g.Cin; p.Cout = g.Cout;
  p.R = pl.R; p.strips = pl.strips;
  p.items = g.N * pl.strips;
  p.ksteps = pl.ksteps;
  p.dy_plane = pl.dy_plane; p.x_chunk = pl.x_chunk;
  // every fp32 accumulation chain in TMEM stays below kWgradChainPixels pixels (round-toward-zero
  // accumulation: the error of a chain grows with its length)
  p.flush_items = std::max(1, kWgradChainPixels / (pl.R * g.W));
  p.part = part;
  // jobs: per output-channel block, the M blocks in groups that fit TMEM (512 columns)
  int boff[5], blbo[5], xoff[5], tap[5][4], nblk;
  wgrad_strip_blocks(p.Wp, subw, boff, blbo, xoff, tap, nblk);
  const int acc_cols = g.Cin <= 32 ? 2 * g.Cin : g.Cin;
  const int per_job = std::min(nblk, 512 / acc_cols);
  const int groups = cdiv(nblk, per_job);
  p.njobs = 0;
  p.period = 0;
  for (int cb = 0; cb < g.Cout / subw; ++cb) {
    for (int gr = 0; gr < groups; ++gr) {
      // split as evenly as the block count allows, the larger groups last
      const int lo = gr * nblk / groups, hi = (gr + 1) * nblk / groups;
      WgStripJob& j = p.job[p.njobs];
      j.co0 = cb * subw;
      j.nblk = hi - lo;
      for (int b = 0; b < 5; ++b) {
        const int src = lo + b < hi ? lo + b : lo;
        j.boff[b] = boff[src]; j.blbo[b] = blbo[src]; j.xoff[b] = xoff[src];
        for (int q = 0; q < 4; ++q) j.tap[b][q] = lo + b < hi ? tap[src][q] : -1;
      }
      // CTAs in proportion to the M blocks of the job
      for (int k = 0; k < j.nblk && p.period < 12; ++k) p.slotjob[p.period++] = p.njobs;
      ++p.njobs;
    }
  }
  EVE_REQUIRE(p.njobs >= 1 && p.njobs <= kWgMaxJobs && p.period >= 1, EVE_ERR_SHAPE,
              "conv_tc_wgrad_strip: %d jobs", p.njobs);
  const long long want = (long long)p.items * p.njobs;
  const int grid = want < kNumSMs ? (int)want : kNumSMs;
  if (grid < p.period) {       // few items: one CTA per job and item, dealt plainly round-robin
    p.period = p.njobs;
    for (int j = 0; j < p.njobs; ++j) p.slotjob[j] = j;
  }
  const size_t full = (size_t)g.Cout * g.K();
  // multi-job launches: the CTAs' partials live behind slot 0, which receives the compacted sum
  if (p.njobs > 1) p.part = part + full;
  CUtensorMap md_hi, md_lo, mx_hi, mx_lo;
  EVE_TRY(make_map_nhwc(&md_hi, d_hi, g.N, g.H, g.W, g.Cout, subw, p.Wp, pl.R + 2, 1));
  EVE_TRY(make_map_nhwc(&md_lo, d_lo, g.N, g.H, g.W, g.Cout, subw, p.Wp, pl.R + 2, 1));
  EVE_TRY(make_map_nhwc(&mx_hi, x_hi, g.N, g.H, g.W, g.Cin, xw, p.Wp, pl.R, 1));
  EVE_TRY(make_map_nhwc(&mx_lo, x_lo, g.N, g.H, g.W, g.Cin, xw, p.Wp, pl.R, 1));
  *splits_out = p.njobs > 1 ? 1 : grid;
  const int key = g.Cin * 100 + subw;
  int rc = EVE_ERR_SHAPE;
  switch (key) {
    case 3264: rc = launch_wgrad_strip<32, 64>(md_hi, md_lo, mx_hi, mx_lo, p, grid, pl.smem, s); break;
    case 6464: rc = launch_wgrad_strip<64, 64>(md_hi, md_lo, mx_hi, mx_lo, p, grid, pl.smem, s); break;
    case 12864: rc = launch_wgrad_strip<128, 64>(md_hi, md_lo, mx_hi, mx_lo, p, grid, pl.smem, s); break;
    case 3232: rc = launch_wgrad_strip<32, 32>(md_hi, md_lo, mx_hi, mx_lo, p, grid, pl.smem, s); break;
    case 12832: rc = launch_wgrad_strip<128, 32>(md_hi, md_lo, mx_hi, mx_lo, p, grid, pl.smem, s); break;
    default:
      EVE_REQUIRE(false, EVE_ERR_SHAPE, "conv_tc_wgrad_strip: %d -> %d channels", g.Cin, g.Cout);
  }
  EVE_TRY(rc);
  if (p.njobs > 1) {
    WgCompactArgs a;
    a.subw = subw; a.Cin = g.Cin;
    for (int cb = 0; cb < 4; ++cb)
      for (int t = 0; t < 9; ++t) a.tapjob[cb][t] = 0;
    for (int j = 0; j < kWgMaxJobs; ++j) a.count[j] = 0;
    for (int j = 0; j < p.njobs; ++j)
      for (int b = 0; b < p.job[j].nblk; ++b)
        for (int q = 0; q < 4; ++q)
          if (p.job[j].tap[b][q] >= 0) a.tapjob[p.job[j].co0 / subw][p.job[j].tap[b][q]] = j;
    // CTAs of a job beyond its item count had nothing to do and wrote nothing
    int rank[kWgMaxJobs] = {0, 0, 0, 0};
    for (int z = 0; z < grid; ++z) {
      const int j = p.slotjob[z % p.period];
      if (rank[j] < p.items && a.count[j] < 160) a.cta[j][a.count[j]++] = (unsigned char)z;
      ++rank[j];
    }
    wgrad_strip_compact_kernel<<<cdiv((long long)full, 256), 256, 0, s>>>(p.part, part, (long long)full, a);
    EVE_LAUNCH_CHECK();
  }
  return EVE_OK;
}

size_t conv_tc_wgrad_partial_floats(const ConvGeom& g) {
  if (!conv_tc_wgrad_supported(g)) return 0;
  TcWgradParams p;
  int mb, nb, sp;
  wgrad_plan(g, p, mb, nb, sp, 3, 8);   // sized for the largest "tc_wgrad_waves" setting
  // the halo-row kernel writes one partial per CTA
  if (row_geometry_ok(g) && (g.Cout == 16 || g.Cout == 32)) sp = std::max(sp, kNumSMs);
  if (wgrad_strip_geometry_ok(g)) sp = std::max(sp, kNumSMs + 1);  // one partial per CTA (+ the compacted one)
  return (size_t)sp * g.Cout * g.K();
}

// part[splits][Cout][KH*KW*Cin] <- per-split partial weight gradients; returns the split count
int conv_tc_wgrad_run(const ConvGeom& g, const void* d_hi, const void* d_lo, const void* x_hi,
                      const void* x_lo, float* part, int npass, int* splits_out,
                      cudaStream_t s, int x_fmt) {
  EVE_REQUIRE(conv_tc_wgrad_supported(g), EVE_ERR_SHAPE, "conv_tc_wgrad: unsupported geometry");
  if (conv_tc_wgrad_row_supported(g) && x_fmt == TC_BF16)
    return conv_tc_wgrad_row_run(g, d_hi, d_lo, x_hi, x_lo, part, npass, splits_out, s);
  if (npass == 3 && x_fmt == TC_BF16 && conv_tc_wgrad_strip_supported(g))
    return conv_tc_wgrad_strip_run(g, d_hi, d_lo, x_hi, x_lo, part, splits_out, s);
  TcWgradParams p;
  int mb, nb, sp;
  wgrad_plan(g, p, mb, nb, sp, npass);
  p.part = part;
  // dy planes are always bf16 (gradient range); x planes are bf16 or the forward's fp16 planes
  p.fmt_m = p.swap ? x_fmt : TC_BF16;
  p.fmt_n = p.swap ? TC_BF16 : x_fmt;
  // md_* = M-side maps, mx_* = N-side maps (see TcWgradParams::swap)
  CUtensorMap md_hi, md_lo, mx_hi, mx_lo;
  if (!p.swap) {
    EVE_TRY(make_map_nhwc(&md_hi, d_hi, g.N, g.OH, g.OW, g.Cout, p.uw, p.bw, p.bh, p.bn));
    EVE_TRY(make_map_nhwc(&mx_hi, x_hi, g.N, g.H, g.W, g.Cin, p.xw, p.bw, p.bh, p.bn, 1, x_fmt));
    if (npass == 3) {
      EVE_TRY(make_map_nhwc(&md_lo, d_lo, g.N, g.OH, g.OW, g.Cout, p.uw, p.bw, p.bh, p.bn));
      EVE_TRY(make_map_nhwc(&mx_lo, x_lo, g.N, g.H, g.W, g.Cin, p.xw, p.bw, p.bh, p.bn, 1, x_fmt));
    }
  } else {
    EVE_TRY(make_map_nhwc(&md_hi, x_hi, g.N, g.H, g.W, g.Cin, p.uw, p.bw, p.bh, p.bn, 2, x_fmt));
    EVE_TRY(make_map_nhwc(&mx_hi, d_hi, g.N, g.OH, g.OW, g.Cout, p.xw, p.bw, p.bh, p.bn));
    if (npass == 3) {
      EVE_TRY(make_map_nhwc(&md_lo, x_lo, g.N, g.H, g.W, g.Cin, p.uw, p.bw, p.bh, p.bn, 2, x_fmt));
      EVE_TRY(make_map_nhwc(&mx_lo, d_lo, g.N, g.OH, g.OW, g.Cout, p.xw, p.bw, p.bh, p.bn));
    }
  }
  if (npass != 3) {
    md_lo = md_hi;
    mx_lo = mx_hi;
  }
  const int smem = p.stages * wgrad_stage_bytes(p, npass) + 1024 + 256;
  if (npass == 3) EVE_TRY(ensure_dynamic_smem((const void*)conv_tc_wgrad_kernel<3>, 227 * 1024));
  else EVE_TRY(ensure_dynamic_smem((const void*)conv_tc_wgrad_kernel<1>, 227 * 1024));
  dim3 grid(mb, nb, sp);
  if (npass == 3)
    conv_tc_wgrad_kernel<3><<<grid, kThreads, smem, s>>>(md_hi, md_lo, mx_hi, mx_lo, p);
  else
    conv_tc_wgrad_kernel<1><<<grid, kThreads, smem, s>>>(md_hi, md_lo, mx_hi, mx_lo, p);
  EVE_LAUNCH_CHECK();
  *splits_out = sp;
  return EVE_OK;
}

}  // namespace eve
