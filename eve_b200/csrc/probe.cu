// Hardware probes (measurement aids, not on the product path): issue rate of tcgen05.mma with
// both operands in shared memory, as a function of the N extent and of the operand start alignment.
#include <cuda.h>

#include "common.cuh"

namespace eve {
namespace {

__device__ __forceinline__ uint32_t p_smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// One CTA per SM, one issuing thread: `reps` x (nmma tcgen05.mma M=128, N=n, K=16, bf16) on
// whatever the shared memory holds, then one commit; returns the clock64 span per CTA.
__global__ void __launch_bounds__(128, 1)
mma_rate_kernel(int n, int nmma, int reps, int a_shift_bytes, int distinct_a, int row_bytes,
                int issuers, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  __shared__ uint64_t bars[4];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;     // fp16 ones / harmless bf16
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(p_smem_u32(&bars[i])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     p_smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot + (uint32_t)warp * 128u;     // each issuing warp: its own accumulator columns
  uint64_t& bar = bars[warp & 3];
  // one elected lane of a converged warp issues (the idiom of the product kernels: a plain
  // `threadIdx.x == 0` branch makes ptxas wrap every UTCHMMA in a BRA.U.ANY convergence loop,
  // which costs ~50 cycles per MMA)
  bool leader = false;
  if (warp < issuers) {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    leader = pred != 0;
  }
  if (leader) {
    // K-major descriptors: row_bytes = 128 / 64 / 32 selects SWIZZLE_128B / 64B / 32B (64 / 32 / 16
    // channels per pixel row, 8-row atoms of 8 * row_bytes bytes)
    const uint64_t layout = row_bytes == 128 ? (2ull << 61) : (row_bytes == 64 ? (4ull << 61) : (6ull << 61));
    const uint64_t hi = (1ull << 16) | ((uint64_t)((8 * row_bytes) >> 4) << 32) | (1ull << 46) | layout;
    const uint32_t a0 = p_smem_u32(smem) + (uint32_t)a_shift_bytes;
    const uint32_t b0 = p_smem_u32(smem) + 96 * 1024;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) |
                           ((uint32_t)(128 >> 4) << 24);
    // descriptors = constant part + (address >> 4); every per-MMA offset below is an immediate
    const uint64_t da0 = hi | (uint64_t)((a0 >> 4) & 0x3FFF);
    const uint64_t db0 = hi | (uint64_t)((b0 >> 4) & 0x3FFF);
    const uint64_t astep = distinct_a ? (uint64_t)(2048 >> 4) : 0ull;
    const long long t0 = clock64();
    uint32_t parity = 0;
    for (int r = 0; r < reps; ++r) {
      for (int i = 0; i < nmma; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint64_t kadv = (uint64_t)((j & (row_bytes / 32 - 1)) * 2);     // K step inside a row
          const uint64_t da = da0 + astep * (uint64_t)j + kadv;
          const uint64_t db = db0 + kadv;
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
              "l"(da), "l"(db), "r"(idesc), "r"(1));
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                       p_smem_u32(&bar)) : "memory");
      uint32_t ok = 0;
      while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(p_smem_u32(&bar)), "r"(parity) : "memory");
      }
      parity ^= 1;
    }
    if (warp == 0) out[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
  }
}


// The same measurement for a CTA PAIR (cta_group::2): M = 256 split over two SMs (128 rows of A in
// each CTA's shared memory), B split in halves of n / 2 rows, one elected lane of the leader CTA
// issues, the accumulator occupies columns [0, n) of both CTAs' TMEM.  Answers: how many cycles does
// each SM spend per M = 128 x N x K = 16 worth of work when its B reads are halved?
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
mma_rate_pair_kernel(int n, int nmma, int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(p_smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     p_smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  bool leader = false;
  if (warp == 0 && rank == 0) {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    leader = pred != 0;
  }
  if (leader) {
    const uint64_t hi = (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
    const uint32_t a0 = p_smem_u32(smem);
    const uint32_t b0 = p_smem_u32(smem) + 96 * 1024;
    // M = 256: the M field (bits 24..28) holds M >> 4 = 16
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) |
                           ((uint32_t)(256 >> 4) << 24);
    const uint64_t da0 = hi | (uint64_t)((a0 >> 4) & 0x3FFF);
    const uint64_t db0 = hi | (uint64_t)((b0 >> 4) & 0x3FFF);
    const long long t0 = clock64();
    uint32_t parity = 0;
    bool failed = false;
    for (int r = 0; r < reps; ++r) {
      for (int i = 0; i < nmma; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint64_t kadv = (uint64_t)((j & 3) * 2);
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
              "l"(da0 + kadv), "l"(db0 + kadv), "r"(idesc), "r"(1));
        }
      }
      asm volatile(
          "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
              p_smem_u32(&bar)), "h"((uint16_t)1) : "memory");
      uint32_t ok = 0;
      long long spins = 0;
      while (!ok && spins < (1ll << 22)) {      // bounded: a protocol mistake must not hang the GPU
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(p_smem_u32(&bar)), "r"(parity) : "memory");
        ++spins;
      }
      if (!ok) {
        failed = true;
        break;
      }
      parity ^= 1;
    }
    out[blockIdx.x >> 1] = failed ? -1 : clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
  }
}

}  // namespace
}  // namespace eve

// cycles_out[grid]: clock64 span of each CTA for reps x nmma MMAs (M=128, N=n, K=16)
extern "C" int eve_probe_mma_rate(int n, int nmma, int reps, int a_shift_bytes, int distinct_a,
                                  int grid, long long* cycles_out, eve_stream_t stream) {
  return eve_probe_mma_rate_swizzle(n, nmma, reps, a_shift_bytes, distinct_a, 128, grid, cycles_out, stream);
}

extern "C" int eve_probe_mma_rate_swizzle(int n, int nmma, int reps, int a_shift_bytes,
                                          int distinct_a, int row_bytes, int grid,
                                          long long* cycles_out, eve_stream_t stream) {
  return eve_probe_mma_rate_issuers(n, nmma, reps, a_shift_bytes, distinct_a, row_bytes, 1, grid,
                                    cycles_out, stream);
}

extern "C" int eve_probe_mma_rate_issuers(int n, int nmma, int reps, int a_shift_bytes,
                                          int distinct_a, int row_bytes, int issuers, int grid,
                                          long long* cycles_out, eve_stream_t stream) {
  using namespace eve;
  EVE_REQUIRE(n >= 16 && n <= 256 && n % 16 == 0 && nmma > 0 && reps > 0 && grid > 0 && cycles_out &&
                  (row_bytes == 128 || row_bytes == 64 || row_bytes == 32) && issuers >= 1 && issuers <= 4 &&
                  (issuers == 1 || n <= 128),
              EVE_ERR_SHAPE, "probe_mma_rate: bad arguments");
  EVE_TRY(ensure_dynamic_smem((const void*)mma_rate_kernel, 200 * 1024));
  mma_rate_kernel<<<grid, 128, 200 * 1024, as_stream(stream)>>>(n, nmma, reps, a_shift_bytes,
                                                               distinct_a, row_bytes, issuers, cycles_out);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}

// cycles_out[pairs]: clock64 span of each leader CTA for reps x nmma MMAs (M=256 over a CTA pair, N=n, K=16)
extern "C" int eve_probe_mma_rate_pair(int n, int nmma, int reps, int pairs, long long* cycles_out,
                                       eve_stream_t stream) {
  using namespace eve;
  EVE_REQUIRE(n >= 32 && n <= 256 && n % 32 == 0 && nmma > 0 && nmma % 8 == 0 && reps > 0 && pairs > 0 &&
                  cycles_out, EVE_ERR_SHAPE, "probe_mma_rate_pair: bad arguments");
  EVE_TRY(ensure_dynamic_smem((const void*)mma_rate_pair_kernel, 200 * 1024));
  mma_rate_pair_kernel<<<2 * pairs, 128, 200 * 1024, as_stream(stream)>>>(n, nmma, reps, cycles_out);
  EVE_LAUNCH_CHECK();
  return EVE_OK;
}
