// Staging of the immutable per-model blob (grid bitmaps, distance tables, reward table)
// from global into shared memory with ONE TMA bulk copy per CTA
// (cp.async.bulk.shared::cluster.global + mbarrier complete_tx; SASS: UBLKCP).
#pragma once
#include <stdint.h>

namespace posg {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// Called by every thread of the CTA.  `bytes` must be a multiple of 16 and both pointers
// 16-byte aligned.  Returns after the blob is visible to all threads.
__device__ __forceinline__ void stage_blob(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                           uint64_t* mbar) {
  const uint32_t bar = smem_u32(mbar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    // make the barrier initialisation visible to the async (TMA) proxy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(bar)
        : "memory");
  }
  __syncthreads();  // barrier init is visible to the waiters
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar)
        : "memory");
  }
}

// The same in two halves, for kernels that have independent work to do while the blob is in flight:
// stage_blob_issue (every thread calls it; thread 0 starts the copy; the caller must __syncthreads() before anybody
// waits) and stage_blob_wait (returns once the blob is visible to the calling thread).
__device__ __forceinline__ void stage_blob_issue(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* mbar) {
  const uint32_t bar = smem_u32(mbar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(bar)
        : "memory");
  }
}
__device__ __forceinline__ void stage_blob_wait(uint64_t* mbar) {
  const uint32_t bar = smem_u32(mbar);
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar)
        : "memory");
  }
}

}  // namespace posg
