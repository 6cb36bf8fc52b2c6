// TMA bulk copy (global -> shared) + mbarrier wrappers, inline PTX for sm_100a.
#pragma once
#include <cuda_runtime.h>

namespace b200pa
{

// ---- TMA bulk copy + mbarrier (PTX; SASS: UBLKCP / SYNCS)
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
   asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
   asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
   asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
   asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra WAIT_%=;\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar)
{
   asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                : "memory");
}

// same copy with an L2 eviction-priority hint (q-data is streamed exactly once per apply: evict_first keeps
// the gathered L-vector, the index streams' neighbours and the slot-order scratch resident instead)
__device__ __forceinline__ unsigned long long l2_policy_evict_first()
{
   unsigned long long pol;
   asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
   return pol;
}
// pol == 0: no hint
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar,
                                             unsigned long long pol)
{
   if (pol == 0ull) { tma_bulk_g2s(dst_smem, src_gmem, bytes, bar); return; }
   asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(dst_smem)),
                "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
                : "memory");
}

// ---- per-thread asynchronous copies global -> shared (LDGSTS): no register is tied up while the data flies
__device__ __forceinline__ void cp_async4(void *dst_smem, const void *src_gmem)
{
   asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
// 8 bytes, or 8 zero bytes when `take` is false (src-size 0: nothing is read, the destination is zero-filled)
__device__ __forceinline__ void cp_async8_zfill(void *dst_smem, const void *src_gmem, bool take)
{
   asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(take ? 8u : 0u) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

} // namespace b200pa
