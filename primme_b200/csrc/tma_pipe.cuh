// tma_pipe.cuh -- minimal mbarrier + bulk-copy (TMA, cp.async.bulk) primitives for sm_100a.
// Raw PTX (no CUTLASS dependency).  A column-major row tile of a tall-skinny matrix is a set of
// short contiguous segments (one per column): each segment is one 1-D bulk copy global -> shared
// that completes on an mbarrier, so the streaming kernels keep several tiles in flight per SM
// without tying up registers or warps (SASS: UBLKCP / SYNCS).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pbtma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
   return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
   asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
   asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
   asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
   asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
   asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
   uint32_t done;
   const uint32_t addr = smem_u32(bar);
   do {
      asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
   } while (!done);
}
// global -> shared bulk copy of `bytes` (multiple of 16; both addresses 16-byte aligned),
// completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
   asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                      smem_u32(dst_smem)),
         "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
         : "memory");
}
// same with an L2 eviction policy for the source lines: a stream that is read exactly once (the CSR arrays of
// the SpMM) is marked evict_first so that it does not push the gathered right-hand-side rows out of the L2
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
   uint64_t pol;
   asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
   return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar,
      uint64_t policy) {
   asm volatile(
         "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
               smem_u32(dst_smem)),
         "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
         : "memory");
}
// global -> shared copy of one box of a 2-D tensor map (cp.async.bulk.tensor, SASS UTMALDG): the box
// [c0, c0+box0) x [c1, c1+box1) lands densely in shared memory, elements outside the tensor are
// zero-filled, and the FULL box byte count completes on `bar`.  dst must be 128-byte aligned.
__device__ __forceinline__ void tensor_g2s_2d(void *dst_smem, const void *tmap, int c0, int c1, uint64_t *bar) {
   asm volatile(
         "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
               smem_u32(dst_smem)),
         "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
         : "memory");
}
// named barrier among the first `nthreads` threads of the CTA (consumer warps only)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
   asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace pbtma
