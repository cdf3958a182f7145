// Bulk asynchronous copy (TMA, cp.async.bulk -> UBLKCP) of one frame's channel values into shared memory, completing
// on an mbarrier: the staging the min-sum kernels of the TM codes use to fetch the NEXT frame while the current one is
// decoded (decode_ms_tm.cu, decode_ms_tm_i16.cu, decode_ms_tm_wide.cu).
#pragma once
#include <cstdint>

namespace ldpc {

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra LAB_WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
// one thread: arm the barrier with the byte count and start the copy (dst, src and bytes are multiples of 16)
__device__ __forceinline__ void bulk_load(void *dst_smem, const void *src_gmem, unsigned bytes, uint64_t *bar) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // order earlier generic-proxy reads of dst
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}

}  // namespace ldpc
