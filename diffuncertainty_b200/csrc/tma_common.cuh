// mbarrier / bulk-copy (1-D TMA) PTX helpers shared by the TMA-pipelined kernels (k1_tma.cu, k1_uni.cu).
#pragma once
#include <cuda_runtime.h>

namespace vu {

// ---- mbarrier / bulk-copy PTX ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
// same with a suspend-time hint: the hardware parks the thread for up to `ns` nanoseconds (or until the phase completes)
// before the probe returns, so a warp that waits long -- the producer on a full ring -- spins with far fewer instructions
__device__ __forceinline__ void mbar_wait_hint(unsigned bar, unsigned parity, unsigned ns) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAITH_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONEH_%=;\n"
        "bra WAITH_%=;\n"
        "DONEH_%=:\n"
        "}\n" ::"r"(bar), "r"(parity), "r"(ns)
        : "memory");
}
// same, for the statistics warps, which are ahead of the pipeline most of the time: sleep between probes instead
// of burning issue slots the consumer warps need
__device__ __forceinline__ void mbar_wait_relaxed(unsigned bar, unsigned parity) {
    unsigned done = 0;
    while (true) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(100);
    }
}
// global -> shared bulk copy, completion counted in bytes on an mbarrier; read-once data: evict-first in L2
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar, unsigned long long policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar), "l"(policy)
                 : "memory");
}

}  // namespace vu
