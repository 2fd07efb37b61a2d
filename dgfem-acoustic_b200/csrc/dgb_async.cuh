// Asynchronous-copy primitives of sm_100a shared by the Bernstein kernels (stage_bb2.cu, stage_bbe.cu): mbarriers with
// transaction counts, TMA bulk copies global <-> shared (cp.async.bulk), L2 prefetch, 16-byte cp.async.
// Under DGB_EMULATE (the g++ builds of the test harnesses in oracle/, never the product) the same names are synchronous host
// stand-ins: a bulk copy is a memcpy that completes its bytes on an emulated mbarrier word, a wait spins on the phase bit.
#pragma once
#include <stdint.h>

namespace dgb {
namespace {

#ifndef DGB_EMULATE

__device__ __forceinline__ uint32_t sAddr2(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit2(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sAddr2(bar)), "r"(count));
}
__device__ __forceinline__ void mbarExpectTx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sAddr2(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait2(unsigned long long* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(sAddr2(bar)), "r"(parity)
                     : "memory");
    }
}
// TMA bulk copy global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulkLoad(void* smemDst, const void* gmemSrc, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sAddr2(smemDst)),
                 "l"(__cvta_generic_to_global(gmemSrc)), "r"(bytes), "r"(sAddr2(bar))
                 : "memory");
}
// TMA bulk copy shared -> global (bulk async-group)
__device__ __forceinline__ void bulkStore(void* gmemDst, const void* smemSrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(__cvta_generic_to_global(gmemDst)), "r"(sAddr2(smemSrc)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulkPrefetchL2(const void* gmemSrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(__cvta_generic_to_global(gmemSrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulkCommit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulkWaitRead() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulkWaitAll() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fenceProxyAsync() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// 16-byte asynchronous copy that bypasses L1 (a trace sector is used once); srcBytes == 0 writes zeros
__device__ __forceinline__ void cpAsync16(void* smemDst, const void* gmemSrc, uint32_t srcBytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sAddr2(smemDst)), "l"(__cvta_generic_to_global(gmemSrc)), "r"(srcBytes) : "memory");
}
__device__ __forceinline__ void cpCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cpWaitAll() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cpWaitAllButOne() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

__device__ __forceinline__ void prefetchL2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(__cvta_generic_to_global(p))); }
__device__ __forceinline__ void mbarInitFence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
#else
// emulated mbarrier word: bits 0..31 = outstanding transaction bytes (signed), bit 32 = phase
inline void emuBarAdd(unsigned long long* bar, long long delta) {
    unsigned long long old = __atomic_load_n(bar, __ATOMIC_SEQ_CST), upd;
    do {
        const int32_t pending = (int32_t)(uint32_t)old + (int32_t)delta;
        unsigned long long phase = (old >> 32) & 1ull;
        if (pending == 0) phase ^= 1ull;
        upd = (phase << 32) | (uint32_t)pending;
    } while (!__atomic_compare_exchange_n(bar, &old, upd, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
}
inline void mbarInit2(unsigned long long* bar, int) { __atomic_store_n(bar, 0ull, __ATOMIC_SEQ_CST); }
inline void mbarInitFence() {}
inline void mbarExpectTx(unsigned long long* bar, uint32_t bytes) { emuBarAdd(bar, (long long)bytes); }
inline void mbarWait2(unsigned long long* bar, uint32_t parity) {
    while (((__atomic_load_n(bar, __ATOMIC_SEQ_CST) >> 32) & 1ull) == (unsigned long long)parity) sched_yield();
}
inline void bulkLoad(void* smemDst, const void* gmemSrc, uint32_t bytes, unsigned long long* bar) {
    memcpy(smemDst, gmemSrc, bytes);
    emuBarAdd(bar, -(long long)bytes);
}
inline void bulkStore(void* gmemDst, const void* smemSrc, uint32_t bytes) { memcpy(gmemDst, smemSrc, bytes); }
inline void bulkPrefetchL2(const void*, uint32_t) {}
inline void prefetchL2(const void*) {}
inline void bulkCommit() {}
inline void bulkWaitRead() {}
inline void bulkWaitAll() {}
inline void fenceProxyAsync() {}
inline void cpAsync16(void* smemDst, const void* gmemSrc, uint32_t srcBytes) {
    if (srcBytes) memcpy(smemDst, gmemSrc, 16); else memset(smemDst, 0, 16);
}
inline void cpCommit() {}
inline void cpWaitAll() {}
inline void cpWaitAllButOne() {}
#endif

}  // namespace
}  // namespace dgb
