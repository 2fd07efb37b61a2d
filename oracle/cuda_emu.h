// Minimal CUDA execution emulation for CPU tests — TEST INFRASTRUCTURE ONLY.
//
// Lets a host compiler build and run a .cu file's kernels as they are written: one OS thread per CUDA thread of a block,
// a barrier for __syncthreads(), blocks one after the other, `__shared__` arrays as function-level statics, dynamic shared
// memory as one buffer per launch. No warps, no memory model, no asynchrony: it checks index logic and arithmetic of a
// kernel, not its CUDA-specific behaviour. The kernel source cooperates through three macros (DGB_EMULATE, DGB_DYNAMIC_SMEM,
// DGB_LAUNCH — see csrc/stage_bb.cu).
#pragma once
#include <cuda_runtime.h>  // host-side declarations; makes __global__ / __device__ / __forceinline__ harmless for g++

#include <pthread.h>
#include <sched.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <thread>
#include <vector>

#undef __shared__
#define __shared__ static  // one copy for all threads; blocks run one after the other
#undef __launch_bounds__
#define __launch_bounds__(...)
#ifndef __restrict__
#define __restrict__
#endif

namespace cuemu {
struct Idx {
    unsigned x = 0, y = 0, z = 0;
};
struct State {
    pthread_barrier_t barrier;
    std::vector<double> smem;
    std::vector<unsigned long long> shfl;  // one 8-byte slot per thread of the block (warp shuffles)
};
inline State*& current() {
    static State* s = nullptr;
    return s;
}
inline void* dynamicSmem() { return current()->smem.data(); }
inline void syncthreads() { pthread_barrier_wait(&current()->barrier); }
}  // namespace cuemu

static thread_local cuemu::Idx threadIdx, blockIdx;
static cuemu::Idx blockDim, gridDim;

#define __syncthreads() cuemu::syncthreads()
// Warp-level primitives for kernels whose block is ONE warp (stage_bb2.cu, stage_bbe.cu): a warp barrier is the block barrier, a shuffle is an
// exchange through one slot per thread between two barriers (every thread of the block must execute it, as the full-mask forms require).
inline void __syncwarp(unsigned = 0xffffffffu) { cuemu::syncthreads(); }
template <typename T>
inline T __shfl_sync(unsigned, T v, int srcLane, int width = 32) {
    static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
    cuemu::State* st = cuemu::current();
    unsigned long long raw = 0;
    memcpy(&raw, &v, sizeof(T));
    st->shfl[threadIdx.x] = raw;
    cuemu::syncthreads();
    const unsigned src = (threadIdx.x / width) * width + (unsigned)(srcLane % width);
    raw = st->shfl[src];
    cuemu::syncthreads();
    T out;
    memcpy(&out, &raw, sizeof(T));
    return out;
}
template <typename T>
inline T __shfl_xor_sync(unsigned mask, T v, int laneMask, int width = 32) { return __shfl_sync(mask, v, (int)((threadIdx.x % width) ^ (unsigned)laneMask), width); }
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
    const unsigned long long both = ((unsigned long long)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) r |= (unsigned)((both >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
}
template <typename T>
inline T __ldcg(const T* p) { return *p; }
template <typename T>
inline void __stcg(T* p, T v) { *p = v; }
inline double __dmul_rn(double a, double b) { return a * b; }
template <typename T>
inline T __ldg(const T* p) { return *p; }
using std::max;
using std::min;

namespace cuemu {
template <typename Kernel, typename... Args>
void launch(Kernel kernel, unsigned grid, unsigned block, size_t smemBytes, Args... args) {
    State st;
    st.smem.assign((smemBytes + sizeof(double) - 1) / sizeof(double) + (smemBytes == 0 ? 1 : 0), 0.0);  // exact size: AddressSanitizer builds see overruns
    st.shfl.assign(block, 0ull);
    pthread_barrier_init(&st.barrier, nullptr, block);
    current() = &st;
    blockDim.x = block;
    gridDim.x = grid;
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < block; ++t)
        pool.emplace_back([=, &st] {
            threadIdx.x = t;
            for (unsigned b = 0; b < grid; ++b) {
                blockIdx.x = b;
                kernel(args...);
                pthread_barrier_wait(&st.barrier);  // the next block reuses the shared memory
            }
        });
    for (auto& th : pool) th.join();
    pthread_barrier_destroy(&st.barrier);
    current() = nullptr;
}
}  // namespace cuemu
