// Launch / dynamic-shared-memory macros shared by the kernels that the CPU tests also run through the CUDA emulation of
// oracle/cuda_emu.h (DGB_EMULATE): under nvcc they are the plain CUDA constructs.
#pragma once
#if defined(DGB_EMULATE) && defined(__CUDACC__)
#error "DGB_EMULATE is for the host-compiled test harnesses under oracle/ only: the product (nvcc) build has no CPU path"
#endif
#ifdef DGB_EMULATE
#define DGB_DYNAMIC_SMEM(type, name) type* name = reinterpret_cast<type*>(cuemu::dynamicSmem())
#define DGB_LAUNCH(kernel, grid, block, smemBytes, stream, ...) cuemu::launch(kernel, grid, block, smemBytes, __VA_ARGS__)
#else
#define DGB_DYNAMIC_SMEM(type, name) extern __shared__ type name[]
#define DGB_LAUNCH(kernel, grid, block, smemBytes, stream, ...) kernel<<<grid, block, smemBytes, stream>>>(__VA_ARGS__)
#endif

#ifndef DGB_EMULATE
#include <cuda_runtime.h>

#include <stdexcept>
#include <string>
namespace dgb {
// thrown by the launchers when a kernel cannot run on the current device; the C ABI turns it into DGB_ERR_UNSUPPORTED
struct UnsupportedError : std::runtime_error {
    using std::runtime_error::runtime_error;
};
// One-time configuration of a kernel PER DEVICE (a process may hold handles on several GPUs): opt-in dynamic shared memory
// and the device's SM count. One KernelConfig per kernel instance (a function-local static of its launcher).
constexpr int kMaxDevices = 64;
struct KernelConfig {
    size_t smem[kMaxDevices] = {};
    int numSm[kMaxDevices] = {};
};
template <typename K>
inline int configureKernel(KernelConfig& kc, K kernel, size_t smemBytes, const char* name) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= kMaxDevices) throw UnsupportedError("device ordinal beyond 63");
    if (kc.numSm[dev] == 0 || kc.smem[dev] < smemBytes) {
        cudaDeviceGetAttribute(&kc.numSm[dev], cudaDevAttrMultiProcessorCount, dev);
        const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes);
        if (e != cudaSuccess) {
            cudaGetLastError();
            kc.numSm[dev] = 0;
            throw UnsupportedError(std::string(name) + ": " + std::to_string(smemBytes) + " bytes of shared memory per CTA are not available on this device (" +
                                   cudaGetErrorString(e) + ")");
        }
        kc.smem[dev] = smemBytes;
    }
    return kc.numSm[dev];
}
}  // namespace dgb
#endif
