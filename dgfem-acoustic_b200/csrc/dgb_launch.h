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
