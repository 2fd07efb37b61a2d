// Host <-> device state transfer of a partitioned handle without a host-side gather (dgb_set_state / dgb_get_state,
// the caller passes the reference's global layout u[eq][el*Np + n], solver.h:15,24).
//
// When the caller's buffer is pinned (page-locked) host memory the GPU reads / writes it directly over PCIe: one thread per
// value walks the rank's elements in LOCAL order and addresses the global array through localToGlobal, so every element is one
// contiguous run of Np doubles on the host side. No staging copy, no host threads: with 8 ranks on one box the host cores and
// the host memory bandwidth are no longer shared by 8 gather loops (profiles/r02: the host gather cost more than the time
// steps). Pageable buffers keep the staged path of dgb_api.cu.
#include "dgb_internal.h"

namespace dgb {
namespace {

// local[q][l*Np + n] = global[q][l2g[l]*Np + n]   (owned + halo elements)
__global__ void gatherStateKernel(const double* __restrict__ global, int64_t Ng, const int32_t* __restrict__ l2g, int K, int Np, double* __restrict__ local,
                                  int64_t stride) {
    const int64_t per = (int64_t)K * Np;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 4 * per; i += (int64_t)gridDim.x * blockDim.x) {
        const int q = (int)(i / per);
        const int64_t r = i - q * per;
        const int l = (int)(r / Np), n = (int)(r - (int64_t)l * Np);
        local[q * stride + r] = global[q * Ng + (int64_t)l2g[l] * Np + n];
    }
}

// global[q][l2g[l]*Np + n] = local[q][l*Np + n]   (owned elements only)
__global__ void scatterStateKernel(double* __restrict__ global, int64_t Ng, const int32_t* __restrict__ l2g, int K, int Np, const double* __restrict__ local,
                                   int64_t stride) {
    const int64_t per = (int64_t)K * Np;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 4 * per; i += (int64_t)gridDim.x * blockDim.x) {
        const int q = (int)(i / per);
        const int64_t r = i - q * per;
        const int l = (int)(r / Np), n = (int)(r - (int64_t)l * Np);
        global[q * Ng + (int64_t)l2g[l] * Np + n] = local[q * stride + r];
    }
}

}  // namespace

void launchGatherState(const double* globalHost, int64_t Ng, const int32_t* l2g, int K, int Np, double* local, int64_t stride, cudaStream_t s) {
    if (K <= 0) return;
    const int64_t tot = 4ll * K * Np;
    gatherStateKernel<<<(unsigned)std::min<int64_t>((tot + 255) / 256, 148 * 16), 256, 0, s>>>(globalHost, Ng, l2g, K, Np, local, stride);
}

void launchScatterState(double* globalHost, int64_t Ng, const int32_t* l2g, int K, int Np, const double* local, int64_t stride, cudaStream_t s) {
    if (K <= 0) return;
    const int64_t tot = 4ll * K * Np;
    scatterStateKernel<<<(unsigned)std::min<int64_t>((tot + 255) / 256, 148 * 16), 256, 0, s>>>(globalHost, Ng, l2g, K, Np, local, stride);
}

}  // namespace dgb

// ---- instrumentation: FP64 issue peak of the current device, measured on the spot (dgb_measure_fp64_tflops) -------------------
// 16 independent DFMA chains per thread, 512 threads per CTA, 2 CTAs per SM: the configuration that reaches the plateau in
// profiles/microbench/r01_fp64_peaks_b200.txt (36.8 TFLOP/s DFMA; the DMMA pipe is the same pipe, 37.1).
namespace dgb {
namespace {
__global__ void __launch_bounds__(512) dfmaPeakKernel(double* out, int iters) {
    double a[16];
    const double x = 1.0000001 + threadIdx.x * 1e-9, y = 0.999999;
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = i * 0.1 + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], x, y);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace

double measureFp64Tflops(cudaStream_t s) {
    int dev = 0, numSm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&numSm, cudaDevAttrMultiProcessorCount, dev);
    double* out = nullptr;
    if (cudaMalloc(&out, sizeof(double)) != cudaSuccess) return 0.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int grid = 2 * numSm, iters = 20000;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {  // the first repetition warms up
        cudaEventRecord(e0, s);
        dfmaPeakKernel<<<grid, 512, 0, s>>>(out, iters);
        cudaEventRecord(e1, s);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double tf = 2.0 * 16 * iters * 512.0 * grid / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return cudaGetLastError() == cudaSuccess ? best : 0.0;
}
}  // namespace dgb
