// Do DFMA (CUDA-core FP64) and DMMA (FP64 tensor) share one pipe on B200? One DMMA warp + one DFMA warp per SM
// sub-partition, timed alone and together.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_mix.bin fp64_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void mix(double* out, int itersMma, int itersFma) {
    const int warp = threadIdx.x >> 5;
    double s = 0;
    if (warp < 4) {
        double c[8][2];
        double a = 1.0 + threadIdx.x * 1e-9, b = 0.5;
#pragma unroll
        for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = i + 0.5; }
        for (int it = 0; it < itersMma; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    } else {
        double a[16];
        double x = 1.0000001 + threadIdx.x * 1e-9, y = 0.999999;
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = i * 0.1 + threadIdx.x;
        for (int it = 0; it < itersFma; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fma(a[i], x, y);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) s += a[i];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static float run(double* out, int nsm, int im, int ifm) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    mix<<<nsm, 256>>>(out, im, ifm); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); mix<<<nsm, 256>>>(out, im, ifm); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    double* out; cudaMalloc(&out, sizeof(double) * nsm * 256);
    const int im = 20000, ifm = 80000;  // 160k DMMA (2.56 M pipe cycles at 16/MMA) vs 1.28 M DFMA warp-instructions per warp
    float a = run(out, nsm, im, 0), b = run(out, nsm, 0, ifm), c = run(out, nsm, im, ifm);
    printf("DMMA warp alone  : %8.3f ms  (%.2f cycles/MMA at 1.965 GHz)\n", a, a * 1.965e6 / (8.0 * im));
    printf("DFMA warp alone  : %8.3f ms  (%.2f cycles/DFMA warp-instruction)\n", b, b * 1.965e6 / (16.0 * ifm));
    printf("both on each SMSP: %8.3f ms  (sum %.3f, max %.3f) -> %s\n", c, a + b, a > b ? a : b,
           c > 0.9 * (a + b) ? "SHARED pipe" : (c < 1.1 * (a > b ? a : b) ? "independent pipes" : "partially shared"));
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
