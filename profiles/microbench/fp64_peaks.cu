// FP64 peak microbenchmarks for B200 (sm_100a): DFMA (CUDA cores) vs DMMA (mma.sync f64 shapes).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peaks fp64_peaks.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters) {
    double a[16];
    double x = 1.0000001 + threadIdx.x * 1e-9, y = 0.999999;
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = i * 0.1 + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], x, y);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NT>
__global__ void dmma884_kernel(double* out, int iters) {
    double c[NT][2];
    double a = 1.0 + threadIdx.x * 1e-9, b = 0.5;
#pragma unroll
    for (int i = 0; i < NT; ++i) { c[i][0] = i; c[i][1] = i + 0.5; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NT; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NT; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NT>
__global__ void dmma16816_kernel(double* out, int iters) {
    double c[NT][4];
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 1.0 + i * 1e-9;
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = 0.5 + i * 1e-9;
#pragma unroll
    for (int i = 0; i < NT; ++i) { c[i][0] = i; c[i][1] = i + 0.5; c[i][2] = i; c[i][3] = 1; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NT; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                           "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NT; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NT>
__global__ void dmma1684_kernel(double* out, int iters) {
    double c[NT][4];
    double a[2] = {1.0 + threadIdx.x * 1e-9, 1.1}, b = 0.5;
#pragma unroll
    for (int i = 0; i < NT; ++i) { c[i][0] = i; c[i][1] = i + 0.5; c[i][2] = i; c[i][3] = 1; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NT; ++i)
            asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a[0]), "d"(a[1]), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NT; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); launch(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    double* out; cudaMalloc(&out, sizeof(double) * nsm * 64 * 1024);
    printf("SMs %d\n", nsm);
    const int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        for (int bps : {1, 2}) {
            int threads = warps * 32, blocks = nsm * bps;
            double ms = time_ms([&] { dfma_kernel<<<blocks, threads>>>(out, iters); }, 5);
            double fl = 2.0 * 16 * iters * (double)threads * blocks;
            printf("DFMA      warps/cta %2d cta/sm %d : %8.3f ms  %7.2f TFLOP/s\n", warps, bps, ms, fl / ms * 1e-9);
        }
    }
    for (int warps : {4, 8, 16}) {
        int threads = warps * 32, blocks = nsm * 2;
        double ms = time_ms([&] { dmma884_kernel<8><<<blocks, threads>>>(out, iters); }, 5);
        double fl = 2.0 * 8 * 8 * 4 * 8 * iters * (double)warps * blocks;
        printf("DMMA m8n8k4   warps/cta %2d cta/sm 2 : %8.3f ms  %7.2f TFLOP/s\n", warps, ms, fl / ms * 1e-9);
        ms = time_ms([&] { dmma1684_kernel<4><<<blocks, threads>>>(out, iters); }, 5);
        fl = 2.0 * 16 * 8 * 4 * 4 * iters * (double)warps * blocks;
        printf("DMMA m16n8k4  warps/cta %2d cta/sm 2 : %8.3f ms  %7.2f TFLOP/s\n", warps, ms, fl / ms * 1e-9);
        ms = time_ms([&] { dmma16816_kernel<4><<<blocks, threads>>>(out, iters); }, 5);
        fl = 2.0 * 16 * 8 * 16 * 4 * iters * (double)warps * blocks;
        printf("DMMA m16n8k16 warps/cta %2d cta/sm 2 : %8.3f ms  %7.2f TFLOP/s\n", warps, ms, fl / ms * 1e-9);
    }
    // sustained DFMA (about 2 s) to see the power-capped clock
    {
        int threads = 512, blocks = nsm * 2;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        for (int r = 0; r < 40; ++r) dfma_kernel<<<blocks, threads>>>(out, iters * 4);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fl = 40.0 * 2.0 * 16 * iters * 4 * (double)threads * blocks;
        printf("DFMA sustained %.0f ms : %7.2f TFLOP/s\n", ms, fl / ms * 1e-9);
        cudaEventRecord(e0);
        for (int r = 0; r < 40; ++r) dmma884_kernel<8><<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        fl = 40.0 * 2.0 * 256 * 8 * iters * (double)(threads / 32) * blocks;
        printf("DMMA m8n8k4 sustained %.0f ms : %7.2f TFLOP/s\n", ms, fl / ms * 1e-9);
    }
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
