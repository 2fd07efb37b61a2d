// DMMA m8n8k4 dependent-chain latency / throughput vs number of independent accumulator chains per warp and warps per SMSP.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_latency.bin dmma_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NT>
__global__ void chains(double* out, int iters, long long* cyc) {
    double c[NT][2];
    double a = 1.0 + threadIdx.x * 1e-9, b = 0.5;
#pragma unroll
    for (int i = 0; i < NT; ++i) { c[i][0] = i; c[i][1] = i + 0.5; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NT; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < NT; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int NT>
void run(int warps, double* out, long long* dcyc) {
    const int iters = 4096;
    chains<NT><<<148, warps * 32>>>(out, iters, dcyc);
    chains<NT><<<148, warps * 32>>>(out, iters, dcyc);
    cudaDeviceSynchronize();
    long long cyc;
    cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost);
    printf("chains/warp %d  warps/SM %2d (per SMSP %d): %7.2f cycles per MMA per warp, %6.2f cycles per MMA per SMSP\n", NT, warps, warps / 4,
           (double)cyc / (iters * NT), (double)cyc / (iters * NT) / (warps / 4.0));
}

int main() {
    double* out; long long* dcyc;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&dcyc, 8);
    for (int warps : {4, 8, 12, 16}) {
        run<1>(warps, out, dcyc); run<2>(warps, out, dcyc); run<3>(warps, out, dcyc); run<4>(warps, out, dcyc); run<6>(warps, out, dcyc); run<8>(warps, out, dcyc);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
