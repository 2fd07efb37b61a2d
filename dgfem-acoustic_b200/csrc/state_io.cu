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
