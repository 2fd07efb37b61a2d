// Generic fused stage kernel (any dimension / order): one thread per (element, node).
//
// Replaces, for one RK stage and all four fields at once, the reference's
//   Mesh::updateFlux (Mesh.cpp:569-674) + solver::numStep (solver.cpp:35-52) [precomputeFlux, getElFlux,
//   getElStiffVector, eigen::minus, eigen::linEq] + the eigen::plusTimes axpys and the final combine
//   (solver.cpp:261-285)
// in the collapsed operator form of SURVEY.md §3.3:
//   rhs_q = sum_u Dw^u (sum_x G_xu F_q,x)  -  sum_lf LIFT_lf (Fscale_lf * flux_q,lf)
// The volume term is evaluated in "derivative of the fields" form (T^u_q = Dw^u q, then combined with the
// per-element constants G and v0 — identical by linearity), which needs no staging of the contravariant fluxes.
//
// This is the HBM-oriented CUDA-core path used at low orders and the fallback for every (dim, order); the
// FP64-bound flagship orders use the tiled DMMA kernel in stage_tiled.cu.
#include "dgb_internal.h"
#include "dgb_device.cuh"
#include "dgb_launch.h"  // launches go through DGB_LAUNCH so that the CPU tests can run this file under oracle/cuda_emu.h

namespace dgb {

namespace {

__host__ __device__ constexpr int npOf(int dim, int p) { return dim == 1 ? p + 1 : dim == 2 ? (p + 1) * (p + 2) / 2 : (p + 1) * (p + 2) * (p + 3) / 6; }
__host__ __device__ constexpr int nfpOf(int dim, int p) { return dim == 1 ? 1 : dim == 2 ? p + 1 : (p + 1) * (p + 2) / 2; }
__host__ __device__ constexpr int nfOf(int dim, int p) { return dim == 1 ? p + 1 : dim + 1; }
__host__ __device__ constexpr int elemsPerCta(int np) { return (256 / np) < 1 ? 1 : (256 / np); }

template <int DIM, int P>
__global__ void __launch_bounds__(256) stageGenericKernel(DeviceMesh M, StageArgs A) {
    constexpr int NP = npOf(DIM, P), NFP = nfpOf(DIM, P), NF = nfOf(DIM, P), E = elemsPerCta(NP);
    constexpr int NFL = NF * NFP;
    __shared__ double sQ[4][E * NP];
    __shared__ double sFl[4][E * NFL];

    const int tid = threadIdx.x;
    const int e0 = A.eBegin + blockIdx.x * E;
    const int nE = min(E, A.eEnd - e0);
    const int64_t S = M.stride;
    const Phys ph = makePhys(M);

    // 1. own nodal values, coalesced
    const bool active = tid < nE * NP;
    const int64_t gidx = (int64_t)e0 * NP + tid;
    if (active) {
#pragma unroll
        for (int q = 0; q < 4; ++q) sQ[q][tid] = A.yin[q * S + gidx];
    }
    __syncthreads();

    // 2. face fluxes: one task per (element, local face, face node)
    for (int w = tid; w < nE * NFL; w += blockDim.x) {
        const int el = w / NFL, r = w - el * NFL, lf = r / NFP, m = r - lf * NFP;
        const int e = e0 + el;
        const int flags = M.fflags[e * NF + lf];
        const int bc = flags & FLAG_BC_MASK;
        const double* fg = M.fgeo + ((int64_t)e * NF + lf) * 4;
        const double n[3] = {fg[0], fg[1], fg[2]};
        const double fscale = fg[3];
        const int own = M.faceNodes[lf * NFP + m];
        double qm[4], qp[4] = {0, 0, 0, 0}, fl[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) qm[q] = sQ[q][el * NP + own];
        if (bc == FACE_INTERIOR) {
            const int nb = M.fnbr[e * NF + lf];
            const int nn = M.nbrMaps[(flags >> FLAG_MAP_SHIFT) * NFP + m];
            const int64_t gi = (int64_t)nb * NP + nn;
#pragma unroll
            for (int q = 0; q < 4; ++q) qp[q] = A.yin[q * S + gi];
        }
        faceFlux(bc, (flags & FLAG_TAU_NEG) ? -1.0 : 1.0, n, ph, qm, qp, fl);
#pragma unroll
        for (int q = 0; q < 4; ++q) sFl[q][el * NFL + r] = fscale * fl[q];
    }
    __syncthreads();
    if (!active) return;

    // 3. volume (12 matvecs in derivative form) + lift
    const int el = tid / NP, i = tid - el * NP;
    const int e = e0 + el;
    double T[DIM][4];
#pragma unroll
    for (int u = 0; u < DIM; ++u)
#pragma unroll
        for (int q = 0; q < 4; ++q) T[u][q] = 0.0;
    const double* q0 = &sQ[0][el * NP];
    const double* q1 = &sQ[1][el * NP];
    const double* q2 = &sQ[2][el * NP];
    const double* q3 = &sQ[3][el * NP];
#pragma unroll 4
    for (int j = 0; j < NP; ++j) {
        const double a0 = q0[j], a1 = q1[j], a2 = q2[j], a3 = q3[j];
#pragma unroll
        for (int u = 0; u < DIM; ++u) {
            const double d = __ldg(&M.DwT[(u * NP + j) * NP + i]);
            T[u][0] = fma(d, a0, T[u][0]);
            T[u][1] = fma(d, a1, T[u][1]);
            T[u][2] = fma(d, a2, T[u][2]);
            T[u][3] = fma(d, a3, T[u][3]);
        }
    }
    double G[DIM][DIM];  // G[x][u]
#pragma unroll
    for (int x = 0; x < DIM; ++x)
#pragma unroll
        for (int u = 0; u < DIM; ++u) G[x][u] = M.Ginv[(int64_t)e * DIM * DIM + x * DIM + u];
    double rhs[4] = {0, 0, 0, 0};
#pragma unroll
    for (int u = 0; u < DIM; ++u) {
        double au = 0.0;
#pragma unroll
        for (int x = 0; x < DIM; ++x) au = fma(G[x][u], ph.v0[x], au);
        double div = 0.0;
#pragma unroll
        for (int x = 0; x < DIM; ++x) div = fma(G[x][u], T[u][1 + x], div);
        rhs[0] += au * T[u][0] + ph.rc2 * div;
#pragma unroll
        for (int x = 0; x < 3; ++x) {
            double r = au * T[u][1 + x];
            if (x < DIM) r = fma(G[x < DIM ? x : 0][u] * ph.invRho, T[u][0], r);
            rhs[1 + x] += r;
        }
    }
    const double* f0 = &sFl[0][el * NFL];
    const double* f1 = &sFl[1][el * NFL];
    const double* f2 = &sFl[2][el * NFL];
    const double* f3 = &sFl[3][el * NFL];
#pragma unroll 4
    for (int l = 0; l < NFL; ++l) {
        const double w = __ldg(&M.nLiftT[l * NP + i]);
        rhs[0] = fma(w, f0[l], rhs[0]);
        rhs[1] = fma(w, f1[l], rhs[1]);
        rhs[2] = fma(w, f2[l], rhs[2]);
        rhs[3] = fma(w, f3[l], rhs[3]);
    }

    // 4. fused RK update (coalesced: gidx is contiguous in tid)
#pragma unroll
    for (int q = 0; q < 4; ++q) rkUpdate(A, q * S + gidx, rhs[q], sQ[q][tid]);
}

template <int DIM, int P>
void launchGeneric(const DeviceMesh& M, const StageArgs& A, cudaStream_t s) {
    constexpr int NP = npOf(DIM, P), E = elemsPerCta(NP);
    const int nEl = A.eEnd - A.eBegin;
    if (nEl <= 0) return;
    const int threads = ((E * NP + 31) / 32) * 32;
    DGB_LAUNCH((stageGenericKernel<DIM, P>), (nEl + E - 1) / E, threads, 0, s, M, A);
}

}  // namespace

StageKernel selectGenericKernel(int dim, int order) {
    StageKernel k;
#define DGB_CASE(D, P) \
    if (dim == D && order == P) { k.launch = &launchGeneric<D, P>; k.name = "stage_generic<" #D "," #P ">"; return k; }
    DGB_CASE(1, 1)
    DGB_CASE(2, 1) DGB_CASE(2, 2) DGB_CASE(2, 3) DGB_CASE(2, 4) DGB_CASE(2, 5) DGB_CASE(2, 6)
    DGB_CASE(3, 1) DGB_CASE(3, 2) DGB_CASE(3, 3) DGB_CASE(3, 4) DGB_CASE(3, 5) DGB_CASE(3, 6)
#undef DGB_CASE
    return k;
}

// ---------------------------------------------------------------------------------------------
// small helper kernels
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void setNodesKernel(double* field, const int32_t* idx, int n, double value) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) field[idx[i]] = value;
}
__global__ void gatherProbesKernel(const double* u, int64_t stride, const int32_t* idx, int n, double* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 4 * n) {
        const int j = i >> 2, q = i & 3;
        const int node = idx[j];
        out[i] = node >= 0 ? u[q * stride + node] : 0.0;
    }
}
// receiver j, field q: sum_n w[j][n] * u[q][el[j]*Np + n]  (Lagrange interpolation inside one element); el < 0 -> 0
__global__ void gatherReceiversKernel(const double* u, int64_t stride, int Np, const int32_t* el, const double* w, int n, double* out, int interleaved) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 4 * n) {
        const int j = i >> 2, q = i & 3;
        const int e = el[j];
        double s = 0.0;
        if (e >= 0) {
            const double* uq = interleaved ? u + (int64_t)e * Np * 4 + q : u + q * stride + (int64_t)e * Np;
            const int cs = interleaved ? 4 : 1;
            const double* wj = w + (int64_t)j * Np;
            for (int nd = 0; nd < Np; ++nd) s += wj[nd] * uq[nd * cs];
        }
        out[i] = s;
    }
}
__global__ void packElementsKernel(const double* y, int64_t stride, int Np, const int32_t* elems, int n, double* buf) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t per = (int64_t)n * Np;
    if (i < 4 * per) {
        const int q = (int)(i / per);
        const int64_t r = i - q * per;
        const int k = (int)(r / Np), nd = (int)(r - (int64_t)k * Np);
        buf[i] = y[q * stride + (int64_t)elems[k] * Np + nd];
    }
}
__global__ void unpackElementsKernel(double* y, int64_t stride, int Np, int firstElem, int n, const double* buf) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t per = (int64_t)n * Np;
    if (i < 4 * per) {
        const int q = (int)(i / per);
        const int64_t r = i - q * per;
        y[q * stride + (int64_t)firstElem * Np + r] = buf[i];
    }
}
}  // namespace

void launchSetNodes(double* field, const int32_t* idx, int n, double value, cudaStream_t s) {
    if (n > 0) DGB_LAUNCH(setNodesKernel, (n + 255) / 256, 256, 0, s, field, idx, n, value);
}
void launchGatherProbes(const double* u, int64_t stride, const int32_t* idx, int n, double* out, cudaStream_t s) {
    if (n > 0) DGB_LAUNCH(gatherProbesKernel, (4 * n + 255) / 256, 256, 0, s, u, stride, idx, n, out);
}
void launchGatherReceivers(const double* u, int64_t stride, int Np, const int32_t* el, const double* w, int n, double* out, cudaStream_t s, int interleaved) {
    if (n > 0) DGB_LAUNCH(gatherReceiversKernel, (4 * n + 127) / 128, 128, 0, s, u, stride, Np, el, w, n, out, interleaved);
}
void launchPackElements(const double* y, int64_t stride, int Np, const int32_t* elems, int n, double* buf, cudaStream_t s) {
    const int64_t tot = 4ll * n * Np;
    if (tot > 0) DGB_LAUNCH(packElementsKernel, (unsigned)((tot + 255) / 256), 256, 0, s, y, stride, Np, elems, n, buf);
}
void launchUnpackElements(double* y, int64_t stride, int Np, int firstElem, int n, const double* buf, cudaStream_t s) {
    const int64_t tot = 4ll * n * Np;
    if (tot > 0) DGB_LAUNCH(unpackElementsKernel, (unsigned)((tot + 255) / 256), 256, 0, s, y, stride, Np, firstElem, n, buf);
}

}  // namespace dgb
