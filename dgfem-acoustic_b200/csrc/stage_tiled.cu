// Tiled FP64 tensor-core (DMMA) stage kernel for tetrahedra of order 3 and 4 — the FP64-bound flagship orders
// (p >= 4 is compute-bound on B200; measured FP64 peak 37 TFLOP/s for both DFMA and DMMA,
// profiles/microbench/r01_fp64_peaks_b200.txt, so DMMA is chosen for its far lower operand/issue traffic).
//
// Same fused operator as stage_generic.cu (updateFlux + numStep + RK axpys of the reference, Mesh.cpp:476-674,
// solver.cpp:35-52, 261-285), organised as small dense contractions on mma.sync.m8n8k4.f64:
//   T^u   = Dw^u  (NPP x NP)  x  Q  (NP x 8 columns)          columns = (element, field) pairs, 2 elements per n-tile
//   rhs   = combine(T^u, G, v0)  -  LIFT (NPP x NFL) x Fl (NFL x 8 columns)
// Persistent CTAs (one per SM); the operators (54 KB at p = 4) are staged once per CTA into shared memory with one
// TMA bulk copy (cp.async.bulk + mbarrier); every WARP owns a unit of 4 consecutive elements end to end (load ->
// neighbour gather / numerical flux -> DMMA contractions -> fused RK update straight from the accumulator
// fragments), so warps only ever __syncwarp() and the memory phases of some warps overlap the tensor phases of
// the others. All operand tiles are K-contiguous with a leading dimension chosen bank-conflict free (tile_cfg.h).
#include <algorithm>

#include "dgb_device.cuh"
#include "dgb_internal.h"
#include "tile_cfg.h"

namespace dgb {

namespace {

constexpr int kMaxMaps = 64;

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int P, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) stageTiledKernel(DeviceMesh M, StageArgs A, int nUnits) {
    using T = TetTile<P>;
    constexpr int NP = T::NP, NFP = T::NFP, NF = T::NF, NFL = T::NFL, MT = T::MT, NPP = T::NPP;
    constexpr int KTQ = T::KTQ, LDQ = T::LDQ, KTF = T::KTF, LDF = T::LDF;

    extern __shared__ __align__(128) unsigned char smemRaw[];
    double* sD = reinterpret_cast<double*>(smemRaw);  // [3][NPP][LDQ]
    double* sL = sD + T::OPD;                         // [NPP][LDF]
    double* sWarpAll = sL + T::OPL;
    int* sFaceNodes = reinterpret_cast<int*>(sWarpAll + WARPS * T::WARP_DOUBLES);
    unsigned char* sMaps = reinterpret_cast<unsigned char*>(sFaceNodes + NFL);
    __shared__ __align__(8) unsigned long long bar;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;

    // ---- operators: one TMA bulk copy per CTA --------------------------------------------------------------
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smemAddr(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        constexpr uint32_t bytes = T::OPS * sizeof(double);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(&bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemAddr(sD)),
                     "l"(M.tiledOps), "r"(bytes), "r"(smemAddr(&bar))
                     : "memory");
    }
    for (int i = tid; i < NFL; i += WARPS * 32) sFaceNodes[i] = M.faceNodes[i];
    const int nMapsS = min(M.nMaps, kMaxMaps);
    for (int i = tid; i < nMapsS * NFP; i += WARPS * 32) sMaps[i] = M.nbrMaps[i];
    double* sQ = sWarpAll + warp * T::WARP_DOUBLES;  // [16 columns][LDQ]
    double* sFl = sQ + 16 * LDQ;                     // [16 columns][LDF]
    for (int i = lane; i < T::WARP_DOUBLES; i += 32) sQ[i] = 0.0;  // padding entries stay zero for ever
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(smemAddr(&bar)), "r"(0u)
                : "memory");
        }
    }
    __syncthreads();

    const Phys ph = makePhys(M);
    const int64_t S = M.stride;
    const int fp = t & 1;        // field pair of this lane's two accumulator columns: (p,vx) or (vy,vz)
    const int elSub = t >> 1;    // which of the n-tile's two elements

    for (int unit = blockIdx.x * WARPS + warp; unit < nUnits; unit += gridDim.x * WARPS) {
        const int e0 = A.eBegin + unit * 4;
        const int nE = min(4, A.eEnd - e0);

        // ---- 1. nodal values of the unit's 4 elements: coalesced per field, stored column-major (K-contiguous) ----
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double* src = A.yin + q * S + (int64_t)e0 * NP;
#pragma unroll
            for (int idx = lane; idx < 4 * NP; idx += 32) {
                const int el = idx / NP, j = idx - el * NP;
                sQ[(el * 4 + q) * LDQ + j] = el < nE ? src[idx] : 0.0;
            }
        }
        __syncwarp();

        // ---- 2. numerical flux at every (element, face, face node): gather the neighbour trace from global ----
        for (int w = lane; w < 4 * NFL; w += 32) {
            const int el = w / NFL, r = w - el * NFL, lf = r / NFP, m = r - lf * NFP;
            double fl[4] = {0, 0, 0, 0};
            double fscale = 0.0;
            if (el < nE) {
                const int e = e0 + el;
                const int flags = M.fflags[e * NF + lf];
                const int bc = flags & FLAG_BC_MASK;
                const double4 fg = *reinterpret_cast<const double4*>(M.fgeo + ((int64_t)e * NF + lf) * 4);
                const double n[3] = {fg.x, fg.y, fg.z};
                fscale = fg.w;
                const int own = sFaceNodes[r];
                double qm[4], qp[4] = {0, 0, 0, 0};
#pragma unroll
                for (int q = 0; q < 4; ++q) qm[q] = sQ[(el * 4 + q) * LDQ + own];
                if (bc == FACE_INTERIOR) {
                    const int nb = M.fnbr[e * NF + lf];
                    const int mapId = flags >> FLAG_MAP_SHIFT;
                    const int nn = mapId < kMaxMaps ? sMaps[mapId * NFP + m] : M.nbrMaps[mapId * NFP + m];
                    const int64_t gi = (int64_t)nb * NP + nn;
#pragma unroll
                    for (int q = 0; q < 4; ++q) qp[q] = A.yin[q * S + gi];
                }
                faceFlux(bc, (flags & FLAG_TAU_NEG) ? -1.0 : 1.0, n, ph, qm, qp, fl);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) sFl[(el * 4 + q) * LDF + r] = fscale * fl[q];
        }

        // per-lane geometric factors of the two elements this lane's columns belong to (one per n-tile)
        double G[2][9];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            const int el = nt * 2 + elSub;
#pragma unroll
            for (int k = 0; k < 9; ++k) G[nt][k] = el < nE ? M.Ginv[(int64_t)(e0 + el) * 9 + k] : 0.0;
        }
        __syncwarp();

        // ---- 3. B fragments of the volume contraction (reused by all 3*MT m-tiles) ----
        double Bq[2][KTQ];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int kt = 0; kt < KTQ; ++kt) Bq[nt][kt] = sQ[(nt * 8 + g) * LDQ + 4 * kt + t];

#pragma unroll 1
        for (int it = 0; it < MT; ++it) {
            // ---- 3a. T^u = Dw^u Q for the 8 rows of this m-tile ----
            double Tacc[3][2][2];
#pragma unroll
            for (int u = 0; u < 3; ++u)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) Tacc[u][nt][0] = Tacc[u][nt][1] = 0.0;
            const double* aRow = sD + (it * 8 + g) * LDQ + t;
#pragma unroll
            for (int kt = 0; kt < KTQ; ++kt) {
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                    const double a = aRow[u * NPP * LDQ + 4 * kt];
                    dmma884(Tacc[u][0][0], Tacc[u][0][1], a, Bq[0][kt]);
                    dmma884(Tacc[u][1][0], Tacc[u][1][1], a, Bq[1][kt]);
                }
            }
            // ---- 3b. combine with the per-element constants: rhs_vol (same fragment layout as the lift accumulators) ----
            double R[2][2], R2[2][2];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const double* Gm = G[nt];  // Gm[x*3+u]
                double au[3];
#pragma unroll
                for (int u = 0; u < 3; ++u) au[u] = Gm[u] * ph.v0[0] + Gm[3 + u] * ph.v0[1] + Gm[6 + u] * ph.v0[2];
                double ownA = 0.0, ownB = 0.0, send0 = 0.0, send1 = 0.0;
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                    const double Ta = Tacc[u][nt][0], Tb = Tacc[u][nt][1];
                    ownA = fma(au[u], Ta, ownA);
                    ownB = fma(au[u], Tb, ownB);
                    if (fp == 0) {  // columns (p, vx)
                        ownA = fma(ph.rc2 * Gm[u], Tb, ownA);
                        ownB = fma(ph.invRho * Gm[u], Ta, ownB);
                        send0 = fma(ph.invRho * Gm[3 + u], Ta, send0);
                        send1 = fma(ph.invRho * Gm[6 + u], Ta, send1);
                    } else {  // columns (vy, vz)
                        send0 = fma(ph.rc2 * Gm[3 + u], Ta, send0);
                        send0 = fma(ph.rc2 * Gm[6 + u], Tb, send0);
                    }
                }
                const double recv0 = __shfl_xor_sync(0xffffffffu, send0, 1);
                const double recv1 = __shfl_xor_sync(0xffffffffu, send1, 1);
                R[nt][0] = ownA + recv0;
                R[nt][1] = fp == 0 ? ownB : ownB + recv1;
                R2[nt][0] = R2[nt][1] = 0.0;
            }
            // ---- 3c. lift: rhs -= LIFT (Fscale * flux), two accumulator chains per n-tile ----
            const double* lRow = sL + (it * 8 + g) * LDF + t;
            const double* fRow0 = sFl + g * LDF + t;
            const double* fRow1 = sFl + (8 + g) * LDF + t;
#pragma unroll
            for (int kt = 0; kt < KTF; kt += 2) {
                const double a0 = lRow[4 * kt];
                dmma884(R[0][0], R[0][1], a0, fRow0[4 * kt]);
                dmma884(R[1][0], R[1][1], a0, fRow1[4 * kt]);
                if (kt + 1 < KTF) {
                    const double a1 = lRow[4 * kt + 4];
                    dmma884(R2[0][0], R2[0][1], a1, fRow0[4 * kt + 4]);
                    dmma884(R2[1][0], R2[1][1], a1, fRow1[4 * kt + 4]);
                }
            }
            // ---- 3d. fused RK update straight from the fragments (8 consecutive nodes per (element, field) = 64 B runs) ----
            const int i = it * 8 + g;
            if (i < NP) {
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    const int el = nt * 2 + elSub;
                    if (el < nE) {
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            const int q = fp * 2 + c;
                            const int64_t gi = q * S + (int64_t)(e0 + el) * NP + i;
                            rkUpdate(A, gi, R[nt][c] + R2[nt][c], sQ[(el * 4 + q) * LDQ + i]);
                        }
                    }
                }
            }
        }
        __syncwarp();  // the next unit overwrites this warp's staging area
    }
}

template <int P, int WARPS>
void launchTiled(const DeviceMesh& M, const StageArgs& A, cudaStream_t s) {
    using T = TetTile<P>;
    const int nEl = A.eEnd - A.eBegin;
    if (nEl <= 0) return;
    static int numSm = 0;
    static bool configured = false;
    const size_t smem = (size_t)(T::OPS + WARPS * T::WARP_DOUBLES) * sizeof(double) + T::NFL * sizeof(int) + kMaxMaps * T::NFP;
    if (!configured) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&numSm, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(stageTiledKernel<P, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    const int nUnits = (nEl + 3) / 4;
    const int grid = std::max(1, std::min(numSm, (nUnits + WARPS - 1) / WARPS));
    stageTiledKernel<P, WARPS><<<grid, WARPS * 32, smem, s>>>(M, A, nUnits);
}

}  // namespace

StageKernel selectTiledKernel(int dim, int order) {
    StageKernel k;
    if (dim == 3 && order == 4) { k.launch = &launchTiled<4, 12>; k.name = "stage_tiled_dmma<3,4>"; }
    if (dim == 3 && order == 3) { k.launch = &launchTiled<3, 12>; k.name = "stage_tiled_dmma<3,3>"; }
    return k;
}

bool tiledLayout(int dim, int order, int* npp, int* ldq, int* ldf) {
    if (dim == 3 && order == 4) { *npp = TetTile<4>::NPP; *ldq = TetTile<4>::LDQ; *ldf = TetTile<4>::LDF; return true; }
    if (dim == 3 && order == 3) { *npp = TetTile<3>::NPP; *ldq = TetTile<3>::LDQ; *ldf = TetTile<3>::LDF; return true; }
    return false;
}

}  // namespace dgb
