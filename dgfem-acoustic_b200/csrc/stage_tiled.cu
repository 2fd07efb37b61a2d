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

// Optional per-phase cycle counters (development aid): compile with -DDGB_TILED_PHASE_TIMERS, read with
// dgbTiledPhaseTimers(). Off in the product build.
#ifdef DGB_TILED_PHASE_TIMERS
__device__ unsigned long long g_phase[8];
#define PHASE_T(var) const long long var = clock64()
#define PHASE_ADD(k, t0, t1) if (lane == 0) atomicAdd(&g_phase[k], (unsigned long long)((t1) - (t0)))
#else
#define PHASE_T(var)
#define PHASE_ADD(k, t0, t1)
#endif

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    // volatile: keeps the issue order written in the source, i.e. the accumulator chains stay interleaved (the DMMA
    // pipe needs ~6 independent chains per warp to stay busy; ptxas otherwise groups the MMAs chain by chain)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
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

        PHASE_T(tp0);
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

        PHASE_T(tp1);
        PHASE_ADD(0, tp0, tp1);
        // ---- 2. numerical flux at every (element, face, face node) ----
        // 2a. face metadata: lane (el*4+lf) loads its face once and pre-multiplies the flux coefficients; the task
        //     lanes fetch them with shuffles (keeps the scarce FP64 pipe for the contractions).
        constexpr int NIT = (4 * NFL + 31) / 32;
        int mflags = FACE_ABSORBING, mnbr = -1;
        double mn0 = 0, mn1 = 0, mn2 = 0, mfs = 0;
        {
            const int mel = (lane & 15) >> 2, mlf = lane & 3;
            if (mel < nE) {
                const int e = e0 + mel;
                mflags = M.fflags[e * NF + mlf];
                mnbr = M.fnbr[e * NF + mlf];
                const double4 fg = *reinterpret_cast<const double4*>(M.fgeo + ((int64_t)e * NF + mlf) * 4);
                mn0 = fg.x; mn1 = fg.y; mn2 = fg.z; mfs = fg.w;
            }
        }
        // per-lane geometric factors of the two elements this lane's columns belong to (one per n-tile); issued here
        // so that their latency overlaps the gathers
        double G[2][9];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            const int el = nt * 2 + elSub;
#pragma unroll
            for (int k = 0; k < 9; ++k) G[nt][k] = el < nE ? M.Ginv[(int64_t)(e0 + el) * 9 + k] : 0.0;
        }
        const double mhf = 0.5 * mfs;
        const double cA = mhf * (ph.v0[0] * mn0 + ph.v0[1] * mn1 + ph.v0[2] * mn2);       // 1/2 Fscale v0.n
        const double cB = mhf * ph.rc2;                                                    // 1/2 Fscale rho0 c0^2
        const double cP = ((mflags & FLAG_TAU_NEG) ? -mhf : mhf) * ph.c0;                  // 1/2 Fscale tau c0
        const double cN0 = mhf * ph.invRho * mn0, cN1 = mhf * ph.invRho * mn1, cN2 = mhf * ph.invRho * mn2;
        // 2b. all neighbour gathers of the unit in flight at once
        double qp[NIT][4];
        int tflags[NIT];
#pragma unroll
        for (int s2 = 0; s2 < NIT; ++s2) {
            const int w = s2 * 32 + lane;
            const int el = min(w / NFL, 3), r = w - (w / NFL) * NFL, lf = r / NFP, m = r - lf * NFP;
            const int src = el * 4 + lf;
            const int flags = __shfl_sync(0xffffffffu, mflags, src);
            const int nb = __shfl_sync(0xffffffffu, mnbr, src);
            tflags[s2] = flags;
            const bool interior = (w < 4 * NFL) && ((flags & FLAG_BC_MASK) == FACE_INTERIOR) && nb >= 0;
            const int mapId = flags >> FLAG_MAP_SHIFT;
            int nn = 0;
            if (interior) nn = mapId < kMaxMaps ? sMaps[mapId * NFP + m] : M.nbrMaps[mapId * NFP + m];
            const int64_t gi = (int64_t)nb * NP + nn;
#pragma unroll
            for (int q = 0; q < 4; ++q) qp[s2][q] = interior ? A.yin[q * S + gi] : 0.0;
        }
        PHASE_T(tp2);
        PHASE_ADD(1, tp1, tp2);
        // 2c. fluxes (Fscale and the 1/2 are folded into the shuffled coefficients)
#pragma unroll
        for (int s2 = 0; s2 < NIT; ++s2) {
            const int w = s2 * 32 + lane;
            const int el = min(w / NFL, 3), r = w - (w / NFL) * NFL, lf = r / NFP;
            const int src = el * 4 + lf;
            const double n0 = __shfl_sync(0xffffffffu, mn0, src), n1 = __shfl_sync(0xffffffffu, mn1, src), n2 = __shfl_sync(0xffffffffu, mn2, src);
            const double fs = __shfl_sync(0xffffffffu, mfs, src);
            const double a = __shfl_sync(0xffffffffu, cA, src), b = __shfl_sync(0xffffffffu, cB, src), pc = __shfl_sync(0xffffffffu, cP, src);
            const double d0 = __shfl_sync(0xffffffffu, cN0, src), d1 = __shfl_sync(0xffffffffu, cN1, src), d2 = __shfl_sync(0xffffffffu, cN2, src);
            if (w < 4 * NFL) {
                const int own = sFaceNodes[r];
                double qm[4], fl[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) qm[q] = sQ[(el * 4 + q) * LDQ + own];
                const int bc = tflags[s2] & FLAG_BC_MASK;
                if (bc == FACE_INTERIOR) {
                    const double ps = qm[0] + qp[s2][0], v0s = qm[1] + qp[s2][1], v1s = qm[2] + qp[s2][2], v2s = qm[3] + qp[s2][3];
                    const double vns = n0 * v0s + n1 * v1s + n2 * v2s;
                    fl[0] = a * ps + b * vns + pc * (qm[0] - qp[s2][0]);
                    fl[1] = a * v0s + d0 * ps + pc * (qm[1] - qp[s2][1]);
                    fl[2] = a * v1s + d1 * ps + pc * (qm[2] - qp[s2][2]);
                    fl[3] = a * v2s + d2 * ps + pc * (qm[3] - qp[s2][3]);
                } else {
                    const double n[3] = {n0, n1, n2};
                    faceFlux(bc, 1.0, n, ph, qm, qp[s2], fl);
#pragma unroll
                    for (int q = 0; q < 4; ++q) fl[q] *= fs;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) sFl[(el * 4 + q) * LDF + r] = fl[q];
            }
        }
        __syncwarp();

        PHASE_T(tp3);
        PHASE_ADD(2, tp2, tp3);
        // ---- 3. B fragments of the volume contraction (reused by all 3*MT m-tiles) ----
        double Bq[2][KTQ];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int kt = 0; kt < KTQ; ++kt) Bq[nt][kt] = sQ[(nt * 8 + g) * LDQ + 4 * kt + t];

#pragma unroll 1
        for (int it = 0; it < MT; ++it) {
            // RK registers of this m-tile's rows: requested now, consumed after the tile's MMAs
            const int i = it * 8 + g;
            const bool rowOk = i < NP;
            int64_t gIdx[2];
            double uPre[2][2], accPre[2][2];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const int el = nt * 2 + elSub;
                gIdx[nt] = (int64_t)(fp * 2) * S + (int64_t)(e0 + el) * NP + i;
                const bool ok = rowOk && el < nE;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int64_t gi = gIdx[nt] + c * S;
                    uPre[nt][c] = (ok && A.mode != MODE_RHS) ? (A.mode == MODE_EULER ? A.yin[gi] : A.u[gi]) : 0.0;
                    accPre[nt][c] = (ok && A.mode >= MODE_RK2 && A.mode <= MODE_RK4) ? A.acc[gi] : 0.0;
                }
            }
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
            double R[2][2], R2[2][2], R3[2][2];
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
                R2[nt][0] = R2[nt][1] = R3[nt][0] = R3[nt][1] = 0.0;
            }
            // ---- 3c. lift: rhs -= LIFT (Fscale * flux), three accumulator chains per n-tile ----
            const double* lRow = sL + (it * 8 + g) * LDF + t;
            const double* fRow0 = sFl + g * LDF + t;
            const double* fRow1 = sFl + (8 + g) * LDF + t;
#pragma unroll
            for (int kt = 0; kt < KTF; kt += 3) {
                const double a0 = lRow[4 * kt];
                dmma884(R[0][0], R[0][1], a0, fRow0[4 * kt]);
                dmma884(R[1][0], R[1][1], a0, fRow1[4 * kt]);
                if (kt + 1 < KTF) {
                    const double a1 = lRow[4 * kt + 4];
                    dmma884(R2[0][0], R2[0][1], a1, fRow0[4 * kt + 4]);
                    dmma884(R2[1][0], R2[1][1], a1, fRow1[4 * kt + 4]);
                }
                if (kt + 2 < KTF) {
                    const double a2 = lRow[4 * kt + 8];
                    dmma884(R3[0][0], R3[0][1], a2, fRow0[4 * kt + 8]);
                    dmma884(R3[1][0], R3[1][1], a2, fRow1[4 * kt + 8]);
                }
            }
            // ---- 3d. fused RK update straight from the fragments (8 consecutive nodes per (element, field) = 64 B runs) ----
            if (rowOk) {
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    if (nt * 2 + elSub < nE) {
#pragma unroll
                        for (int c = 0; c < 2; ++c) rkApply(A, gIdx[nt] + c * S, (R[nt][c] + R2[nt][c]) + R3[nt][c], uPre[nt][c], accPre[nt][c]);
                    }
                }
            }
        }
        __syncwarp();  // the next unit overwrites this warp's staging area
        PHASE_T(tp4);
        PHASE_ADD(3, tp3, tp4);
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

#ifdef DGB_TILED_PHASE_TIMERS
extern "C" void dgbTiledPhaseTimers(unsigned long long* out, int reset) {
    cudaMemcpyFromSymbol(out, g_phase, sizeof(unsigned long long) * 8);
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_phase, z, sizeof(z)); }
}
#endif

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
