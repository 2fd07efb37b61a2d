// Tiled FP64 tensor-core (DMMA) stage kernel for tetrahedra of order 3 and 4 — the FP64-bound flagship orders
// (p >= 4 is compute-bound on B200; measured FP64 peak 37 TFLOP/s for both DFMA and DMMA,
// profiles/microbench/r01_fp64_peaks_b200.txt, so DMMA is chosen for its far lower operand/issue traffic).
//
// Same fused operator as stage_generic.cu (updateFlux + numStep + RK axpys of the reference, Mesh.cpp:476-674,
// solver.cpp:35-52, 261-285), organised as small dense contractions on mma.sync.m8n8k4.f64:
//   lift  = -LIFT (NPP x NFL) x Fl (NFL x 8 columns)           columns = (element, field) pairs, 2 elements per n-tile
//   T^u   =  Dw^u (NPP x NP)  x Q  (NP  x 8 columns)
//   rhs   = combine(T^u, G, v0) + lift
//
// Structure (one persistent CTA per SM, 8 warps):
//  * the operators (54 KB at p = 4) are staged once per CTA into shared memory with one TMA bulk copy
//    (cp.async.bulk + mbarrier);
//  * every WARP owns a unit of 4 consecutive elements end to end and only ever __syncwarp()s, so the memory phases of
//    some warps overlap the tensor phases of the others;
//  * software pipeline over units: while the DMMAs of unit k run, the nodal values, the neighbour traces (a gather
//    through the face-node maps), the face geometry and the inverse Jacobians of unit k+1 stream into shared memory
//    with cp.async (LDGSTS); the numerical flux of unit k+1 is then computed in place from shared memory;
//  * the k = dt L(y) values go through shared memory once more so that the RK registers (u, acc, y) are read and written
//    fully coalesced, with all loads of the unit in flight together.
// All MMA operand tiles are K-contiguous with a leading dimension chosen bank-conflict free (tile_cfg.h).
#include <algorithm>

#include "dgb_device.cuh"
#include "dgb_internal.h"
#include "dgb_launch.h"
#include "tile_cfg.h"

namespace dgb {

namespace {

constexpr int kMaxMaps = 64;
#ifndef DGB_TILED_WARPS
#define DGB_TILED_WARPS 8
#endif
constexpr int kWarps = DGB_TILED_WARPS;

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
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 8-byte asynchronous copy global -> shared (LDGSTS); valid == false writes zeros without touching global memory
__device__ __forceinline__ void cpAsync8(void* dst, const void* src, bool valid) {
    const int srcSize = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smemAddr(dst)), "l"(src), "r"(srcSize) : "memory");
}
__device__ __forceinline__ void cpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cpAsyncWaitAll() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int P>
struct WarpLayout {
    using T = TetTile<P>;
    // per-warp staging (doubles): nodal values, two flux/trace buffers, inverse Jacobians and face geometry (x2)
    // 16 columns of (element, field); 4 extra doubles after every element's 4 columns, so that the four elements of a
    // unit start in different bank groups (the flux phase works with one lane pair per (element, face))
    static constexpr int Q = 16 * T::LDQ + 16, FL = 16 * T::LDF + 16, GEO = 4 * 9 + 16 * 4;
    static constexpr int DOUBLES = Q + 2 * FL + 2 * GEO;
};

// FLOW == false is the v0 == 0 specialisation: the velocity columns are pre-combined with the inverse Jacobian
// (c_u = sum_x G_xu v_x, in place in shared memory), so operator Dw^u only has to see the 8 columns
// (p of 4 elements | c_u of 4 elements): 135 instead of 270 volume MMAs per unit and no cross-lane exchange.
template <int P, bool FLOW>
__global__ void __launch_bounds__(kWarps * 32, 1) stageTiledKernel(DeviceMesh M, StageArgs A, int nUnits) {
    using T = TetTile<P>;
    using W = WarpLayout<P>;
    constexpr int NP = T::NP, NFP = T::NFP, NF = T::NF, NFL = T::NFL, MT = T::MT, NPP = T::NPP;
    constexpr int KTQ = T::KTQ, LDQ = T::LDQ, KTF = T::KTF, LDF = T::LDF;
    constexpr int NIT = (4 * NFL + 31) / 32;  // flux tasks per lane
    auto colQ = [](int c) { return c * LDQ + (c >> 2) * 4; };  // start of column c = (element*4 + field) in sQ
    auto colF = [](int c) { return c * LDF + (c >> 2) * 4; };  // ... and in the flux / staging buffers

    extern __shared__ __align__(128) unsigned char smemRaw[];
    double* sD = reinterpret_cast<double*>(smemRaw);  // [3][NPP][LDQ]
    double* sL = sD + T::OPD;                         // [NPP][LDF]
    double* sWarpAll = sL + T::OPL;
    int* sFaceNodes = reinterpret_cast<int*>(sWarpAll + kWarps * W::DOUBLES);
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
    for (int i = tid; i < NFL; i += kWarps * 32) sFaceNodes[i] = M.faceNodes[i];
    const int nMapsS = min(M.nMaps, kMaxMaps);
    for (int i = tid; i < nMapsS * NFP; i += kWarps * 32) sMaps[i] = M.nbrMaps[i];
    double* sQ = sWarpAll + warp * W::DOUBLES;  // [16 columns][LDQ]
    double* sFlBuf = sQ + W::Q;                 // 2 x [16 columns][LDF]
    double* sGeoBuf = sFlBuf + 2 * W::FL;       // 2 x { Ginv [4][9], fgeo [16][4] }
    for (int i = lane; i < W::DOUBLES; i += 32) sQ[i] = 0.0;  // the K padding of sQ stays zero for ever
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
    // k = dt * L(y) comes straight out of the contractions: the staged operators are scaled by dt once per CTA
    // (dt == 1 for MODE_RHS), which removes one FP64 multiply per unknown from the FP64 pipe the DMMAs need.
    {
        const double dtScale = A.mode == MODE_RHS ? 1.0 : A.dt;
        for (int i = tid; i < T::OPS; i += kWarps * 32) sD[i] *= dtScale;
    }
    __syncthreads();

    const Phys ph = makePhys(M);
    const int64_t S = M.stride;
    const int fp = t & 1;      // field pair of this lane's two accumulator columns: (p,vx) or (vy,vz)
    const int elSub = t >> 1;  // which of the n-tile's two elements
    const int unitStride = gridDim.x * kWarps;
    const int mel = (lane & 15) >> 2, mlf = lane & 3;  // the face this lane looks after in the metadata step

    // Face metadata of a unit: lane (el*4+lf) holds flags / neighbour of its face (lanes 16..31 mirror 0..15).
    auto loadFaceIds = [&](int unit, int& flags, int& nbr) {
        flags = FACE_ABSORBING;
        nbr = -1;
        if (unit < nUnits) {
            const int e = A.eBegin + unit * 4 + mel;
            if (e < A.eEnd) { flags = M.fflags[e * NF + mlf]; nbr = M.fnbr[e * NF + mlf]; }
        }
    };
    // Everything unit `unit` needs, streamed into shared memory: nodal values, neighbour traces, geometry.
    auto prefetch = [&](int unit, int buf, int flags, int nbr) {
        if (unit >= nUnits) return;
        const int e0 = A.eBegin + unit * 4;
        const int nE = min(4, A.eEnd - e0);
        double* sFl = sFlBuf + buf * W::FL;
        double* sGeo = sGeoBuf + buf * W::GEO;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double* src = A.yin + q * S + (int64_t)e0 * NP;
#pragma unroll
            for (int idx = lane; idx < 4 * NP; idx += 32) {
                const int el = idx / NP, j = idx - el * NP;
                cpAsync8(&sQ[colQ(el * 4 + q) + j], el < nE ? src + idx : A.yin, el < nE);
            }
        }
#pragma unroll
        for (int s2 = 0; s2 < NIT; ++s2) {
            const int w = s2 * 32 + lane;
            const int el = min(w / NFL, 3), r = w - (w / NFL) * NFL, lf = r / NFP, m = r - lf * NFP;
            const int src = el * 4 + lf;
            const int fl = __shfl_sync(0xffffffffu, flags, src);
            const int nb = __shfl_sync(0xffffffffu, nbr, src);
            if (w < 4 * NFL) {
                const bool interior = ((fl & FLAG_BC_MASK) == FACE_INTERIOR) && nb >= 0;
                const int mapId = fl >> FLAG_MAP_SHIFT;
                int nn = 0;
                if (interior) nn = mapId < kMaxMaps ? sMaps[mapId * NFP + m] : M.nbrMaps[mapId * NFP + m];
                const int64_t gi = interior ? (int64_t)nb * NP + nn : 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) cpAsync8(&sFl[colF(el * 4 + q) + r], A.yin + q * S + gi, interior);
            }
        }
        for (int i = lane; i < 36; i += 32) cpAsync8(&sGeo[i], i < nE * 9 ? M.Ginv + (int64_t)e0 * 9 + i : M.Ginv, i < nE * 9);
        for (int i = lane; i < 64; i += 32) cpAsync8(&sGeo[36 + i], i < nE * 16 ? M.fgeo + (int64_t)e0 * 16 + i : M.fgeo, i < nE * 16);
    };

    int unit = blockIdx.x * kWarps + warp;
    int flagsCur, nbrCur, flagsNext = FACE_ABSORBING, nbrNext = -1;
    loadFaceIds(unit, flagsCur, nbrCur);
    prefetch(unit, 0, flagsCur, nbrCur);
    cpAsyncCommit();

    for (int iter = 0; unit < nUnits; unit += unitStride, ++iter) {
        const int buf = iter & 1;
        const int e0 = A.eBegin + unit * 4;
        const int nE = min(4, A.eEnd - e0);
        double* sFl = sFlBuf + buf * W::FL;
        const double* sGeo = sGeoBuf + buf * W::GEO;
        PHASE_T(tp0);
        loadFaceIds(unit + unitStride, flagsNext, nbrNext);  // consumed by the prefetch below, after the flux phase
        cpAsyncWaitAll();
        __syncwarp();
        PHASE_T(tp1);
        PHASE_ADD(0, tp0, tp1);

        // ---- 1. numerical flux at every (element, face, face node), in place over the gathered neighbour traces ----
        // Two lanes per face (lane & 15 = el*4+lf; lane >> 4 picks the even / odd face nodes): every coefficient of the
        // face is lane-local, no shuffles; Fscale and the 1/2 of the central flux are folded into the coefficients.
        {
            const double* fg = sGeo + 36 + (mel * 4 + mlf) * 4;
            const double n0 = fg[0], n1 = fg[1], n2 = fg[2], fs = fg[3];
            const double hf = 0.5 * fs;
            const double cA = hf * (ph.v0[0] * n0 + ph.v0[1] * n1 + ph.v0[2] * n2);  // 1/2 Fscale v0.n
            const double cP = ((flagsCur & FLAG_TAU_NEG) ? -hf : hf) * ph.c0;         // 1/2 Fscale tau c0
            const double cm = cA + cP, cp = cA - cP;                                  // weights of the own / neighbour value
            const double cB = hf * ph.rc2, cR = hf * ph.invRho;
            const double g0 = cB * n0, g1 = cB * n1, g2 = cB * n2;                    // 1/2 Fscale rho0 c0^2 n
            const double d0 = cR * n0, d1 = cR * n1, d2 = cR * n2;                    // 1/2 Fscale n / rho0
            const int bc = flagsCur & FLAG_BC_MASK;
            const double* qBase = sQ + colQ(mel * 4);
            double* fBase = sFl + colF(mel * 4) + mlf * NFP;
            const int* fn = sFaceNodes + mlf * NFP;
            // all face nodes of the lane in flight together (loads first, then arithmetic, then stores)
            constexpr int NM = (NFP + 1) / 2;
            double qm[NM][4], qp[NM][4];
#pragma unroll
            for (int s2 = 0; s2 < NM; ++s2) {
                const int m = min(2 * s2 + (lane >> 4), NFP - 1);
                const int own = fn[m];
#pragma unroll
                for (int q = 0; q < 4; ++q) { qm[s2][q] = qBase[q * LDQ + own]; qp[s2][q] = fBase[q * LDF + m]; }
            }
#pragma unroll
            for (int s2 = 0; s2 < NM; ++s2) {
                const int m = 2 * s2 + (lane >> 4);
                double fl[4];
                if (bc == FACE_INTERIOR) {
                    const double ps = qm[s2][0] + qp[s2][0];
                    fl[0] = cm * qm[s2][0] + cp * qp[s2][0] + g0 * (qm[s2][1] + qp[s2][1]) + g1 * (qm[s2][2] + qp[s2][2]) + g2 * (qm[s2][3] + qp[s2][3]);
                    fl[1] = cm * qm[s2][1] + cp * qp[s2][1] + d0 * ps;
                    fl[2] = cm * qm[s2][2] + cp * qp[s2][2] + d1 * ps;
                    fl[3] = cm * qm[s2][3] + cp * qp[s2][3] + d2 * ps;
                } else {
                    const double n[3] = {n0, n1, n2};
                    faceFlux(bc, 1.0, n, ph, qm[s2], qp[s2], fl);
#pragma unroll
                    for (int q = 0; q < 4; ++q) fl[q] *= fs;
                }
                if (m < NFP) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) fBase[q * LDF + m] = fl[q];
                }
            }
        }
        __syncwarp();
        PHASE_T(tp2);
        PHASE_ADD(1, tp1, tp2);

        // ---- 2. B fragments of the volume contraction, then release sQ to the prefetch of the next unit ----
        double Bq[FLOW ? 2 : 3][KTQ];
        if constexpr (FLOW) {
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int kt = 0; kt < KTQ; ++kt) Bq[nt][kt] = sQ[colQ(nt * 8 + g) + 4 * kt + t];
        } else {
            // contravariant velocity in place: columns (vx,vy,vz) of every element become (c_0,c_1,c_2)
#pragma unroll
            for (int idx = lane; idx < 4 * NP; idx += 32) {
                const int el = idx / NP, j = idx - el * NP;
                const double* Gm = sGeo + el * 9;  // Gm[x*3+u] = du_u/dx_x
                double* v = sQ + colQ(el * 4 + 1) + j;
                const double vx = v[0], vy = v[LDQ], vz = v[2 * LDQ];
#pragma unroll
                for (int u = 0; u < 3; ++u) v[u * LDQ] = Gm[u] * vx + Gm[3 + u] * vy + Gm[6 + u] * vz;
            }
            __syncwarp();
            // operator u sees the columns (p of elements 0..3 | c_u of elements 0..3)
            if (g < 4) {
#pragma unroll
                for (int kt = 0; kt < KTQ; ++kt) Bq[0][kt] = Bq[1][kt] = Bq[2][kt] = sQ[colQ(g * 4) + 4 * kt + t];
            } else {
#pragma unroll
                for (int u = 0; u < 3; ++u)
#pragma unroll
                    for (int kt = 0; kt < KTQ; ++kt) Bq[u][kt] = sQ[colQ((g - 4) * 4 + 1 + u) + 4 * kt + t];
            }
        }
        __syncwarp();
        prefetch(unit + unitStride, buf ^ 1, flagsNext, nbrNext);
        cpAsyncCommit();
        PHASE_T(tp3);
        PHASE_ADD(2, tp2, tp3);

        // ---- 3. lift for all m-tiles: R[it] = -LIFT (Fscale * flux); operand fragments are fetched one k-step ahead ----
        {
            double R[MT][2][2];
#pragma unroll
            for (int it = 0; it < MT; ++it)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) R[it][nt][0] = R[it][nt][1] = 0.0;
            const double* lRow = sL + g * LDF + t;
            const double* fRow0 = sFl + colF(g) + t;
            const double* fRow1 = sFl + colF(8 + g) + t;
            double aN[MT], b0N = fRow0[0], b1N = fRow1[0];
#pragma unroll
            for (int it = 0; it < MT; ++it) aN[it] = lRow[it * 8 * LDF];
#pragma unroll
            for (int kt = 0; kt < KTF; ++kt) {
                double a[MT];
                const double b0 = b0N, b1 = b1N;
#pragma unroll
                for (int it = 0; it < MT; ++it) a[it] = aN[it];
                if (kt + 1 < KTF) {
                    b0N = fRow0[4 * (kt + 1)];
                    b1N = fRow1[4 * (kt + 1)];
#pragma unroll
                    for (int it = 0; it < MT; ++it) aN[it] = lRow[it * 8 * LDF + 4 * (kt + 1)];
                }
#pragma unroll
                for (int it = 0; it < MT; ++it) {
                    dmma884(R[it][0][0], R[it][0][1], a[it], b0);
                    dmma884(R[it][1][0], R[it][1][1], a[it], b1);
                }
            }
            __syncwarp();  // every lane is done reading the fluxes: sFl becomes the output staging area
#pragma unroll
            for (int it = 0; it < MT; ++it) {
                const int i = it * 8 + g;
                if (i < NP) {
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) {
                        const int col = (nt * 2 + elSub) * 4 + fp * 2;
                        sFl[colF(col) + i] = R[it][nt][0];
                        sFl[colF(col + 1) + i] = R[it][nt][1];
                    }
                }
            }
        }
        PHASE_T(tp4);
        PHASE_ADD(3, tp3, tp4);

        // ---- 4. volume term per m-tile, combined with the per-element constants and added to the staged lift ----
        if constexpr (FLOW) {
            // Branch-free combine: the coefficient set of a lane depends on whether its two columns are (p,vx) or (vy,vz).
            //   own0 += cf0*Ta + cf1*Tb ; own1 += cf0*Tb + cf2*Ta ; send0 += cf3*Ta + cf4*Tb ; send1 += cf5*Ta
            double cf[2][3][6];
    #pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const double* Gm = sGeo + (nt * 2 + elSub) * 9;  // Gm[x*3+u] = du_u/dx_x
    #pragma unroll
                for (int u = 0; u < 3; ++u) {
                    cf[nt][u][0] = Gm[u] * ph.v0[0] + Gm[3 + u] * ph.v0[1] + Gm[6 + u] * ph.v0[2];
                    cf[nt][u][1] = fp == 0 ? ph.rc2 * Gm[u] : 0.0;
                    cf[nt][u][2] = fp == 0 ? ph.invRho * Gm[u] : 0.0;
                    cf[nt][u][3] = fp == 0 ? ph.invRho * Gm[3 + u] : ph.rc2 * Gm[3 + u];
                    cf[nt][u][4] = fp == 0 ? 0.0 : ph.rc2 * Gm[6 + u];
                    cf[nt][u][5] = fp == 0 ? ph.invRho * Gm[6 + u] : 0.0;
                }
            }
            auto volumeMma = [&](int it, double (&Tacc)[3][2][2]) {
    #pragma unroll
                for (int u = 0; u < 3; ++u)
    #pragma unroll
                    for (int nt = 0; nt < 2; ++nt) Tacc[u][nt][0] = Tacc[u][nt][1] = 0.0;
                const double* aRow = sD + (it * 8 + g) * LDQ + t;
                double aN[3];
    #pragma unroll
                for (int u = 0; u < 3; ++u) aN[u] = aRow[u * NPP * LDQ];
    #pragma unroll
                for (int kt = 0; kt < KTQ; ++kt) {
                    double a[3];
    #pragma unroll
                    for (int u = 0; u < 3; ++u) a[u] = aN[u];
                    if (kt + 1 < KTQ) {
    #pragma unroll
                        for (int u = 0; u < 3; ++u) aN[u] = aRow[u * NPP * LDQ + 4 * (kt + 1)];
                    }
    #pragma unroll
                    for (int u = 0; u < 3; ++u) {
                        dmma884(Tacc[u][0][0], Tacc[u][0][1], a[u], Bq[0][kt]);
                        dmma884(Tacc[u][1][0], Tacc[u][1][1], a[u], Bq[1][kt]);
                    }
                }
            };
            auto combine = [&](int it, const double (&Tacc)[3][2][2]) {
                const int i = it * 8 + g;
    #pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    double own0 = 0.0, own1 = 0.0, send0 = 0.0, send1 = 0.0;
    #pragma unroll
                    for (int u = 0; u < 3; ++u) {
                        const double Ta = Tacc[u][nt][0], Tb = Tacc[u][nt][1];
                        const double* c = cf[nt][u];
                        own0 = fma(c[0], Ta, own0);
                        own0 = fma(c[1], Tb, own0);
                        own1 = fma(c[0], Tb, own1);
                        own1 = fma(c[2], Ta, own1);
                        send0 = fma(c[3], Ta, send0);
                        send0 = fma(c[4], Tb, send0);
                        send1 = fma(c[5], Ta, send1);
                    }
                    const double recv0 = __shfl_xor_sync(0xffffffffu, send0, 1);
                    const double recv1 = __shfl_xor_sync(0xffffffffu, send1, 1);
                    // (p,vx) lanes: p += partner's divergence part, vx complete; (vy,vz) lanes: each gets its pressure-gradient part
                    if (i < NP) {
                        const int col = (nt * 2 + elSub) * 4 + fp * 2;
                        sFl[colF(col) + i] += own0 + recv0;
                        sFl[colF(col + 1) + i] += fp == 0 ? own1 : own1 + recv1;
                    }
                }
            };
            {
                double Ta0[3][2][2], Ta1[3][2][2];
                volumeMma(0, Ta0);
    #pragma unroll 1
                for (int it = 1; it + 1 < MT; it += 2) {  // two tiles per trip so that the two accumulator sets keep their names
                    volumeMma(it, Ta1);
                    combine(it - 1, Ta0);
                    volumeMma(it + 1, Ta0);
                    combine(it, Ta1);
                }
                if ((MT & 1) == 0) {
                    volumeMma(MT - 1, Ta1);
                    combine(MT - 2, Ta0);
                    combine(MT - 1, Ta1);
                } else {
                    combine(MT - 1, Ta0);
                }
            }
        } else {
            // lanes t < 2 hold T^u_p of elements (2t, 2t+1): they finish the three velocity equations;
            // lanes t >= 2 hold Dw^u c_u of elements (2(t-2), 2(t-2)+1): they finish the pressure equation.
            const int eA = (t & 1) * 2;
            double gw[2][9];
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int k = 0; k < 9; ++k) gw[c][k] = ph.invRho * sGeo[(eA + c) * 9 + k];
            auto volumeMma0 = [&](int it, double (&Tacc)[3][2]) {
#pragma unroll
                for (int u = 0; u < 3; ++u) Tacc[u][0] = Tacc[u][1] = 0.0;
                const double* aRow = sD + (it * 8 + g) * LDQ + t;
                double aN[3];
#pragma unroll
                for (int u = 0; u < 3; ++u) aN[u] = aRow[u * NPP * LDQ];
#pragma unroll
                for (int kt = 0; kt < KTQ; ++kt) {
                    double a[3];
#pragma unroll
                    for (int u = 0; u < 3; ++u) a[u] = aN[u];
                    if (kt + 1 < KTQ) {
#pragma unroll
                        for (int u = 0; u < 3; ++u) aN[u] = aRow[u * NPP * LDQ + 4 * (kt + 1)];
                    }
#pragma unroll
                    for (int u = 0; u < 3; ++u) dmma884(Tacc[u][0], Tacc[u][1], a[u], Bq[u][kt]);
                }
            };
            auto combine0 = [&](int it, const double (&Tacc)[3][2]) {
                const int i = it * 8 + g;
                if (i < NP) {
                    if (t < 2) {
#pragma unroll
                        for (int c = 0; c < 2; ++c)
#pragma unroll
                            for (int x = 0; x < 3; ++x) {
                                const double r = gw[c][x * 3] * Tacc[0][c] + gw[c][x * 3 + 1] * Tacc[1][c] + gw[c][x * 3 + 2] * Tacc[2][c];
                                sFl[colF((eA + c) * 4 + 1 + x) + i] += r;
                            }
                    } else {
#pragma unroll
                        for (int c = 0; c < 2; ++c) sFl[colF((eA + c) * 4) + i] += ph.rc2 * ((Tacc[0][c] + Tacc[1][c]) + Tacc[2][c]);
                    }
                }
            };
            double Ta0[3][2], Ta1[3][2];
            volumeMma0(0, Ta0);
#pragma unroll 1
            for (int it = 1; it + 1 < MT; it += 2) {
                volumeMma0(it, Ta1);
                combine0(it - 1, Ta0);
                volumeMma0(it + 1, Ta0);
                combine0(it, Ta1);
            }
            if ((MT & 1) == 0) {
                volumeMma0(MT - 1, Ta1);
                combine0(MT - 2, Ta0);
                combine0(MT - 1, Ta1);
            } else {
                combine0(MT - 1, Ta0);
            }
        }
        __syncwarp();
        PHASE_T(tp5);
        PHASE_ADD(4, tp4, tp5);

        // ---- 5. fused RK update, coalesced: all RK-register loads of the unit in flight, then the stores ----
        {
            constexpr int NV = (4 * NP + 31) / 32;
            const int nValid = nE * NP;
            double uv[4][NV], av[4][NV];
            const bool needU = A.mode != MODE_RHS;
            const bool needAcc = A.mode >= MODE_RK2 && A.mode <= MODE_RK4;
            const double* uSrc = A.mode == MODE_EULER ? A.yin : A.u;
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int s2 = 0; s2 < NV; ++s2) {
                    const int idx = s2 * 32 + lane;
                    const int64_t gi = q * S + (int64_t)e0 * NP + idx;
                    const bool ok = idx < nValid;
                    uv[q][s2] = (ok && needU) ? uSrc[gi] : 0.0;
                    av[q][s2] = (ok && needAcc) ? A.acc[gi] : 0.0;
                }
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int s2 = 0; s2 < NV; ++s2) {
                    const int idx = s2 * 32 + lane;
                    if (idx < nValid) {
                        const int el = idx / NP, i = idx - el * NP;
                        rkApplyK(A, q * S + (int64_t)e0 * NP + idx, sFl[colF(el * 4 + q) + i], uv[q][s2], av[q][s2]);
                    }
                }
        }
        __syncwarp();
        PHASE_T(tp6);
        PHASE_ADD(5, tp5, tp6);
        flagsCur = flagsNext;
        nbrCur = nbrNext;
    }
    cpAsyncWaitAll();
}

template <int P, bool FLOW>
void launchTiledImpl(const DeviceMesh& M, const StageArgs& A, cudaStream_t s) {
    using T = TetTile<P>;
    const int nEl = A.eEnd - A.eBegin;
    if (nEl <= 0) return;
    static KernelConfig kc;
    const size_t smem = (size_t)(T::OPS + kWarps * WarpLayout<P>::DOUBLES) * sizeof(double) + T::NFL * sizeof(int) + kMaxMaps * T::NFP;
    const int numSm = configureKernel(kc, stageTiledKernel<P, FLOW>, smem, "stage_tiled_dmma");
    const int nUnits = (nEl + 3) / 4;
    const int grid = std::max(1, std::min(numSm - std::min(A.smReserve, numSm / 2), (nUnits + kWarps - 1) / kWarps));
    stageTiledKernel<P, FLOW><<<grid, kWarps * 32, smem, s>>>(M, A, nUnits);
}

template <int P>
void launchTiled(const DeviceMesh& M, const StageArgs& A, cudaStream_t s) {
    if (M.v0[0] == 0.0 && M.v0[1] == 0.0 && M.v0[2] == 0.0) launchTiledImpl<P, false>(M, A, s);
    else launchTiledImpl<P, true>(M, A, s);
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
    if (dim == 3 && order == 4) { k.launch = &launchTiled<4>; k.name = "stage_tiled_dmma<3,4>"; }
    if (dim == 3 && order == 3) { k.launch = &launchTiled<3>; k.name = "stage_tiled_dmma<3,3>"; }
    return k;
}

bool tiledLayout(int dim, int order, int* npp, int* ldq, int* ldf) {
    if (dim == 3 && order == 4) { *npp = TetTile<4>::NPP; *ldq = TetTile<4>::LDQ; *ldf = TetTile<4>::LDF; return true; }
    if (dim == 3 && order == 3) { *npp = TetTile<3>::NPP; *ldq = TetTile<3>::LDQ; *ldf = TetTile<3>::LDF; return true; }
    return false;
}

}  // namespace dgb
