// Warp-specialised FP64 tensor-core (DMMA) stage kernel for tetrahedra of order 3 and 4, zero mean flow.
//
// Same fused operator as stage_generic.cu / stage_tiled.cu (updateFlux + numStep + RK axpys of the reference,
// Mesh.cpp:476-674, solver.cpp:35-52, 261-285). What changes is who does what inside the persistent CTA:
//
//   * 4 MMA warps, one per SM sub-partition. Each keeps ITS SLICE OF THE OPERATORS IN REGISTERS for the whole launch
//     (role r < 3: Dw^r and k-tiles of -LIFT; role 3: the remaining k-tiles of -LIFT), so the tensor pipe is fed with one
//     shared-memory load per ~5 DMMAs instead of one per DMMA: the element data are the A fragments (8 elements x 4 k),
//     the register-resident operator the B fragments (4 k x 8 nodes), accumulators are (8 elements x 8 nodes).
//       role r < 3 :  T^r   = Dw^r p                                   (feeds the three velocity equations)
//                     P_r   = Dw^r (rho0 c0^2 c_r) - LIFT[:, slice r] F_p        c_r = sum_x G_xr v_x
//                     V_r,x =                      - LIFT[:, slice r] F_vx      x = 0..2
//       role 3     :  P_3, V_3,x  with the last slice of LIFT
//     The K-split partial sums are added up by the epilogue.
//   * 8 service warps, warp s owns element s of every tile of 8 elements: cp.async gathers of the nodal values, the
//     neighbour traces (through the face-node maps) and the geometry of tile t+1; numerical flux in place and the
//     contravariant velocities of tile t; combine + fused, coalesced RK update of tile t-1.
//   * mbarrier hand-off in both directions (full / empty for the 3-deep input ring and the 2-deep output ring); no
//     __syncthreads in the steady state. Registers are re-balanced with setmaxnreg (MMA warps up, service warps down).
//
// Shared memory rows are K-contiguous per element with a leading dimension == 8 (mod 16) doubles: the 128-bit fragment
// loads (two k-tiles per load; the operator registers are permuted to match) and the 128-bit accumulator stores are
// bank-conflict free.
#include <algorithm>
#include <type_traits>

#include "dgb_device.cuh"
#include "dgb_internal.h"

namespace dgb {

namespace {

constexpr int kMaxMapsWs = 64;
constexpr int kMmaWarps = 4, kFrontWarps = 8, kBackWarps = 4, kThreadsWs = (kMmaWarps + kFrontWarps + kBackWarps) * 32;
constexpr int kTileEl = 8;       // elements per tile = rows of one m8n8k4 A fragment
constexpr int kInStages = 2, kOutStages = 2, kGeoStages = 8;  // powers of two; geometry must outlive the back warps' lag
constexpr int kRegsMma = 184, kRegsFront = 112, kRegsBack = 104;  // 128*184 + 256*112 + 128*104 = 65536

__host__ __device__ constexpr int padTo8mod16(int n) {
    int ld = (n + 1) / 2 * 2;
    while (ld % 16 != 8) ld += 2;
    return ld;
}

template <int P>
struct WsCfg {
    static constexpr int NP = (P + 1) * (P + 2) * (P + 3) / 6, NFP = (P + 1) * (P + 2) / 2, NF = 4, NFL = NF * NFP;
    static constexpr int NT = (NP + 7) / 8, NPP = NT * 8;    // node tiles (n of the MMA)
    static constexpr int KTQ = (NP + 3) / 4, KQ = KTQ * 4;   // k-tiles of a volume block, padded block length
    // lift k-tiles owned by each of the MMA roles 0..2 (role r: [r*NSD, (r+1)*NSD)); role 3 owns the rest
    static constexpr int NSD = P == 4 ? 3 : 2;
    // input row of one element (doubles): p | c_0 | c_1 | c_2 | F_p | F_vx | F_vy | F_vz
    static constexpr int OFF_P = 0, OFF_C = KQ, OFF_F = 4 * KQ;
    static constexpr int LDI = padTo8mod16(4 * KQ + 4 * NFL);
    static constexpr int LDS_ = 3 * KQ;       // staging of the raw velocities [3][KQ] for the flux gather
    static constexpr int LDG_ = 32;           // geometry: Ginv[9] at 0, fgeo[16] at 16
    static constexpr int LDO = padTo8mod16(NPP);  // output row of one element in one panel
    static constexpr int NPANEL = 3 + 16;     // T^0..2, then part[role][field]
    static constexpr int IN_TILE = kTileEl * LDI, STG_TILE = kTileEl * LDS_, GEO_TILE = kTileEl * LDG_;
    static constexpr int OUT_TILE = NPANEL * kTileEl * LDO;
    static constexpr int SMEM_DOUBLES = kInStages * IN_TILE + STG_TILE + kGeoStages * GEO_TILE + kOutStages * OUT_TILE;
    // Position of face node (lf, m) inside a flux block: the first 8 nodes of the four faces, then the remaining ones, so
    // that the 8-lanes-per-face mapping of the flux phase touches consecutive addresses (the lift operator columns are
    // permuted to match when they are loaded into registers)
    __host__ __device__ static constexpr int faceSlot(int lf, int m) { return m < 8 ? lf * 8 + m : 32 + lf * (NFP - 8) + (m - 8); }
    __host__ __device__ static constexpr int slotFaceNode(int k) {  // inverse: lift column lf*NFP + m of slot k
        return k < 32 ? (k >> 3) * NFP + (k & 7) : ((k - 32) / (NFP - 8)) * NFP + 8 + (k - 32) % (NFP - 8);
    }
    static_assert(NFL % 4 == 0, "lift contraction length must be a multiple of 4");
};

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ uint32_t sAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cpAsync8z(void* dst, const void* src, bool valid) {
    const int srcSize = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(sAddr(dst)), "l"(src), "r"(srcSize) : "memory");
}
__device__ __forceinline__ void mbarInit(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sAddr(bar)), "r"(count));
}
__device__ __forceinline__ void mbarArrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sAddr(bar)) : "memory");
}
__device__ __forceinline__ void mbarWait(unsigned long long* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(sAddr(bar)), "r"(parity)
                     : "memory");
    }
}

// k index (inside a block that starts at a multiple of 4) that lane t feeds into k-tile kt of a block of KT k-tiles:
// k-tiles are loaded in pairs with one 128-bit access per lane (doubles 8*pair + 2t, 2t + 1), a last odd k-tile with a
// 64-bit access (4*kt + t).
__host__ __device__ constexpr int kOfTile(int kt, int KT, int t) {
    return kt < (KT / 2) * 2 ? 8 * (kt / 2) + 2 * t + (kt & 1) : 4 * kt + t;
}
// A fragments (element data) of KT consecutive k-tiles starting at `base` (this lane's element row + block offset)
template <int KT>
__device__ __forceinline__ void loadFrags(double (&a)[KT], const double* base, int t) {
#pragma unroll
    for (int kp = 0; kp < KT / 2; ++kp) {
        const double2 v = *reinterpret_cast<const double2*>(base + 8 * kp + 2 * t);
        a[2 * kp] = v.x;
        a[2 * kp + 1] = v.y;
    }
    if (KT & 1) a[KT - 1] = base[4 * (KT - 1) + t];
}

struct WsSmem {
    double* in;    // [kInStages][8][LDI]
    double* stg;   // [8][3][KQ]
    double* geo;   // [kGeoStages][8][32]
    double* out;   // [kOutStages][NPANEL][8][LDO]
    int* faceNodes;
    unsigned char* maps;
    unsigned long long* bars;  // full[3], inEmpty[3], outFull[2], outEmpty[2]
};

// ------------------------------------------------------------------------------------------------------------------
// MMA warps. Roles 0..2 run the SAME instruction stream (only register contents and a few offsets differ) and the
// repeated jobs are real loops: the per-SM instruction working set has to stay inside the 32 KB L1.5 instruction cache,
// the loop of one MMA warp inside the ~6 KB L0 of its sub-partition (a DMMA costs 32 B of code with its pacing NOP).
// ------------------------------------------------------------------------------------------------------------------
struct MmaBars {
    unsigned long long *full, *inEmpty, *outFull, *outEmpty;
};
__device__ __forceinline__ MmaBars mmaBars(const WsSmem& sm) {
    return {sm.bars, sm.bars + kInStages, sm.bars + 2 * kInStages, sm.bars + 2 * kInStages + kOutStages};
}

// roles 0..2: T^r = Dw^r p ; P_r = Dw^r (rho c^2 c_r) + (-LIFT slice r) F_p ; V_r,x = (-LIFT slice r) F_vx
template <int P>
__device__ __forceinline__ void mmaWarpD(const DeviceMesh& M, const StageArgs& A, const WsSmem& sm, int nIt, int role, int lane) {
    using C = WsCfg<P>;
    constexpr int NP = C::NP, NT = C::NT, KTQ = C::KTQ, LDI = C::LDI, LDO = C::LDO, NFL = C::NFL, NS = C::NSD;
    constexpr int PS = kTileEl * LDO;
    const int g = lane >> 2, t = lane & 3;
    const double scale = A.mode == MODE_RHS ? 1.0 : A.dt;  // k = dt L(y) leaves the tensor pipe directly
    const int kt0 = role * NS;

    // register-resident operator slice, as B fragments: lane (g, t) holds Op[node 8*nt + g][k(kt, t)]
    double D[NT][KTQ], L[NT][NS];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int i = nt * 8 + g;
#pragma unroll
        for (int kt = 0; kt < KTQ; ++kt) {
            const int k = kOfTile(kt, KTQ, t);
            D[nt][kt] = (i < NP && k < NP) ? scale * M.DwT[((size_t)role * NP + k) * NP + i] : 0.0;
        }
#pragma unroll
        for (int kt = 0; kt < NS; ++kt) {
            const int k = 4 * kt0 + kOfTile(kt, NS, t);
            L[nt][kt] = i < NP ? scale * M.nLiftT[(size_t)C::slotFaceNode(k) * NP + i] : 0.0;
        }
    }
    const MmaBars bar = mmaBars(sm);
    const int rowOff = g * LDI, offC = C::OFF_C + role * C::KQ, offF = C::OFF_F + 4 * kt0;
    const int outOff = g * LDO + 2 * t, panelT = role * PS, panelP = (3 + role * 4) * PS;

    for (int it = 0; it < nIt; ++it) {
        const int b = it & (kInStages - 1), ob = it & (kOutStages - 1);
        const double* row = sm.in + b * C::IN_TILE + rowOff;
        double* out = sm.out + ob * C::OUT_TILE + outOff;
        mbarWait(&bar.full[b], (it / kInStages) & 1);
        mbarWait(&bar.outEmpty[ob], ((it / kOutStages) & 1) ^ 1);

        auto store = [&](double* dst, const double (&acc)[NT][2]) {
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) *reinterpret_cast<double2*>(dst + nt * 8) = make_double2(acc[nt][0], acc[nt][1]);
        };
        double aP[KTQ], aC[KTQ], aV[NS];
        loadFrags<KTQ>(aP, row + C::OFF_P, t);
        loadFrags<KTQ>(aC, row + offC, t);
        {
            double acc[NT][2];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
#pragma unroll
            for (int kt = 0; kt < KTQ; ++kt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) dmma(acc[nt], aP[kt], D[nt][kt]);
            loadFrags<NS>(aV, row + offF, t);
            store(out + panelT, acc);
        }
        {
            double acc[NT][2];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
#pragma unroll
            for (int kt = 0; kt < KTQ; ++kt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) dmma(acc[nt], aC[kt], D[nt][kt]);
#pragma unroll
            for (int kt = 0; kt < NS; ++kt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) dmma(acc[nt], aV[kt], L[nt][kt]);
            loadFrags<NS>(aV, row + offF + NFL, t);
            store(out + panelP, acc);
        }
#pragma unroll 1
        for (int x = 1; x < 4; ++x) {
            double acc[NT][2];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
#pragma unroll
            for (int kt = 0; kt < NS; ++kt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) dmma(acc[nt], aV[kt], L[nt][kt]);
            loadFrags<NS>(aV, row + offF + (x < 3 ? x + 1 : 3) * NFL, t);  // next field (the last request is a dummy)
            store(out + panelP + x * PS, acc);
        }
        __syncwarp();
        if (lane == 0) {
            mbarArrive(&bar.inEmpty[b]);
            mbarArrive(&bar.outFull[ob]);
        }
    }
}

// role 3: P_3 and V_3,x with the last slice of -LIFT
template <int P>
__device__ __forceinline__ void mmaWarpL(const DeviceMesh& M, const StageArgs& A, const WsSmem& sm, int nIt, int lane) {
    using C = WsCfg<P>;
    constexpr int NP = C::NP, NT = C::NT, LDI = C::LDI, LDO = C::LDO, NFL = C::NFL, KT0 = 3 * C::NSD, NS = NFL / 4 - KT0;
    constexpr int PS = kTileEl * LDO;
    const int g = lane >> 2, t = lane & 3;
    const double scale = A.mode == MODE_RHS ? 1.0 : A.dt;
    double L[NT][NS];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int i = nt * 8 + g;
#pragma unroll
        for (int kt = 0; kt < NS; ++kt) {
            const int k = 4 * KT0 + kOfTile(kt, NS, t);
            L[nt][kt] = i < NP ? scale * M.nLiftT[(size_t)C::slotFaceNode(k) * NP + i] : 0.0;
        }
    }
    const MmaBars bar = mmaBars(sm);
    for (int it = 0; it < nIt; ++it) {
        const int b = it & (kInStages - 1), ob = it & (kOutStages - 1);
        const double* row = sm.in + b * C::IN_TILE + g * LDI + C::OFF_F + 4 * KT0;
        double* out = sm.out + ob * C::OUT_TILE + g * LDO + 2 * t + 15 * PS;
        mbarWait(&bar.full[b], (it / kInStages) & 1);
        mbarWait(&bar.outEmpty[ob], ((it / kOutStages) & 1) ^ 1);
        double aV[NS];
        loadFrags<NS>(aV, row, t);
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
            double acc[NT][2];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
#pragma unroll
            for (int kt = 0; kt < NS; ++kt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) dmma(acc[nt], aV[kt], L[nt][kt]);
            loadFrags<NS>(aV, row + (q < 3 ? q + 1 : 3) * NFL, t);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) *reinterpret_cast<double2*>(out + q * PS + nt * 8) = make_double2(acc[nt][0], acc[nt][1]);
        }
        __syncwarp();
        if (lane == 0) {
            mbarArrive(&bar.inEmpty[b]);
            mbarArrive(&bar.outFull[ob]);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// front warps: warp s owns element s of every tile (loads, contravariant velocities, numerical flux)
// ------------------------------------------------------------------------------------------------------------------
template <int P>
__device__ __forceinline__ void frontWarp(const DeviceMesh& M, const StageArgs& A, const WsSmem& sm, int nIt, int s, int lane) {
    using C = WsCfg<P>;
    constexpr int NP = C::NP, NFP = C::NFP, NF = C::NF, NFL = C::NFL, KQ = C::KQ, LDI = C::LDI;
    constexpr int NA = NP < 32 ? NP : 32, NB = NP - NA;  // nodes handled by lane j (pass A) and by lanes 0..NB-1 as node 32+lane (pass B)
    constexpr int NFB = NFP - 8;                         // face nodes 8.. of a face are handled in the second flux pass
    static_assert(NB >= 0 && NB <= 32 && NFB > 0 && NFB <= 8, "lane mappings assume NP <= 64, 8 < NFP <= 16");
    const Phys ph = makePhys(M);
    const int64_t S = M.stride;
    // flux lanes: face lf, face node ml (first pass) and 8 + ml (second pass, ml < NFB); all per-lane constants
    const int lf = lane >> 3, ml = lane & 7;
    const bool act1 = ml < NFB;
    const int own0 = sm.faceNodes[lf * NFP + ml], own1 = sm.faceNodes[lf * NFP + (act1 ? 8 + ml : 0)];
    const int slot0 = C::faceSlot(lf, ml), slot1 = C::faceSlot(lf, act1 ? 8 + ml : 8);
    const bool actA = lane < NA, actQB = lane < NB;
    // the four fields of the stage input
    const double* const y0 = A.yin;
    const double* const y1 = y0 + S;
    const double* const y2 = y1 + S;
    const double* const y3 = y2 + S;

    unsigned long long* full = sm.bars;
    unsigned long long* inEmpty = sm.bars + kInStages;
    double* const stg = sm.stg + s * C::LDS_;
    const int eStep = (int)gridDim.x * kTileEl;
    const int eLast = A.eEnd - 1;
    int eNext = A.eBegin + (int)blockIdx.x * kTileEl + s;  // element of the tile whose loads are issued next

    // Face metadata of element e (clamped: rows past the end of the range are computed on a copy of the last element and
    // never stored)
    auto loadMeta = [&](int e, int& flags, int& nbr) {
        const int ec = min(e, eLast);
        flags = M.fflags[ec * NF + lf];
        nbr = M.fnbr[ec * NF + lf];
    };
    // Everything the element needs: nodal values and neighbour traces into registers, geometry into shared memory.
    // Boundary faces read their own element instead of a neighbour (the value is ignored by the boundary fluxes).
    auto issueLoads = [&](int e, int geoSlot, int flags, int nbr, double (&qv)[2][4], double (&tr)[2][4]) {
        const int ec = min(e, eLast);
        const unsigned off = (unsigned)ec * NP + lane;
        if (actA) { qv[0][0] = y0[off]; qv[0][1] = y1[off]; qv[0][2] = y2[off]; qv[0][3] = y3[off]; }
        if (NB > 0 && actQB) { qv[1][0] = y0[off + 32]; qv[1][1] = y1[off + 32]; qv[1][2] = y2[off + 32]; qv[1][3] = y3[off + 32]; }
        const bool interior = ((flags & FLAG_BC_MASK) == FACE_INTERIOR) && nbr >= 0;
        const int mapId = flags >> FLAG_MAP_SHIFT;
        int nn0 = own0, nn1 = own1;
        if (interior) {
            const unsigned char* mp = mapId < kMaxMapsWs ? sm.maps + mapId * NFP : M.nbrMaps + mapId * NFP;
            nn0 = mp[ml];
            nn1 = mp[act1 ? 8 + ml : 0];
        }
        const unsigned tb = (unsigned)(interior ? nbr : ec) * NP;
        const unsigned t0 = tb + nn0, t1 = tb + nn1;
        tr[0][0] = y0[t0]; tr[0][1] = y1[t0]; tr[0][2] = y2[t0]; tr[0][3] = y3[t0];
        tr[1][0] = y0[t1]; tr[1][1] = y1[t1]; tr[1][2] = y2[t1]; tr[1][3] = y3[t1];
        double* geo = sm.geo + geoSlot * C::GEO_TILE + s * C::LDG_;
        if (lane < 9) cpAsync8z(&geo[lane], M.Ginv + (size_t)ec * 9 + lane, true);
        else if (lane >= 16) cpAsync8z(&geo[lane], M.fgeo + (size_t)ec * 16 + (lane - 16), true);
    };

    // contravariant velocities + numerical flux into the input row of element s
    auto prepAndFlux = [&](double* inRow, const double* geo, int flags, const double (&qv)[2][4], const double (&tr)[2][4]) {
        {
            // p and c_u = rho0 c0^2 sum_x G_xu v_x   (Ginv[x*3+u] = du_u/dx_x); raw velocities staged for the flux gather
            double Gs[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) Gs[k] = ph.rc2 * geo[k];
#pragma unroll
            for (int v = 0; v < (NB > 0 ? 2 : 1); ++v) {
                const int j = v * 32 + lane;
                if (v == 0 ? actA : actQB) {
                    const double vx = qv[v][1], vy = qv[v][2], vz = qv[v][3];
                    inRow[C::OFF_P + j] = qv[v][0];
#pragma unroll
                    for (int u = 0; u < 3; ++u) inRow[C::OFF_C + u * KQ + j] = Gs[u] * vx + Gs[3 + u] * vy + Gs[6 + u] * vz;
                    stg[j] = vx; stg[KQ + j] = vy; stg[2 * KQ + j] = vz;
                }
            }
        }
        __syncwarp();
        {
            const double2 nA = *reinterpret_cast<const double2*>(geo + 16 + lf * 4);
            const double2 nB = *reinterpret_cast<const double2*>(geo + 16 + lf * 4 + 2);
            const double n0 = nA.x, n1 = nA.y, n2 = nB.x, fs = nB.y;
            const double hf = 0.5 * fs;
            const double cA = hf * (ph.v0[0] * n0 + ph.v0[1] * n1 + ph.v0[2] * n2);  // 1/2 Fscale v0.n
            const double cP = ((flags & FLAG_TAU_NEG) ? -hf : hf) * ph.c0;            // 1/2 Fscale tau c0
            const double cm = cA + cP, cp = cA - cP;
            const double cB = hf * ph.rc2, cR = hf * ph.invRho;
            const double g0 = cB * n0, g1 = cB * n1, g2 = cB * n2;
            const double d0 = cR * n0, d1 = cR * n1, d2 = cR * n2;
            const int bc = flags & FLAG_BC_MASK;
            double qm[2][4];
#pragma unroll
            for (int v = 0; v < 2; ++v) {
                const int own = v == 0 ? own0 : own1;
                qm[v][0] = inRow[C::OFF_P + own];
#pragma unroll
                for (int x = 0; x < 3; ++x) qm[v][1 + x] = stg[x * KQ + own];
            }
#pragma unroll
            for (int v = 0; v < 2; ++v) {
                double fl[4];
                if (bc == FACE_INTERIOR) {
                    const double ps = qm[v][0] + tr[v][0];
                    fl[0] = cm * qm[v][0] + cp * tr[v][0] + g0 * (qm[v][1] + tr[v][1]) + g1 * (qm[v][2] + tr[v][2]) + g2 * (qm[v][3] + tr[v][3]);
                    fl[1] = cm * qm[v][1] + cp * tr[v][1] + d0 * ps;
                    fl[2] = cm * qm[v][2] + cp * tr[v][2] + d1 * ps;
                    fl[3] = cm * qm[v][3] + cp * tr[v][3] + d2 * ps;
                } else {
                    const double n[3] = {n0, n1, n2};
                    faceFlux(bc, 1.0, n, ph, qm[v], tr[v], fl);
#pragma unroll
                    for (int q = 0; q < 4; ++q) fl[q] *= fs;
                }
                if (v == 0 || act1) {
                    double* dst = inRow + C::OFF_F + (v == 0 ? slot0 : slot1);
#pragma unroll
                    for (int q = 0; q < 4; ++q) dst[q * NFL] = fl[q];
                }
            }
        }
    };

    int flagsCur, nbrCur, flagsNext, nbrNext;
    double qv[2][4], tr[2][4];
    loadMeta(eNext, flagsCur, nbrCur);
    issueLoads(eNext, 0, flagsCur, nbrCur, qv, tr);
    asm volatile("cp.async.commit_group;" ::: "memory");
    eNext += eStep;
    loadMeta(eNext, flagsNext, nbrNext);

    for (int it = 0; it < nIt; ++it) {
        const int b = it & (kInStages - 1);
        // tile it: its input slot is free once the MMA warps are done with tile it - kInStages
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        mbarWait(&inEmpty[b], ((it / kInStages) & 1) ^ 1);
        prepAndFlux(sm.in + b * C::IN_TILE + s * LDI, sm.geo + (it & (kGeoStages - 1)) * C::GEO_TILE + s * C::LDG_, flagsCur, qv, tr);
        __syncwarp();
        if (lane == 0) mbarArrive(&full[b]);
        // loads of tile it+1 (the registers are free again), face metadata of tile it+2
        flagsCur = flagsNext; nbrCur = nbrNext;
        if (it + 1 < nIt) {
            issueLoads(eNext, (it + 1) & (kGeoStages - 1), flagsCur, nbrCur, qv, tr);
            asm volatile("cp.async.commit_group;" ::: "memory");
            eNext += eStep;
            loadMeta(eNext, flagsNext, nbrNext);
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------------------------
// back warps: warp q owns field q of every tile (K-split partial sums, velocity combine, fused RK update)
// ------------------------------------------------------------------------------------------------------------------
// k = dt L(y) combined with the RK registers; MODE is a compile-time constant here (see rkApplyK in dgb_device.cuh)
template <int MODE>
__device__ __forceinline__ void rkStore(double* __restrict__ u, double* __restrict__ acc, double* __restrict__ yout, double k, double uval,
                                        double accval) {
    if constexpr (MODE == MODE_RK1) { *acc = k; *yout = fma(0.5, k, uval); }
    else if constexpr (MODE == MODE_RK2) { *acc = fma(2.0, k, accval); *yout = fma(0.5, k, uval); }
    else if constexpr (MODE == MODE_RK3) { *acc = fma(2.0, k, accval); *yout = uval + k; }
    else if constexpr (MODE == MODE_RK4) { *u = fma(accval + k, 1.0 / 6.0, uval); }
    else if constexpr (MODE == MODE_EULER) { *yout = uval + k; }
    else { *yout = k; }
}

template <int P, int MODE>
__device__ __forceinline__ void backWarpImpl(const DeviceMesh& M, const StageArgs& A, const WsSmem& sm, int nIt, int q, int lane) {
    using C = WsCfg<P>;
    constexpr int NP = C::NP, LDO = C::LDO;
    constexpr int PS = kTileEl * LDO;                      // output panel stride
    constexpr int NV = (kTileEl * NP + 31) / 32;           // passes over the 8*NP values of one field of a tile
    constexpr bool needU = MODE != MODE_RHS, needAcc = MODE >= MODE_RK2 && MODE <= MODE_RK4;
    const Phys ph = makePhys(M);
    const int64_t S = M.stride;
    unsigned long long* outFull = sm.bars + 2 * kInStages;
    unsigned long long* outEmpty = sm.bars + 2 * kInStages + kOutStages;
    // per-pass lane constants: value idx = v*32 + lane of the tile <-> (element el, node i)
    int oOff[NV], gOff[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const int idx = min(v * 32 + lane, kTileEl * NP - 1);
        const int el = idx / NP, i = idx - el * NP;
        oOff[v] = el * LDO + i;
        gOff[v] = el * C::LDG_ + (q > 0 ? (q - 1) * 3 : 0);
    }
    const double* uSrc = (MODE == MODE_EULER ? A.yin : A.u) + q * S;
    double* const uDst = A.u + q * S;
    double* const accP = A.acc + q * S;
    double* const youtP = A.yout + q * S;
    const int eStep = (int)gridDim.x * kTileEl;
    int e0 = A.eBegin + (int)blockIdx.x * kTileEl;

    for (int it = 0; it < nIt; ++it, e0 += eStep) {
        const int nVal = min(kTileEl, A.eEnd - e0) * NP;  // valid values of this tile
        const size_t base = (size_t)e0 * NP + lane;
        double uv[NV], av[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const bool ok = v * 32 + lane < nVal;
            uv[v] = (needU && ok) ? uSrc[base + v * 32] : 0.0;
            av[v] = (needAcc && ok) ? accP[base + v * 32] : 0.0;
        }
        const int ob = it & (kOutStages - 1);
        mbarWait(&outFull[ob], (it / kOutStages) & 1);
        const double* out = sm.out + ob * C::OUT_TILE;
        const double* geo = sm.geo + (it & (kGeoStages - 1)) * C::GEO_TILE;
        double k[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const double* o = out + oOff[v];
            k[v] = (o[(3 + q) * PS] + o[(7 + q) * PS]) + (o[(11 + q) * PS] + o[(15 + q) * PS]);
            if (q > 0) {
                const double* Gx = geo + gOff[v];
                k[v] += ph.invRho * (Gx[0] * o[0] + Gx[1] * o[PS] + Gx[2] * o[2 * PS]);
            }
        }
        __syncwarp();
        if (lane == 0) mbarArrive(&outEmpty[ob]);
#pragma unroll
        for (int v = 0; v < NV; ++v)
            if (v * 32 + lane < nVal) rkStore<MODE>(uDst + base + v * 32, accP + base + v * 32, youtP + base + v * 32, k[v], uv[v], av[v]);
    }
}

template <int P>
__device__ __forceinline__ void backWarp(const DeviceMesh& M, const StageArgs& A, const WsSmem& sm, int nIt, int q, int lane) {
    switch (A.mode) {
        case MODE_RK1: backWarpImpl<P, MODE_RK1>(M, A, sm, nIt, q, lane); break;
        case MODE_RK2: backWarpImpl<P, MODE_RK2>(M, A, sm, nIt, q, lane); break;
        case MODE_RK3: backWarpImpl<P, MODE_RK3>(M, A, sm, nIt, q, lane); break;
        case MODE_RK4: backWarpImpl<P, MODE_RK4>(M, A, sm, nIt, q, lane); break;
        case MODE_EULER: backWarpImpl<P, MODE_EULER>(M, A, sm, nIt, q, lane); break;
        default: backWarpImpl<P, MODE_RHS>(M, A, sm, nIt, q, lane); break;
    }
}

template <int P>
__global__ void __launch_bounds__(kThreadsWs, 1) stageWsKernel(DeviceMesh M, StageArgs A, int nTiles) {
    using C = WsCfg<P>;
    extern __shared__ __align__(128) unsigned char smemRaw[];
    WsSmem sm;
    sm.in = reinterpret_cast<double*>(smemRaw);
    sm.stg = sm.in + kInStages * C::IN_TILE;
    sm.geo = sm.stg + C::STG_TILE;
    sm.out = sm.geo + kGeoStages * C::GEO_TILE;
    sm.bars = reinterpret_cast<unsigned long long*>(sm.out + kOutStages * C::OUT_TILE);
    sm.faceNodes = reinterpret_cast<int*>(sm.bars + 16);
    sm.maps = reinterpret_cast<unsigned char*>(sm.faceNodes + C::NFL);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < kInStages; ++i) { mbarInit(&sm.bars[i], kFrontWarps); mbarInit(&sm.bars[kInStages + i], kMmaWarps); }
        for (int i = 0; i < kOutStages; ++i) { mbarInit(&sm.bars[2 * kInStages + i], kMmaWarps); mbarInit(&sm.bars[2 * kInStages + kOutStages + i], kBackWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < C::NFL; i += kThreadsWs) sm.faceNodes[i] = M.faceNodes[i];
    const int nMapsS = min(M.nMaps, kMaxMapsWs);
    for (int i = tid; i < nMapsS * C::NFP; i += kThreadsWs) sm.maps[i] = M.nbrMaps[i];
    // the K padding of the input rows must hold finite values (zeros) for ever
    for (int i = tid; i < kInStages * C::IN_TILE + C::STG_TILE; i += kThreadsWs) sm.in[i] = 0.0;
    __syncthreads();

    const int nIt = ((int)blockIdx.x < nTiles) ? (nTiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    if (warp < kMmaWarps) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsMma));
        if (warp < 3) mmaWarpD<P>(M, A, sm, nIt, warp, lane);
        else mmaWarpL<P>(M, A, sm, nIt, lane);
    } else if (warp < kMmaWarps + kFrontWarps) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsFront));
        frontWarp<P>(M, A, sm, nIt, warp - kMmaWarps, lane);
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsBack));
        backWarp<P>(M, A, sm, nIt, warp - kMmaWarps - kFrontWarps, lane);
    }
}

template <int P>
void launchWs(const DeviceMesh& M, const StageArgs& A, cudaStream_t s) {
    using C = WsCfg<P>;
    const int nEl = A.eEnd - A.eBegin;
    if (nEl <= 0) return;
    static int numSm = 0;
    static bool configured = false;
    const size_t smem = (size_t)C::SMEM_DOUBLES * sizeof(double) + 16 * sizeof(unsigned long long) + C::NFL * sizeof(int) + kMaxMapsWs * C::NFP;
    if (!configured) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&numSm, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(stageWsKernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    const int nTiles = (nEl + kTileEl - 1) / kTileEl;
    const int grid = std::max(1, std::min(numSm, nTiles));
    stageWsKernel<P><<<grid, kThreadsWs, smem, s>>>(M, A, nTiles);
}

}  // namespace

// Zero-mean-flow instances only; the caller keeps the tiled kernel for v0 != 0.
StageKernel selectWsKernel(int dim, int order) {
    StageKernel k;
    if (dim == 3 && order == 4) { k.launch = &launchWs<4>; k.name = "stage_ws_dmma<3,4>"; }
    if (dim == 3 && order == 3) { k.launch = &launchWs<3>; k.name = "stage_ws_dmma<3,3>"; }
    return k;
}

}  // namespace dgb
