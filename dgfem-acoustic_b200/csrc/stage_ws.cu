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
constexpr int kMmaWarps = 4, kSvcWarps = 8, kThreadsWs = (kMmaWarps + kSvcWarps) * 32;
constexpr int kTileEl = 8;       // elements per tile = rows of one m8n8k4 A fragment
constexpr int kInStages = 2, kOutStages = 2, kGeoStages = 3;
constexpr int kRegsMma = 216, kRegsSvc = 144;  // 128*216 + 256*144 = 64512 <= 65536

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
    // first lift k-tile of every MMA role (role r owns [SLICE[r], SLICE[r+1]))
    __host__ __device__ static constexpr int slice(int r) {
        if (P == 4) { constexpr int s[5] = {0, 3, 6, 8, 15}; return s[r]; }
        constexpr int s[5] = {0, 2, 4, 6, 10};
        return s[r];
    }
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
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
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
// MMA warps
// ------------------------------------------------------------------------------------------------------------------
template <int P, int ROLE>
__device__ __forceinline__ void mmaWarp(const DeviceMesh& M, const StageArgs& A, const WsSmem& sm, int nIt, int lane) {
    using C = WsCfg<P>;
    constexpr int NP = C::NP, NT = C::NT, KTQ = C::KTQ, LDI = C::LDI, LDO = C::LDO, NFL = C::NFL;
    constexpr int KT0 = C::slice(ROLE), NS = C::slice(ROLE + 1) - C::slice(ROLE);
    constexpr bool HAS_D = ROLE < 3;
    const int g = lane >> 2, t = lane & 3;
    const double scale = A.mode == MODE_RHS ? 1.0 : A.dt;  // k = dt L(y) leaves the tensor pipe directly

    // register-resident operator slice, as B fragments: lane (g, t) holds Op[node 8*nt + g][k(kt, t)]
    double D[HAS_D ? NT : 1][HAS_D ? KTQ : 1];
    double L[NT][NS];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int i = nt * 8 + g;
        if constexpr (HAS_D) {
#pragma unroll
            for (int kt = 0; kt < KTQ; ++kt) {
                const int k = kOfTile(kt, KTQ, t);
                D[nt][kt] = (i < NP && k < NP) ? scale * M.DwT[((size_t)ROLE * NP + k) * NP + i] : 0.0;
            }
        }
#pragma unroll
        for (int kt = 0; kt < NS; ++kt) {
            const int k = 4 * KT0 + kOfTile(kt, NS, t);
            L[nt][kt] = i < NP ? scale * M.nLiftT[(size_t)C::slotFaceNode(k) * NP + i] : 0.0;
        }
    }

    unsigned long long* full = sm.bars;
    unsigned long long* inEmpty = sm.bars + kInStages;
    unsigned long long* outFull = sm.bars + 2 * kInStages;
    unsigned long long* outEmpty = sm.bars + 2 * kInStages + kOutStages;

    for (int it = 0; it < nIt; ++it) {
        const int b3 = it % kInStages, b2 = it % kOutStages;
        const double* row = sm.in + b3 * C::IN_TILE + g * LDI;
        double* out = sm.out + b2 * C::OUT_TILE + g * LDO + 2 * t;  // + panel * 8 * LDO + nt * 8
        mbarWait(&full[b3], (it / kInStages) & 1);
        mbarWait(&outEmpty[b2], ((it / kOutStages) & 1) ^ 1);

        auto store = [&](int panel, const double (&acc)[NT][2]) {
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
                *reinterpret_cast<double2*>(out + panel * (kTileEl * LDO) + nt * 8) = make_double2(acc[nt][0], acc[nt][1]);
        };

        if constexpr (HAS_D) {
            // T^r = Dw^r p, then P_r = Dw^r (rho c^2 c_r) + (-LIFT slice) F_p; the fragments of the next job are requested
            // before the DMMAs of the current one
            double aP[KTQ], aC[KTQ], aF[NS];
            loadFrags<KTQ>(aP, row + C::OFF_P, t);
            loadFrags<KTQ>(aC, row + C::OFF_C + ROLE * C::KQ, t);
            double accT[NT][2], accP[NT][2];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) accT[nt][0] = accT[nt][1] = accP[nt][0] = accP[nt][1] = 0.0;
#pragma unroll
            for (int kt = 0; kt < KTQ; ++kt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) dmma(accT[nt], aP[kt], D[nt][kt]);
            loadFrags<NS>(aF, row + C::OFF_F + 4 * KT0, t);
#pragma unroll
            for (int kt = 0; kt < KTQ; ++kt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) dmma(accP[nt], aC[kt], D[nt][kt]);
            store(ROLE, accT);
            double aV[NS];
            loadFrags<NS>(aV, row + C::OFF_F + NFL + 4 * KT0, t);
#pragma unroll
            for (int kt = 0; kt < NS; ++kt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) dmma(accP[nt], aF[kt], L[nt][kt]);
            store(3 + ROLE * 4, accP);
#pragma unroll
            for (int x = 0; x < 3; ++x) {
                double accV[NT][2];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) accV[nt][0] = accV[nt][1] = 0.0;
                double aN[NS];
                if (x < 2) loadFrags<NS>(aN, row + C::OFF_F + (x + 2) * NFL + 4 * KT0, t);
#pragma unroll
                for (int kt = 0; kt < NS; ++kt)
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) dmma(accV[nt], aV[kt], L[nt][kt]);
                store(3 + ROLE * 4 + 1 + x, accV);
                if (x < 2) {
#pragma unroll
                    for (int kt = 0; kt < NS; ++kt) aV[kt] = aN[kt];
                }
            }
        } else {
            double aV[NS];
            loadFrags<NS>(aV, row + C::OFF_F + 4 * KT0, t);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                double acc[NT][2];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
                double aN[NS];
                if (q < 3) loadFrags<NS>(aN, row + C::OFF_F + (q + 1) * NFL + 4 * KT0, t);
#pragma unroll
                for (int kt = 0; kt < NS; ++kt)
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) dmma(acc[nt], aV[kt], L[nt][kt]);
                store(3 + ROLE * 4 + q, acc);
                if (q < 3) {
#pragma unroll
                    for (int kt = 0; kt < NS; ++kt) aV[kt] = aN[kt];
                }
            }
        }
        __syncwarp();
        if (lane == 0) {
            mbarArrive(&inEmpty[b3]);
            mbarArrive(&outFull[b2]);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// service warps
// ------------------------------------------------------------------------------------------------------------------
// k = dt L(y) combined with the RK registers; MODE is a compile-time constant here (see rkApplyK in dgb_device.cuh)
template <int MODE>
__device__ __forceinline__ void rkStore(const StageArgs& A, int64_t gi, double k, double uval, double accval) {
    if constexpr (MODE == MODE_RK1) { A.acc[gi] = k; A.yout[gi] = fma(0.5, k, uval); }
    else if constexpr (MODE == MODE_RK2) { A.acc[gi] = fma(2.0, k, accval); A.yout[gi] = fma(0.5, k, uval); }
    else if constexpr (MODE == MODE_RK3) { A.acc[gi] = fma(2.0, k, accval); A.yout[gi] = uval + k; }
    else if constexpr (MODE == MODE_RK4) { A.u[gi] = fma(accval + k, 1.0 / 6.0, uval); }
    else if constexpr (MODE == MODE_EULER) { A.yout[gi] = uval + k; }
    else { A.yout[gi] = k; }
}

template <int P>
__device__ __forceinline__ void serviceWarp(const DeviceMesh& M, const StageArgs& A, const WsSmem& sm, int nIt, int s, int lane) {
    using C = WsCfg<P>;
    constexpr int NP = C::NP, NFP = C::NFP, NF = C::NF, NFL = C::NFL, KQ = C::KQ, LDI = C::LDI, LDO = C::LDO;
    constexpr int NA = NP < 32 ? NP : 32, NB = NP - NA;  // nodes handled by lane j (pass A) and by lanes 0..NB-1 as node 32+lane (pass B)
    constexpr int NFB = NFP - 8;                         // face nodes 8.. of a face are handled in the second flux pass
    constexpr int PS = kTileEl * LDO;                    // output panel stride
    static_assert(NB >= 0 && 4 * NB <= 32 && NFB > 0 && NFB <= 8, "lane mappings assume 20 <= NP <= 40, 8 < NFP <= 16");
    const Phys ph = makePhys(M);
    const int64_t S = M.stride;
    // flux lanes: face lf, face node ml (first pass) and 8 + ml (second pass, ml < NFB); all per-lane constants
    const int lf = lane >> 3, ml = lane & 7;
    const bool act1 = ml < NFB;
    const int own0 = sm.faceNodes[lf * NFP + ml], own1 = sm.faceNodes[lf * NFP + (act1 ? 8 + ml : 0)];
    const int slot0 = C::faceSlot(lf, ml), slot1 = C::faceSlot(lf, act1 ? 8 + ml : 8);
    // epilogue pass B lanes: (field qB, node iB)
    const int qB = NB > 0 ? min(lane / (NB > 0 ? NB : 1), 3) : 0, iB = NB > 0 ? 32 + lane % (NB > 0 ? NB : 1) : 0;
    const bool actB = NB > 0 && lane < 4 * NB;
    const bool actA = lane < NA, actQB = lane < NB;

    unsigned long long* full = sm.bars;
    unsigned long long* inEmpty = sm.bars + kInStages;
    unsigned long long* outFull = sm.bars + 2 * kInStages;
    unsigned long long* outEmpty = sm.bars + 2 * kInStages + kOutStages;
    double* const stg = sm.stg + s * C::LDS_;

    auto elemOf = [&](int it) { return A.eBegin + (int)(blockIdx.x + (unsigned)it * gridDim.x) * kTileEl + s; };
    auto loadMeta = [&](int it, int& flags, int& nbr) {
        flags = FACE_ABSORBING;
        nbr = -1;
        if (it < nIt) {
            const int e = elemOf(it);
            if (e < A.eEnd) { flags = M.fflags[e * NF + lf]; nbr = M.fnbr[e * NF + lf]; }
        }
    };
    // Everything element s of tile `it` needs: nodal values and neighbour traces into registers, geometry into shared memory
    auto issueLoads = [&](int it, int flags, int nbr, double (&qv)[2][4], double (&tr)[2][4]) {
        const int e = elemOf(it);
        const bool valid = e < A.eEnd;
        const double* src = A.yin + (int64_t)(valid ? e : 0) * NP + lane;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            qv[0][q] = (valid && actA) ? src[q * S] : 0.0;
            qv[1][q] = (valid && actQB) ? src[q * S + 32] : 0.0;
        }
        const bool interior = valid && ((flags & FLAG_BC_MASK) == FACE_INTERIOR) && nbr >= 0;
        const int mapId = flags >> FLAG_MAP_SHIFT;
        int nn0 = 0, nn1 = 0;
        if (interior) {
            const unsigned char* mp = mapId < kMaxMapsWs ? sm.maps + mapId * NFP : M.nbrMaps + mapId * NFP;
            nn0 = mp[ml];
            nn1 = mp[act1 ? 8 + ml : 0];
        }
        const double* t0 = A.yin + (int64_t)(interior ? nbr : 0) * NP;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            tr[0][q] = interior ? t0[q * S + nn0] : 0.0;
            tr[1][q] = (interior && act1) ? t0[q * S + nn1] : 0.0;
        }
        double* geo = sm.geo + (it % kGeoStages) * C::GEO_TILE + s * C::LDG_;
        if (lane < 9) cpAsync8z(&geo[lane], M.Ginv + (int64_t)(valid ? e : 0) * 9 + lane, valid);
        else if (lane >= 16) cpAsync8z(&geo[lane], M.fgeo + (int64_t)(valid ? e : 0) * 16 + (lane - 16), valid);
    };

    // contravariant velocities + numerical flux of tile `it` into the input row of element s
    auto prepAndFlux = [&](int it, int flags, const double (&qv)[2][4], const double (&tr)[2][4]) {
        double* inRow = sm.in + (it % kInStages) * C::IN_TILE + s * LDI;
        const double* geo = sm.geo + (it % kGeoStages) * C::GEO_TILE + s * C::LDG_;
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

    const bool needU = A.mode != MODE_RHS;
    const bool needAcc = A.mode >= MODE_RK2 && A.mode <= MODE_RK4;
    const double* uSrc = A.mode == MODE_EULER ? A.yin : A.u;
    // RK registers of element s of tile `it`: [0..3] node `lane` of the four fields, [4] the pass-B entry of this lane
    auto loadRk = [&](int it, double (&uv)[5], double (&av)[5]) {
        const int e = elemOf(it);
        const bool valid = e < A.eEnd;
        const int64_t base = (int64_t)(valid ? e : 0) * NP;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const bool ok = valid && actA;
            uv[q] = (ok && needU) ? uSrc[base + q * S + lane] : 0.0;
            av[q] = (ok && needAcc) ? A.acc[base + q * S + lane] : 0.0;
        }
        const bool okB = valid && actB;
        uv[4] = (okB && needU) ? uSrc[base + qB * S + iB] : 0.0;
        av[4] = (okB && needAcc) ? A.acc[base + qB * S + iB] : 0.0;
    };
    // K-split partial sums + velocity combine + fused RK update of element s of tile `it`
    auto epilogueImpl = [&](auto modeTag, int it, const double (&uv)[5], const double (&av)[5]) {
        constexpr int MODE = decltype(modeTag)::value;
        const int e = elemOf(it);
        const bool valid = e < A.eEnd;
        const double* out = sm.out + (it % kOutStages) * C::OUT_TILE + s * LDO;
        const double* geo = sm.geo + (it % kGeoStages) * C::GEO_TILE + s * C::LDG_;
        const int64_t base = (int64_t)(valid ? e : 0) * NP;
        if (actA) {
            const double* o = out + lane;
            double k[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) k[q] = (o[(3 + q) * PS] + o[(7 + q) * PS]) + (o[(11 + q) * PS] + o[(15 + q) * PS]);
            const double T0 = o[0], T1 = o[PS], T2 = o[2 * PS];
#pragma unroll
            for (int x = 0; x < 3; ++x) k[1 + x] += ph.invRho * (geo[x * 3] * T0 + geo[x * 3 + 1] * T1 + geo[x * 3 + 2] * T2);
            if (valid) {
#pragma unroll
                for (int q = 0; q < 4; ++q) rkStore<MODE>(A, base + q * S + lane, k[q], uv[q], av[q]);
            }
        }
        if (NB > 0 && actB) {
            const double* o = out + iB;
            double k = (o[(3 + qB) * PS] + o[(7 + qB) * PS]) + (o[(11 + qB) * PS] + o[(15 + qB) * PS]);
            if (qB > 0) {
                const double* Gx = geo + (qB - 1) * 3;
                k += ph.invRho * (Gx[0] * o[0] + Gx[1] * o[PS] + Gx[2] * o[2 * PS]);
            }
            if (valid) rkStore<MODE>(A, base + qB * S + iB, k, uv[4], av[4]);
        }
    };
    auto epilogue = [&](int it, const double (&uv)[5], const double (&av)[5]) {
        switch (A.mode) {
            case MODE_RK1: epilogueImpl(std::integral_constant<int, MODE_RK1>{}, it, uv, av); break;
            case MODE_RK2: epilogueImpl(std::integral_constant<int, MODE_RK2>{}, it, uv, av); break;
            case MODE_RK3: epilogueImpl(std::integral_constant<int, MODE_RK3>{}, it, uv, av); break;
            case MODE_RK4: epilogueImpl(std::integral_constant<int, MODE_RK4>{}, it, uv, av); break;
            case MODE_EULER: epilogueImpl(std::integral_constant<int, MODE_EULER>{}, it, uv, av); break;
            default: epilogueImpl(std::integral_constant<int, MODE_RHS>{}, it, uv, av); break;
        }
    };

    int flagsCur, nbrCur, flagsNext, nbrNext;
    double qv[2][4], tr[2][4], uv[5], av[5];
    loadMeta(0, flagsCur, nbrCur);
    issueLoads(0, flagsCur, nbrCur, qv, tr);
    asm volatile("cp.async.commit_group;" ::: "memory");
    loadMeta(1, flagsNext, nbrNext);

    for (int it = 0; it < nIt; ++it) {
        const int b = it % kInStages;
        // 1. tile it: its input slot is free once the MMA warps are done with tile it - kInStages
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        mbarWait(&inEmpty[b], ((it / kInStages) & 1) ^ 1);
        prepAndFlux(it, flagsCur, qv, tr);
        __syncwarp();
        if (lane == 0) mbarArrive(&full[b]);
        // 2. loads of tile it+1 (registers are free again), face metadata of tile it+2
        flagsCur = flagsNext; nbrCur = nbrNext;
        if (it + 1 < nIt) issueLoads(it + 1, flagsCur, nbrCur, qv, tr);
        asm volatile("cp.async.commit_group;" ::: "memory");
        loadMeta(it + 2, flagsNext, nbrNext);
        // 3. tile it-1 leaves: partial sums -> RK update
        if (it >= 1) {
            const int ob = (it - 1) % kOutStages;
            mbarWait(&outFull[ob], ((it - 1) / kOutStages) & 1);
            epilogue(it - 1, uv, av);
            __syncwarp();
            if (lane == 0) mbarArrive(&outEmpty[ob]);
        }
        loadRk(it, uv, av);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (nIt >= 1) {
        const int ob = (nIt - 1) % kOutStages;
        mbarWait(&outFull[ob], ((nIt - 1) / kOutStages) & 1);
        epilogue(nIt - 1, uv, av);
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
        for (int i = 0; i < kInStages; ++i) { mbarInit(&sm.bars[i], kSvcWarps); mbarInit(&sm.bars[kInStages + i], kMmaWarps); }
        for (int i = 0; i < kOutStages; ++i) { mbarInit(&sm.bars[2 * kInStages + i], kMmaWarps); mbarInit(&sm.bars[2 * kInStages + kOutStages + i], kSvcWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < C::NFL; i += kThreadsWs) sm.faceNodes[i] = M.faceNodes[i];
    const int nMapsS = min(M.nMaps, kMaxMapsWs);
    for (int i = tid; i < nMapsS * C::NFP; i += kThreadsWs) sm.maps[i] = M.nbrMaps[i];
    // the K padding of the input rows and of the staging area must hold finite values (zeros) for ever
    for (int i = tid; i < kInStages * C::IN_TILE + C::STG_TILE; i += kThreadsWs) sm.in[i] = 0.0;
    __syncthreads();

    const int nIt = ((int)blockIdx.x < nTiles) ? (nTiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    if (warp < kMmaWarps) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsMma));
        if (warp == 0) mmaWarp<P, 0>(M, A, sm, nIt, lane);
        else if (warp == 1) mmaWarp<P, 1>(M, A, sm, nIt, lane);
        else if (warp == 2) mmaWarp<P, 2>(M, A, sm, nIt, lane);
        else mmaWarp<P, 3>(M, A, sm, nIt, lane);
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsSvc));
        serviceWarp<P>(M, A, sm, nIt, warp - kMmaWarps, lane);
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
