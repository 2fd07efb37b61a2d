// Warp-specialised FP64 tensor-core (DMMA) stage kernel for tetrahedra of order 3 and 4, zero mean flow.
//
// Same fused operator as stage_generic.cu / stage_tiled.cu (updateFlux + numStep + RK axpys of the reference,
// Mesh.cpp:476-674, solver.cpp:35-52, 261-285). What changes is who does what inside the persistent CTA of 16 warps
// that works through tiles of 8 consecutive elements:
//
//   * 4 MMA warps, one per SM sub-partition. Each keeps ITS SLICE OF THE OPERATORS IN REGISTERS for the whole launch, so
//     the tensor pipe is fed with one shared-memory load per ~5 DMMAs instead of one per DMMA: the element data are the
//     A fragments (8 elements x 4 k), the register-resident operator the B fragments (4 k x 8 nodes), the accumulators
//     are (8 elements x 8 nodes).
//       role D_r, r = 0..2 :  T^r    = Dw^r (p / rho0)                            (feeds the three velocity equations)
//                             S_p   += Dw^r c_r - LIFT[:, slice r] F_p / (rho0 c0^2)        c_r = sum_x G_xr v_x
//                             S_vx  +=          - LIFT[:, slice r] F_vx                     x = 0..2
//       role L             :  S_q   += - LIFT[:, last slice] F_q
//     The K-split partial sums are chained in pairs through the DMMA accumulator input (L -> D0 and D1 -> D2: the head
//     of a pair starts from zero, the tail loads the head's sums as its initial accumulators): an FP64 add inside a DMMA
//     is free, an FP64 add in another warp is not, because DFMA/DADD and DMMA share one pipe
//     (profiles/microbench/r01_fp64_mix_b200.txt). A tile is ONE straight-line block of code per MMA warp with a single
//     wait in front: ptxas hides the loads and stores between the DMMAs only inside a basic block.
//   * 8 front warps, warp s owns element s of every tile: cp.async ring for its nodal values and geometry, neighbour
//     traces through the face-node maps into registers (both two tiles ahead), contravariant velocities and numerical
//     flux into the input ring.
//   * 4 back warps, warp q owns field q of every tile: adds the two chain sums, combines the velocity equations with the
//     inverse Jacobian, fused RK update with fully coalesced global accesses.
//   * mbarrier hand-off everywhere (input ring, chain, output ring); no __syncthreads in the steady state. Registers
//     are re-balanced with setmaxnreg (MMA warps up, the others down).
//
// Shared memory rows are K-contiguous per element with a leading dimension == 8 (mod 16) doubles: the 128-bit fragment
// loads (two k-tiles per load; the operator registers are permuted to match) and the 128-bit accumulator loads/stores
// are bank-conflict free.
#include <algorithm>
#include <type_traits>

#include "dgb_device.cuh"
#include "dgb_internal.h"
#include "dgb_launch.h"

namespace dgb {

namespace {

constexpr int kMaxMapsWs = 48;
constexpr int kMmaWarps = 4, kFrontWarps = 8, kBackWarps = 4, kThreadsWs = (kMmaWarps + kFrontWarps + kBackWarps) * 32;
constexpr int kTileEl = 8;  // elements per tile = rows of one m8n8k4 A fragment
// Ring depths. Input and output rings are 3 deep because the head of a chain pair runs one tile ahead of its tail; the
// inverse-Jacobian ring (a power of two) must outlive the back warps' lag behind the front warps (<= 8 tiles); the
// face-geometry and nodal-value rings only serve the front warps (prefetch distance 2). The total is kept below 195 KB so
// that the SM is configured with the 196 KB carve-out and ~60 KB of L1 remain for the neighbour-trace gathers.
#ifndef DGB_WS_OUTSTAGES
#define DGB_WS_OUTSTAGES 3
#endif
#ifndef DGB_WS_STGSTAGES
#define DGB_WS_STGSTAGES 3
#endif
constexpr int kInStages = 3, kOutStages = DGB_WS_OUTSTAGES, kGeoStages = 8, kFgStages = 4, kStgStages = DGB_WS_STGSTAGES;
constexpr int kOwnAhead = kStgStages - 1;  // prefetch distance (tiles) of the nodal values / geometry
// setmaxnreg only moves registers inside the CTA's launch allocation: 512 threads x 128 registers = 65536
constexpr int kRegsLaunch = 128, kRegsMma = 184, kRegsFront = 104, kRegsBack = 120;
static_assert(128 * kRegsMma + 256 * kRegsFront + 128 * kRegsBack <= kThreadsWs * kRegsLaunch, "register budget of the CTA");

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
    // lift k-tiles owned by each of the MMA roles D_0..2 (role r: [r*NSD, (r+1)*NSD)); role L owns the rest
#ifndef DGB_WS_NSD4
#define DGB_WS_NSD4 3
#endif
    static constexpr int NSD = P == 4 ? DGB_WS_NSD4 : 2;
    // input row of one element (doubles): p / rho0 | c_0 | c_1 | c_2 | F_p / (rho0 c0^2) | F_vx | F_vy | F_vz
    static constexpr int OFF_P = 0, OFF_C = KQ, OFF_F = 4 * KQ;
    static constexpr int LDI = padTo8mod16(4 * KQ + 4 * NFL);
    static constexpr int LDS_ = 4 * KQ;           // raw nodal values [4][KQ] of one element (cp.async ring)
    static constexpr int LDG_ = 10;               // inverse Jacobian Ginv[9] of one element
    static constexpr int LDFG = 16;               // face geometry of one element: 4 x (normal, Fscale)
    static constexpr int LDO = padTo8mod16(NPP);  // output row of one element in one panel
    static constexpr int NPANEL = 3 + 8;          // T^0..2, sums of chain a (4 fields), sums of chain b (4 fields)
    static constexpr int IN_TILE = kTileEl * LDI, STG_TILE = kTileEl * LDS_, GEO_TILE = kTileEl * LDG_, FG_TILE = kTileEl * LDFG;
    static constexpr int OUT_TILE = NPANEL * kTileEl * LDO;
    static constexpr int SMEM_DOUBLES = kInStages * IN_TILE + kStgStages * STG_TILE + kGeoStages * GEO_TILE + kFgStages * FG_TILE + kOutStages * OUT_TILE;
    // Position of face node (lf, m) inside a flux block: the first 8 nodes of the four faces, then the remaining ones, so
    // that the 8-lanes-per-face mapping of the flux phase touches consecutive addresses (the lift operator columns are
    // permuted to match when they are loaded into registers)
    __host__ __device__ static constexpr int faceSlot(int lf, int m) { return m < 8 ? lf * 8 + m : 32 + lf * (NFP - 8) + (m - 8); }
    __host__ __device__ static constexpr int slotFaceNode(int k) {  // inverse: lift column lf*NFP + m of slot k
        return k < 32 ? (k >> 3) * NFP + (k & 7) : ((k - 32) / (NFP - 8)) * NFP + 8 + (k - 32) % (NFP - 8);
    }
    static_assert(NFL % 4 == 0, "lift contraction length must be a multiple of 4");
    // everything the CTA asks for: rings + mbarriers + face-node table + face-pairing byte maps
    static constexpr size_t SMEM_BYTES = (size_t)SMEM_DOUBLES * 8 + 32 * 8 + NFL * 4 + kMaxMapsWs * NFP;
    // 196 KB carve-out minus the 1 KB the system reserves per CTA: above it the SM falls back to the 228 KB carve-out and
    // the kernel loses half of its L1 (measured: 2.65 -> 3.15 ms per stage, profiles/r01_ws_ablation.txt)
    static_assert(SMEM_BYTES <= 195 * 1024, "shared memory must stay under the 196 KB carve-out");
};

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ uint32_t sAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cpAsync8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sAddr(dst)), "l"(src) : "memory");
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

// Global accesses of the service warps. The L1 and the shared memory are one data array, and what is left of it as L1
// (~60 KB with the 196 KB carve-out this kernel fits under) is what serves the 8-byte neighbour-trace gathers: measured on
// B200, a build with 230 KB of shared memory (28 KB of L1) is 19 % slower, and gathers that bypass L1 (DGB_WS_LDHINT=2) are
// 22 % slower (profiles/r01_ws_ablation.txt). Streaming data (RK registers, results) is therefore marked evict-first
// (DGB_WS_LDHINT >= 1), the gathers keep the default caching.
#ifndef DGB_WS_LDHINT
#define DGB_WS_LDHINT 1
#endif
__device__ __forceinline__ double ldStream(const double* p) {
#if DGB_WS_LDHINT >= 1
    double v;
    asm volatile("ld.global.cs.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
#else
    return *p;
#endif
}
__device__ __forceinline__ void stStream(double* p, double v) {
#if DGB_WS_LDHINT >= 1
    asm volatile("st.global.cs.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
#else
    *p = v;
#endif
}
__device__ __forceinline__ double ldGather(const double* p) {
#if DGB_WS_LDHINT >= 2
    double v;
    asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
#else
    return *p;
#endif
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
    double* stg;   // [kStgStages][8][4][KQ]
    double* geo;   // [kGeoStages][8][LDG_]
    double* fg;    // [kFgStages][8][16]
    double* out;   // [kOutStages][NPANEL][8][LDO]
    int* faceNodes;
    unsigned char* maps;
    unsigned long long* bars;  // see WsBars
};

// Barrier block: full[kInStages] (front -> MMA), inEmpty[kInStages] (MMA -> front), chain[2][kOutStages] (head -> tail of
// the pairs L -> D0 and D1 -> D2), outFull[kOutStages] (both tails -> back), outEmpty[kOutStages] (back -> both heads)
struct WsBars {
    unsigned long long *full, *inEmpty, *chain, *outFull, *outEmpty;
};
__device__ __forceinline__ WsBars wsBars(const WsSmem& sm) {
    unsigned long long* b = sm.bars;
    return {b, b + kInStages, b + 2 * kInStages, b + 2 * kInStages + 2 * kOutStages, b + 2 * kInStages + 3 * kOutStages};
}
constexpr int kNumBars = 2 * kInStages + 4 * kOutStages;
// Tiles of a CTA: chunks of DGB_WS_CHUNK consecutive tiles, the chunks interleaved over the grid (chunk j goes to CTA
// j % gridDim). Interleaving keeps the SMs on neighbouring parts of the mesh at the same time (L2 hits for the neighbour
// traces; one contiguous range per CTA measured 12 % slower); consecutive tiles on one SM find their neighbours in its L1.
#ifndef DGB_WS_CHUNK
#define DGB_WS_CHUNK 1
#endif
struct TileMap {
    int b, G, count;  // CTA, grid size, number of tiles of this CTA
    // element offset (relative to the first element of the launch) of the it-th tile of this CTA; may run past the end
    __device__ __forceinline__ int elem(int it) const {
        constexpr int c = DGB_WS_CHUNK;
        return (c == 1 ? it * G + b : ((it / c) * G + b) * c + it % c) * kTileEl;
    }
};
__device__ __forceinline__ TileMap tileMap(int nTiles) {
    constexpr int c = DGB_WS_CHUNK;
    const int b = (int)blockIdx.x, G = (int)gridDim.x;
    const int nChunks = (nTiles + c - 1) / c;
    int count = 0;
    if (b < nChunks) {
        const int mine = (nChunks - b + G - 1) / G;          // chunks of this CTA
        const int last = b + (mine - 1) * G;                 // its last chunk may be a partial one
        count = (mine - 1) * c + min(c, nTiles - last * c);
    }
    return {b, G, count};
}
// position and phase parity in an N-deep ring (N need not be a power of two)
template <int N>
struct Ring {
    int b = 0;
    uint32_t phase = 0;
    __device__ __forceinline__ void next() { if (++b == N) { b = 0; phase ^= 1; } }
};

// ------------------------------------------------------------------------------------------------------------------
// MMA warps. Roles D_0..2 run the SAME instruction stream (only register contents and a few offsets differ): the
// per-SM instruction working set has to stay inside the 32 KB L1.5 instruction cache (a DMMA costs 32 B of code with
// its pacing NOP).
// ------------------------------------------------------------------------------------------------------------------
template <int P>
__device__ __forceinline__ void mmaWarpD(const DeviceMesh& M, const StageArgs& A, const WsSmem& sm, int nIt, int role, int lane) {
    using C = WsCfg<P>;
    constexpr int NP = C::NP, NT = C::NT, KTQ = C::KTQ, LDI = C::LDI, LDO = C::LDO, NFL = C::NFL, NS = C::NSD;
    constexpr int PS = kTileEl * LDO;
    const int g = lane >> 2, t = lane & 3;
    const double scale = A.mode == MODE_RHS ? 1.0 : A.dt;  // k = dt L(y) leaves the tensor pipe directly
    const int kt0 = role * NS;
    const bool tail = role != 1;  // D0 continues L's sums (chain a), D2 continues D1's (chain b); D1 starts chain b

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
    const WsBars bar = wsBars(sm);
    // what this role waits for before it may touch the output slot, and whom it tells when its sums are in place
    unsigned long long* const waitBar = role == 0 ? bar.chain : role == 1 ? bar.outEmpty : bar.chain + kOutStages;
    unsigned long long* const doneBar = role == 1 ? bar.chain + kOutStages : bar.outFull;
    const int rowOff = g * LDI, offC = C::OFF_C + role * C::KQ, offF = C::OFF_F + 4 * kt0;
    const int outOff = g * LDO + 2 * t, panelT = role * PS, panelS = (role == 0 ? 3 : 7) * PS;

    Ring<kInStages> in;
    Ring<kOutStages> ou;
    for (int it = 0; it < nIt; ++it, in.next(), ou.next()) {
        const double* row = sm.in + in.b * C::IN_TILE + rowOff;
        double* out = sm.out + ou.b * C::OUT_TILE + outOff;
        double* sum = out + panelS;
        mbarWait(&bar.full[in.b], in.phase);
        mbarWait(&waitBar[ou.b], role == 1 ? ou.phase ^ 1 : ou.phase);

        // ---- one basic block: 150 DMMAs; ptxas schedules the fragment / accumulator loads and the stores between them ----
#ifndef DGB_WS_NOMMA
        auto initAcc = [&](double (&acc)[NT][2], const double* src) {  // the head of a chain starts from zero
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                double2 v = make_double2(0.0, 0.0);
                if (tail) v = *reinterpret_cast<const double2*>(src + nt * 8);
                acc[nt][0] = v.x;
                acc[nt][1] = v.y;
            }
        };
        auto store = [&](double* dst, const double (&acc)[NT][2]) {
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) *reinterpret_cast<double2*>(dst + nt * 8) = make_double2(acc[nt][0], acc[nt][1]);
        };
        double aP[KTQ], aC[KTQ];
        loadFrags<KTQ>(aP, row + C::OFF_P, t);
        loadFrags<KTQ>(aC, row + offC, t);
        {   // T^r
            double acc[NT][2];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
#pragma unroll
            for (int kt = 0; kt < KTQ; ++kt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) dmma(acc[nt], aP[kt], D[nt][kt]);
            store(out + panelT, acc);
        }
        {   // S_p
            double acc[NT][2], aV[NS];
            initAcc(acc, sum);
            loadFrags<NS>(aV, row + offF, t);
#pragma unroll
            for (int kt = 0; kt < KTQ; ++kt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) dmma(acc[nt], aC[kt], D[nt][kt]);
#pragma unroll
            for (int kt = 0; kt < NS; ++kt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) dmma(acc[nt], aV[kt], L[nt][kt]);
            store(sum, acc);
        }
#pragma unroll
        for (int x = 1; x < 4; ++x) {  // S_vx
            double acc[NT][2], aV[NS];
            initAcc(acc, sum + x * PS);
            loadFrags<NS>(aV, row + offF + x * NFL, t);
#pragma unroll
            for (int kt = 0; kt < NS; ++kt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) dmma(acc[nt], aV[kt], L[nt][kt]);
            store(sum + x * PS, acc);
        }
#endif
        __syncwarp();
        if (lane == 0) {
            mbarArrive(&bar.inEmpty[in.b]);
            mbarArrive(&doneBar[ou.b]);
        }
    }
}

// role L: starts chain a, S_q = (-LIFT, last slice) F_q
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
    const WsBars bar = wsBars(sm);
    const int rowOff = g * LDI + C::OFF_F + 4 * KT0;
    Ring<kInStages> in;
    Ring<kOutStages> ou;
    for (int it = 0; it < nIt; ++it, in.next(), ou.next()) {
        const double* row = sm.in + in.b * C::IN_TILE + rowOff;
        double* sum = sm.out + ou.b * C::OUT_TILE + g * LDO + 2 * t + 3 * PS;
        mbarWait(&bar.full[in.b], in.phase);
        mbarWait(&bar.outEmpty[ou.b], ou.phase ^ 1);
#ifndef DGB_WS_NOMMA
        double aV[2][NS];
        loadFrags<NS>(aV[0], row, t);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (q < 3) loadFrags<NS>(aV[(q + 1) & 1], row + (q + 1) * NFL, t);
            double acc[NT][2];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
#pragma unroll
            for (int kt = 0; kt < NS; ++kt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) dmma(acc[nt], aV[q & 1][kt], L[nt][kt]);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) *reinterpret_cast<double2*>(sum + q * PS + nt * 8) = make_double2(acc[nt][0], acc[nt][1]);
        }
#endif
        __syncwarp();
        if (lane == 0) {
            mbarArrive(&bar.inEmpty[in.b]);
            mbarArrive(&bar.chain[ou.b]);
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
    constexpr int NA = NP < 32 ? NP : 32, NB = NP - NA;  // nodes handled by lane j (pass A); the NB remaining ones in pass B
    constexpr int NFB = NFP - 8;                         // face nodes 8.. of a face are handled in the second flux pass
    static_assert(NB >= 0 && 4 * NB <= 32 && NFB > 0 && NFB <= 8, "lane mappings assume NP <= 40, 8 < NFP <= 16");
    const Phys ph = makePhys(M);
    const int64_t S = M.stride;
    // flux lanes: face lf, face node ml (first pass) and 8 + ml (second pass, ml < NFB); all per-lane constants
    const int lf = lane >> 3, ml = lane & 7;
    const bool act1 = ml < NFB;
    const int own0 = sm.faceNodes[lf * NFP + ml], own1 = sm.faceNodes[lf * NFP + (act1 ? 8 + ml : 0)];
    const int slot0 = C::faceSlot(lf, ml), slot1 = C::faceSlot(lf, act1 ? 8 + ml : 8);
    const bool actA = lane < NA;
    // pass B lanes: copies move (field lane / NB, node 32 + lane % NB); the contravariant velocities use (u = lane / NB, same node)
    const bool actB = NB > 0 && lane < 4 * NB, actCB = NB > 0 && lane < 3 * NB;
    const int qB = NB > 0 ? min(lane / (NB > 0 ? NB : 1), 3) : 0, jB = NB > 0 ? 32 + lane % (NB > 0 ? NB : 1) : 0;
    // the four fields of the stage input
    const double* const y0 = A.yin;
    const double* const y1 = y0 + S;
    const double* const y2 = y1 + S;
    const double* const y3 = y2 + S;
    const double* const yB = y0 + qB * S + jB;
    const double invRc2 = 1.0 / ph.rc2;

    const WsBars bar = wsBars(sm);
    const TileMap tm = tileMap((A.eEnd - A.eBegin + kTileEl - 1) / kTileEl);
    const int eLast = A.eEnd - 1;
    auto elemOf = [&](int it) { return A.eBegin + tm.elem(it) + s; };  // element s of the it-th tile of this CTA

    // Face metadata of element e (clamped: rows past the end of the range are computed on a copy of the last element and
    // never stored)
    auto loadMeta = [&](int e, int& flags, int& nbr) {
        const int ec = min(e, eLast);
        flags = M.fflags[ec * NF + lf];
        nbr = M.fnbr[ec * NF + lf];
    };
    // Own nodal values [4][KQ] and geometry of element e, asynchronously into their rings
    auto issueOwn = [&](int e, int slot) {
        const int ec = min(e, eLast);
        const unsigned off = (unsigned)ec * NP + lane;
        double* stg = sm.stg + (slot % kStgStages) * C::STG_TILE + s * C::LDS_;
        if (actA) {
            cpAsync8(&stg[lane], y0 + off);
            cpAsync8(&stg[KQ + lane], y1 + off);
            cpAsync8(&stg[2 * KQ + lane], y2 + off);
            cpAsync8(&stg[3 * KQ + lane], y3 + off);
        }
        if (actB) cpAsync8(&stg[qB * KQ + jB], yB + (unsigned)ec * NP);
        double* geo = sm.geo + (slot & (kGeoStages - 1)) * C::GEO_TILE + s * C::LDG_;
        double* fg = sm.fg + (slot & (kFgStages - 1)) * C::FG_TILE + s * C::LDFG;
        if (lane < 9) cpAsync8(&geo[lane], M.Ginv + (size_t)ec * 9 + lane);
        else if (lane >= 16) cpAsync8(&fg[lane - 16], M.fgeo + (size_t)ec * 16 + (lane - 16));
    };
    // Neighbour traces of the lane's two face nodes, into registers. Boundary faces read their own element instead (the
    // value is ignored by the boundary fluxes).
    auto loadTraces = [&](int e, int flags, int nbr, double (&tr)[2][4]) {
        const int ec = min(e, eLast);
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
        tr[0][0] = ldGather(y0 + t0); tr[0][1] = ldGather(y1 + t0); tr[0][2] = ldGather(y2 + t0); tr[0][3] = ldGather(y3 + t0);
        tr[1][0] = ldGather(y0 + t1); tr[1][1] = ldGather(y1 + t1); tr[1][2] = ldGather(y2 + t1); tr[1][3] = ldGather(y3 + t1);
    };

    // contravariant velocities + numerical flux into the input row of element s
    auto prepAndFlux = [&](double* inRow, const double* stg, const double* geo, const double* fg, int flags, const double (&tr)[2][4]) {
        {
            // p / rho0 and c_u = sum_x G_xu v_x   (Ginv[x*3+u] = du_u/dx_x)
            if (actA) {
                const double vx = stg[KQ + lane], vy = stg[2 * KQ + lane], vz = stg[3 * KQ + lane];
                inRow[C::OFF_P + lane] = ph.invRho * stg[lane];
#pragma unroll
                for (int u = 0; u < 3; ++u) inRow[C::OFF_C + u * KQ + lane] = geo[u] * vx + geo[3 + u] * vy + geo[6 + u] * vz;
            }
            if (NB > 0) {
                if (actCB) inRow[C::OFF_C + qB * KQ + jB] = geo[qB] * stg[KQ + jB] + geo[3 + qB] * stg[2 * KQ + jB] + geo[6 + qB] * stg[3 * KQ + jB];
                else if (actB) inRow[C::OFF_P + jB] = ph.invRho * stg[jB];
            }
        }
        {
            const double2 nA = *reinterpret_cast<const double2*>(fg + lf * 4);
            const double2 nB = *reinterpret_cast<const double2*>(fg + lf * 4 + 2);
            const double n0 = nA.x, n1 = nA.y, n2 = nB.x, fs = nB.y;
            const double hf = 0.5 * fs;
            // the pressure flux is divided by rho0 c0^2 here (the back warps multiply the finished pressure sum by it)
            const double cP = ((flags & FLAG_TAU_NEG) ? -hf : hf) * ph.c0;  // 1/2 Fscale tau c0
            const double cPp = cP * invRc2;
            const double g0 = hf * n0, g1 = hf * n1, g2 = hf * n2;          // 1/2 Fscale n      (x rho0 c0^2 / rho0 c0^2)
            const double cR = hf * ph.invRho;
            const double d0 = cR * n0, d1 = cR * n1, d2 = cR * n2;          // 1/2 Fscale n / rho0
            const int bc = flags & FLAG_BC_MASK;
            double qm[2][4];
#pragma unroll
            for (int v = 0; v < 2; ++v) {
                const int own = v == 0 ? own0 : own1;
#pragma unroll
                for (int q = 0; q < 4; ++q) qm[v][q] = stg[q * KQ + own];
            }
#pragma unroll
            for (int v = 0; v < 2; ++v) {
                double fl[4];
                if (bc == FACE_INTERIOR) {
                    // zero mean flow: 1/2 n.(F(q-)+F(q+)) + 1/2 tau c0 (q- - q+)
                    const double ps = qm[v][0] + tr[v][0];
                    fl[0] = cPp * (qm[v][0] - tr[v][0]) + g0 * (qm[v][1] + tr[v][1]) + g1 * (qm[v][2] + tr[v][2]) + g2 * (qm[v][3] + tr[v][3]);
                    fl[1] = cP * (qm[v][1] - tr[v][1]) + d0 * ps;
                    fl[2] = cP * (qm[v][2] - tr[v][2]) + d1 * ps;
                    fl[3] = cP * (qm[v][3] - tr[v][3]) + d2 * ps;
                } else {
                    const double n[3] = {n0, n1, n2};
                    faceFlux(bc, 1.0, n, ph, qm[v], tr[v], fl);
                    fl[0] *= fs * invRc2;
#pragma unroll
                    for (int q = 1; q < 4; ++q) fl[q] *= fs;
                }
                if (v == 0 || act1) {
                    double* dst = inRow + C::OFF_F + (v == 0 ? slot0 : slot1);
#pragma unroll
                    for (int q = 0; q < 4; ++q) dst[q * NFL] = fl[q];
                }
            }
        }
    };

    // Prefetch: own values / geometry kOwnAhead tiles ahead in the cp.async rings, traces two tiles ahead in two register
    // sets. The loop is unrolled by two so that each set keeps its registers (even / odd tiles): a rotating copy would make
    // the compiler wait for the loads it has just issued.
    Ring<kInStages> in;
    int it = 0;
    auto step = [&](int& flagsX, int& nbrX, double (&trX)[2][4], int& flagsP, int& nbrP) {
        asm volatile("cp.async.wait_group %0;" ::"n"(kOwnAhead - 1) : "memory");  // own values / geometry of tile it have landed
        __syncwarp();
        mbarWait(&bar.inEmpty[in.b], in.phase ^ 1);            // the MMA warps are done with tile it - kInStages
#ifndef DGB_WS_NOFRONT
        prepAndFlux(sm.in + in.b * C::IN_TILE + s * LDI, sm.stg + (it % kStgStages) * C::STG_TILE + s * C::LDS_,
                    sm.geo + (it & (kGeoStages - 1)) * C::GEO_TILE + s * C::LDG_, sm.fg + (it & (kFgStages - 1)) * C::FG_TILE + s * C::LDFG,
                    flagsX, trX);
#endif
        __syncwarp();
        if (lane == 0) mbarArrive(&bar.full[in.b]);
        // refill: own values of tile it + kOwnAhead, traces of tile it + 2 (same register set), face metadata of tile it + 3
        issueOwn(elemOf(it + kOwnAhead), it + kOwnAhead);
        asm volatile("cp.async.commit_group;" ::: "memory");
        flagsX = flagsP; nbrX = nbrP;
#ifndef DGB_WS_NOTRACES
        loadTraces(elemOf(it + 2), flagsX, nbrX, trX);
#endif
        loadMeta(elemOf(it + 3), flagsP, nbrP);
        in.next();
        ++it;
    };
    int flagsA, nbrA, flagsB, nbrB, flagsP, nbrP;
    double trA[2][4], trB[2][4];
    loadMeta(elemOf(0), flagsA, nbrA);
    loadMeta(elemOf(1), flagsB, nbrB);
#pragma unroll
    for (int d = 0; d < kOwnAhead; ++d) {
        issueOwn(elemOf(d), d);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    loadTraces(elemOf(0), flagsA, nbrA, trA);
    loadTraces(elemOf(1), flagsB, nbrB, trB);
    loadMeta(elemOf(2), flagsP, nbrP);
    while (it < nIt) {
        step(flagsA, nbrA, trA, flagsP, nbrP);
        if (it < nIt) step(flagsB, nbrB, trB, flagsP, nbrP);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------------------------
// back warps: warp q owns field q of every tile (chain sums, velocity combine, fused RK update)
// ------------------------------------------------------------------------------------------------------------------
// k = dt L(y) combined with the RK registers; MODE is a compile-time constant here (see rkApplyK in dgb_device.cuh)
template <int MODE>
__device__ __forceinline__ void rkStore(double* __restrict__ u, double* __restrict__ acc, double* __restrict__ yout, double k, double uval,
                                        double accval) {
    if constexpr (MODE == MODE_RK1) { stStream(acc, k); stStream(yout, fma(0.5, k, uval)); }
    else if constexpr (MODE == MODE_RK2) { stStream(acc, fma(2.0, k, accval)); stStream(yout, fma(0.5, k, uval)); }
    else if constexpr (MODE == MODE_RK3) { stStream(acc, fma(2.0, k, accval)); stStream(yout, uval + k); }
    else if constexpr (MODE == MODE_RK4) { stStream(u, fma(accval + k, 1.0 / 6.0, uval)); }
    else if constexpr (MODE == MODE_EULER) { stStream(yout, uval + k); }
    else { stStream(yout, k); }
}

template <int P, int MODE>
__device__ __forceinline__ void backWarpImpl(const DeviceMesh& M, const StageArgs& A, const WsSmem& sm, int nIt, int w, int lane) {
    using C = WsCfg<P>;
    constexpr int NP = C::NP, LDO = C::LDO;
    constexpr int PS = kTileEl * LDO;              // output panel stride
    constexpr int NV = (kTileEl * NP + 31) / 32;   // passes over the 8*NP values of one field of a tile
    constexpr bool needU = MODE != MODE_RHS, needAcc = MODE >= MODE_RK2 && MODE <= MODE_RK4;
    const int q = w;
    const Phys ph = makePhys(M);
    const int64_t S = M.stride;
    const WsBars bar = wsBars(sm);
    // per-pass lane constants: value idx = v*32 + lane of the tile <-> (element el, node i)
    int oOff[NV];  // (element << 16) | offset of (element, node) inside a panel
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const int idx = min(v * 32 + lane, kTileEl * NP - 1);
        const int el = idx / NP, i = idx - el * NP;
        oOff[v] = (el << 16) | (el * LDO + i);
    }
    const double* uSrc = (MODE == MODE_EULER ? A.yin : A.u) + q * S;
    double* const uDst = A.u + q * S;
    double* const accP = A.acc + q * S;
    double* const youtP = A.yout + q * S;
    const TileMap tm = tileMap((A.eEnd - A.eBegin + kTileEl - 1) / kTileEl);

    Ring<kOutStages> ou;
    for (int it = 0; it < nIt; ++it, ou.next()) {
        const int e0 = A.eBegin + tm.elem(it);
        const int nVal = min(kTileEl, A.eEnd - e0) * NP;  // valid values of this tile
        const size_t base = (size_t)e0 * NP + lane;
        double uv[NV], av[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const bool ok = v * 32 + lane < nVal;
#ifdef DGB_WS_NOBACKMEM
            uv[v] = av[v] = 0.0;
#else
            uv[v] = (needU && ok) ? ldStream(uSrc + base + v * 32) : 0.0;
            av[v] = (needAcc && ok) ? ldStream(accP + base + v * 32) : 0.0;
#endif
        }
        mbarWait(&bar.outFull[ou.b], ou.phase);
        const double* out = sm.out + ou.b * C::OUT_TILE;
        const double* geo = sm.geo + (it & (kGeoStages - 1)) * C::GEO_TILE;
        double k[NV];
#ifdef DGB_WS_NOBACK
#pragma unroll
        for (int v = 0; v < NV; ++v) k[v] = out[oOff[v] & 0xffff];
#else
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const double* o = out + (oOff[v] & 0xffff);
            k[v] = o[(3 + q) * PS] + o[(7 + q) * PS];  // chain a + chain b
            if (q > 0) {
                const double* Gx = geo + (oOff[v] >> 16) * C::LDG_ + (q - 1) * 3;
                k[v] += Gx[0] * o[0] + Gx[1] * o[PS] + Gx[2] * o[2 * PS];  // the 1/rho0 is already in T^u
            } else {
                k[v] *= ph.rc2;  // the pressure equation was assembled divided by rho0 c0^2
            }
        }
#endif
        __syncwarp();
        if (lane == 0) mbarArrive(&bar.outEmpty[ou.b]);
#pragma unroll
        for (int v = 0; v < NV; ++v)
#ifdef DGB_WS_NOBACKMEM
            if (k[v] == 123.456) rkStore<MODE>(
#else
            if (v * 32 + lane < nVal) rkStore<MODE>(
#endif
               uDst + base + v * 32, accP + base + v * 32, youtP + base + v * 32, k[v], uv[v], av[v]);
    }
}

template <int P>
__device__ __forceinline__ void backWarp(const DeviceMesh& M, const StageArgs& A, const WsSmem& sm, int nIt, int w, int lane) {
    switch (A.mode) {
        case MODE_RK1: backWarpImpl<P, MODE_RK1>(M, A, sm, nIt, w, lane); break;
        case MODE_RK2: backWarpImpl<P, MODE_RK2>(M, A, sm, nIt, w, lane); break;
        case MODE_RK3: backWarpImpl<P, MODE_RK3>(M, A, sm, nIt, w, lane); break;
        case MODE_RK4: backWarpImpl<P, MODE_RK4>(M, A, sm, nIt, w, lane); break;
        case MODE_EULER: backWarpImpl<P, MODE_EULER>(M, A, sm, nIt, w, lane); break;
        default: backWarpImpl<P, MODE_RHS>(M, A, sm, nIt, w, lane); break;
    }
}

template <int P>
__global__ void __launch_bounds__(kThreadsWs, 1) stageWsKernel(DeviceMesh M, StageArgs A, int nTiles) {
    using C = WsCfg<P>;
    extern __shared__ __align__(128) unsigned char smemRaw[];
    WsSmem sm;
    sm.in = reinterpret_cast<double*>(smemRaw);
    sm.stg = sm.in + kInStages * C::IN_TILE;
    sm.geo = sm.stg + kStgStages * C::STG_TILE;
    sm.fg = sm.geo + kGeoStages * C::GEO_TILE;
    sm.out = sm.fg + kFgStages * C::FG_TILE;
    sm.bars = reinterpret_cast<unsigned long long*>(sm.out + kOutStages * C::OUT_TILE);
    sm.faceNodes = reinterpret_cast<int*>(sm.bars + kNumBars);
    sm.maps = reinterpret_cast<unsigned char*>(sm.faceNodes + C::NFL);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        const WsBars bar = wsBars(sm);
        for (int i = 0; i < kInStages; ++i) { mbarInit(&bar.full[i], kFrontWarps); mbarInit(&bar.inEmpty[i], kMmaWarps); }
        for (int i = 0; i < kOutStages; ++i) {
            mbarInit(&bar.chain[i], 1);
            mbarInit(&bar.chain[kOutStages + i], 1);
            mbarInit(&bar.outFull[i], 2);
            mbarInit(&bar.outEmpty[i], kBackWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < C::NFL; i += kThreadsWs) sm.faceNodes[i] = M.faceNodes[i];
    const int nMapsS = min(M.nMaps, kMaxMapsWs);
    for (int i = tid; i < nMapsS * C::NFP; i += kThreadsWs) sm.maps[i] = M.nbrMaps[i];
    // the K padding of the input rows must hold finite values (zeros) for ever
    for (int i = tid; i < kInStages * C::IN_TILE + kStgStages * C::STG_TILE; i += kThreadsWs) sm.in[i] = 0.0;
    __syncthreads();

    const int nIt = tileMap(nTiles).count;
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
    static KernelConfig kc;
    static_assert(kNumBars <= 32, "barrier block");
    const size_t smem = C::SMEM_BYTES;
    const int numSm = configureKernel(kc, stageWsKernel<P>, smem, "stage_ws_dmma");
    const int nTiles = (nEl + kTileEl - 1) / kTileEl;
    const int grid = std::max(1, std::min(numSm - std::min(A.smReserve, numSm / 2), nTiles));
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
