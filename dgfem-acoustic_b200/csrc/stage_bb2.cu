// Bernstein-Bezier stage kernel, second generation (dgb_set_option("kernel", 6)): the sparse operators of bb_ops.h behind an
// asynchronous, warp-private pipeline built on the Blackwell copy engines.
//
// Same fused operator as every other stage kernel (updateFlux + numStep + RK axpys of the reference, Mesh.cpp:476-674,
// solver.cpp:35-52, 261-285) and the same arithmetic as stage_bb.cu (strong form on Bernstein coefficients). What changes is
// the data layout and who moves the data. The first-generation kernels (stage_bb.cu) were measured on B200 at 2.43 ms per
// stage on the 1.43 M-tetrahedron order-4 mesh (profiles/r02/): 12 warps per SM, each walking load -> wait -> compute ->
// gather -> wait -> ... with every memory latency exposed (long-scoreboard stalls 5 per issue) and the L1 data pipe 57 %
// busy with 8-byte gathers (9.2 sectors per request). Here:
//
//   * STATE LAYOUT [element][canonical coefficient][4 fields] (32 B per coefficient). In Bernstein mode the device layout
//     is free (dgb_set_state / dgb_get_state convert anyway), so the four fields of a coefficient share one 32-byte
//     sector: a neighbour-trace gather fetches ONE fully used sector instead of four sectors used at 25 %, a tile of 8
//     consecutive elements is ONE contiguous run of 8*Np*32 bytes in each state array, and the canonical coefficient
//     order of bb_ops.h removes every permutation look-up from the kernel.
//   * ONE WARP = ONE PERSISTENT CTA that owns tiles of 8 elements end to end (thread = (element, field), as in
//     stage_bb_seq). No CTA barrier exists; the warp's own copies are all asynchronous:
//       - the tile of the stage input, and the RK registers u / acc of the tile, arrive by TMA bulk copies
//         (cp.async.bulk, mbarrier complete_tx) — no LSU instruction, no register, no L1 wavefront;
//       - the results leave by TMA bulk stores straight from the shared-memory tiles they were combined in;
//       - the neighbour traces are gathered face by face with 16-byte cp.async (zero-filled on boundary faces), lane
//         (element, slot) fetching its own element's trace (the per-element address work is done once per face), into one
//         or two trace buffers: with two, the traces of face s + 2 travel while faces s and s + 1 are lifted;
//       - the next tile's stage input is requested as soon as the stores have read the tile buffers (it sits in L2 by then:
//         prefetched at the start of the tile), its face metadata and inverse Jacobian travel in registers one tile ahead.
//   * 8 warps per SM at order 4 (27.7 KB of shared memory each, 242 registers): the latency that occupancy hid badly is
//     hidden by the copies in flight instead.
//   * Tetrahedra and triangles of orders 1..6 (bb::Simplex<DIM, N>; tetrahedra of order 6 — 84 coefficients per thread — spill ~300 bytes); the halo exchange of partitioned handles is
//     part of the kernel (bulk stores into the peers' halo slots, last CTA signals, border tiles wait).
//
// Shared memory of a warp: stage-input / u tile | acc tile (2 x 8*Np*32 B, padded where Np*4 != 4 mod 8) | one or two faces'
// traces | face coefficients | 3 mbarriers.
// The file also runs on the CPU: oracle/bb2_emulate.cpp includes it behind oracle/cuda_emu.h (DGB_EMULATE; test infrastructure).
#include <type_traits>

#include "bb_ops.h"
#include "dgb_async.cuh"
#include "dgb_device.cuh"
#include "dgb_internal.h"
#include "dgb_launch.h"

namespace dgb {

namespace {

constexpr int kTE2 = 8;  // elements per tile: 8 elements x 4 fields = the 32 lanes of the warp
#ifndef DGB_BB2_ACC_GLOBAL
#define DGB_BB2_ACC_GLOBAL 0  // 1 (experiment): acc is read / written straight from / to global memory in the epilogue (L2-prefetched); its tile buffer receives
                              // u at the start of the tile and the stage-input buffer the NEXT tile's input after the last face: no wait for either.
                              // Measured (profiles/r02/t_*): the 70 8-byte global accesses per thread cost far more than the two waits they remove —
                              // tetrahedra of order 4: 1.85 vs 1.46 ms, order 5: 0.99 vs 0.77; triangles of order 3 / 4: 2..5 % faster. Off.
#endif
#ifndef DGB_BB2_OWN_UNROLLED
#define DGB_BB2_OWN_UNROLLED 0  // 1 (experiment): own-trace offsets as immediates, one copy of the input loop per face. Measured (profiles/r02/ab_*): the tile loop grows
                                // from 4 472 to 6 832 instructions and the kernel slows from 1.46 to 1.63 ms at order 4 (instruction caches of 8 warps in different
                                // phases) — the table look-up stays
#endif
#ifndef DGB_BB2_TRACE_BUFFERS
#define DGB_BB2_TRACE_BUFFERS 0  // 0: per configuration (BB2Cfg::NTB); 1 / 2: forced (experiments)
#endif

template <int DIM, int P>
struct BB2Cfg {
    typedef bb::Simplex<DIM, P> SX;
    static constexpr int NP = SX::NP, NFP = SX::NFP, NF = DIM + 1;
    // Element stride of a state tile in shared memory (doubles). Lane (element, field) reads coefficient i of its element at
    // el*ES + i*4 + field: the four elements of a half-warp must start 4 (mod 8) doubles apart to hit disjoint 8-word bank windows. Np*4 does
    // at order 4 (140) — the tile is then one contiguous run, ONE bulk copy — but not at orders 3 and 5 (80 and 224 doubles: every
    // element on the same banks, measured 4-way conflicts) or 2 (40: 2-way): those tiles are padded and travel as one bulk copy
    // per element, issued by lanes 0..7.
    static constexpr int padFor(int n) { int pad = 0; while ((n + pad) % 8 != 4) pad += 4; return pad; }
    static constexpr int PAD = padFor(NP * 4);
    static constexpr int ES = NP * 4 + PAD;
    static constexpr int TILE = kTE2 * ES;             // doubles of one state tile in shared memory
    static constexpr int TRS = NFP * 4;                // element stride of the trace buffer: the 64-bit reads of lane (element, field) are conflict free
    static constexpr int NIT = (NFP + 1) / 2;          // gather instructions per face: lane (element, slot) moves the 16-byte chunk 4*it + slot of its element's trace
    static constexpr int RS = (NFP + 15) / 16 * 16;    // row stride (bytes) of DeviceMesh::bbNbr16
    static constexpr int FCS = 8;                      // doubles per (local face, element): app, aps, b, c, d, n (the quads of a quarter-warp read two records: no conflict)
    // trace buffers: the traces of face s + NTB are in flight while face s is lifted. Measured (profiles/r02/s_*, u_*): two buffers win at
    // tetrahedra of order 3 / 4 (0.441 -> 0.434, 1.60 -> 1.48 ms) and triangles of order 4 / 5 / 6 (1..3 %), one wins where the second costs a
    // resident warp or the lift is too short to matter (tetrahedra of order 1 / 2 / 5, triangles of order 1..3)
    static constexpr int NTB = DGB_BB2_TRACE_BUFFERS != 0 ? DGB_BB2_TRACE_BUFFERS : ((DIM == 3 && (P == 3 || P == 4)) || (DIM == 2 && P >= 4)) ? 2 : 1;
    static constexpr size_t SMEM = (size_t)(2 * TILE + NTB * kTE2 * TRS + 32 * FCS) * sizeof(double) + 4 * sizeof(unsigned long long);
    static_assert((TILE * 8) % 128 == 0 && (TRS * 8) % 16 == 0, "bulk-copy and 128-bit alignment of the shared-memory tiles");
};

// Volume term of field q of one element whose coefficients lie interleaved ([coefficient][4 fields], canonical order) at
// col: the same arithmetic as bb::fieldVolume, for tetrahedra and triangles (on triangles v_z has no coupling term).
template <int DIM, int N>
__device__ __forceinline__ void fieldVolumeInterleaved(int q, const double* col, const double (&gl)[4][3], const double (&v0)[3], bool flow, double rc2,
                                                       double invRho, double (&out)[bb::Simplex<DIM, N>::NP]) {
    typedef bb::Simplex<DIM, N> SX;
    constexpr int NP = SX::NP, ND = SX::ND, NV = SX::NV;
    double t[ND];
#pragma unroll
    for (int i = 0; i < ND; ++i) t[i] = 0.0;
    const int nCoupling = q == 0 ? DIM : q <= DIM ? 1 : 0;
    const int nPass = nCoupling + (flow ? 1 : 0);
    for (int pass = 0; pass < nPass; ++pass) {  // run-time trip count: one copy of the body
        int field;
        double w[NV];
        if (pass < nCoupling) {
            field = q == 0 ? 1 + pass : 0;
            const int x = q == 0 ? pass : q - 1;
            const double s = q == 0 ? rc2 : invRho;
#pragma unroll
            for (int j = 0; j < NV; ++j) w[j] = s * (x == 0 ? gl[j][0] : (x == 1 || DIM == 2) ? gl[j][1] : gl[j][2]);
        } else {
            field = q;
#pragma unroll
            for (int j = 0; j < NV; ++j) w[j] = DIM == 3 ? v0[0] * gl[j][0] + v0[1] * gl[j][1] + v0[2] * gl[j][2] : v0[0] * gl[j][0] + v0[1] * gl[j][1];
        }
        double cc[NP];
#pragma unroll
        for (int i = 0; i < NP; ++i) cc[i] = col[i * 4 + field];
        SX::dirDerivAcc(cc, w, t);
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) out[i] = 0.0;
    SX::elevate(t, -1.0, out);
}

// canonical volume index of coefficient b (canonical face order) of canonical face J — what DeviceMesh::bbOwn tabulates
template <int DIM, int N, int J>
__host__ __device__ constexpr int bb2OwnIndex(int b) {
    if (DIM == 2) return bb::d2::layerIdx<N, (J < 3 ? J : 0)>(0, b);
    for (int b1 = 0; b1 <= N; ++b1)
        for (int b2 = 0; b2 <= N - b1; ++b2)
            if (bb::fidx(N, b1, b2) == b) return bb::layerIdx<N, J>(0, b1, b2);
    return 0;
}

// resident warps per SM the register allocation aims at (the shared memory of a warp allows about as many)
// measured (profiles/r02/ac_*, ad_*): at Np = 10 tetrahedra (order 2) want 16 warps (0.60; 12: 0.54, 20: 0.56), triangles (order 3) 20 (0.72; 16: 0.69)
#ifndef DGB_BB2_W2D_NP15
#define DGB_BB2_W2D_NP15 16  // triangles of order 4: 0.80 -> 0.87 with 16 instead of 12 warps
#endif
#ifndef DGB_BB2_W2D_NP21
#define DGB_BB2_W2D_NP21 12  // triangles of order 5: 0.87 (14 warps: 0.81, spills)
#endif
#ifndef DGB_BB2_W2D_NP28
#define DGB_BB2_W2D_NP28 10  // triangles of order 6: 0.80 -> 0.84 with 10 instead of 8
#endif
#ifndef DGB_BB2_W3D_NP20
#define DGB_BB2_W3D_NP20 12  // tetrahedra of order 3: 0.71 (10 warps: 0.69, 14: 0.63)
#endif
#define DGB_BB2_WARPS(DIM, NP)                                                                                                                          \
    ((NP) <= 3 ? 24 : (NP) < 10 ? 16 : (NP) == 10 ? ((DIM) == 2 ? 20 : 16) : ((DIM) == 2 && (NP) == 15) ? DGB_BB2_W2D_NP15 : ((DIM) == 2 && (NP) == 21) ? DGB_BB2_W2D_NP21 \
     : ((DIM) == 2 && (NP) == 28) ? DGB_BB2_W2D_NP28 : ((DIM) == 3 && (NP) == 20) ? DGB_BB2_W3D_NP20 : (NP) <= 21 ? 12 : (NP) <= 35 ? 8 : (NP) <= 56 ? 6 : 4)
__host__ __device__ constexpr int bb2WarpsPerSm(int dim, int np) { return DGB_BB2_WARPS(dim, np); }

template <int DIM, int P>
__global__ void __launch_bounds__(32, bb2WarpsPerSm(DIM, BB2Cfg<DIM, P>::NP)) stageBB2Kernel(DeviceMesh M, StageArgs A, int nTiles) {
    using C = BB2Cfg<DIM, P>;
    using SX = typename C::SX;
    constexpr int NP = C::NP, NFP = C::NFP, TRS = C::TRS, ES = C::ES, NF = C::NF;
    constexpr bool kAccGlobal = DGB_BB2_ACC_GLOBAL != 0;
    constexpr bool kPerElement = C::PAD != 0;  // padded tiles: one bulk copy per element instead of one per tile
    DGB_DYNAMIC_SMEM(double2, smemRaw2);  // 16-byte aligned: bulk copies, cp.async, 128-bit accesses, 8-byte mbarriers
    double* const sY = reinterpret_cast<double*>(smemRaw2);  // [8][NP][4] stage input of the tile; after the last face: u, combined in place, stored
    double* const sA = sY + C::TILE;                         // acc tile: loaded, combined in place, stored
    double* const sT = sA + C::TILE;                         // [NTB][8][TRS] traces of the faces in flight
    double* const sFc = sT + C::NTB * kTE2 * TRS;            // [4][8][FCS] face coefficients of the tile
    unsigned long long* const bars = reinterpret_cast<unsigned long long*>(sFc + 32 * C::FCS);  // [0] stage input, [1] u, [2] acc

    const int lane = threadIdx.x, el = lane >> 2, q = lane & 3;
    const unsigned FULL = 0xffffffffu;
    const Phys ph = makePhys(M);
    const bool flow = ph.v0[0] != 0.0 || ph.v0[1] != 0.0 || ph.v0[2] != 0.0;
    const int mode = A.mode;
    // u is reloaded over the stage-input tile after the last face; in the first RK stage and in an Euler step the stage input IS u: the tile stays
    const bool loadU = kAccGlobal ? mode != MODE_RHS : (mode == MODE_RK2 || mode == MODE_RK3 || mode == MODE_RK4);
    const bool loadA = mode == MODE_RK2 || mode == MODE_RK3 || mode == MODE_RK4;
    const bool storeA = mode == MODE_RK1 || mode == MODE_RK2 || mode == MODE_RK3;
    const double* const uSrc = mode == MODE_EULER ? A.yin : A.u;
    double* const uDst = mode == MODE_RK4 ? A.u : A.yout;

    // fused halo exchange (partitioned handles): the tiles that hold cut-adjacent elements come last (interior-first
    // numbering); they read halo values — wait for the peers' flags of the previous stage first — and their results go
    // straight into the peers' halo slots
    const FusedHalo* const fx = A.fx;
    bool haloReady = fx == nullptr;
    auto touchesBorder = [&](int tt) { return A.eBegin + (tt + 1) * kTE2 > fx->Kinterior; };
    auto waitPeers = [&]() {
#ifndef DGB_EMULATE  // partitioned handles exist on devices only
        if (lane < fx->nPeers) {
            const unsigned long long* f = fx->myFlags + fx->waitRank[lane];
            unsigned long long t0, t1, v;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            for (;;) {
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
                if (v >= A.fxEpochWait) break;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > fx->timeoutNs) {
                    *fx->err = 1 + fx->waitRank[lane];
                    __threadfence_system();
                    break;
                }
                __nanosleep(100);
            }
        }
#endif
        __syncwarp();
        haloReady = true;
    };

    int t = blockIdx.x;
    if (t >= nTiles) return;
    if (lane == 0) {
        mbarInit2(&bars[0], 1);
        mbarInit2(&bars[1], 1);
        mbarInit2(&bars[2], 1);
        mbarInitFence();
    }
    __syncwarp();

    // face metadata (lane = (element, local face)) and inverse Jacobian (lane = (element, *)) of tile tt, clamped at the range end
    auto loadMeta = [&](int tt, int& flags, int& nbr, double (&fg)[4], double (&G)[DIM * DIM]) {
        const int e = min(A.eBegin + tt * kTE2 + el, A.eEnd - 1);
        const int ef = e * NF + min(q, NF - 1);  // triangles: the fourth lane of an element repeats face 2 (never selected)
        flags = M.fflags[ef];
        nbr = M.fnbr[ef];
        const double2 f0 = *reinterpret_cast<const double2*>(M.fgeo + (int64_t)ef * 4);
        const double2 f1 = *reinterpret_cast<const double2*>(M.fgeo + (int64_t)ef * 4 + 2);
        fg[0] = f0.x; fg[1] = f0.y; fg[2] = f1.x; fg[3] = f1.y * SX::FACE_SCALE;
#pragma unroll
        for (int j = 0; j < DIM * DIM; ++j) G[j] = M.Ginv[(int64_t)e * (DIM * DIM) + j];
    };
    auto tileBytes = [&](int tt) { return (uint32_t)(min(kTE2, A.eEnd - (A.eBegin + tt * kTE2)) * NP * 32); };
    // Tile copies (warp-collective). Contiguous tiles: lane 0 moves the whole tile; padded tiles: lane l moves element l.
    auto loadTile = [&](double* dst, const double* src, uint32_t bytes, unsigned long long* bar) {
        if (lane == 0) mbarExpectTx(bar, bytes);
        if constexpr (kPerElement) {
            __syncwarp();
            if (lane * (NP * 32) < (int)bytes) bulkLoad(dst + lane * ES, src + lane * (NP * 4), NP * 32, bar);
        } else {
            if (lane == 0) bulkLoad(dst, src, bytes, bar);
        }
    };
    auto storeTile = [&](double* dst, const double* src, uint32_t bytes) {  // the caller commits
        if constexpr (kPerElement) {
            if (lane * (NP * 32) < (int)bytes) bulkStore(dst + lane * (NP * 4), src + lane * ES, NP * 32);
        } else {
            if (lane == 0) bulkStore(dst, src, bytes);
        }
    };
    auto issueY = [&](int tt) { loadTile(sY, A.yin + (int64_t)(A.eBegin + tt * kTE2) * NP * 4, tileBytes(tt), &bars[0]); };
    // Traces of canonical face J for the tile whose face metadata (lane = (element, local face)) are flagsT / nbrT. Lane
    // (element, slot) moves the 16-byte chunks 4*it + slot of ITS element's trace: chunk j is half j&1 of the 32-byte record
    // (4 fields) of trace coefficient j>>1. Everything that depends on the element — neighbour, pairing map, base address — is
    // formed once per face; an iteration is one byte extraction, one 64-bit multiply-add and the copy. Boundary faces copy zeros.
    auto issueTraces = [&](int J, int flagsT, int nbrT, double* buf) {
        const int srcLane = (lane & ~3) | M.bbFaceLf[J];
        const int fl = __shfl_sync(FULL, flagsT, srcLane), nb = __shfl_sync(FULL, nbrT, srcLane);
        const bool interior = (fl & FLAG_BC_MASK) == FACE_INTERIOR && nb >= 0;
        uint32_t w[C::RS / 4];
        {
            const uint4* row = reinterpret_cast<const uint4*>(M.bbNbr16 + (size_t)((interior ? (fl >> FLAG_MAP_SHIFT) : 0) * 4 + J) * C::RS);
#pragma unroll
            for (int k = 0; k < C::RS / 16; ++k) {
                const uint4 v = row[k];
                w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
            }
        }
        const int hb = (lane >> 1) & 1;
        const uint32_t selLo = 0x4440u | hb, selHi = 0x4440u | (2 + hb);  // byte 2*it + hb of the row, zero-extended
        const double* const base = A.yin + (int64_t)(interior ? nb : 0) * (NP * 4) + (lane & 1) * 2;
        double* const dst = buf + el * TRS + (lane & 3) * 2;
        const uint32_t sz = interior ? 16u : 0u;
#pragma unroll
        for (int it = 0; it < C::NIT; ++it) {
            const uint32_t ci = __byte_perm(w[it >> 1], 0u, (it & 1) ? selHi : selLo);
            if (2 * it + 1 < NFP || hb == 0) cpAsync16(dst + it * 8, base + ci * 4, sz);
        }
        cpCommit();
    };

    int flags, nbr;
    double fg[4], G[DIM * DIM];
    loadMeta(t, flags, nbr, fg, G);
    issueY(t);
    if (!haloReady && touchesBorder(t)) waitPeers();
    issueTraces(0, flags, nbr, sT);
    if constexpr (C::NTB == 2) issueTraces(1, flags, nbr, sT + kTE2 * TRS);
    int tb = 0;  // which trace buffer the next face reads
    uint32_t phY = 0, phU = 0, phA = 0;

    for (;;) {
        const int tn = t + (int)gridDim.x;
        const bool more = tn < nTiles;
        const int e0 = A.eBegin + t * kTE2;
        const uint32_t bytes = tileBytes(t);

        // face coefficients of the tile (lane = (element, local face)), barycentric gradients of the lane's element
        {
            const int bc = flags & FLAG_BC_MASK;
            const double v0n = ph.v0[0] * fg[0] + ph.v0[1] * fg[1] + ph.v0[2] * fg[2];
            const bb::FaceCoef k = bb::faceCoef(bc, (flags & FLAG_TAU_NEG) ? -1.0 : 1.0, fg[3], v0n, ph.c0, ph.rho0);
            double2* fc = reinterpret_cast<double2*>(sFc + (q * 8 + el) * C::FCS);
            fc[0] = make_double2(k.app, k.aps);
            fc[1] = make_double2(k.b, k.c);
            fc[2] = make_double2(k.d, fg[0]);
            fc[3] = make_double2(fg[1], fg[2]);
        }
        double gl[4][3];
#pragma unroll
        for (int x = 0; x < DIM; ++x) {
            double s = 0.0;
#pragma unroll
            for (int u = 0; u < DIM; ++u) {
                gl[1 + u][x] = G[x * DIM + u];
                s = s + G[x * DIM + u];
            }
            gl[0][x] = -s;
        }
        // metadata of the next tile: requested now, consumed during the last face of this tile
        int flagsN = 0, nbrN = -1;
        double fgN[4] = {0, 0, 0, 0}, GN[DIM * DIM] = {};
        if (more) loadMeta(tn, flagsN, nbrN, fgN, GN);
        // acc is needed by the epilogue only; the stores of the previous tile have read sA (waited for before its stage input was requested)
        if constexpr (kAccGlobal) {
            if (loadU) loadTile(sA, uSrc + (int64_t)e0 * NP * 4, bytes, &bars[1]);  // the previous tile's store has read the buffer (waited for at its end)
            if (lane == 0 && loadA) bulkPrefetchL2(A.acc + (int64_t)e0 * NP * 4, bytes);
        } else {
            if (loadA) loadTile(sA, A.acc + (int64_t)e0 * NP * 4, bytes, &bars[2]);
        }
        if (lane == 0 && more) bulkPrefetchL2(A.yin + (int64_t)(A.eBegin + tn * kTE2) * NP * 4, tileBytes(tn));  // the request at the tile boundary will be an L2 hit
        // u can only be requested when the last face has read the stage-input tile it replaces (one lift before it is needed):
        // bring it to L2 now, so that the request finds it there
#ifdef DGB_BB2_UPREFETCH  // measured: 1.615 ms per stage with it, 1.585 without (config 5) — off
        if (lane == 0 && loadU && mode != MODE_RK1) bulkPrefetchL2(uSrc + (int64_t)e0 * NP * 4, bytes);
#endif
        mbarWait2(&bars[0], phY);
        phY ^= 1;
        __syncwarp();

        double out[NP];
        fieldVolumeInterleaved<DIM, P>(q, sY + el * ES, gl, ph.v0, flow, ph.rc2, ph.invRho, out);

#pragma unroll 1
        for (int J = 0; J < NF; ++J) {
            const int lf = M.bbFaceLf[J];
            // lift inputs of face J for this lane's (element, field), straight into registers. With a = own - neighbour
            // (boundary: the neighbour trace is zero), S = n . a_v, the reference's fluxes read (bb_ops.h: faceInput)
            //   x_p = app a_p + aps S,   x_v = b a_v + n_v (c a_p + d S):
            // every lane forms a for its own field; the quad needs two scalars per trace coefficient, S (the pressure lane) and g = c a_p + d S (the
            // velocity lanes), and gets them in TWO exchange rounds: with m = mu a (mu = 1 for p, n_v for the velocities)
            //   round 1 (xor 1): lane p <-> v_x, v_y <-> v_z            r1 = the partner's m
            //   round 2 (xor 2): send = P m + Q r1 with (P, Q) = (c, d) for p, (d, c) for v_x, (1, 1) for v_y, v_z
            //                    p receives m_vy + m_vz (S = r1 + r2); v_y, v_z receive c a_p + d m_vx; v_x receives m_vy + m_vz
            //   x = A1' a + E1 r1 + E2 r2  with lane-dependent coefficients (the own term d n_v m of g is folded into A1')
            // — one instruction stream for interior, absorbing and reflecting faces, the same arithmetic for an element wherever it sits in a tile.
            double A1, E1, E2, mu, Pq, Qq;
            {
                const double2* fc = reinterpret_cast<const double2*>(sFc + (lf * 8 + el) * C::FCS);
                const double2 c0 = fc[0], c1 = fc[1], c2 = fc[2], c3 = fc[3];  // (app, aps) (b, c) (d, n0) (n1, n2)
                const double nv = q == 1 ? c2.y : q == 2 ? c3.x : c3.y;
                const double nd = nv * c2.x;
                mu = q == 0 ? 1.0 : nv;
                Pq = q == 0 ? c1.y : q == 1 ? c2.x : 1.0;
                Qq = q == 0 ? c2.x : q == 1 ? c1.y : 1.0;
                E1 = q == 0 ? c0.y : q == 1 ? nv * c1.y : nd;
                E2 = q == 0 ? c0.y : q == 1 ? nd : nv;
                A1 = q == 0 ? c0.x : c1.x + nd * nv;
            }
            if constexpr (C::NTB == 2) cpWaitAllButOne(); else cpWaitAll();  // the traces of this face have landed (the next face's may still travel)
            __syncwarp();
            double x[NFP];
            const double* const own = sY + el * ES + q;
            double* const tbuf = sT + tb * (kTE2 * TRS);
            const double* const tr = tbuf + el * TRS + q;
#if DGB_BB2_OWN_UNROLLED
            // the own-trace offsets as immediates: one copy of the loop per face (no table look-up, no address arithmetic; more code)
            auto inputs = [&](auto Jc) {
                constexpr int JJ = decltype(Jc)::value;
#pragma unroll
                for (int b = 0; b < NFP; ++b) {
                    const double a = own[bb2OwnIndex<DIM, P, (JJ < NF ? JJ : 0)>(b) * 4] - tr[b * 4];
                    const double m = mu * a;
                    const double r1 = __shfl_xor_sync(FULL, m, 1);
                    const double r2 = __shfl_xor_sync(FULL, Pq * m + Qq * r1, 2);
                    x[b] = A1 * a + (E1 * r1 + E2 * r2);
                }
            };
            switch (J) {  // warp-uniform
                case 0: inputs(std::integral_constant<int, 0>{}); break;
                case 1: inputs(std::integral_constant<int, 1>{}); break;
                case 2: inputs(std::integral_constant<int, 2>{}); break;
                default: inputs(std::integral_constant<int, 3>{}); break;
            }
#else
#pragma unroll
            for (int b = 0; b < NFP; ++b) {
                const double a = own[M.bbOwn[J][b] * 4] - tr[b * 4];
                const double m = mu * a;
                const double r1 = __shfl_xor_sync(FULL, m, 1);
                const double r2 = __shfl_xor_sync(FULL, Pq * m + Qq * r1, 2);
                x[b] = A1 * a + (E1 * r1 + E2 * r2);
            }
#endif
            __syncwarp();  // this trace buffer is free: the traces of face (this + NTB) travel while this face and the next are lifted
            if constexpr (C::NTB == 2) {
                tb ^= 1;
                const int Jn = J + 2;  // face index in the sequence of this tile's faces followed by the next tile's
                if (Jn < NF) issueTraces(Jn, flags, nbr, tbuf);
                else if (more) {
                    if (Jn == NF && !haloReady && touchesBorder(tn)) waitPeers();
                    issueTraces(Jn - NF, flagsN, nbrN, tbuf);
                } else cpCommit();  // an empty group keeps "all but the most recent group" uniform to the last face
                // every read of the stage-input tile is done: the tile now receives u (or, with acc in global memory, the next tile's input)
                if (J == NF - 1) {
                    if constexpr (kAccGlobal) { if (more) issueY(tn); }
                    else if (loadU) loadTile(sY, uSrc + (int64_t)e0 * NP * 4, bytes, &bars[1]);
                }
            } else {
                if (J < NF - 1) issueTraces(J + 1, flags, nbr, tbuf);
                else {
                    // every read of the stage-input tile is done: the tile now receives u; the next tile's first traces start
                    if constexpr (kAccGlobal) { if (more) issueY(tn); }
                    else if (loadU) loadTile(sY, uSrc + (int64_t)e0 * NP * 4, bytes, &bars[1]);
                    if (more) {
                        if (!haloReady && touchesBorder(tn)) waitPeers();
                        issueTraces(0, flagsN, nbrN, tbuf);
                    }
                }
            }
            double zl[NP];
            SX::liftLocal(x, zl);
            switch (J) {  // warp-uniform
                case 0: SX::template scatterAdd<0>(zl, out); break;
                case 1: SX::template scatterAdd<1>(zl, out); break;
                case 2: SX::template scatterAdd<2>(zl, out); break;
                default: SX::template scatterAdd<3>(zl, out); break;
            }
        }

        // fused RK update in the shared-memory tiles of the RK registers, then bulk stores
        if constexpr (kAccGlobal) {
            if (loadU) { mbarWait2(&bars[1], phU); phU ^= 1; }
            double* const pu = sA + el * ES + q;
            const bool valid = e0 + el < A.eEnd;
            double* const ga = A.acc + ((int64_t)min(e0 + el, A.eEnd - 1) * NP) * 4 + q;
            const double dt = A.dt;
            switch (mode) {
                case MODE_RK1:
#pragma unroll
                    for (int i = 0; i < NP; ++i) { const double k = __dmul_rn(dt, out[i]); if (valid) __stcg(ga + i * 4, k); pu[i * 4] = pu[i * 4] + 0.5 * k; }
                    break;
                case MODE_RK2:
#pragma unroll
                    for (int i = 0; i < NP; ++i) { const double k = __dmul_rn(dt, out[i]); const double a = __ldcg(ga + i * 4) + 2 * k; if (valid) __stcg(ga + i * 4, a); pu[i * 4] = pu[i * 4] + 0.5 * k; }
                    break;
                case MODE_RK3:
#pragma unroll
                    for (int i = 0; i < NP; ++i) { const double k = __dmul_rn(dt, out[i]); const double a = __ldcg(ga + i * 4) + 2 * k; if (valid) __stcg(ga + i * 4, a); pu[i * 4] = pu[i * 4] + k; }
                    break;
                case MODE_RK4:
#pragma unroll
                    for (int i = 0; i < NP; ++i) { const double k = __dmul_rn(dt, out[i]); pu[i * 4] = fma(__ldcg(ga + i * 4) + k, 1.0 / 6.0, pu[i * 4]); }
                    break;
                case MODE_EULER:
#pragma unroll
                    for (int i = 0; i < NP; ++i) pu[i * 4] = pu[i * 4] + __dmul_rn(dt, out[i]);
                    break;
                default:
#pragma unroll
                    for (int i = 0; i < NP; ++i) pu[i * 4] = out[i];
                    break;
            }
        } else {
        // Two passes: what needs only acc first (k in place of out, the acc tile updated and on its way back), then — u is requested last and may
        // still be travelling — what needs u. The arithmetic is that of the one-pass form (RK4: fma(acc + k, 1/6, u)).
        if (loadA) { mbarWait2(&bars[2], phA); phA ^= 1; }
        {
            double* const pu = sY + el * ES + q;
            double* const pa = sA + el * ES + q;
            const double dt = A.dt;
            if (mode == MODE_RK1) {
#pragma unroll
                for (int i = 0; i < NP; ++i) { const double k = __dmul_rn(dt, out[i]); out[i] = k; pa[i * 4] = k; }
            } else if (mode == MODE_RK2 || mode == MODE_RK3) {
#pragma unroll
                for (int i = 0; i < NP; ++i) { const double k = __dmul_rn(dt, out[i]); out[i] = k; pa[i * 4] = pa[i * 4] + 2 * k; }
            } else if (mode == MODE_RK4) {
#pragma unroll
                for (int i = 0; i < NP; ++i) out[i] = pa[i * 4] + __dmul_rn(dt, out[i]);
            } else if (mode == MODE_EULER) {
#pragma unroll
                for (int i = 0; i < NP; ++i) out[i] = __dmul_rn(dt, out[i]);
            }
            if (storeA) {
                fenceProxyAsync();
                __syncwarp();
                storeTile(A.acc + (int64_t)e0 * NP * 4, sA, bytes);
                bulkCommit();
            }
            if (loadU) { mbarWait2(&bars[1], phU); phU ^= 1; }
            if (mode == MODE_RK1 || mode == MODE_RK2) {
#pragma unroll
                for (int i = 0; i < NP; ++i) pu[i * 4] = pu[i * 4] + 0.5 * out[i];
            } else if (mode == MODE_RK3 || mode == MODE_EULER) {
#pragma unroll
                for (int i = 0; i < NP; ++i) pu[i * 4] = pu[i * 4] + out[i];
            } else if (mode == MODE_RK4) {
#pragma unroll
                for (int i = 0; i < NP; ++i) pu[i * 4] = fma(out[i], 1.0 / 6.0, pu[i * 4]);
            } else {
#pragma unroll
                for (int i = 0; i < NP; ++i) pu[i * 4] = out[i];
            }
        }
        }
        fenceProxyAsync();
        __syncwarp();
        double* const sUt = kAccGlobal ? sA : sY;  // the tile that holds the updated u
        storeTile(uDst + (int64_t)e0 * NP * 4, sUt, bytes);
        if (fx != nullptr && touchesBorder(t)) {
            // lane l ships element l of the tile to every peer that holds it as a halo element: one bulk store per target,
            // straight from the shared-memory tile over NVLink
            const int k = e0 + lane - fx->Kinterior;
            if (lane < (int)(bytes / (NP * 32)) && k >= 0) {
                const int p1 = fx->pushOff[k + 1];
                for (int p = fx->pushOff[k]; p < p1; ++p)
                    bulkStore(fx->arr[A.fxWhich][fx->pushPeer[p]] + (int64_t)fx->pushSlot[p] * NP * 4, sUt + lane * ES, NP * 32);
            }
        }
        bulkCommit();  // (the acc tile left in the first pass of the epilogue)
        if (more) {
            bulkWaitRead();  // the tiles have been read by the stores (every lane waits for the copies it issued)
            __syncwarp();
            if constexpr (!kAccGlobal) issueY(tn);
        }
        if (!more) break;
        t = tn;
        flags = flagsN;
        nbr = nbrN;
#pragma unroll
        for (int j = 0; j < 4; ++j) fg[j] = fgN[j];
#pragma unroll
        for (int j = 0; j < DIM * DIM; ++j) G[j] = GN[j];
    }
    bulkWaitAll();  // every store of this warp, local and remote, is complete
#ifndef DGB_EMULATE
    if (fx != nullptr) {
        // the last CTA to get here raises this rank's flag at every peer (release at system scope after all pushes)
        __syncwarp();
        unsigned int last = 0;
        if (lane == 0) {
            __threadfence_system();
            last = atomicAdd(fx->doneCounter, 1u) == gridDim.x - 1 ? 1u : 0u;
        }
        last = __shfl_sync(FULL, last, 0);
        if (last) {
            if (lane == 0) *fx->doneCounter = 0;
            __threadfence_system();
            if (lane < fx->nPeers) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(fx->peerFlag[lane]), "l"(A.fxEpochSignal) : "memory");
        }
    }
#endif
}

template <int DIM, int P>
void launchBB2(const DeviceMesh& M, const StageArgs& A, cudaStream_t s) {
    using C = BB2Cfg<DIM, P>;
    const int nEl = A.eEnd - A.eBegin;
    if (nEl <= 0) return;
    const int nTiles = (nEl + kTE2 - 1) / kTE2;
#ifdef DGB_EMULATE
    const int grid = std::max(1, std::min(nTiles, 3));  // a few persistent "CTAs": every warp walks several tiles
#else
    static KernelConfig kc;
    static int perSm[kMaxDevices] = {};
    const int numSm = configureKernel(kc, stageBB2Kernel<DIM, P>, C::SMEM, "stage_bb2");
    int dev = 0;
    cudaGetDevice(&dev);
    if (perSm[dev] == 0) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, stageBB2Kernel<DIM, P>, 32, C::SMEM) != cudaSuccess || n < 1) {
            cudaGetLastError();
            throw UnsupportedError("stage_bb2: the kernel does not fit an SM of this device");
        }
        perSm[dev] = n;
    }
    const int grid = std::max(1, std::min(nTiles, (numSm - std::min(A.smReserve, numSm / 2)) * perSm[dev]));
#endif
    DGB_LAUNCH((stageBB2Kernel<DIM, P>), grid, 32, C::SMEM, s, M, A, nTiles);
}

// nodal <-> Bernstein conversion between the field-major nodal layout u[q][el*Np + n] (the reference's, the C ABI's) and the
// interleaved coefficient layout c[(el*Np + i)*4 + q]. mat is [Np][Np] row-major: toBB: c_i = sum_n mat[i][n] u_n (rows of
// V^-1 in canonical order); fromBB: u_n = sum_i mat[n][i] c_i (columns of V in canonical order). in != out.
__global__ void __launch_bounds__(256) convertBB2Kernel(const double* __restrict__ in, double* __restrict__ out, int64_t stride, int Np, int K,
                                                        const double* __restrict__ mat, int toBB) {
    DGB_DYNAMIC_SMEM(double, sx2);  // [4][E*Np]
    const int E = 256 / Np;
    const int e0 = blockIdx.x * E;
    const int nE = min(E, K - e0);
    const int tid = threadIdx.x;
    const int cnt = nE * Np;
    if (toBB) {
        if (tid < cnt)
            for (int q = 0; q < 4; ++q) sx2[q * E * Np + tid] = in[q * stride + (int64_t)e0 * Np + tid];
    } else {
        for (int i = tid; i < 4 * cnt; i += 256) sx2[(i & 3) * E * Np + (i >> 2)] = in[(int64_t)e0 * Np * 4 + i];
    }
    __syncthreads();
    if (tid >= cnt) return;
    const int el = tid / Np, r = tid - el * Np;
    double acc[4] = {0, 0, 0, 0};
    for (int m = 0; m < Np; ++m) {
        const double a = mat[r * Np + m];
        for (int q = 0; q < 4; ++q) acc[q] = fma(a, sx2[q * E * Np + el * Np + m], acc[q]);
    }
    if (toBB) {
        double2* o = reinterpret_cast<double2*>(out + ((int64_t)e0 * Np + tid) * 4);
        o[0] = make_double2(acc[0], acc[1]);
        o[1] = make_double2(acc[2], acc[3]);
    } else {
        for (int q = 0; q < 4; ++q) out[q * stride + (int64_t)e0 * Np + tid] = acc[q];
    }
}

// whole elements of an interleaved state array: out[k] = y[elems[k]] (pack, elems != nullptr) — 4*Np doubles per element
__global__ void packElementsBB2Kernel(const double* __restrict__ y, int per, const int32_t* __restrict__ elems, int n, double* __restrict__ buf) {
    const int64_t tot = (int64_t)n * per;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(i / per), r = (int)(i - (int64_t)k * per);
        buf[i] = y[(int64_t)elems[k] * per + r];
    }
}

}  // namespace

StageKernel selectBB2Kernel(int dim, int order) {
    StageKernel k;
#define DGB_BB2(D, P) if (dim == D && order == P) { k.launch = &launchBB2<D, P>; k.name = "stage_bb2<" #D "," #P ">"; }
    DGB_BB2(3, 1) DGB_BB2(3, 2) DGB_BB2(3, 3) DGB_BB2(3, 4) DGB_BB2(3, 5) DGB_BB2(3, 6)
    DGB_BB2(2, 1) DGB_BB2(2, 2) DGB_BB2(2, 3) DGB_BB2(2, 4) DGB_BB2(2, 5) DGB_BB2(2, 6)
#undef DGB_BB2
    return k;
}

void launchConvertBB2(const double* in, double* out, int64_t stride, int Np, int K, const double* mat, bool toBB, cudaStream_t s) {
    if (K <= 0) return;
    const int E = 256 / Np;
    DGB_LAUNCH(convertBB2Kernel, (unsigned)((K + E - 1) / E), 256, (size_t)4 * E * Np * sizeof(double), s, in, out, stride, Np, K, mat, toBB ? 1 : 0);
}

void launchPackElementsBB2(const double* y, int Np, const int32_t* elems, int n, double* buf, cudaStream_t s) {
    const int64_t tot = 4ll * n * Np;
    if (tot <= 0) return;
    const unsigned blocks = (unsigned)std::min<int64_t>((tot + 255) / 256, 148 * 8);
    DGB_LAUNCH(packElementsBB2Kernel, blocks, 256, 0, s, y, 4 * Np, elems, n, buf);
}

}  // namespace dgb
