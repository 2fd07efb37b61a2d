// Bernstein-Bezier stage kernel for the LOWEST orders (dgb_set_option("kernel", 7)): one thread = one element, all four fields.
//
// Same fused operator as every other stage kernel (updateFlux + numStep + RK axpys of the reference, Mesh.cpp:476-674,
// solver.cpp:35-52, 261-285), same arithmetic (bb_ops.h) and the same state layout as stage_bb2.cu — interleaved canonical
// Bernstein coefficients c[(el*Np + i)*4 + field], representation 2 — so a handle can switch between the two without a
// conversion. What differs is the mapping. With 3..10 coefficients per element and field (triangles of order 1..3, tetrahedra
// of order 1) the (element, field) threads of stage_bb2 spend most of their instructions on per-tile bookkeeping, quad shuffles
// and one-face-at-a-time gathers of a few bytes; measured on B200 they reach 0.37..0.51 of the HBM roof there. Here
//   * a warp is a persistent CTA that owns tiles of 32 consecutive elements; lane l is element l of the tile and keeps the
//     element's 4*Np coefficients and its 4*Np results in registers — no shuffle, no shared-memory face buffers, every index
//     a compile-time constant (the faces are unrolled);
//   * the three streams of the tile (stage input, u, acc) arrive by TMA bulk copies into lane-strided shared-memory rows and
//     leave the same way (one copy per tile where the bank conflicts of unpadded rows are cheap, one per element into padded
//     rows otherwise — see BBECfg); the next tile's stage input is requested as soon as the coefficients sit in registers, u / acc as soon as the
//     previous tile's stores have read their buffers;
//   * neighbour traces are gathered straight into registers: two 128-bit loads per trace coefficient (all four fields, one
//     fully used 32-byte sector), all faces in flight at once.
#include "bb_ops.h"
#include "dgb_async.cuh"
#include "dgb_device.cuh"
#include "dgb_internal.h"
#include "dgb_launch.h"

namespace dgb {

namespace {

constexpr int kTEE = 32;  // elements per tile = lanes
#ifndef DGB_BBE_CONTIG
#define DGB_BBE_CONTIG(NP) ((NP) == 3 || (NP) == 6 || (NP) == 10)
#endif
#ifndef DGB_BBE_PADDED_TRANSPORT
#define DGB_BBE_PADDED_TRANSPORT 2
#endif
#ifndef DGB_BBE_WARPS
// resident warps per SM the register allocation aims at; measured (profiles/r02/x_*): triangles of order 1 0.82 -> 0.89 with 16 instead of 12 warps, tetrahedra of
// order 1 0.62 -> 0.66 with 12 instead of 10, triangles of order 2 0.71 -> 0.60 with 10 instead of 8 (spills)
#define DGB_BBE_WARPS(NP) ((NP) <= 3 ? 16 : (NP) <= 4 ? 12 : (NP) <= 6 ? 8 : 6)
#endif

template <int DIM, int P>
struct BBECfg {
    typedef bb::Simplex<DIM, P> SX;
    static constexpr int NP = SX::NP, NFP = SX::NFP, NF = DIM + 1;
    // Measured on B200 (profiles/r02/): with one bulk copy per element (padded, conflict-free rows) the kernel is bound by the
    // REQUEST RATE of the copy engine (160 requests of 96..192 bytes per tile: 0.58 / 0.59 of the HBM roof on triangles of order
    // 1 / 2); with ONE bulk copy per tile and array into unpadded rows the 128-bit accesses of the 32 lanes conflict 2-way
    // (Np = 3) or 4-way (Np = 6, 10) and the kernel reaches 0.83 / 0.71 (Np = 10: 0.56 -> 0.60, tetrahedra of order 2: 0.47 either way —
    // stage_bb2 is the better kernel at Np = 10). Np = 4 (tetrahedra of order 1: 128-byte rows, 8-way conflicts, 0.46 unpadded) keeps padded rows,
    // filled by warp-cooperative cp.async instead of per-element bulk copies (TRANSPORT below: 0.57 -> 0.62).
    static constexpr bool CONTIG = DGB_BBE_CONTIG(NP);
    // (The cooperative transport below for EVERY row size loses to the single bulk copy with conflicts: 0.72 vs 0.90 at Np = 3, 0.67 vs 0.72 at Np = 6,
    // 0.56 vs 0.60 at Np = 10 — profiles/r02/ae_*.)
    // how padded rows travel: 1 = one TMA bulk copy per element (request-rate bound, see above), 2 = warp-cooperative 16-byte cp.async into the padded
    // rows (32 lanes x 16 B = 512 contiguous bytes of global memory per instruction) and 128-bit shared loads + coalesced 128-bit global stores back
    static constexpr int TRANSPORT = CONTIG ? 0 : DGB_BBE_PADDED_TRANSPORT;
    static constexpr int CPR = NP * 2;               // 16-byte chunks per element
    static constexpr int ES = CONTIG ? NP * 4 : NP * 4 + 2;  // element stride in shared memory (doubles); padded: ES/2 odd
    static constexpr int TILE = kTEE * ES;
    static constexpr int RS = (NFP + 15) / 16 * 16;  // row stride (bytes) of DeviceMesh::bbNbr16
    static constexpr size_t SMEM = (size_t)3 * TILE * sizeof(double) + 4 * sizeof(unsigned long long);
    static constexpr int WARPS = DGB_BBE_WARPS(NP);  // what shared memory and registers allow per SM
};

// canonical volume index of coefficient b (canonical face order) of canonical face J
template <int DIM, int N, int J>
__host__ __device__ constexpr int ownIndex(int b) {
    if (DIM == 2) return bb::d2::layerIdx<N, J < 3 ? J : 0>(0, b);
    for (int b1 = 0; b1 <= N; ++b1)
        for (int b2 = 0; b2 <= N - b1; ++b2)
            if (bb::fidx(N, b1, b2) == b) return bb::layerIdx<N, J>(0, b1, b2);
    return 0;
}

template <int DIM, int P>
__global__ void __launch_bounds__(32, BBECfg<DIM, P>::WARPS) stageBBEKernel(DeviceMesh M, StageArgs A, int nTiles) {
    using C = BBECfg<DIM, P>;
    using SX = typename C::SX;
    constexpr int NP = C::NP, NFP = C::NFP, NF = C::NF, ES = C::ES, NV = SX::NV, ND = SX::ND;
    DGB_DYNAMIC_SMEM(double2, smemRawE);  // 16-byte aligned: bulk copies, 128-bit accesses, 8-byte mbarriers
    double* const sY = reinterpret_cast<double*>(smemRawE);  // stage input of the tile, [32][ES]
    double* const sU = sY + C::TILE;                          // u: loaded, combined in place, stored
    double* const sA = sU + C::TILE;                          // acc likewise
    unsigned long long* const bars = reinterpret_cast<unsigned long long*>(sA + C::TILE);  // [0] stage input, [1] u, [2] acc

    const int lane = threadIdx.x;
    const Phys ph = makePhys(M);
    const bool flow = ph.v0[0] != 0.0 || ph.v0[1] != 0.0 || (DIM == 3 && ph.v0[2] != 0.0);
    const int mode = A.mode;
    const bool loadU = mode != MODE_RHS, loadA = mode == MODE_RK2 || mode == MODE_RK3 || mode == MODE_RK4;
    const bool storeA = mode == MODE_RK1 || mode == MODE_RK2 || mode == MODE_RK3;
    const double* const uSrc = mode == MODE_EULER ? A.yin : A.u;
    double* const uDst = mode == MODE_RK4 ? A.u : A.yout;

    int t = blockIdx.x;
    if (t >= nTiles) return;
    if (lane == 0) {
        mbarInit2(&bars[0], 1);
        mbarInit2(&bars[1], 1);
        mbarInit2(&bars[2], 1);
        mbarInitFence();
    }
    __syncwarp();

    auto nElems = [&](int tt) { return min(kTEE, A.eEnd - (A.eBegin + tt * kTEE)); };
    // tile copies. Unpadded rows: lane 0 moves the tile; padded rows: lane l moves element l (Np*32 contiguous bytes)
    auto loadTile = [&](double* dst, const double* src, int n, unsigned long long* bar) {
        if (lane == 0) mbarExpectTx(bar, (uint32_t)(n * NP * 32));
        if constexpr (C::CONTIG) {
            if (lane == 0) bulkLoad(dst, src, (uint32_t)(n * NP * 32), bar);
        } else {
            __syncwarp();
            if (lane < n) bulkLoad(dst + lane * ES, src + (int64_t)lane * (NP * 4), NP * 32, bar);
        }
    };
    auto storeTile = [&](double* dst, const double* src, int n) {
        if constexpr (C::CONTIG) {
            if (lane == 0) bulkStore(dst, src, (uint32_t)(n * NP * 32));
        } else {
            if (lane < n) bulkStore(dst + (int64_t)lane * (NP * 4), src + lane * ES, NP * 32);
        }
    };

    // cooperative transport of padded rows (TRANSPORT == 2): chunk g of the tile = chunk g % CPR of element g / CPR
    constexpr int kCoop = (kTEE * C::CPR + 31) / 32;
    auto coopLoad = [&](double* dst, const double* src, int n) {
        const int total = n * C::CPR;
#pragma unroll
        for (int k = 0; k < kCoop; ++k) {
            const int g = k * 32 + lane;
            if (g < total) {
                const int er = g / C::CPR, c = g - er * C::CPR;
                cpAsync16(dst + er * ES + c * 2, src + (int64_t)g * 2, 16u);
            }
        }
    };
    auto coopStore = [&](double* dst, const double* src, int n) {
        const int total = n * C::CPR;
#pragma unroll
        for (int k = 0; k < kCoop; ++k) {
            const int g = k * 32 + lane;
            if (g < total) {
                const int er = g / C::CPR, c = g - er * C::CPR;
                *reinterpret_cast<double2*>(dst + (int64_t)g * 2) = *reinterpret_cast<const double2*>(src + er * ES + c * 2);
            }
        }
    };
    constexpr bool kCoopMode = C::TRANSPORT == 2;

    if constexpr (kCoopMode) {
        coopLoad(sY, A.yin + (int64_t)(A.eBegin + t * kTEE) * NP * 4, nElems(t));
        cpCommit();
    } else {
        loadTile(sY, A.yin + (int64_t)(A.eBegin + t * kTEE) * NP * 4, nElems(t), &bars[0]);
    }
    uint32_t phY = 0, phU = 0, phA = 0;

    for (;;) {
        const int tn = t + (int)gridDim.x;
        const bool more = tn < nTiles;
        const int n = nElems(t);
        const int e0 = A.eBegin + t * kTEE;
        const int e = min(e0 + lane, A.eEnd - 1);  // lanes beyond the range repeat its last element (results not stored)
        if constexpr (kCoopMode) {  // one cp.async group for u and acc (possibly empty: the group count per tile stays uniform)
            if (loadU) coopLoad(sU, uSrc + (int64_t)e0 * NP * 4, n);
            if (loadA) coopLoad(sA, A.acc + (int64_t)e0 * NP * 4, n);
            cpCommit();
        } else {
            if (loadU) loadTile(sU, uSrc + (int64_t)e0 * NP * 4, n, &bars[1]);
            if (loadA) loadTile(sA, A.acc + (int64_t)e0 * NP * 4, n, &bars[2]);
        }

#ifdef DGB_BBE_META_PREFETCH  // measured (profiles/r02/y_*): 0.87 vs 0.90 (triangles, order 1), 0.73 vs 0.71 (order 2), 0.66 vs 0.66 (tetrahedra, order 1) — off
        // the next tile's face metadata and inverse Jacobian: to L2 now (no register is spent), so that the dependent chain metadata -> neighbour
        // gathers of the next tile starts from L2 hits
        if (more) {
            const int en = min(A.eBegin + tn * kTEE + lane, A.eEnd - 1);
            prefetchL2(M.fflags + (int64_t)en * NF);
            prefetchL2(M.fnbr + (int64_t)en * NF);
            prefetchL2(M.fgeo + (int64_t)en * NF * 4);
            prefetchL2(M.fgeo + (int64_t)en * NF * 4 + (NF * 4 - 1));
            prefetchL2(M.Ginv + (int64_t)en * (DIM * DIM));
        }
#endif
        // geometry of the lane's element: barycentric gradients
        double gl[4][3];
        {
            double G[DIM * DIM];
#pragma unroll
            for (int j = 0; j < DIM * DIM; ++j) G[j] = M.Ginv[(int64_t)e * (DIM * DIM) + j];
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
        }
        // faces: metadata (the traces themselves are requested face by face below; the scheduler hoists those loads as far as
        // the register budget allows)
        int flOf[NF], nbOf[NF];
        double fgOf[NF][4];
        auto loadFace = [&](auto Jc) {
            constexpr int J = decltype(Jc)::value;
            if constexpr (J < NF) {
                const int ef = e * NF + M.bbFaceLf[J];
                flOf[J] = M.fflags[ef];
                nbOf[J] = M.fnbr[ef];
                const double2 f0 = *reinterpret_cast<const double2*>(M.fgeo + (int64_t)ef * 4);
                const double2 f1 = *reinterpret_cast<const double2*>(M.fgeo + (int64_t)ef * 4 + 2);
                fgOf[J][0] = f0.x; fgOf[J][1] = f0.y; fgOf[J][2] = f1.x; fgOf[J][3] = f1.y * SX::FACE_SCALE;
            }
        };
        loadFace(std::integral_constant<int, 0>{});
        loadFace(std::integral_constant<int, 1>{});
        loadFace(std::integral_constant<int, 2>{});
        loadFace(std::integral_constant<int, 3>{});

        // own coefficients: shared memory -> registers, then the buffer is free for the next tile's stage input
        if constexpr (kCoopMode) {
            cpWaitAllButOne();  // the stage input (older than the u / acc group) has landed
            __syncwarp();
        } else {
            mbarWait2(&bars[0], phY);
            phY ^= 1;
        }
        double c[4][NP];
        {
            const double* row = sY + lane * ES;
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                const double2 lo = *reinterpret_cast<const double2*>(row + i * 4);
                const double2 hi = *reinterpret_cast<const double2*>(row + i * 4 + 2);
                c[0][i] = lo.x; c[1][i] = lo.y; c[2][i] = hi.x; c[3][i] = hi.y;
            }
        }
        __syncwarp();
        if constexpr (kCoopMode) {
            if (more) coopLoad(sY, A.yin + (int64_t)(A.eBegin + tn * kTEE) * NP * 4, nElems(tn));
            cpCommit();
        } else {
            if (more) loadTile(sY, A.yin + (int64_t)(A.eBegin + tn * kTEE) * NP * 4, nElems(tn), &bars[0]);
        }

        // volume term: out_q = -elevate( sum of directional derivatives )
        double out[4][NP];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            double td[ND];
#pragma unroll
            for (int i = 0; i < ND; ++i) td[i] = 0.0;
            if (q == 0) {
#pragma unroll
                for (int x = 0; x < DIM; ++x) {
                    double w[NV];
#pragma unroll
                    for (int j = 0; j < NV; ++j) w[j] = ph.rc2 * gl[j][x];
                    SX::dirDerivAcc(c[1 + x], w, td);
                }
            } else if (q <= DIM) {
                double w[NV];
#pragma unroll
                for (int j = 0; j < NV; ++j) w[j] = ph.invRho * gl[j][q - 1];
                SX::dirDerivAcc(c[0], w, td);
            }
            if (flow) {
                double w[NV];
#pragma unroll
                for (int j = 0; j < NV; ++j) w[j] = DIM == 3 ? ph.v0[0] * gl[j][0] + ph.v0[1] * gl[j][1] + ph.v0[2] * gl[j][2] : ph.v0[0] * gl[j][0] + ph.v0[1] * gl[j][1];
                SX::dirDerivAcc(c[q], w, td);
            }
#pragma unroll
            for (int i = 0; i < NP; ++i) out[q][i] = 0.0;
            SX::elevate(td, -1.0, out[q]);
        }

        // faces: neighbour traces straight into registers (all four fields of a coefficient: two 128-bit loads, one sector), lift
        // inputs from own - neighbour (boundary: the neighbour trace is zero), lift, scatter
        auto doFace = [&](auto Jc) {
            constexpr int J = decltype(Jc)::value;
            if constexpr (J < NF) {
                const int fl = flOf[J], nb = nbOf[J];
                const int bc = fl & FLAG_BC_MASK;
                const bool interior = bc == FACE_INTERIOR && nb >= 0;
                uint32_t w[C::RS / 4];
                const uint4* row = reinterpret_cast<const uint4*>(M.bbNbr16 + (size_t)((interior ? (fl >> FLAG_MAP_SHIFT) : 0) * 4 + J) * C::RS);
#pragma unroll
                for (int k = 0; k < C::RS / 16; ++k) {
                    const uint4 v = __ldg(row + k);
                    w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
                }
                const double* const base = A.yin + (int64_t)(interior ? nb : 0) * (NP * 4);
                double tr[NFP][4];
#pragma unroll
                for (int b = 0; b < NFP; ++b) {
                    const uint32_t ci = (w[b >> 2] >> (8 * (b & 3))) & 0xffu;
                    double2 lo = make_double2(0.0, 0.0), hi = make_double2(0.0, 0.0);
                    if (interior) {
                        lo = *reinterpret_cast<const double2*>(base + ci * 4);
                        hi = *reinterpret_cast<const double2*>(base + ci * 4 + 2);
                    }
                    tr[b][0] = lo.x; tr[b][1] = lo.y; tr[b][2] = hi.x; tr[b][3] = hi.y;
                }
                const double nrm[3] = {fgOf[J][0], fgOf[J][1], fgOf[J][2]};
                const double v0n = ph.v0[0] * nrm[0] + ph.v0[1] * nrm[1] + ph.v0[2] * nrm[2];
                const bb::FaceCoef k = bb::faceCoef(bc, (fl & FLAG_TAU_NEG) ? -1.0 : 1.0, fgOf[J][3], v0n, ph.c0, ph.rho0);
                double x[4][NFP];
#pragma unroll
                for (int b = 0; b < NFP; ++b) {
                    const int oi = ownIndex<DIM, P, J>(b);
                    const double a[4] = {c[0][oi] - tr[b][0], c[1][oi] - tr[b][1], c[2][oi] - tr[b][2], c[3][oi] - tr[b][3]};
                    double x4[4];
                    bb::faceInput(k, nrm, a, x4);
                    x[0][b] = x4[0]; x[1][b] = x4[1]; x[2][b] = x4[2]; x[3][b] = x4[3];
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    double zl[NP];
                    SX::liftLocal(x[q], zl);
                    SX::template scatterAdd<J>(zl, out[q]);
                }
            }
        };
        doFace(std::integral_constant<int, 0>{});
        doFace(std::integral_constant<int, 1>{});
        doFace(std::integral_constant<int, 2>{});
        doFace(std::integral_constant<int, 3>{});

        // fused RK update in the shared-memory rows of the RK registers, then bulk stores
        if constexpr (kCoopMode) {
            cpWaitAllButOne();  // u and acc have landed (the next tile's stage input may still travel)
            __syncwarp();
        } else {
            if (loadU) { mbarWait2(&bars[1], phU); phU ^= 1; }
            if (loadA) { mbarWait2(&bars[2], phA); phA ^= 1; }
        }
        {
            double* const pu = sU + lane * ES;
            double* const pa = sA + lane * ES;
            const double dt = A.dt;
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                double uv[4] = {0.0, 0.0, 0.0, 0.0}, av[4] = {0.0, 0.0, 0.0, 0.0};
                if (loadU) {
                    const double2 lo = *reinterpret_cast<const double2*>(pu + i * 4), hi = *reinterpret_cast<const double2*>(pu + i * 4 + 2);
                    uv[0] = lo.x; uv[1] = lo.y; uv[2] = hi.x; uv[3] = hi.y;
                }
                if (loadA) {
                    const double2 lo = *reinterpret_cast<const double2*>(pa + i * 4), hi = *reinterpret_cast<const double2*>(pa + i * 4 + 2);
                    av[0] = lo.x; av[1] = lo.y; av[2] = hi.x; av[3] = hi.y;
                }
                double un[4], an[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double k = __dmul_rn(dt, out[q][i]);
                    switch (mode) {
                        case MODE_RK1: an[q] = k; un[q] = uv[q] + 0.5 * k; break;
                        case MODE_RK2: an[q] = av[q] + 2 * k; un[q] = uv[q] + 0.5 * k; break;
                        case MODE_RK3: an[q] = av[q] + 2 * k; un[q] = uv[q] + k; break;
                        case MODE_RK4: an[q] = 0.0; un[q] = fma(av[q] + k, 1.0 / 6.0, uv[q]); break;
                        case MODE_EULER: an[q] = 0.0; un[q] = uv[q] + k; break;
                        default: an[q] = 0.0; un[q] = out[q][i]; break;
                    }
                }
                *reinterpret_cast<double2*>(pu + i * 4) = make_double2(un[0], un[1]);
                *reinterpret_cast<double2*>(pu + i * 4 + 2) = make_double2(un[2], un[3]);
                if (storeA) {
                    *reinterpret_cast<double2*>(pa + i * 4) = make_double2(an[0], an[1]);
                    *reinterpret_cast<double2*>(pa + i * 4 + 2) = make_double2(an[2], an[3]);
                }
            }
        }
        if constexpr (kCoopMode) {
            __syncwarp();
            coopStore(uDst + (int64_t)e0 * NP * 4, sU, n);
            if (storeA) coopStore(A.acc + (int64_t)e0 * NP * 4, sA, n);
            if (!more) break;
            __syncwarp();  // every lane has read the rows it stores before the next tile's copies overwrite them
        } else {
            fenceProxyAsync();
            __syncwarp();
            storeTile(uDst + (int64_t)e0 * NP * 4, sU, n);
            if (storeA) storeTile(A.acc + (int64_t)e0 * NP * 4, sA, n);
            bulkCommit();
            if (!more) break;
            bulkWaitRead();  // the stores have read sU / sA (every lane waits for the copies it issued)
            __syncwarp();
        }
        t = tn;
    }
    if constexpr (kCoopMode) cpWaitAll(); else bulkWaitAll();
}

template <int DIM, int P>
void launchBBE(const DeviceMesh& M, const StageArgs& A, cudaStream_t s) {
    using C = BBECfg<DIM, P>;
    const int nEl = A.eEnd - A.eBegin;
    if (nEl <= 0) return;
    const int nTiles = (nEl + kTEE - 1) / kTEE;
#ifdef DGB_EMULATE
    const int grid = std::max(1, std::min(nTiles, 3));  // a few persistent "CTAs": every warp walks several tiles
#else
    static KernelConfig kc;
    static int perSm[kMaxDevices] = {};
    const int numSm = configureKernel(kc, stageBBEKernel<DIM, P>, C::SMEM, "stage_bbe");
    int dev = 0;
    cudaGetDevice(&dev);
    if (perSm[dev] == 0) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, stageBBEKernel<DIM, P>, 32, C::SMEM) != cudaSuccess || n < 1) {
            cudaGetLastError();
            throw UnsupportedError("stage_bbe: the kernel does not fit an SM of this device");
        }
        perSm[dev] = n;
    }
    const int grid = std::max(1, std::min(nTiles, (numSm - std::min(A.smReserve, numSm / 2)) * perSm[dev]));
#endif
    DGB_LAUNCH((stageBBEKernel<DIM, P>), grid, 32, C::SMEM, s, M, A, nTiles);
}

}  // namespace

StageKernel selectBBEKernel(int dim, int order) {
    StageKernel k;
#define DGB_BBE(D, P) if (dim == D && order == P) { k.launch = &launchBBE<D, P>; k.name = "stage_bbe<" #D "," #P ">"; }
    DGB_BBE(2, 1) DGB_BBE(2, 2) DGB_BBE(2, 3) DGB_BBE(3, 1) DGB_BBE(3, 2)
#undef DGB_BBE
    return k;
}

}  // namespace dgb
