// Bernstein-Bezier stage kernel for straight-sided tetrahedra (opt-in: dgb_set_option("kernel", 4)).
//
// The state arrays hold the BERNSTEIN COEFFICIENTS of the four fields (same layout u[eq][el*Np+n]; coefficient n belongs
// to the Bernstein function whose index is N times the barycentric position of the mesh's node n; dgb_set_state /
// dgb_get_state convert with V^-1 / V). In that basis the reference's dense per-element operators — the stiffness
// quadrature of Mesh::getElStiffVector (Mesh.cpp:476-489), the inverse mass of eigen::linEq (utils.cpp:118-123) and the
// face integrals of Mesh::precomputeFlux / getElFlux (Mesh.cpp:500-557) — collapse to the sparse closed forms of
// bb_ops.h: ~6.6 k FP64 instructions per element and stage at order 4 instead of ~36 kFLOP of dense contractions, which moves the
// order-4 stage from the FP64 roof to the HBM roof. CUDA cores only: there is nothing dense left.
//
// One CTA takes TE consecutive elements:
//   1. coalesced load of the four fields' coefficients into shared memory;
//   2. one task per (element, local face, face node): neighbour coefficient through the face-node map (the maps of the
//      nodal scheme apply unchanged), numerical flux exactly as in the nodal kernels (dgb_device.cuh: faceFlux), stored as
//      Fscale * (n.F(u-) - flux*) — the strong form, equal to the reference's weak form under exact quadrature;
//   3. one thread per (element, FIELD): directional derivatives, elevation and the four face lifts of bb_ops.h entirely
//      in registers (every index is a compile-time constant after unrolling), result back to shared memory;
//   4. coalesced fused RK update, as in the other kernels.
#include "dgb_internal.h"
#include "dgb_device.cuh"
#include "bb_ops.h"

// The CPU tests run THIS file through a small CUDA emulation (oracle/cuda_emu.h, oracle/bb_emulate.cpp: one OS thread per
// CUDA thread, a barrier for __syncthreads); the few constructs a host compiler cannot take go through dgb_launch.h.
#include "dgb_launch.h"

namespace dgb {

namespace {

__constant__ bb::Tables c_bbTables[bb::MAX_ORDER + 1];  // indexed by the order

// smallest stride >= n (in doubles) for which the 16 lanes (4 elements x 4 fields) of a half-warp hit 16 distinct 8-byte
// banks when lane (e, q) reads  q*stride + e*elemStride + const
__host__ __device__ constexpr int conflictFreeStride(int n, int elemStride) {
    for (int pad = 0; pad < 16; ++pad) {
        const int s = n + pad;
        bool ok = true;
        unsigned seen = 0;
        for (int q = 0; q < 4; ++q)
            for (int e = 0; e < 4; ++e) {
                const int b = (q * s + e * elemStride) % 16;
                if (seen & (1u << b)) ok = false;
                seen |= 1u << b;
            }
        if (ok) return s;
    }
    return n;
}

template <int P, int TE_>
struct BBCfg {
    static constexpr int NP = bb::tet(P), NFP = bb::tri(P), NFL = 4 * NFP;
    static constexpr int TE = TE_;          // elements per CTA: 32 (4 warps), 16, or 8 (one warp: the barriers become warp-local)
    static constexpr int THREADS = 4 * TE;  // == (element, face) pairs == (element, field) pairs of a tile
    static constexpr int SQ = conflictFreeStride(TE * NP, NP);    // field stride of the coefficient tile
    static constexpr int SF = conflictFreeStride(TE * NFL, NFL);  // field stride of the face-input tile
    static constexpr int FC = 8;                                  // doubles per (element, face): 5 coefficients + normal
    static constexpr size_t SMEM = (size_t)(4 * (SQ + SF) + THREADS * FC) * sizeof(double) + (size_t)2 * THREADS * sizeof(int);
};

__device__ __forceinline__ void cpAsync8(double* smemDst, const double* gmemSrc) {
#ifdef DGB_EMULATE
    *smemDst = *gmemSrc;
#else
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smemDst);
    const size_t src = __cvta_generic_to_global(gmemSrc);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
#endif
}
__device__ __forceinline__ void cpAsyncCommit() {
#ifndef DGB_EMULATE
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
__device__ __forceinline__ void cpAsyncWaitAll() {
#ifndef DGB_EMULATE
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
}

template <int P, int TE_>
__global__ void __launch_bounds__(BBCfg<P, TE_>::THREADS) stageBBKernel(DeviceMesh M, StageArgs A) {
    using C = BBCfg<P, TE_>;
    constexpr int NP = C::NP, NFP = C::NFP, NFL = C::NFL, TE = C::TE;
    static_assert(bb::BC_INTERIOR == FACE_INTERIOR && bb::BC_ABSORBING == FACE_ABSORBING && bb::BC_REFLECTING == FACE_REFLECTING, "face codes");
    DGB_DYNAMIC_SMEM(double, smem);
    double* sQ = smem;                       // [4][SQ]   coefficients of the tile
    double* sFl = smem + 4 * C::SQ;          // [4][SF]   face inputs of the lift, later the result
    double* sFc = sFl + 4 * C::SF;           // [TE*4][8] per (element, face): app, aps, b, c, d, n
    int* sNbr = reinterpret_cast<int*>(sFc + C::THREADS * C::FC);  // [TE*4] first coefficient of the neighbour element, -1 on the boundary
    int* sMap = sNbr + C::THREADS;                                 // [TE*4] offset of the face-node map

    const int tid = threadIdx.x;
    const int e0 = A.eBegin + blockIdx.x * TE;
    const int nE = min(TE, A.eEnd - e0);
    const int64_t S = M.stride;
    const Phys ph = makePhys(M);

    // 1. own coefficients: asynchronous copies, all in flight at once (coalesced, 8 bytes each: an element starts on an 8-byte boundary only)
    for (int i = tid; i < nE * NP; i += C::THREADS) {
        const int64_t g = (int64_t)e0 * NP + i;
#pragma unroll
        for (int q = 0; q < 4; ++q) cpAsync8(&sQ[q * C::SQ + i], &A.yin[q * S + g]);
    }
    cpAsyncCommit();

    // 1b. one thread per (element, local face): the face-constant coefficients of the lift input (bb_ops.h), once per face
    if (tid < 4 * nE) {
        const int e = e0 + (tid >> 2), lf = tid & 3;
        const int flags = M.fflags[e * 4 + lf];
        const int bc = flags & FLAG_BC_MASK;
        const double* fg = M.fgeo + ((int64_t)e * 4 + lf) * 4;
        const double n0 = fg[0], n1 = fg[1], n2 = fg[2];
        const double v0n = ph.v0[0] * n0 + ph.v0[1] * n1 + ph.v0[2] * n2;
        const bb::FaceCoef k = bb::faceCoef(bc, (flags & FLAG_TAU_NEG) ? -1.0 : 1.0, fg[3], v0n, ph.c0, ph.rho0);
        double* fc = sFc + tid * C::FC;
        fc[0] = k.app; fc[1] = k.aps; fc[2] = k.b; fc[3] = k.c; fc[4] = k.d; fc[5] = n0; fc[6] = n1; fc[7] = n2;
        sNbr[tid] = bc == FACE_INTERIOR ? M.fnbr[e * 4 + lf] * NP : -1;
        sMap[tid] = (flags >> FLAG_MAP_SHIFT) * NFP;
    }
    cpAsyncWaitAll();
    __syncthreads();

    // 2. face inputs of the lift, one task per (element, local face, face node): jump against the neighbour's coefficient
    //    (same face-node maps as the nodal scheme) or the own trace on a boundary face; unrolled so that the gathers of
    //    several tasks are in flight together
#pragma unroll 3
    for (int w = tid; w < nE * NFL; w += C::THREADS) {
        const int el = w / NFL, r = w - el * NFL, lf = r / NFP, m = r - lf * NFP;
        const int ef = el * 4 + lf;
        const double* fc = sFc + ef * C::FC;
        const bb::FaceCoef k = {fc[0], fc[1], fc[2], fc[3], fc[4]};
        const double n[3] = {fc[5], fc[6], fc[7]};
        const int own = M.faceNodes[lf * NFP + m];
        double a[4], x[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) a[q] = sQ[q * C::SQ + el * NP + own];
        const int nb = sNbr[ef];
        if (nb >= 0) {
            const int64_t gi = (int64_t)nb + M.nbrMaps[sMap[ef] + m];
#pragma unroll
            for (int q = 0; q < 4; ++q) a[q] -= A.yin[q * S + gi];
        }
        bb::faceInput(k, n, a, x);
#pragma unroll
        for (int q = 0; q < 4; ++q) sFl[q * C::SF + el * NFL + r] = x[q];
    }
    __syncthreads();

    // 3. one thread per (element, field): sparse Bernstein operators in registers
    {
        const int el = tid >> 2, q = tid & 3;
        if (el < nE) {
            const int e = e0 + el;
            const double* G = M.Ginv + (int64_t)e * 9;  // G[x*3+u] = d u_u / d x_x ; lambda_0 = 1 - u_0 - u_1 - u_2
            double gl[4][3];
#pragma unroll
            for (int x = 0; x < 3; ++x) {
                const double g0 = G[x * 3 + 0], g1 = G[x * 3 + 1], g2 = G[x * 3 + 2];
                gl[0][x] = -(g0 + g1 + g2);
                gl[1][x] = g0;
                gl[2][x] = g1;
                gl[3][x] = g2;
            }
            double* mine = sFl + q * C::SF + el * NFL;
            const bool flow = ph.v0[0] != 0.0 || ph.v0[1] != 0.0 || ph.v0[2] != 0.0;
            double out[NP];
            const bb::Tables& T = c_bbTables[P];
            bb::fieldRhs<P>(q, sQ + el * NP, C::SQ, mine, T, gl, ph.v0, flow, ph.rc2, ph.invRho, out);
            // the face inputs of this (element, field) are consumed: its slot now carries the result, mesh node order
#pragma unroll
            for (int i = 0; i < NP; ++i) mine[T.permC2G[i]] = out[i];
        }
    }
    __syncthreads();

    // 4. fused RK update, coalesced; per field, all RK-register loads of a thread are issued before the first use (the tile's
    //    latency is paid once, not once per entry)
    {
        constexpr int NIT = (TE * NP + C::THREADS - 1) / C::THREADS;
        const bool rkRegs = A.mode == MODE_RK2 || A.mode == MODE_RK3 || A.mode == MODE_RK4;
        const int cnt = nE * NP;
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
            const int64_t gq = q * S + (int64_t)e0 * NP;
            double uv[NIT], av[NIT];
#pragma unroll
            for (int k = 0; k < NIT; ++k) {
                const int i = tid + k * C::THREADS;
                uv[k] = av[k] = 0.0;
                if (i < cnt) {
                    uv[k] = rkRegs ? A.u[gq + i] : sQ[q * C::SQ + i];  // first stage / Euler: the stage input is the base state
                    if (rkRegs) av[k] = A.acc[gq + i];
                }
            }
#pragma unroll
            for (int k = 0; k < NIT; ++k) {
                const int i = tid + k * C::THREADS;
                if (i < cnt) {
                    const int el = i / NP, nd = i - el * NP;
                    rkApply(A, gq + i, sFl[q * C::SF + el * NFL + nd], uv[k], av[k]);
                }
            }
        }
    }
}

// ---- variant B: the four faces one after the other --------------------------------------------------------------------------
// Same arithmetic, different schedule: the face inputs of ONE face at a time live in shared memory (15 instead of 60 values
// per element and field at order 4), so a tile needs about half the shared memory and more CTAs fit an SM; the lift body is
// shared by the four faces (a run-time loop; only the scatter into the element's coefficients is face-specific), which also
// shrinks the code. Costs two more barriers per face. dgb_set_option("bb_variant", 1).
template <int P, int TE_>
struct BBSeqCfg {
    static constexpr int NP = bb::tet(P), NFP = bb::tri(P);
    static constexpr int TE = TE_, THREADS = 4 * TE;
    static constexpr int SQ = conflictFreeStride(TE * NP, NP);
    static constexpr int SX = conflictFreeStride(TE * NFP, NFP);  // field stride of the one-face input tile
    static constexpr int FC = 8;
    static constexpr size_t SMEM = (size_t)(4 * (SQ + SX) + THREADS * FC) * sizeof(double) + (size_t)2 * THREADS * sizeof(int);
};

template <int P, int TE_>
__global__ void __launch_bounds__(BBSeqCfg<P, TE_>::THREADS) stageBBSeqKernel(DeviceMesh M, StageArgs A) {
    using C = BBSeqCfg<P, TE_>;
    constexpr int NP = C::NP, NFP = C::NFP, TE = C::TE;
    DGB_DYNAMIC_SMEM(double, smem);
    double* sQ = smem;                       // [4][SQ]   coefficients of the tile, at the end the result
    double* sX = smem + 4 * C::SQ;           // [4][SX]   lift inputs of the current face, canonical 2D order
    double* sFc = sX + 4 * C::SX;            // [TE*4][8] per (element, local face): app, aps, b, c, d, n
    int* sNbr = reinterpret_cast<int*>(sFc + C::THREADS * C::FC);
    int* sMap = sNbr + C::THREADS;

    const int tid = threadIdx.x;
    const int e0 = A.eBegin + blockIdx.x * TE;
    const int nE = min(TE, A.eEnd - e0);
    const int64_t S = M.stride;
    const Phys ph = makePhys(M);
    const bb::Tables& T = c_bbTables[P];

    for (int i = tid; i < nE * NP; i += C::THREADS) {
        const int64_t g = (int64_t)e0 * NP + i;
#pragma unroll
        for (int q = 0; q < 4; ++q) cpAsync8(&sQ[q * C::SQ + i], &A.yin[q * S + g]);
    }
    cpAsyncCommit();
    if (tid < 4 * nE) {
        const int e = e0 + (tid >> 2), lf = tid & 3;
        const int flags = M.fflags[e * 4 + lf];
        const int bc = flags & FLAG_BC_MASK;
        const double* fg = M.fgeo + ((int64_t)e * 4 + lf) * 4;
        const double n0 = fg[0], n1 = fg[1], n2 = fg[2];
        const double v0n = ph.v0[0] * n0 + ph.v0[1] * n1 + ph.v0[2] * n2;
        const bb::FaceCoef k = bb::faceCoef(bc, (flags & FLAG_TAU_NEG) ? -1.0 : 1.0, fg[3], v0n, ph.c0, ph.rho0);
        double* fc = sFc + tid * C::FC;
        fc[0] = k.app; fc[1] = k.aps; fc[2] = k.b; fc[3] = k.c; fc[4] = k.d; fc[5] = n0; fc[6] = n1; fc[7] = n2;
        sNbr[tid] = bc == FACE_INTERIOR ? M.fnbr[e * 4 + lf] * NP : -1;
        sMap[tid] = (flags >> FLAG_MAP_SHIFT) * NFP;
    }
    cpAsyncWaitAll();
    __syncthreads();

    // volume term of this thread's (element, field), in registers for the rest of the tile
    const int elT = tid >> 2, qT = tid & 3;
    const bool mineActive = elT < nE;
    double out[NP];
    if (mineActive) {
        const double* G = M.Ginv + (int64_t)(e0 + elT) * 9;
        double gl[4][3];
#pragma unroll
        for (int x = 0; x < 3; ++x) {
            const double g0 = G[x * 3 + 0], g1 = G[x * 3 + 1], g2 = G[x * 3 + 2];
            gl[0][x] = -(g0 + g1 + g2);
            gl[1][x] = g0;
            gl[2][x] = g1;
            gl[3][x] = g2;
        }
        const bool flow = ph.v0[0] != 0.0 || ph.v0[1] != 0.0 || ph.v0[2] != 0.0;
        bb::fieldVolume<P>(qT, sQ + elT * NP, C::SQ, T, gl, ph.v0, flow, ph.rc2, ph.invRho, out);
    }

#pragma unroll 1
    for (int J = 0; J < 4; ++J) {
        const int lf = T.faceLf[J];
        // lift inputs of face J: one task per (element, canonical face index b)
        for (int w = tid; w < nE * NFP; w += C::THREADS) {
            const int el = w / NFP, b = w - el * NFP;
            const int m = T.facePos[J][b];  // position in the mesh's face-node list of local face lf
            const int ef = el * 4 + lf;
            const double* fc = sFc + ef * C::FC;
            const bb::FaceCoef k = {fc[0], fc[1], fc[2], fc[3], fc[4]};
            const double n[3] = {fc[5], fc[6], fc[7]};
            const int own = M.faceNodes[lf * NFP + m];
            double a[4], x[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) a[q] = sQ[q * C::SQ + el * NP + own];
            const int nb = sNbr[ef];
            if (nb >= 0) {
                const int64_t gi = (int64_t)nb + M.nbrMaps[sMap[ef] + m];
#pragma unroll
                for (int q = 0; q < 4; ++q) a[q] -= A.yin[q * S + gi];
            }
            bb::faceInput(k, n, a, x);
#pragma unroll
            for (int q = 0; q < 4; ++q) sX[q * C::SX + el * NFP + b] = x[q];
        }
        __syncthreads();
        if (mineActive) {
            double x[NFP], zl[NP];
            const double* mine = sX + qT * C::SX + elT * NFP;
#pragma unroll
            for (int b = 0; b < NFP; ++b) x[b] = mine[b];
            bb::liftFaceLocal<P>(x, zl);
            switch (J) {  // warp-uniform
                case 0: bb::scatterAddFace<P, 0>(zl, out); break;
                case 1: bb::scatterAddFace<P, 1>(zl, out); break;
                case 2: bb::scatterAddFace<P, 2>(zl, out); break;
                default: bb::scatterAddFace<P, 3>(zl, out); break;
            }
        }
        __syncthreads();  // sX is rewritten by the next face; after the last face nobody reads the coefficient tile any more
    }

    // result into the thread's own column of the coefficient tile (mesh node order)
    if (mineActive) {
        double* col = sQ + qT * C::SQ + elT * NP;
#pragma unroll
        for (int i = 0; i < NP; ++i) col[T.permC2G[i]] = out[i];
    }
    __syncthreads();

    // fused RK update, coalesced; per field, all global loads of a thread are issued before the first use. The stage input
    // (overwritten in the tile by the result) is read again only where the update needs it: first RK stage, Euler.
    {
        constexpr int NIT = (TE * NP + C::THREADS - 1) / C::THREADS;
        const bool rkRegs = A.mode == MODE_RK2 || A.mode == MODE_RK3 || A.mode == MODE_RK4;
        const bool needY = A.mode == MODE_RK1 || A.mode == MODE_EULER;
        const int cnt = nE * NP;
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
            const int64_t gq = q * S + (int64_t)e0 * NP;
            double uv[NIT], av[NIT];
#pragma unroll
            for (int k = 0; k < NIT; ++k) {
                const int i = tid + k * C::THREADS;
                uv[k] = av[k] = 0.0;
                if (i < cnt) {
                    if (rkRegs) { uv[k] = A.u[gq + i]; av[k] = A.acc[gq + i]; }
                    else if (needY) uv[k] = A.yin[gq + i];
                }
            }
#pragma unroll
            for (int k = 0; k < NIT; ++k) {
                const int i = tid + k * C::THREADS;
                if (i < cnt) rkApply(A, gq + i, sQ[q * C::SQ + i], uv[k], av[k]);
            }
        }
    }
}

template <int P, int TE_>
void launchBBSeq(const DeviceMesh& M, const StageArgs& A, cudaStream_t s) {
    using C = BBSeqCfg<P, TE_>;
    const int nEl = A.eEnd - A.eBegin;
    if (nEl <= 0) return;
#ifndef DGB_EMULATE
    static KernelConfig kc;
    configureKernel(kc, stageBBSeqKernel<P, TE_>, C::SMEM, "stage_bb_seq");
#endif
    DGB_LAUNCH((stageBBSeqKernel<P, TE_>), (nEl + C::TE - 1) / C::TE, C::THREADS, C::SMEM, s, M, A);
}

template <int P, int TE_>
void launchBB(const DeviceMesh& M, const StageArgs& A, cudaStream_t s) {
    using C = BBCfg<P, TE_>;
    const int nEl = A.eEnd - A.eBegin;
    if (nEl <= 0) return;
#ifndef DGB_EMULATE
    static KernelConfig kc;
    configureKernel(kc, stageBBKernel<P, TE_>, C::SMEM, "stage_bb");
#endif
    DGB_LAUNCH((stageBBKernel<P, TE_>), (nEl + C::TE - 1) / C::TE, C::THREADS, C::SMEM, s, M, A);
}

// y = Mat x per element and field (nodal <-> Bernstein conversion of a whole state array); in and out may alias
__global__ void __launch_bounds__(256) elementMatrixKernel(const double* in, double* out, int64_t stride, int Np, int K, const double* __restrict__ mat) {
    DGB_DYNAMIC_SMEM(double, sx);  // [4][E*Np]
    const int E = 256 / Np;
    const int e0 = blockIdx.x * E;
    const int nE = min(E, K - e0);
    const int tid = threadIdx.x;
    const bool active = tid < nE * Np;
    const int64_t g = (int64_t)e0 * Np + tid;
    if (active)
        for (int q = 0; q < 4; ++q) sx[q * E * Np + tid] = in[q * stride + g];
    __syncthreads();
    if (!active) return;
    const int el = tid / Np, n = tid - el * Np;
    double acc[4] = {0, 0, 0, 0};
    for (int m = 0; m < Np; ++m) {
        const double a = mat[n * Np + m];
        for (int q = 0; q < 4; ++q) acc[q] = fma(a, sx[q * E * Np + el * Np + m], acc[q]);
    }
    for (int q = 0; q < 4; ++q) out[q * stride + g] = acc[q];
}

// Hard source in Bernstein mode (solver.cpp:248-256 sets NODAL values): one CTA per element that carries source nodes.
// With delta_n = value - (V c)_n at the element's source nodes n (all taken from the coefficients before the update, the
// overwrites of different nodes commute),  c += sum_n Vinv[:, n] delta_n  makes (V c)_n = value there and leaves every other
// nodal value of the element unchanged.
__global__ void __launch_bounds__(64) setNodesBBKernel(double* field, int Np, const int32_t* elList, const int32_t* nodeOff, const int32_t* nodeLocal,
                                                       double value, const double* __restrict__ V, const double* __restrict__ Vinv, int coefStride,
                                                       const uint8_t* __restrict__ perm) {
    __shared__ double c[bb::MAX_NP], delta[bb::MAX_NP];
    const int b = blockIdx.x, t = threadIdx.x;
    const int64_t base = (int64_t)elList[b] * Np;
    const int n0 = nodeOff[b], nn = nodeOff[b + 1] - n0;
    auto at = [&](int m) -> double& { return field[(base + (perm ? perm[m] : m)) * coefStride]; };
    for (int m = t; m < Np; m += blockDim.x) c[m] = at(m);
    __syncthreads();
    for (int k = t; k < nn; k += blockDim.x) {
        const int n = nodeLocal[n0 + k];
        double cur = 0.0;
        for (int m = 0; m < Np; ++m) cur = fma(V[n * Np + m], c[m], cur);
        delta[k] = value - cur;
    }
    __syncthreads();
    for (int m = t; m < Np; m += blockDim.x) {
        double s = 0.0;
        for (int k = 0; k < nn; ++k) s = fma(Vinv[m * Np + nodeLocal[n0 + k]], delta[k], s);
        at(m) = c[m] + s;
    }
}

}  // namespace

StageKernel selectBBKernel(int dim, int order, int variant, int tile) {
    StageKernel k;
    if (dim != 3 || (tile != 8 && tile != 16 && tile != 32)) return k;
#define DGB_CASE_T(P, T)                                                                                        \
    if (order == P && tile == T) {                                                                              \
        if (variant == 1) { k.launch = &launchBBSeq<P, T>; k.name = "stage_bb_seq<3," #P ">/" #T; }             \
        else { k.launch = &launchBB<P, T>; k.name = "stage_bb<3," #P ">/" #T; }                                 \
        return k;                                                                                               \
    }
#define DGB_CASE(P) DGB_CASE_T(P, 32) DGB_CASE_T(P, 16) DGB_CASE_T(P, 8)
    DGB_CASE(2) DGB_CASE(3) DGB_CASE(4) DGB_CASE(5)
#undef DGB_CASE
#undef DGB_CASE_T
    return k;
}

void setBBTables(int order, const bb::Tables& T) {
    if (order < 0 || order > bb::MAX_ORDER) return;
#ifdef DGB_EMULATE
    c_bbTables[order] = T;
#else
    cudaMemcpyToSymbol(c_bbTables, &T, sizeof(T), (size_t)order * sizeof(bb::Tables), cudaMemcpyHostToDevice);
#endif
}

void launchSetNodesBB(double* field, int Np, const int32_t* elList, const int32_t* nodeOff, const int32_t* nodeLocal, int nEl, double value,
                      const double* V, const double* Vinv, cudaStream_t s, int coefStride, const uint8_t* perm) {
    if (nEl > 0) DGB_LAUNCH(setNodesBBKernel, nEl, 64, 0, s, field, Np, elList, nodeOff, nodeLocal, value, V, Vinv, coefStride, perm);
}

void launchElementMatrix(const double* in, double* out, int64_t stride, int Np, int K, const double* mat, cudaStream_t s) {
    if (K <= 0) return;
    const int E = 256 / Np;
    DGB_LAUNCH(elementMatrixKernel, (K + E - 1) / E, 256, (size_t)4 * E * Np * sizeof(double), s, in, out, stride, Np, K, mat);
}

}  // namespace dgb
