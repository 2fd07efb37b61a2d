// Stage kernel for curved (non-affine) elements, any dimension / order.
//
// On curved elements nothing of the reference's scheme collapses (the quadrature is not exact, every element has its own mass
// matrix), so this kernel evaluates the reference's own loops, one CTA per element and all four fields at once:
//   Mesh::updateFlux        Mesh.cpp:569-674   nodal physical flux; ghost flux at the face integration points (both BCs)
//   Mesh::getElStiffVector  Mesh.cpp:476-489   S_i = sum_g w_g detJ(g) grad phi_i(g) . (sum_j phi_j(g) Flux_j)   (regrouped)
//   Mesh::precomputeFlux    Mesh.cpp:500-539   face integrals with the normal and surface Jacobian of every point
//   Mesh::getElFlux         Mesh.cpp:548-557   scatter with the orientation sign
//   eigen::minus / linEq    utils.cpp:118-136  k = dt * M_el^-1 (S - F), then the fused RK update of the other kernels
// on the reference's own tables (one Jacobian / normal per integration point, CurvedMesh in curved_setup.h). A face shared
// by two elements is integrated by both with the same arithmetic (the face's up / down sides, not the element's), so both
// see the same value. Correctness first: this path exists for the thin layer of curved elements along curved boundaries;
// straight-sided elements keep the collapsed kernels. The CPU tests run this file through oracle/cuda_emu.h.
#include "curved_setup.h"
#include "dgb_internal.h"
#include "dgb_device.cuh"

#include "dgb_launch.h"

namespace dgb {

namespace {

#ifdef DGB_EMULATE
constexpr int CURVED_THREADS = 8;  // the emulation pays per OS thread and barrier; the task loops do not care
#else
constexpr int CURVED_THREADS = 128;
#endif

// physical flux of field q (0 p, 1..3 velocity) in direction x at one node (Mesh.cpp:577-591)
__device__ __forceinline__ double nodalFlux(int q, int x, const double u[4], const double v0[3], double rc2, double rho0) {
    if (q == 0) return v0[x] * u[0] + rc2 * u[1 + x];
    return q - 1 == x ? v0[x] * u[q] + u[0] / rho0 : v0[x] * u[q];
}

__global__ void __launch_bounds__(CURVED_THREADS) stageCurvedKernel(CurvedMesh C, StageArgs A) {
    DGB_DYNAMIC_SMEM(double, smem);
    const int Np = C.Np, Nfp = C.Nfp, Nf = C.Nf, nG = C.nG, nGf = C.nGf, dim = C.dim;
    double* sU = smem;                    // [4][Np]      stage input of the element
    double* sFx = sU + 4 * Np;            // [4][Np][3]   nodal physical flux
    double* sJinv = sFx + 12 * Np;        // [nG][9]      (dx/du)^-1 at the integration points, index x*3+u
    double* sWd = sJinv + 9 * nG;         // [nG]         w_g detJ(g)
    double* sFg = sWd + nG;               // [nG][12]     flux interpolated to the integration points, index q*3+x
    double* sS = sFg + 12 * nG;           // [4][Np]      S - F
    double* sSF = sS + 4 * Np;            // [Nfp][4][4]  per face node and field: (F_up + F_dn)[0..2], u_up - u_dn
    double* sFI = sSF + 16 * Nfp;         // [nGf][4]     FIntPts per field

    const int tid = threadIdx.x;
    const int el = A.eBegin + blockIdx.x;
    const int64_t S = C.stride;
    const double rc2 = C.rho0 * C.c0 * C.c0;

    for (int i = tid; i < 4 * Np; i += CURVED_THREADS) {
        const int q = i / Np, n = i - q * Np;
        sU[i] = A.yin[q * S + (int64_t)el * Np + n];
    }
    for (int g = tid; g < nG; g += CURVED_THREADS) {
        const double* J = C.elJac + ((int64_t)el * nG + g) * 9;  // J[u*3+x] = dx_x/du_u ; the reference solves the dim x dim block
        double B[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};               // B[x*3+u] = du_u/dx_x
        if (dim == 1) B[0] = 1.0 / J[0];
        else if (dim == 2) {
            const double det = J[0] * J[4] - J[1] * J[3];
            B[0] = J[4] / det; B[1] = -J[1] / det; B[3] = -J[3] / det; B[4] = J[0] / det;
        } else {
            const double det = J[0] * (J[4] * J[8] - J[5] * J[7]) - J[1] * (J[3] * J[8] - J[5] * J[6]) + J[2] * (J[3] * J[7] - J[4] * J[6]);
            B[0] = (J[4] * J[8] - J[5] * J[7]) / det; B[1] = (J[2] * J[7] - J[1] * J[8]) / det; B[2] = (J[1] * J[5] - J[2] * J[4]) / det;
            B[3] = (J[5] * J[6] - J[3] * J[8]) / det; B[4] = (J[0] * J[8] - J[2] * J[6]) / det; B[5] = (J[2] * J[3] - J[0] * J[5]) / det;
            B[6] = (J[3] * J[7] - J[4] * J[6]) / det; B[7] = (J[1] * J[6] - J[0] * J[7]) / det; B[8] = (J[0] * J[4] - J[1] * J[3]) / det;
        }
        for (int k = 0; k < 9; ++k) sJinv[g * 9 + k] = B[k];
        sWd[g] = C.elWeight[g] * C.elDet[(int64_t)el * nG + g];
    }
    __syncthreads();

    // nodal physical flux of the element
    for (int i = tid; i < 12 * Np; i += CURVED_THREADS) {
        const int q = i / (3 * Np), r = i - q * 3 * Np, n = r / 3, x = r - n * 3;
        const double u[4] = {sU[n], sU[Np + n], sU[2 * Np + n], sU[3 * Np + n]};
        sFx[(q * Np + n) * 3 + x] = nodalFlux(q, x, u, C.v0, rc2, C.rho0);
    }
    __syncthreads();

    // flux at the integration points, then S_i = sum_g w_g detJ(g) grad phi_i(g) . Flux(g)
    for (int i = tid; i < 12 * nG; i += CURVED_THREADS) {
        const int g = i / 12, c = i - g * 12, q = c / 3, x = c - q * 3;
        double s = 0.0;
        for (int j = 0; j < Np; ++j) s += C.elBasis[g * Np + j] * sFx[(q * Np + j) * 3 + x];
        sFg[i] = s;
    }
    __syncthreads();
    for (int i = tid; i < 4 * Np; i += CURVED_THREADS) {
        const int q = i / Np, n = i - q * Np;
        double s = 0.0;
        for (int g = 0; g < nG; ++g) {
            const double* ug = C.elUGrad + ((int64_t)g * Np + n) * 3;
            const double* B = sJinv + g * 9;
            double dot = 0.0;
            for (int x = 0; x < dim; ++x) {
                double gx = 0.0;
                for (int u = 0; u < dim; ++u) gx += B[x * 3 + u] * ug[u];
                dot += gx * sFg[g * 12 + q * 3 + x];
            }
            s += sWd[g] * dot;
        }
        sS[i] = s;
    }
    __syncthreads();

    // faces, one after the other (all threads take the same path: the face data are per element)
    for (int lf = 0; lf < Nf; ++lf) {
        const int f = C.elFId[(int64_t)el * Nf + lf];
        const int up = C.fNbrElId[2 * (int64_t)f], dn = C.fNbrElId[2 * (int64_t)f + 1];
        const int side = up == el ? 0 : 1;
        const double orient = C.elFOrientation[(int64_t)el * Nf + lf];
        const bool boundary = C.fIsBoundary[f] != 0;
        if (!boundary) {
            for (int i = tid; i < Nfp; i += CURVED_THREADS) {
                const int64_t nu = (int64_t)up * Np + C.fNToElNId[((int64_t)f * Nfp + i) * 2];
                const int64_t nd = (int64_t)dn * Np + C.fNToElNId[((int64_t)f * Nfp + i) * 2 + 1];
                double uu[4], ud[4];
                for (int q = 0; q < 4; ++q) { uu[q] = A.yin[q * S + nu]; ud[q] = A.yin[q * S + nd]; }
                for (int q = 0; q < 4; ++q) {
                    for (int x = 0; x < 3; ++x) sSF[(i * 4 + q) * 4 + x] = nodalFlux(q, x, uu, C.v0, rc2, C.rho0) + nodalFlux(q, x, ud, C.v0, rc2, C.rho0);
                    sSF[(i * 4 + q) * 4 + 3] = uu[q] - ud[q];
                }
            }
            __syncthreads();
            for (int t = tid; t < 4 * nGf; t += CURVED_THREADS) {
                const int g = t >> 2, q = t & 3;
                const double* n = C.fNormal + ((int64_t)f * nGf + g) * 3;
                double s = 0.0;
                for (int i = 0; i < Nfp; ++i) {
                    const double* sf = sSF + (i * 4 + q) * 4;
                    double dot = 0.0;
                    for (int x = 0; x < 3; ++x) dot += n[x] * (0.5 * (sf[x] + C.fc * C.c0 * n[x] * sf[3]));  // n . Fnum, Mesh.cpp:519-527
                    s += dot * C.fBasis[g * Nfp + i];
                }
                sFI[t] = s;
            }
        } else {
            for (int g = tid; g < nGf; g += CURVED_THREADS) {
                double ug[4] = {0, 0, 0, 0};
                for (int n = 0; n < Nfp; ++n) {
                    const int node = C.fNToElNId[((int64_t)f * Nfp + n) * 2];
                    const double b = C.fBasis[g * Nfp + n];
                    for (int q = 0; q < 4; ++q) ug[q] += sU[q * Np + node] * b;
                }
                const double* n = C.fNormal + ((int64_t)f * nGf + g) * 3;
                if (C.fBC[f] == 1) {  // reflecting: physical flux of the wall-tangent ghost state, projected on n (Mesh.cpp:616-648)
                    const double dot = n[0] * ug[1] + n[1] * ug[2] + n[2] * ug[3];
                    ug[1] -= dot * n[0]; ug[2] -= dot * n[1]; ug[3] -= dot * n[2];
                    for (int q = 0; q < 4; ++q)
                        sFI[g * 4 + q] = n[0] * nodalFlux(q, 0, ug, C.v0, rc2, C.rho0) + n[1] * nodalFlux(q, 1, ug, C.v0, rc2, C.rho0) +
                                         n[2] * nodalFlux(q, 2, ug, C.v0, rc2, C.rho0);
                } else {  // absorbing: RKR rows (Mesh.cpp:391-418, 652-667)
                    const double vn = n[0] * ug[1] + n[1] * ug[2] + n[2] * ug[3];
                    sFI[g * 4] = 0.25 * C.c0 * ug[0] + 0.25 * C.c0 * C.c0 * C.rho0 * vn;
                    for (int x = 0; x < 3; ++x) sFI[g * 4 + 1 + x] = 0.25 * n[x] / C.rho0 * ug[0] + 0.25 * C.c0 * n[x] * vn;
                }
            }
        }
        __syncthreads();
        for (int t = tid; t < 4 * Nfp; t += CURVED_THREADS) {
            const int n = t >> 2, q = t & 3;
            double s = 0.0;
            for (int g = 0; g < nGf; ++g) s += C.fWeight[g] * C.fBasis[g * Nfp + n] * sFI[g * 4 + q] * C.fDet[(int64_t)f * nGf + g];
            const int node = C.fNToElNId[((int64_t)f * Nfp + n) * 2 + side];
            sS[q * Np + node] -= orient * s;  // eigen::minus of getElFlux's scatter; one task per (face node, field): no two tasks share an entry
        }
        __syncthreads();
    }

    // k = dt * M_el^-1 (S - F) (the row-major inverse read column-major like the reference's Eigen::Map), fused RK update
    const double* Mi = C.Minv + (int64_t)(el - C.firstCurved) * Np * Np;
    for (int i = tid; i < 4 * Np; i += CURVED_THREADS) {
        const int q = i / Np, n = i - q * Np;
        double s = 0.0;
        for (int j = 0; j < Np; ++j) s += Mi[j * Np + n] * sS[q * Np + j];
        rkUpdate(A, q * S + (int64_t)el * Np + n, s, sU[i]);
    }
}

}  // namespace

size_t curvedSmemBytes(const CurvedMesh& C) {
    return sizeof(double) * ((size_t)4 * C.Np + 12 * C.Np + 9 * C.nG + C.nG + 12 * C.nG + 4 * C.Np + 16 * C.Nfp + 4 * C.nGf);
}

void launchCurved(const CurvedMesh& C, const StageArgs& A, cudaStream_t s) {
    const int nEl = A.eEnd - A.eBegin;
    if (nEl <= 0) return;
    const size_t bytes = curvedSmemBytes(C);
#ifndef DGB_EMULATE
    static KernelConfig kc;
    configureKernel(kc, stageCurvedKernel, bytes, "stage_curved");
#endif
    DGB_LAUNCH(stageCurvedKernel, nEl, CURVED_THREADS, bytes, s, C, A);
}

}  // namespace dgb
