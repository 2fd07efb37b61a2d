// Curved (non-affine) elements, SURVEY.md §8 f3: what the curved stage kernel (stage_curved.cu) reads. On curved elements the
// reference's quadrature is not exact and the element mass matrices are not multiples of the reference one, so nothing
// collapses: the kernel evaluates the reference's own loops (Mesh.cpp:476-489, 500-557, 569-674, utils.cpp:118-123) on the
// reference's own tables — one Jacobian / normal per integration point — plus the per-element inverse mass matrices that
// Mesh::precomputeMassMatrix (Mesh.cpp:440-466) builds. Plain C++: used by dgb_api.cu and by the CPU emulation harness.
#pragma once
#include <stdint.h>

#include <cmath>
#include <stdexcept>
#include <vector>

#include "../../include/dgb.h"

namespace dgb {

// device (or, under emulation, host) pointers + sizes, passed to the kernel by value
struct CurvedMesh {
    int dim, Np, Nfp, Nf, K, F, nG, nGf, fc;
    const double *elBasis, *elUGrad, *elWeight, *fBasis, *fWeight;  // [nG][Np], [nG][Np][3], [nG], [nGf][Nfp], [nGf]
    const double *elJac, *elDet;                                    // [K][nG][9] (u*3+x), [K][nG]
    const double *fNormal, *fDet;                                   // [F][nGf][3], [F][nGf]
    const int32_t *elFId, *elFOrientation, *fNbrElId, *fNToElNId;   // [K][Nf], [K][Nf], [F][2], [F][Nfp][2]
    const uint8_t* fIsBoundary;                                     // [F]
    const int32_t* fBC;                                             // [F]
    const double* Minv;                                             // [K - firstCurved][Np][Np] inverse element mass matrices
    int firstCurved;                                                // elements >= firstCurved take this path (0: all of them)
    double c0, rho0, v0[3];
    int64_t stride;                                                 // K*Np
};

// true if some element or face carries different geometry at different integration points
inline bool isCurved(const dgb_desc* d) {
    if (d->nGeomEl > 1)
        for (int el = 0; el < d->K; ++el)
            for (int g = 1; g < d->nG; ++g)
                for (int k = 0; k < 9; ++k) {
                    const double a = d->elJacobian[((size_t)el * d->nG) * 9 + k], b = d->elJacobian[((size_t)el * d->nG + g) * 9 + k];
                    if (std::fabs(a - b) > 1e-11 * (std::fabs(a) + std::fabs(b) + 1e-300) + 1e-13) return true;
                }
    if (d->nGeomF > 1)
        for (int f = 0; f < d->F; ++f)
            for (int g = 1; g < d->nGf; ++g)
                for (int k = 0; k < 3; ++k)
                    if (std::fabs(d->fNormal[((size_t)f * d->nGf) * 3 + k] - d->fNormal[((size_t)f * d->nGf + g) * 3 + k]) > 1e-11) return true;
    return false;
}

// per element: does its Jacobian vary over its integration points, or the normal / surface Jacobian over one of its faces?
inline std::vector<uint8_t> curvedElements(const dgb_desc* d) {
    std::vector<uint8_t> flag(d->K, 0), fflag(d->F, 0);
    if (d->nGeomF > 1)
        for (int f = 0; f < d->F; ++f)
            for (int g = 1; g < d->nGf && !fflag[f]; ++g) {
                for (int k = 0; k < 3; ++k)
                    if (std::fabs(d->fNormal[((size_t)f * d->nGf) * 3 + k] - d->fNormal[((size_t)f * d->nGf + g) * 3 + k]) > 1e-11) fflag[f] = 1;
                const double a = d->fJacobianDet[(size_t)f * d->nGf], b = d->fJacobianDet[(size_t)f * d->nGf + g];
                if (std::fabs(a - b) > 1e-11 * (std::fabs(a) + std::fabs(b))) fflag[f] = 1;
            }
    for (int el = 0; el < d->K; ++el) {
        if (d->nGeomEl > 1)
            for (int g = 1; g < d->nG && !flag[el]; ++g)
                for (int k = 0; k < 9; ++k) {
                    const double a = d->elJacobian[((size_t)el * d->nG) * 9 + k], b = d->elJacobian[((size_t)el * d->nG + g) * 9 + k];
                    if (std::fabs(a - b) > 1e-11 * (std::fabs(a) + std::fabs(b) + 1e-300) + 1e-13) flag[el] = 1;
                }
        for (int lf = 0; lf < d->Nf; ++lf) if (fflag[d->elFId[(size_t)el * d->Nf + lf]]) flag[el] = 1;
    }
    return flag;
}

// Straight-sided elements may keep the collapsed kernels if the curved ones form a suffix of the numbering (the front end of
// this repository orders them so): returns the first curved element of that suffix, or 0 (everything through the curved
// kernel) when curved and straight-sided elements interleave.
inline int curvedSuffixStart(const std::vector<uint8_t>& flag) {
    int first = (int)flag.size();
    while (first > 0 && flag[first - 1]) --first;
    for (int el = 0; el < first; ++el) if (flag[el]) return 0;
    return first;
}

// inverse element mass matrices of the elements >= first, M_ij = sum_g phi_i(g) phi_j(g) w_g detJ(el, g)  (Mesh.cpp:440-466),
// extended precision
inline std::vector<double> curvedInverseMass(const dgb_desc* d, int first = 0) {
    const int Np = d->Np, nG = d->nG, K = d->K;
    if (d->nGeomEl != nG) throw std::runtime_error("curved elements need one Jacobian per integration point (nGeomEl == nG)");
    std::vector<double> out((size_t)(K - first) * Np * Np);
#pragma omp parallel for schedule(static)
    for (int el = first; el < K; ++el) {
        std::vector<long double> A((size_t)Np * Np, 0), B((size_t)Np * Np, 0);
        for (int g = 0; g < nG; ++g) {
            const long double wd = (long double)d->elWeight[g] * d->elJacobianDet[(size_t)el * nG + g];
            for (int i = 0; i < Np; ++i) {
                const long double wi = wd * d->elBasisFct[(size_t)g * Np + i];
                for (int j = 0; j < Np; ++j) A[(size_t)i * Np + j] += wi * d->elBasisFct[(size_t)g * Np + j];
            }
        }
        for (int i = 0; i < Np; ++i) B[(size_t)i * Np + i] = 1;
        for (int c = 0; c < Np; ++c) {  // Gauss-Jordan with partial pivoting
            int piv = c;
            for (int r = c + 1; r < Np; ++r) if (fabsl(A[(size_t)r * Np + c]) > fabsl(A[(size_t)piv * Np + c])) piv = r;
            if (piv != c)
                for (int k = 0; k < Np; ++k) { std::swap(A[(size_t)c * Np + k], A[(size_t)piv * Np + k]); std::swap(B[(size_t)c * Np + k], B[(size_t)piv * Np + k]); }
            const long double dd = 1 / A[(size_t)c * Np + c];
            for (int k = 0; k < Np; ++k) { A[(size_t)c * Np + k] *= dd; B[(size_t)c * Np + k] *= dd; }
            for (int r = 0; r < Np; ++r) {
                if (r == c) continue;
                const long double f = A[(size_t)r * Np + c];
                if (f == 0) continue;
                for (int k = 0; k < Np; ++k) { A[(size_t)r * Np + k] -= f * A[(size_t)c * Np + k]; B[(size_t)r * Np + k] -= f * B[(size_t)c * Np + k]; }
            }
        }
        for (size_t i = 0; i < (size_t)Np * Np; ++i) out[(size_t)(el - first) * Np * Np + i] = (double)B[i];
    }
    return out;
}

}  // namespace dgb
