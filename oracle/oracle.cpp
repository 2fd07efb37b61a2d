// oracle.cpp — CPU restatement of the hot path of povanberg/DGFEM-Acoustic.
//
// TEST INFRASTRUCTURE. This file is the parity checker and the CPU baseline; the product
// (dgfem-acoustic_b200/) never links, loads or calls it.
//
// Parity status: PINNED against the reference's own sources compiled here (oracle/_ref, built by
// oracle/Makefile from /root/reference/src/*.cpp on top of the Gmsh/Eigen stand-ins in oracle/shim/);
// tests/test_oracle_vs_reference.py compares the two step by step. The reference ships no tests or golden
// vectors of its own, and Gmsh itself is unavailable, so everything that comes out of Gmsh
// (basis values, quadrature points, node ordering) is provided by the stand-in for both.
//
// Two modes, same inputs (a dgb_desc, include/dgb.h, i.e. the arrays of the reference's Mesh object):
//   mode 0 "faithful": the reference's loop nests, storage and OpenMP placement, function by function:
//        precomputeMassMatrix/getElMassMatrix  src/Mesh.cpp:440-466   + eigen::inverse src/utils.cpp:105-109
//        getElStiffVector                      src/Mesh.cpp:476-489
//        precomputeFlux                        src/Mesh.cpp:500-539
//        getElFlux                             src/Mesh.cpp:548-557
//        updateFlux (nodal + ghost/BC), RKR    src/Mesh.cpp:569-674, 391-418
//        numStep                               src/solver.cpp:35-52   + eigen::minus/linEq src/utils.cpp:118-136
//        rungeKutta / forwardEuler loops       src/solver.cpp:171-292 / 61-161
//   mode 1 "operator": the same discrete operator collapsed on affine elements:
//        rhs = sum_u Dw^u (sum_x G_xu F_x) - M^-1 E Mf (Fscale * flux), all loops OpenMP-parallel.
#include <omp.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../include/dgb.h"

namespace {

thread_local std::string g_err;

// Dense inverse with partial pivoting (stands in for Eigen's MatrixXd::inverse(), utils.cpp:105-109).
void invertInPlace(double* A, int n) {
    std::vector<double> W((size_t)n * 2 * n);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) { W[(size_t)i * 2 * n + j] = A[(size_t)i * n + j]; W[(size_t)i * 2 * n + n + j] = (i == j); }
    for (int c = 0; c < n; ++c) {
        int piv = c;
        for (int r = c + 1; r < n; ++r) if (std::fabs(W[(size_t)r * 2 * n + c]) > std::fabs(W[(size_t)piv * 2 * n + c])) piv = r;
        if (W[(size_t)piv * 2 * n + c] == 0.0) throw std::runtime_error("singular mass matrix");
        if (piv != c) for (int k = 0; k < 2 * n; ++k) std::swap(W[(size_t)c * 2 * n + k], W[(size_t)piv * 2 * n + k]);
        const double d = 1.0 / W[(size_t)c * 2 * n + c];
        for (int k = 0; k < 2 * n; ++k) W[(size_t)c * 2 * n + k] *= d;
        for (int r = 0; r < n; ++r) {
            if (r == c) continue;
            const double f = W[(size_t)r * 2 * n + c];
            if (f == 0.0) continue;
            for (int k = 0; k < 2 * n; ++k) W[(size_t)r * 2 * n + k] -= f * W[(size_t)c * 2 * n + k];
        }
    }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) A[(size_t)i * n + j] = W[(size_t)i * 2 * n + n + j];
}

void invertLong(std::vector<long double>& A, int n) {
    std::vector<long double> W((size_t)n * 2 * n);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) { W[(size_t)i * 2 * n + j] = A[(size_t)i * n + j]; W[(size_t)i * 2 * n + n + j] = (i == j); }
    for (int c = 0; c < n; ++c) {
        int piv = c;
        for (int r = c + 1; r < n; ++r) if (fabsl(W[(size_t)r * 2 * n + c]) > fabsl(W[(size_t)piv * 2 * n + c])) piv = r;
        if (piv != c) for (int k = 0; k < 2 * n; ++k) std::swap(W[(size_t)c * 2 * n + k], W[(size_t)piv * 2 * n + k]);
        const long double d = 1.0L / W[(size_t)c * 2 * n + c];
        for (int k = 0; k < 2 * n; ++k) W[(size_t)c * 2 * n + k] *= d;
        for (int r = 0; r < n; ++r) {
            if (r == c) continue;
            const long double f = W[(size_t)r * 2 * n + c];
            for (int k = 0; k < 2 * n; ++k) W[(size_t)r * 2 * n + k] -= f * W[(size_t)c * 2 * n + k];
        }
    }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) A[(size_t)i * n + j] = W[(size_t)i * 2 * n + n + j];
}

// Solve A g = b with A(r,c) = jac[r*3+c] restricted to dim x dim  (Mesh.cpp:56-72: J^T grad = ugrad)
void solveJT(const double* jac, int dim, const double* b, double* g) {
    double A[3][4];
    for (int r = 0; r < dim; ++r) { for (int c = 0; c < dim; ++c) A[r][c] = jac[r * 3 + c]; A[r][dim] = b[r]; }
    for (int c = 0; c < dim; ++c) {
        int piv = c;
        for (int r = c + 1; r < dim; ++r) if (std::fabs(A[r][c]) > std::fabs(A[piv][c])) piv = r;
        if (piv != c) for (int k = 0; k <= dim; ++k) std::swap(A[c][k], A[piv][k]);
        for (int r = c + 1; r < dim; ++r) {
            const double f = A[r][c] / A[c][c];
            for (int k = c; k <= dim; ++k) A[r][k] -= f * A[c][k];
        }
    }
    g[0] = g[1] = g[2] = 0.0;
    for (int r = dim - 1; r >= 0; --r) {
        double s = A[r][dim];
        for (int c = r + 1; c < dim; ++c) s -= A[r][c] * g[c];
        g[r] = s / A[r][r];
    }
}

struct Oracle {
    // ---- copy of the Mesh arrays (dgb_desc) ----
    int dim, order, Np, Nfp, Nf, K, F, nG, nGf, nGeomEl, nGeomF, fc;
    std::vector<double> elBasisFct, elUGradBasisFct, elWeight, fBasisFct, fWeight;
    std::vector<double> elJacobian, elJacobianDet, fNormalA, fJacobianDetA;
    std::vector<int32_t> elFIdA, elFOrientationA, fNbrElIdA, fNToElNIdA, fBC;
    std::vector<uint8_t> fIsBoundary;
    double c0, rho0, v0[3], dt;
    int threads;
    size_t N;

    // ---- accessors with the reference's names (include/Mesh.h:31-108) ----
    double elJac(int el, int g, int x, int u) const { return elJacobian[((size_t)el * nGeomEl + (nGeomEl == 1 ? 0 : g)) * 9 + u * 3 + x]; }
    double elJacobianDetAt(int el, int g) const { return elJacobianDet[(size_t)el * nGeomEl + (nGeomEl == 1 ? 0 : g)]; }
    double elBasis(int g, int i) const { return elBasisFct[(size_t)g * Np + i]; }
    double fBasis(int g, int i) const { return fBasisFct[(size_t)g * Nfp + i]; }
    const double* fNormal(int f, int g) const { return &fNormalA[((size_t)f * nGeomF + (nGeomF == 1 ? 0 : g)) * 3]; }
    double fJacobianDet(int f, int g) const { return fJacobianDetA[(size_t)f * nGeomF + (nGeomF == 1 ? 0 : g)]; }
    int elFId(int el, int lf) const { return elFIdA[(size_t)el * Nf + lf]; }
    int fNbrElId(int f, int s) const { return fNbrElIdA[2 * (size_t)f + s]; }
    int fNToElNId(int f, int nf, int s) const { return fNToElNIdA[((size_t)f * Nfp + nf) * 2 + s]; }
    int elFOrientation(int el, int lf) const { return elFOrientationA[(size_t)el * Nf + lf]; }

    // ---- faithful-mode storage (names follow the reference) ----
    bool faithfulReady = false;
    std::vector<double> m_elGradBasisFcts;  // [K][nG][Np][3]   Mesh.cpp:57
    std::vector<double> m_elMassMatrices;   // [K][Np][Np] inverse mass, row-major   Mesh.cpp:441
    std::vector<double> m_fFlux;            // [F][Nfp]
    std::vector<double> RKR;                // [F*nGf][16]       Mesh.cpp:391-418
    std::vector<double> uGhost[4];          // [F*nGf]
    std::vector<double> FluxGhost[4];       // [F*nGf][3]
    std::vector<double> Flux[4];            // [N][3]            solver.cpp:180

    // ---- operator-mode storage ----
    bool operatorReady = false;
    std::vector<double> Dw;        // [dim][Np][Np]
    std::vector<double> MinvRef;   // [Np][Np]
    std::vector<double> Mf;        // [Nfp][Nfp]
    std::vector<double> LiftRef;   // [Np][Nf*Nfp] (element-local face-node order, for orc_get_operators)
    std::vector<double> Ginv;      // [K][3][3]  Ginv[x][u] = du_u/dx_x
    std::vector<int> faceNodesRef; // [Nf][Nfp]

    // ---- sources ----
    std::vector<int32_t> srcOff, srcIdx;
    // receivers (a new capability next to the path): field values interpolated inside an element
    std::vector<int32_t> rcvEl;
    std::vector<double> rcvW, rcvRec;  // [nrcv][Np] Lagrange weights; record [step][nrcv][4]
    std::vector<double> srcAmp, srcFreq, srcPhase, srcDur;

    explicit Oracle(const dgb_desc& d, int nthreads) {
        dim = d.dim; order = d.order; Np = d.Np; Nfp = d.Nfp; Nf = d.Nf; K = d.K; F = d.F; nG = d.nG; nGf = d.nGf;
        nGeomEl = d.nGeomEl; nGeomF = d.nGeomF; fc = d.fc;
        c0 = d.c0; rho0 = d.rho0; v0[0] = d.v0[0]; v0[1] = d.v0[1]; v0[2] = d.v0[2]; dt = d.dt;
        threads = nthreads > 0 ? nthreads : omp_get_max_threads();
        N = (size_t)K * Np;
        if (!(nGeomEl == 1 || nGeomEl == nG) || !(nGeomF == 1 || nGeomF == nGf)) throw std::runtime_error("bad nGeomEl/nGeomF");
        elBasisFct.assign(d.elBasisFct, d.elBasisFct + (size_t)nG * Np);
        elUGradBasisFct.assign(d.elUGradBasisFct, d.elUGradBasisFct + (size_t)nG * Np * 3);
        elWeight.assign(d.elWeight, d.elWeight + nG);
        fBasisFct.assign(d.fBasisFct, d.fBasisFct + (size_t)nGf * Nfp);
        fWeight.assign(d.fWeight, d.fWeight + nGf);
        elJacobian.assign(d.elJacobian, d.elJacobian + (size_t)K * nGeomEl * 9);
        elJacobianDet.assign(d.elJacobianDet, d.elJacobianDet + (size_t)K * nGeomEl);
        fNormalA.assign(d.fNormal, d.fNormal + (size_t)F * nGeomF * 3);
        fJacobianDetA.assign(d.fJacobianDet, d.fJacobianDet + (size_t)F * nGeomF);
        elFIdA.assign(d.elFId, d.elFId + (size_t)K * Nf);
        elFOrientationA.assign(d.elFOrientation, d.elFOrientation + (size_t)K * Nf);
        fNbrElIdA.assign(d.fNbrElId, d.fNbrElId + (size_t)F * 2);
        fNToElNIdA.assign(d.fNToElNId, d.fNToElNId + (size_t)F * Nfp * 2);
        fIsBoundary.assign(d.fIsBoundary, d.fIsBoundary + F);
        fBC.assign(d.fBC, d.fBC + F);
    }

    // =========================================================================================
    // Faithful mode
    // =========================================================================================
    void prepareFaithful() {
        if (faithfulReady) return;
        // physical gradients of the basis functions, Mesh.cpp:56-72
        m_elGradBasisFcts.assign((size_t)K * nG * Np * 3, 0.0);
#pragma omp parallel for schedule(static) num_threads(threads)
        for (int el = 0; el < K; ++el)
            for (int g = 0; g < nG; ++g) {
                double jac[9];
                for (int u = 0; u < 3; ++u) for (int x = 0; x < 3; ++x) jac[u * 3 + x] = elJac(el, g, x, u);
                for (int f = 0; f < Np; ++f)
                    solveJT(jac, dim, &elUGradBasisFct[((size_t)g * Np + f) * 3], &m_elGradBasisFcts[(((size_t)el * nG + g) * Np + f) * 3]);
            }
        // RKR, Mesh.cpp:391-418
        RKR.assign((size_t)F * nGf * 16, 0.0);
        for (int f = 0; f < F; ++f) {
            if (fBC[f] != 0) continue;
            for (int g = 0; g < nGf; ++g) {
                double* R = &RKR[((size_t)f * nGf + g) * 16];
                const double* n = fNormal(f, g);
                R[0] = 0.25 * c0;
                R[1] = 0.25 * c0 * c0 * rho0 * n[0];
                R[2] = 0.25 * c0 * c0 * rho0 * n[1];
                R[3] = 0.25 * c0 * c0 * rho0 * n[2];
                for (int i = 0; i < 3; ++i) {
                    R[4 + 4 * i] = 0.25 * n[i] / rho0;
                    R[5 + 4 * i] = 0.25 * c0 * n[i] * n[0];
                    R[6 + 4 * i] = 0.25 * c0 * n[i] * n[1];
                    R[7 + 4 * i] = 0.25 * c0 * n[i] * n[2];
                }
            }
        }
        m_fFlux.assign((size_t)F * Nfp, 0.0);
        for (int q = 0; q < 4; ++q) {
            uGhost[q].assign((size_t)F * nGf, 0.0);
            FluxGhost[q].assign((size_t)F * nGf * 3, 0.0);
            Flux[q].assign(N * 3, 0.0);
        }
        // precomputeMassMatrix, Mesh.cpp:440-466 (the inverse is stored)
        m_elMassMatrices.assign((size_t)K * Np * Np, 0.0);
#pragma omp parallel for schedule(static) num_threads(threads)
        for (int el = 0; el < K; ++el) {
            double* M = &m_elMassMatrices[(size_t)el * Np * Np];
            for (int i = 0; i < Np; ++i)
                for (int j = 0; j < Np; ++j) {
                    M[i * Np + j] = 0.0;
                    for (int g = 0; g < nG; ++g) M[i * Np + j] += elBasis(g, i) * elBasis(g, j) * elWeight[g] * elJacobianDetAt(el, g);
                }
            invertInPlace(M, Np);
        }
        faithfulReady = true;
    }

    // Mesh::updateFlux, Mesh.cpp:569-674 (serial in the reference)
    void updateFlux(const double* u) {
        const double* U[4] = {u, u + N, u + 2 * N, u + 3 * N};
        for (int el = 0; el < K; ++el) {
            for (int n = 0; n < Np; ++n) {
                const size_t i = (size_t)el * Np + n;
                double* Fp = &Flux[0][3 * i]; double* Fx = &Flux[1][3 * i]; double* Fy = &Flux[2][3 * i]; double* Fz = &Flux[3][3 * i];
                Fp[0] = v0[0] * U[0][i] + rho0 * c0 * c0 * U[1][i];
                Fp[1] = v0[1] * U[0][i] + rho0 * c0 * c0 * U[2][i];
                Fp[2] = v0[2] * U[0][i] + rho0 * c0 * c0 * U[3][i];
                Fx[0] = v0[0] * U[1][i] + U[0][i] / rho0; Fx[1] = v0[1] * U[1][i]; Fx[2] = v0[2] * U[1][i];
                Fy[0] = v0[0] * U[2][i]; Fy[1] = v0[1] * U[2][i] + U[0][i] / rho0; Fy[2] = v0[2] * U[2][i];
                Fz[0] = v0[0] * U[3][i]; Fz[1] = v0[1] * U[3][i]; Fz[2] = v0[2] * U[3][i] + U[0][i] / rho0;
            }
            for (int lf = 0; lf < Nf; ++lf) {
                const int fId = elFId(el, lf);
                if (!fIsBoundary[fId]) continue;
                for (int g = 0; g < nGf; ++g) {
                    const size_t gId = (size_t)fId * nGf + g;
                    double ug[4] = {0, 0, 0, 0};
                    for (int n = 0; n < Nfp; ++n) {
                        const size_t nId = (size_t)el * Np + fNToElNId(fId, n, 0);
                        for (int q = 0; q < 4; ++q) ug[q] += U[q][nId] * fBasis(g, n);
                    }
                    const double* nrm = fNormal(fId, g);
                    if (fBC[fId] == 1) {
                        const double dot = nrm[0] * ug[1] + nrm[1] * ug[2] + nrm[2] * ug[3];
                        ug[1] -= dot * nrm[0]; ug[2] -= dot * nrm[1]; ug[3] -= dot * nrm[2];
                        double FG[4][3] = {
                            {v0[0] * ug[0] + rho0 * c0 * c0 * ug[1], v0[1] * ug[0] + rho0 * c0 * c0 * ug[2], v0[2] * ug[0] + rho0 * c0 * c0 * ug[3]},
                            {v0[0] * ug[1] + ug[0] / rho0, v0[1] * ug[1], v0[2] * ug[1]},
                            {v0[0] * ug[2], v0[1] * ug[2] + ug[0] / rho0, v0[2] * ug[2]},
                            {v0[0] * ug[3], v0[1] * ug[3], v0[2] * ug[3] + ug[0] / rho0}};
                        for (int q = 0; q < 4; ++q) {
                            double* FGq = &FluxGhost[q][3 * gId];
                            FGq[1] = FG[q][1]; FGq[2] = FG[q][2];
                            FGq[0] = nrm[0] * FG[q][0] + nrm[1] * FG[q][1] + nrm[2] * FG[q][2];
                        }
                    } else {
                        const double* R = &RKR[gId * 16];
                        for (int q = 0; q < 4; ++q)
                            FluxGhost[q][3 * gId] = R[4 * q] * ug[0] + R[4 * q + 1] * ug[1] + R[4 * q + 2] * ug[2] + R[4 * q + 3] * ug[3];
                    }
                    for (int q = 0; q < 4; ++q) uGhost[q][gId] = ug[q];
                }
            }
        }
    }

    // Mesh::precomputeFlux, Mesh.cpp:500-539. The reference's nested parallel region makes every thread run all
    // faces; the result is that of a serial loop, which is what is done here.
    void precomputeFlux(const double* u, const std::vector<double>& FluxEq, int eq) {
        std::vector<double> FIntPts(nGf);
        double Fnum[3];
        for (int f = 0; f < F; ++f) {
            std::fill(FIntPts.begin(), FIntPts.end(), 0.0);
            if (fIsBoundary[f]) {
                for (int g = 0; g < nGf; ++g) FIntPts[g] = FluxGhost[eq][3 * ((size_t)f * nGf + g)];
            } else {
                for (int i = 0; i < Nfp; ++i) {
                    const size_t elUp = (size_t)fNbrElId(f, 0) * Np + fNToElNId(f, i, 0);
                    const size_t elDn = (size_t)fNbrElId(f, 1) * Np + fNToElNId(f, i, 1);
                    for (int g = 0; g < nGf; ++g) {
                        const double* nrm = fNormal(f, g);
                        for (int x = 0; x < 3; ++x)
                            Fnum[x] = 0.5 * ((FluxEq[3 * elUp + x] + FluxEq[3 * elDn + x]) + fc * c0 * nrm[x] * (u[elUp] - u[elDn]));
                        FIntPts[g] += (nrm[0] * Fnum[0] + nrm[1] * Fnum[1] + nrm[2] * Fnum[2]) * fBasis(g, i);
                    }
                }
            }
            for (int n = 0; n < Nfp; ++n) {
                double s = 0;
                for (int g = 0; g < nGf; ++g) s += fWeight[g] * fBasis(g, n) * FIntPts[g] * fJacobianDet(f, g);
                m_fFlux[(size_t)f * Nfp + n] = s;
            }
        }
    }

    // Mesh::getElFlux, Mesh.cpp:548-557
    void getElFlux(int el, double* Fv) const {
        std::fill(Fv, Fv + Np, 0.0);
        for (int lf = 0; lf < Nf; ++lf) {
            const int f = elFId(el, lf);
            const int i = (el == fNbrElId(f, 0)) ? 0 : 1;
            for (int nf = 0; nf < Nfp; ++nf) Fv[fNToElNId(f, nf, i)] += elFOrientation(el, lf) * m_fFlux[(size_t)f * Nfp + nf];
        }
    }

    // Mesh::getElStiffVector, Mesh.cpp:476-489
    void getElStiffVector(int el, const std::vector<double>& FluxEq, double* S) const {
        for (int i = 0; i < Np; ++i) {
            S[i] = 0.0;
            for (int j = 0; j < Np; ++j) {
                const double* Fj = &FluxEq[3 * ((size_t)el * Np + j)];
                for (int g = 0; g < nG; ++g) {
                    const double* gr = &m_elGradBasisFcts[(((size_t)el * nG + g) * Np + i) * 3];
                    S[i] += (Fj[0] * gr[0] + Fj[1] * gr[1] + Fj[2] * gr[2]) * elBasis(g, j) * elWeight[g] * elJacobianDetAt(el, g);
                }
            }
        }
    }

    // solver::numStep, solver.cpp:35-52 : u <- beta*u + dt*M^-1 (S - F), in place, equation by equation
    void numStep(double* u, double beta) {
        for (int eq = 0; eq < 4; ++eq) {
            double* ueq = u + (size_t)eq * N;
            precomputeFlux(ueq, Flux[eq], eq);
#pragma omp parallel num_threads(threads)
            {
                std::vector<double> elFluxV(Np), elStiff(Np), y(Np);
#pragma omp for schedule(static)
                for (int el = 0; el < K; ++el) {
                    getElFlux(el, elFluxV.data());
                    getElStiffVector(el, Flux[eq], elStiff.data());
                    for (int i = 0; i < Np; ++i) elStiff[i] -= elFluxV[i];  // eigen::minus, utils.cpp:132
                    // eigen::linEq, utils.cpp:118-123: the row-major inverse is read through a column-major Map,
                    // i.e. A(i,j) = data[j*Np+i]
                    const double* A = &m_elMassMatrices[(size_t)el * Np * Np];
                    double* Y = ueq + (size_t)el * Np;
                    for (int i = 0; i < Np; ++i) {
                        double s = 0.0;
                        for (int j = 0; j < Np; ++j) s += A[(size_t)j * Np + i] * elStiff[j];
                        y[i] = beta * Y[i] + dt * s;
                    }
                    for (int i = 0; i < Np; ++i) Y[i] = y[i];
                }
            }
        }
    }

    // =========================================================================================
    // Operator mode
    // =========================================================================================
    void prepareOperator() {
        if (operatorReady) return;
        // affinity check
        for (int el = 0; el < K && nGeomEl > 1; ++el)
            for (int g = 1; g < nG; ++g)
                for (int k = 0; k < 9; ++k) {
                    const double a = elJacobian[((size_t)el * nG) * 9 + k], b = elJacobian[((size_t)el * nG + g) * 9 + k];
                    if (std::fabs(a - b) > 1e-12 * (1.0 + std::fabs(a))) throw std::runtime_error("operator mode needs affine elements");
                }
        std::vector<long double> M((size_t)Np * Np, 0.0L);
        for (int i = 0; i < Np; ++i)
            for (int j = 0; j < Np; ++j)
                for (int g = 0; g < nG; ++g) M[(size_t)i * Np + j] += (long double)elBasis(g, i) * elBasis(g, j) * elWeight[g];
        invertLong(M, Np);
        MinvRef.resize((size_t)Np * Np);
        for (size_t k = 0; k < M.size(); ++k) MinvRef[k] = (double)M[k];
        Dw.assign((size_t)dim * Np * Np, 0.0);
        for (int u = 0; u < dim; ++u) {
            std::vector<long double> Ku((size_t)Np * Np, 0.0L);
            for (int i = 0; i < Np; ++i)
                for (int j = 0; j < Np; ++j)
                    for (int g = 0; g < nG; ++g)
                        Ku[(size_t)i * Np + j] += (long double)elUGradBasisFct[((size_t)g * Np + i) * 3 + u] * elBasis(g, j) * elWeight[g];
            for (int i = 0; i < Np; ++i)
                for (int j = 0; j < Np; ++j) {
                    long double s = 0.0L;
                    for (int k = 0; k < Np; ++k) s += M[(size_t)i * Np + k] * Ku[(size_t)k * Np + j];
                    Dw[((size_t)u * Np + i) * Np + j] = (double)s;
                }
        }
        std::vector<long double> MfL((size_t)Nfp * Nfp, 0.0L);
        for (int n = 0; n < Nfp; ++n)
            for (int i = 0; i < Nfp; ++i)
                for (int g = 0; g < nGf; ++g) MfL[(size_t)n * Nfp + i] += (long double)fWeight[g] * fBasis(g, n) * fBasis(g, i);
        Mf.resize(MfL.size());
        for (size_t k = 0; k < MfL.size(); ++k) Mf[k] = (double)MfL[k];
        // element 0 is the first owner of all of its faces: its maps give the local face-node table
        faceNodesRef.resize((size_t)Nf * Nfp);
        for (int lf = 0; lf < Nf; ++lf)
            for (int m = 0; m < Nfp; ++m) faceNodesRef[lf * Nfp + m] = fNToElNId(elFId(0, lf), m, 0);
        LiftRef.assign((size_t)Np * Nf * Nfp, 0.0);
        for (int i = 0; i < Np; ++i)
            for (int lf = 0; lf < Nf; ++lf)
                for (int m = 0; m < Nfp; ++m) {
                    long double s = 0.0L;
                    for (int n = 0; n < Nfp; ++n) s += M[(size_t)i * Np + faceNodesRef[lf * Nfp + n]] * MfL[(size_t)n * Nfp + m];
                    LiftRef[(size_t)i * Nf * Nfp + lf * Nfp + m] = (double)s;
                }
        Ginv.assign((size_t)K * 9, 0.0);
        for (int el = 0; el < K; ++el) {
            double jac[9];
            for (int u = 0; u < 3; ++u) for (int x = 0; x < 3; ++x) jac[u * 3 + x] = elJac(el, 0, x, u);
            for (int u = 0; u < dim; ++u) {  // column u of A^-1: solve A g = e_u
                double e[3] = {0, 0, 0}, g[3];
                e[u] = 1.0;
                solveJT(jac, dim, e, g);
                for (int x = 0; x < dim; ++x) Ginv[(size_t)el * 9 + x * 3 + u] = g[x];
            }
        }
        operatorReady = true;
    }

    // rhs = L(u)  (no dt), operator form
    void rhsOperator(const double* u, double* rhs) const {
        const double* U[4] = {u, u + N, u + 2 * N, u + 3 * N};
        const double rc2 = rho0 * c0 * c0;
#pragma omp parallel num_threads(threads)
        {
            std::vector<double> cf((size_t)4 * dim * Np), flux((size_t)4 * Nfp), tmp((size_t)4 * Nfp), acc((size_t)4 * Np);
#pragma omp for schedule(static)
            for (int el = 0; el < K; ++el) {
                const double* G = &Ginv[(size_t)el * 9];
                // contravariant fluxes c^u_q[j] = sum_x Ginv[x][u] F_q,x(j)
                for (int j = 0; j < Np; ++j) {
                    const size_t i = (size_t)el * Np + j;
                    const double p = U[0][i], v[3] = {U[1][i], U[2][i], U[3][i]};
                    double Fq[4][3];
                    for (int x = 0; x < 3; ++x) {
                        Fq[0][x] = v0[x] * p + rc2 * v[x];
                        for (int q = 1; q < 4; ++q) Fq[q][x] = v0[x] * v[q - 1] + (x == q - 1 ? p / rho0 : 0.0);
                    }
                    for (int q = 0; q < 4; ++q)
                        for (int uu = 0; uu < dim; ++uu) {
                            double s = 0.0;
                            for (int x = 0; x < dim; ++x) s += G[x * 3 + uu] * Fq[q][x];
                            cf[((size_t)q * dim + uu) * Np + j] = s;
                        }
                }
                for (int q = 0; q < 4; ++q)
                    for (int i = 0; i < Np; ++i) {
                        double s = 0.0;
                        for (int uu = 0; uu < dim; ++uu) {
                            const double* D = &Dw[((size_t)uu * Np + i) * Np];
                            const double* c = &cf[((size_t)q * dim + uu) * Np];
                            for (int j = 0; j < Np; ++j) s += D[j] * c[j];
                        }
                        acc[(size_t)q * Np + i] = s;
                    }
                // faces
                for (int lf = 0; lf < Nf; ++lf) {
                    const int f = elFId(el, lf);
                    const int side = (el == fNbrElId(f, 0)) ? 0 : 1;
                    const double o = elFOrientation(el, lf);
                    const double* n = fNormal(f, 0);
                    const double v0n = v0[0] * n[0] + v0[1] * n[1] + v0[2] * n[2];
                    for (int m = 0; m < Nfp; ++m) {
                        double fl[4];
                        if (fIsBoundary[f]) {
                            const size_t i = (size_t)el * Np + fNToElNId(f, m, 0);
                            const double p = U[0][i];
                            double v[3] = {U[1][i], U[2][i], U[3][i]};
                            const double vn = n[0] * v[0] + n[1] * v[1] + n[2] * v[2];
                            if (fBC[f] == 1) {
                                for (int x = 0; x < 3; ++x) v[x] -= vn * n[x];
                                fl[0] = v0n * p + rc2 * (n[0] * v[0] + n[1] * v[1] + n[2] * v[2]);
                                for (int x = 0; x < 3; ++x) fl[1 + x] = v0n * v[x] + n[x] * p / rho0;
                            } else {
                                fl[0] = 0.25 * c0 * p + 0.25 * c0 * c0 * rho0 * vn;
                                for (int x = 0; x < 3; ++x) fl[1 + x] = 0.25 * n[x] / rho0 * p + 0.25 * c0 * n[x] * vn;
                            }
                        } else {
                            const size_t iu = (size_t)fNbrElId(f, 0) * Np + fNToElNId(f, m, 0);
                            const size_t id = (size_t)fNbrElId(f, 1) * Np + fNToElNId(f, m, 1);
                            const double ps = U[0][iu] + U[0][id];
                            const double vs[3] = {U[1][iu] + U[1][id], U[2][iu] + U[2][id], U[3][iu] + U[3][id]};
                            const double vns = n[0] * vs[0] + n[1] * vs[1] + n[2] * vs[2];
                            fl[0] = 0.5 * (v0n * ps + rc2 * vns) + 0.5 * fc * c0 * (U[0][iu] - U[0][id]);
                            for (int x = 0; x < 3; ++x)
                                fl[1 + x] = 0.5 * (v0n * vs[x] + n[x] * ps / rho0) + 0.5 * fc * c0 * (U[1 + x][iu] - U[1 + x][id]);
                        }
                        for (int q = 0; q < 4; ++q) flux[(size_t)q * Nfp + m] = o * fl[q];
                    }
                    const double Fscale = fJacobianDet(f, 0) / elJacobianDetAt(el, 0);
                    for (int q = 0; q < 4; ++q)
                        for (int n2 = 0; n2 < Nfp; ++n2) {
                            double s = 0.0;
                            for (int m = 0; m < Nfp; ++m) s += Mf[(size_t)n2 * Nfp + m] * flux[(size_t)q * Nfp + m];
                            tmp[(size_t)q * Nfp + n2] = Fscale * s;
                        }
                    for (int q = 0; q < 4; ++q)
                        for (int i = 0; i < Np; ++i) {
                            double s = 0.0;
                            for (int n2 = 0; n2 < Nfp; ++n2) s += MinvRef[(size_t)i * Np + fNToElNId(f, n2, side)] * tmp[(size_t)q * Nfp + n2];
                            acc[(size_t)q * Np + i] -= s;
                        }
                }
                for (int q = 0; q < 4; ++q)
                    for (int i = 0; i < Np; ++i) rhs[(size_t)q * N + (size_t)el * Np + i] = acc[(size_t)q * Np + i];
            }
        }
    }

    // =========================================================================================
    // One "stage": k <- beta*k + dt*L(k) in place (what updateFlux + numStep do together)
    // =========================================================================================
    void stageInPlace(int mode, double* k, double beta, std::vector<double>& scratch) {
        if (mode == 0) {
            updateFlux(k);
            numStep(k, beta);
        } else {
            scratch.resize(4 * N);
            rhsOperator(k, scratch.data());
#pragma omp parallel for schedule(static) num_threads(threads)
            for (size_t i = 0; i < 4 * N; ++i) k[i] = beta * k[i] + dt * scratch[i];
        }
    }

    void applySources(double* u, double t) const {
        for (size_t s = 0; s < srcAmp.size(); ++s)
            if (t < srcDur[s]) {
                const double val = srcAmp[s] * sin(2 * M_PI * srcFreq[s] * t + srcPhase[s]);  // solver.cpp:255
                for (int k = srcOff[s]; k < srcOff[s + 1]; ++k) u[srcIdx[k]] = val;
            }
    }

    // solver::rungeKutta (solver.cpp:216-286) / solver::forwardEuler (solver.cpp:105-155) for nsteps iterations
    double run(int mode, int integrator, double* u, double t, int nsteps, int nprobe, const int32_t* probeIdx, double* probeOut) {
        if (mode == 0) prepareFaithful(); else prepareOperator();
        std::vector<double> k1, k2, k3, k4, scratch;
        for (int step = 0; step < nsteps; ++step, t += dt) {
            for (int j = 0; j < nprobe; ++j)
                for (int q = 0; q < 4; ++q) probeOut[((size_t)step * nprobe + j) * 4 + q] = u[(size_t)q * N + probeIdx[j]];
            for (size_t j = 0; j < rcvEl.size(); ++j)  // same instant as the probes / the reference's snapshots (solver.cpp:222)
                for (int q = 0; q < 4; ++q) {
                    double s = 0;
                    for (int n = 0; n < Np; ++n) s += rcvW[j * Np + n] * u[(size_t)q * N + (size_t)rcvEl[j] * Np + n];
                    rcvRec.push_back(s);
                }
            applySources(u, t);
            if (integrator == DGB_EULER1) {
                stageInPlace(mode, u, 1.0, scratch);
                continue;
            }
            k1.assign(u, u + 4 * N); k2 = k1; k3 = k1; k4 = k1;  // solver.cpp:261
            stageInPlace(mode, k1.data(), 0.0, scratch);
            for (size_t i = 0; i < 4 * N; ++i) k2[i] += 0.5 * k1[i];  // eigen::plusTimes, solver.cpp:266
            stageInPlace(mode, k2.data(), 0.0, scratch);
            for (size_t i = 0; i < 4 * N; ++i) k3[i] += 0.5 * k2[i];
            stageInPlace(mode, k3.data(), 0.0, scratch);
            for (size_t i = 0; i < 4 * N; ++i) k4[i] += 1 * k3[i];
            stageInPlace(mode, k4.data(), 0.0, scratch);
            for (size_t i = 0; i < 4 * N; ++i) u[i] += (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]) / 6.0;  // solver.cpp:283
        }
        return t;
    }
};

}  // namespace

extern "C" {

const char* orc_last_error(void) { return g_err.c_str(); }

void* orc_create(const dgb_desc* d, int threads) {
    try {
        return new Oracle(*d, threads);
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}

void orc_destroy(void* h) { delete static_cast<Oracle*>(h); }

int orc_set_sources(void* h, int nsrc, const int32_t* offsets, const int32_t* nodeIdx, const double* amp, const double* freq,
                    const double* phase, const double* duration) {
    auto* o = static_cast<Oracle*>(h);
    o->srcOff.assign(offsets, offsets + nsrc + 1);
    o->srcIdx.assign(nodeIdx, nodeIdx + offsets[nsrc]);
    o->srcAmp.assign(amp, amp + nsrc);
    o->srcFreq.assign(freq, freq + nsrc);
    o->srcPhase.assign(phase, phase + nsrc);
    o->srcDur.assign(duration, duration + nsrc);
    return 0;
}

int orc_set_receivers(void* h, int nrcv, const int32_t* el, const double* weights) {
    auto* o = static_cast<Oracle*>(h);
    o->rcvEl.assign(el, el + nrcv);
    o->rcvW.assign(weights, weights + (size_t)nrcv * o->Np);
    o->rcvRec.clear();
    return 0;
}

// out[step][receiver][4]; returns the number of recorded steps and clears the record
int orc_get_receivers(void* h, double* out, int capacity_steps) {
    auto* o = static_cast<Oracle*>(h);
    const size_t per = o->rcvEl.size() * 4;
    const int n = per ? (int)std::min<size_t>(o->rcvRec.size() / per, (size_t)capacity_steps) : 0;
    std::copy(o->rcvRec.begin(), o->rcvRec.begin() + (size_t)n * per, out);
    o->rcvRec.clear();
    return n;
}

int orc_run(void* h, int mode, int integrator, double* u, double t_start, int nsteps, int nprobe, const int32_t* probeIdx,
            double* probeOut, double* t_end) {
    try {
        double t = static_cast<Oracle*>(h)->run(mode, integrator, u, t_start, nsteps, nprobe, probeIdx, probeOut);
        if (t_end) *t_end = t;
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

// rhs = L(u): mode 0 evaluates it through the faithful stage with dt = 1, beta = 0
int orc_eval_rhs(void* h, int mode, const double* u, double* rhs) {
    try {
        auto* o = static_cast<Oracle*>(h);
        if (mode == 0) {
            o->prepareFaithful();
            std::copy(u, u + 4 * o->N, rhs);
            const double keep = o->dt;
            o->dt = 1.0;
            o->updateFlux(rhs);
            o->numStep(rhs, 0.0);
            o->dt = keep;
        } else {
            o->prepareOperator();
            o->rhsOperator(u, rhs);
        }
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

int orc_get_operators(void* h, double* Dw, double* Lift) {
    try {
        auto* o = static_cast<Oracle*>(h);
        o->prepareOperator();
        std::copy(o->Dw.begin(), o->Dw.end(), Dw);
        std::copy(o->LiftRef.begin(), o->LiftRef.end(), Lift);
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

}  // extern "C"
