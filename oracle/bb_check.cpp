// CPU checker of the Bernstein-Bezier path — TEST INFRASTRUCTURE ONLY (like everything under oracle/).
//
// Runs the product's own operator code (dgfem-acoustic_b200/csrc/bb_ops.h, bb_setup.h: the very templates and face
// arithmetic the CUDA kernel stage_bb.cu instantiates) on the host, element by element, with the connectivity taken
// straight from the desc (the reference's fNbrElId / fNToElNId), so that tests can compare  V * rhs_Bernstein(V^-1 u)
// with the oracle's L(u) without a GPU. Nothing in the product links or loads this file.
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../dgfem-acoustic_b200/csrc/bb_setup.h"

using namespace dgb::bb;

namespace {
thread_local std::string g_err;

// right-hand side of one field of one element through the Simplex<DIM, N> interface (the calls stage_bb2.cu makes)
template <int DIM, int N>
void fieldRhsSimplex(int q, const double* col0, int colStride, const double* dphi, const Tables& T, const double (&gl)[4][3], const double (&v0)[3],
                     bool flow, double rc2, double invRho, double (&out)[Simplex<DIM, N>::NP]) {
    typedef Simplex<DIM, N> SX;
    constexpr int NP = SX::NP, NFP = SX::NFP, NV = SX::NV, ND = SX::ND;
    double t[ND];
    for (int i = 0; i < ND; ++i) t[i] = 0.0;
    const int nCoupling = q == 0 ? DIM : q <= DIM ? 1 : 0;
    const int nPass = nCoupling + (flow ? 1 : 0);
    for (int pass = 0; pass < nPass; ++pass) {
        int field;
        double w[NV];
        if (pass < nCoupling) {
            field = q == 0 ? 1 + pass : 0;
            const int x = q == 0 ? pass : q - 1;
            const double s = q == 0 ? rc2 : invRho;
            for (int j = 0; j < NV; ++j) w[j] = s * gl[j][x];
        } else {
            field = q;
            for (int j = 0; j < NV; ++j) {
                w[j] = 0;
                for (int x = 0; x < DIM; ++x) w[j] += v0[x] * gl[j][x];
            }
        }
        double cc[NP];
        for (int i = 0; i < NP; ++i) cc[i] = col0[field * (size_t)colStride + T.permC2G[i]];
        SX::dirDerivAcc(cc, w, t);
    }
    for (int i = 0; i < NP; ++i) out[i] = 0.0;
    SX::elevate(t, -1.0, out);
    auto face = [&](auto Jc) {
        constexpr int J = decltype(Jc)::value;
        if constexpr (J < NV) {
            double x[NFP], zl[NP];
            for (int b = 0; b < NFP; ++b) x[b] = dphi[T.faceLf[J] * NFP + T.facePos[J][b]];
            SX::liftLocal(x, zl);
            SX::template scatterAdd<J>(zl, out);
        }
    };
    face(std::integral_constant<int, 0>{}); face(std::integral_constant<int, 1>{}); face(std::integral_constant<int, 2>{}); face(std::integral_constant<int, 3>{});
}

template <int DIM, int N, bool SIMPLEX>
void evalRhs(const dgb_desc& d, const Setup& S, const double* u, double* rhs) {
    constexpr int NP = Simplex<DIM, N>::NP, NFP = Simplex<DIM, N>::NFP, NV = DIM + 1;
    const int K = d.K;
    const size_t Ntot = (size_t)K * NP;
    const double rc2 = d.rho0 * d.c0 * d.c0, invRho = 1.0 / d.rho0;
    const bool flow = d.v0[0] != 0.0 || d.v0[1] != 0.0 || d.v0[2] != 0.0;
    // nodal -> Bernstein, mesh node order
    std::vector<double> c(4 * Ntot), out(4 * Ntot);
    for (int q = 0; q < 4; ++q)
        for (int el = 0; el < K; ++el)
            for (int m = 0; m < NP; ++m) {
                double s = 0;
                for (int n = 0; n < NP; ++n) s += S.Vinv[(size_t)m * NP + n] * u[q * Ntot + (size_t)el * NP + n];
                c[q * Ntot + (size_t)el * NP + m] = s;
            }
    const int gE = d.nGeomEl, gF = d.nGeomF;
    for (int el = 0; el < K; ++el) {
        // d lambda_j / d x from the element Jacobian (index u*3+x = dx_x/du_u): rows of its inverse, lambda_0 = 1 - sum
        const double* Jm = &d.elJacobian[(size_t)el * gE * 9];
        double A[3][3], B[3][3] = {};
        for (int r = 0; r < 3; ++r) for (int cc = 0; cc < 3; ++cc) A[r][cc] = Jm[r * 3 + cc];
        if (DIM == 3) {
            const double det = A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) +
                               A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
            for (int r = 0; r < 3; ++r)
                for (int cc = 0; cc < 3; ++cc) {
                    const int r1 = (cc + 1) % 3, r2 = (cc + 2) % 3, c1 = (r + 1) % 3, c2 = (r + 2) % 3;
                    B[r][cc] = (A[r1][c1] * A[r2][c2] - A[r1][c2] * A[r2][c1]) / det;  // B[x][u] = du_u/dx_x
                }
        } else {  // leading 2 x 2 block: A[u][x] = dx_x/du_u, B = A^-1 (indices [x][u])
            const double det = A[0][0] * A[1][1] - A[0][1] * A[1][0];
            B[0][0] = A[1][1] / det; B[0][1] = -A[0][1] / det;
            B[1][0] = -A[1][0] / det; B[1][1] = A[0][0] / det;
        }
        double gl[4][3] = {};
        for (int x = 0; x < DIM; ++x) {
            double s = 0;
            for (int j = 1; j < NV; ++j) { gl[j][x] = B[x][j - 1]; s += B[x][j - 1]; }
            gl[0][x] = -s;
        }
        // face inputs Fscale * (n.F(u-) - flux*) per field, mesh (local face, face node) order
        double dphi[4][4 * NFP];
        for (int lf = 0; lf < NV; ++lf) {
            const int f = d.elFId[(size_t)el * NV + lf];
            const int side = d.fNbrElId[2 * (size_t)f] == el ? 0 : 1;
            const double o = d.elFOrientation[(size_t)el * NV + lf];
            double n[3];
            for (int x = 0; x < 3; ++x) n[x] = o * d.fNormal[(size_t)f * gF * 3 + x];  // outward
            const double Fs = (SIMPLEX ? Simplex<DIM, N>::FACE_SCALE : 1.0) * d.fJacobianDet[(size_t)f * gF] / d.elJacobianDet[(size_t)el * gE];
            const double v0n = d.v0[0] * n[0] + d.v0[1] * n[1] + d.v0[2] * n[2];
            int pos[MAX_NP];
            for (int m = 0; m < NFP; ++m) pos[S.faceNodes[(size_t)lf * NFP + m]] = m;
            const double tau = d.fc * o * (side == 0 ? 1.0 : -1.0);
            const int bc = !d.fIsBoundary[f] ? BC_INTERIOR : d.fBC[f] == 1 ? BC_REFLECTING : BC_ABSORBING;
            const FaceCoef fk = faceCoef(bc, tau, Fs, v0n, d.c0, d.rho0);  // the kernel's own face arithmetic
            for (int k = 0; k < NFP; ++k) {
                const int own = d.fNToElNId[((size_t)f * NFP + k) * 2 + side];
                const int m = pos[own];
                double a[4], x4[4];
                for (int q = 0; q < 4; ++q) a[q] = c[q * Ntot + (size_t)el * NP + own];
                if (bc == BC_INTERIOR) {
                    const int nb = d.fNbrElId[2 * (size_t)f + (1 - side)];
                    const int nbn = d.fNToElNId[((size_t)f * NFP + k) * 2 + (1 - side)];
                    for (int q = 0; q < 4; ++q) a[q] -= c[q * Ntot + (size_t)nb * NP + nbn];
                }
                faceInput(fk, n, a, x4);
                for (int q = 0; q < 4; ++q) dphi[q][lf * NFP + m] = x4[q];
            }
        }
        const double v0[3] = {d.v0[0], d.v0[1], d.v0[2]};
        for (int q = 0; q < 4; ++q) {
            double r[NP];
            if constexpr (DIM == 3 && !SIMPLEX) fieldRhs<N>(q, &c[(size_t)el * NP], (int)Ntot, dphi[q], S.T, gl, v0, flow, rc2, invRho, r);  // the first-generation kernels' entry
            else fieldRhsSimplex<DIM, N>(q, &c[(size_t)el * NP], (int)Ntot, dphi[q], S.T, gl, v0, flow, rc2, invRho, r);
            for (int i = 0; i < NP; ++i) out[q * Ntot + (size_t)el * NP + S.T.permC2G[i]] = r[i];
        }
    }
    // Bernstein -> nodal
    for (int q = 0; q < 4; ++q)
        for (int el = 0; el < K; ++el)
            for (int n = 0; n < NP; ++n) {
                double s = 0;
                for (int m = 0; m < NP; ++m) s += S.V[(size_t)n * NP + m] * out[q * Ntot + (size_t)el * NP + m];
                rhs[q * Ntot + (size_t)el * NP + n] = s;
            }
}
int evalImpl(const dgb_desc* d, const double* u, double* rhs, double* liftDev, int32_t* alphaOut, bool simplexPath) {
    try {
        const Setup S = buildSetup(d);
        if (alphaOut) for (size_t i = 0; i < S.alpha.size(); ++i) alphaOut[i] = S.alpha[i];
        double dev = 0;
#define BBC_CASE(P)                                                           \
    case P:                                                                   \
        dev = liftDeviation<P>(S);                                            \
        if (d->dim == 2) evalRhs<2, P, true>(*d, S, u, rhs);                  \
        else if (simplexPath) evalRhs<3, P, true>(*d, S, u, rhs);             \
        else evalRhs<3, P, false>(*d, S, u, rhs);                             \
        break;
        switch (d->order) {
            BBC_CASE(1) BBC_CASE(2) BBC_CASE(3) BBC_CASE(4) BBC_CASE(5) BBC_CASE(6)
            default: throw std::runtime_error("order out of range");
        }
#undef BBC_CASE
        if (liftDev) *liftDev = dev;
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}
}  // namespace

extern "C" {

const char* bbc_last_error(void) { return g_err.c_str(); }

// rhs = L(u) through the Bernstein path, nodal in / nodal out; *liftDev (optional) = deviation of the closed-form lift
// from the dense nodal one; alphaOut (optional, [Np][4]) = the recovered Bernstein index of every node.
// Tetrahedra: the first-generation kernels' entry bb::fieldRhs; triangles: the Simplex<2, N> interface of stage_bb2.cu.
int bbc_eval_rhs(const dgb_desc* d, const double* u, double* rhs, double* liftDev, int32_t* alphaOut) {
    return evalImpl(d, u, rhs, liftDev, alphaOut, false);
}
// the same through the Simplex<DIM, N> interface in both dimensions
int bbc_eval_rhs_simplex(const dgb_desc* d, const double* u, double* rhs, double* liftDev, int32_t* alphaOut) {
    return evalImpl(d, u, rhs, liftDev, alphaOut, true);
}
}
