// Host-side set-up of the Bernstein-Bezier path (bb_ops.h): recovers the reference nodes from the tables of the
// reference's Mesh (dgb_desc), builds the nodal <-> Bernstein conversion, the permutations between the mesh's node
// order and the canonical index order of bb_ops.h, and checks the closed-form sparse operators against the dense
// nodal ones. Plain C++ (no CUDA): compiled into libdgb.so (dgb_api.cu) and into the CPU checker (oracle/bb_check.cpp).
#pragma once
#include <stdint.h>

#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/dgb.h"
#include "bb_ops.h"

namespace dgb {
namespace bb {

struct Setup {
    int dim = 3, nV = 4;             // tetrahedra (4 vertices / faces) or triangles (3)
    int N = 0, Np = 0, Nfp = 0;
    Tables T{};
    std::vector<int> alpha;          // [Np][4] Bernstein index of every mesh node (node n sits at alpha/N; triangles: alpha[3] = 0)
    std::vector<int32_t> faceNodes;  // [nV][Nfp] element-local node of face node m (first-owner order, as dgb_api.cu)
    std::vector<int> ownIdx;         // [nV][Nfp] canonical volume index of coefficient b (canonical face order) of canonical face J
    std::vector<double> V, Vinv;     // [Np][Np] row-major, mesh node order on both sides: u_n = sum_m V[n][m] c_m
    std::vector<double> liftNodal;   // [Np][nV*Nfp]  Mref^-1[:, faceNodes(lf)] Mf  (dense, for the self-check)
};

typedef long double ld;

inline void invertLD(std::vector<ld>& A, int n) {
    std::vector<ld> B((size_t)n * n, 0);
    for (int i = 0; i < n; ++i) B[(size_t)i * n + i] = 1;
    for (int c = 0; c < n; ++c) {
        int piv = c;
        for (int r = c + 1; r < n; ++r) if (fabsl(A[(size_t)r * n + c]) > fabsl(A[(size_t)piv * n + c])) piv = r;
        if (A[(size_t)piv * n + c] == 0) throw std::runtime_error("Bernstein set-up: singular matrix");
        if (piv != c)
            for (int k = 0; k < n; ++k) { std::swap(A[(size_t)c * n + k], A[(size_t)piv * n + k]); std::swap(B[(size_t)c * n + k], B[(size_t)piv * n + k]); }
        const ld dd = 1 / A[(size_t)c * n + c];
        for (int k = 0; k < n; ++k) { A[(size_t)c * n + k] *= dd; B[(size_t)c * n + k] *= dd; }
        for (int r = 0; r < n; ++r) {
            if (r == c) continue;
            const ld f = A[(size_t)r * n + c];
            if (f == 0) continue;
            for (int k = 0; k < n; ++k) { A[(size_t)r * n + k] -= f * A[(size_t)c * n + k]; B[(size_t)r * n + k] -= f * B[(size_t)c * n + k]; }
        }
    }
    A.swap(B);
}

inline int layerIdxRt(int N, int J, int l, int b1, int b2) {
    const int b0 = N - l - b1 - b2;
    const int a1 = J == 0 ? b0 : J == 1 ? l : b1;
    const int a2 = J == 0 ? b1 : J == 1 ? b1 : J == 2 ? l : b2;
    const int a3 = J == 0 ? b2 : J == 1 ? b2 : J == 2 ? b2 : l;
    return vidx(N, a1, a2, a3);
}

inline int layerIdxRt2(int N, int J, int l, int b1) {
    const int b0 = N - l - b1;
    const int a1 = J == 0 ? b0 : J == 1 ? l : b1;
    const int a2 = J == 0 ? b1 : J == 1 ? b1 : l;
    return fidx(N, a1, a2);
}

inline Setup buildSetup(const dgb_desc* d) {
    if (!((d->dim == 3 && d->Nf == 4) || (d->dim == 2 && d->Nf == 3))) throw std::runtime_error("Bernstein path: tetrahedra and triangles only");
    const int dim = d->dim, nV = dim + 1;
    const int N = d->order, Np = d->Np, Nfp = d->Nfp, nG = d->nG, nGf = d->nGf;
    if (N < 1 || N > MAX_ORDER || Np != (dim == 3 ? tet(N) : tri(N)) || Nfp != (dim == 3 ? tri(N) : N + 1))
        throw std::runtime_error("Bernstein path: unexpected element sizes");
    Setup S;
    S.dim = dim; S.nV = nV;
    S.N = N; S.Np = Np; S.Nfp = Nfp;
    // 1. parametric coordinates of the nodes. x_u = sum_n X_un phi_n reproduces the coordinate function u, so
    //    sum_n X_un dphi_n/du'(g) = delta_uu' at every quadrature point and sum_n X_un int(phi_n) = int(u) = 1/24 on the unit
    //    tetrahedron (1/6 on the unit triangle): least squares over all points (the desc carries no node coordinates).
    ld wsum = 0;
    for (int g = 0; g < nG; ++g) wsum += d->elWeight[g];
    if (fabsl(wsum - (dim == 3 ? 1.0L / 6 : 0.5L)) > 1e-10L) throw std::runtime_error("Bernstein path: the reference element is not the unit simplex");
    const ld cInt = dim == 3 ? 24 : 6;  // 1 / int(u)
    std::vector<ld> AtA((size_t)Np * Np, 0), Atb((size_t)3 * Np, 0), phiInt(Np, 0);
    for (int g = 0; g < nG; ++g)
        for (int n = 0; n < Np; ++n) phiInt[n] += (ld)d->elWeight[g] * d->elBasisFct[(size_t)g * Np + n];
    for (int g = 0; g < nG; ++g)
        for (int up = 0; up < dim; ++up)
            for (int n = 0; n < Np; ++n) {
                const ld a = d->elUGradBasisFct[((size_t)g * Np + n) * 3 + up];
                for (int m = 0; m < Np; ++m) AtA[(size_t)n * Np + m] += a * d->elUGradBasisFct[((size_t)g * Np + m) * 3 + up];
                Atb[(size_t)up * Np + n] += a;  // rhs delta_{u,up}: row of coordinate u = up picks these
            }
    const ld rowScale = nG;  // weight of the integral condition relative to the gradient rows
    for (int n = 0; n < Np; ++n)
        for (int m = 0; m < Np; ++m) AtA[(size_t)n * Np + m] += rowScale * phiInt[n] * phiInt[m] * cInt * cInt;  // (cInt phiInt).(cInt phiInt)
    invertLD(AtA, Np);
    std::vector<ld> X((size_t)3 * Np, 0);
    for (int u = 0; u < dim; ++u)
        for (int n = 0; n < Np; ++n) {
            ld s = 0;
            for (int m = 0; m < Np; ++m) s += AtA[(size_t)n * Np + m] * (Atb[(size_t)u * Np + m] + rowScale * cInt * phiInt[m]);
            X[(size_t)u * Np + n] = s;
        }
    S.alpha.assign((size_t)Np * 4, 0);
    for (int n = 0; n < Np; ++n) {
        int sum = 0;
        for (int u = 0; u < dim; ++u) {
            const ld v = X[(size_t)u * Np + n] * N;
            const long r = lroundl(v);
            if (fabsl(v - r) > 1e-6L || r < 0 || r > N) throw std::runtime_error("Bernstein path: the element's nodes are not equispaced");
            S.alpha[(size_t)n * 4 + 1 + u] = (int)r;
            sum += (int)r;
        }
        if (sum > N) throw std::runtime_error("Bernstein path: node outside the reference simplex");
        S.alpha[(size_t)n * 4] = N - sum;
    }
    // 2. canonical <-> mesh node order
    std::vector<int> seen(Np, -1);
    for (int n = 0; n < Np; ++n) {
        const int* a = &S.alpha[(size_t)n * 4];
        const int i = dim == 3 ? vidx(N, a[1], a[2], a[3]) : fidx(N, a[1], a[2]);
        if (seen[i] >= 0) throw std::runtime_error("Bernstein path: two nodes share a position");
        seen[i] = n;
        S.T.permC2G[i] = (uint8_t)n;
    }
    // 3. V, V^-1
    auto fact = [](int k) { ld f = 1; for (int i = 2; i <= k; ++i) f *= i; return f; };
    std::vector<ld> V((size_t)Np * Np);
    for (int n = 0; n < Np; ++n)
        for (int m = 0; m < Np; ++m) {
            const int* a = &S.alpha[(size_t)m * 4];
            ld v = fact(N);
            for (int j = 0; j < nV; ++j) {
                v /= fact(a[j]);
                const ld lam = (ld)S.alpha[(size_t)n * 4 + j] / N;
                for (int k = 0; k < a[j]; ++k) v *= lam;
            }
            V[(size_t)n * Np + m] = v;
        }
    std::vector<ld> Vi = V;
    invertLD(Vi, Np);
    S.V.resize(V.size()); S.Vinv.resize(V.size());
    for (size_t i = 0; i < V.size(); ++i) { S.V[i] = (double)V[i]; S.Vinv[i] = (double)Vi[i]; }
    // 4. faces: the mesh's local face-node lists (element 0 is the first owner of its faces), the canonical face of each
    S.faceNodes.resize((size_t)nV * Nfp);
    for (int lf = 0; lf < nV; ++lf) {
        const int f = d->elFId[lf];
        if (d->fNbrElId[2 * (size_t)f] != 0) throw std::runtime_error("Bernstein path: element 0 must be the first owner of its faces");
        for (int m = 0; m < Nfp; ++m) S.faceNodes[(size_t)lf * Nfp + m] = d->fNToElNId[((size_t)f * Nfp + m) * 2];
    }
    S.ownIdx.assign((size_t)nV * Nfp, 0);
    bool haveJ[4] = {false, false, false, false};
    for (int lf = 0; lf < nV; ++lf) {
        int J = -1;
        for (int j = 0; j < nV; ++j) {
            bool zero = true;
            for (int m = 0; m < Nfp; ++m) zero = zero && S.alpha[(size_t)S.faceNodes[(size_t)lf * Nfp + m] * 4 + j] == 0;
            if (zero) J = j;
        }
        if (J < 0 || haveJ[J]) throw std::runtime_error("Bernstein path: face-node lists do not describe the faces of the simplex");
        haveJ[J] = true;
        S.T.faceLf[J] = (uint8_t)lf;
        auto place = [&](int b, int vol) {
            S.ownIdx[(size_t)J * Nfp + b] = vol;
            const int node = S.T.permC2G[vol];
            int pos = -1;
            for (int m = 0; m < Nfp; ++m) if (S.faceNodes[(size_t)lf * Nfp + m] == node) pos = m;
            if (pos < 0) throw std::runtime_error("Bernstein path: face node list misses a face node");
            S.T.facePos[J][b] = (uint8_t)pos;
        };
        if (dim == 3) {
            for (int b1 = 0; b1 <= N; ++b1)
                for (int b2 = 0; b2 <= N - b1; ++b2) place(fidx(N, b1, b2), layerIdxRt(N, J, 0, b1, b2));
        } else {
            for (int b1 = 0; b1 <= N; ++b1) place(b1, layerIdxRt2(N, J, 0, b1));
        }
    }
    // 5. dense nodal lift (for the self-check): Mref^-1[:, faceNodes(lf)] Mf
    std::vector<ld> Minv((size_t)Np * Np, 0), Mf((size_t)Nfp * Nfp, 0);
    for (int g = 0; g < nG; ++g)
        for (int i = 0; i < Np; ++i) {
            const ld wi = (ld)d->elWeight[g] * d->elBasisFct[(size_t)g * Np + i];
            for (int j = 0; j < Np; ++j) Minv[(size_t)i * Np + j] += wi * d->elBasisFct[(size_t)g * Np + j];
        }
    invertLD(Minv, Np);
    for (int g = 0; g < nGf; ++g)
        for (int a = 0; a < Nfp; ++a) {
            const ld wa = (ld)d->fWeight[g] * d->fBasisFct[(size_t)g * Nfp + a];
            for (int b = 0; b < Nfp; ++b) Mf[(size_t)a * Nfp + b] += wa * d->fBasisFct[(size_t)g * Nfp + b];
        }
    S.liftNodal.assign((size_t)Np * nV * Nfp, 0.0);
    for (int i = 0; i < Np; ++i)
        for (int lf = 0; lf < nV; ++lf)
            for (int m = 0; m < Nfp; ++m) {
                ld s = 0;
                for (int n = 0; n < Nfp; ++n) s += Minv[(size_t)i * Np + S.faceNodes[(size_t)lf * Nfp + n]] * Mf[(size_t)n * Nfp + m];
                S.liftNodal[(size_t)i * nV * Nfp + lf * Nfp + m] = (double)s;
            }
    return S;
}

// largest deviation between the closed-form sparse lift of bb_ops.h and V^-1 LIFT_nodal V_face (relative to the largest entry)
template <int DIM, int N, int J>
inline double liftDeviationFace(const Setup& S) {
    typedef Simplex<DIM, N> SX;
    constexpr int NP = SX::NP, NFP = SX::NFP, NV = SX::NV;
    const int lf = S.T.faceLf[J];
    double worst = 0, scale = 0;
    for (int b = 0; b < NFP; ++b) {
        double x[NFP] = {}, zl[NP] = {}, out[NP] = {};
        x[b] = 1.0;
        SX::liftLocal(x, zl);
        SX::template scatterAdd<J>(zl, out);
        // dense: unit Bernstein coefficient b of the face -> nodal trace values V[faceNode m][its volume node] -> lift -> V^-1
        const int nodeB = S.T.permC2G[S.ownIdx[(size_t)J * NFP + b]];
        double nod[NP] = {};
        for (int i = 0; i < NP; ++i)
            for (int m = 0; m < NFP; ++m) {
                const int fnode = S.faceNodes[(size_t)lf * NFP + m];
                nod[i] += S.liftNodal[(size_t)i * NV * NFP + lf * NFP + m] * S.V[(size_t)fnode * NP + nodeB];
            }
        for (int ci = 0; ci < NP; ++ci) {
            double ref = 0;
            const int nodeC = S.T.permC2G[ci];
            for (int i = 0; i < NP; ++i) ref += S.Vinv[(size_t)nodeC * NP + i] * nod[i];
            worst = std::max(worst, std::fabs(ref - SX::FACE_SCALE * out[ci]));
            scale = std::max(scale, std::fabs(ref));
        }
    }
    return worst / (scale > 0 ? scale : 1.0);
}

template <int DIM, int N>
inline double liftDeviationDim(const Setup& S) {
    double w = std::max(std::max(liftDeviationFace<DIM, N, 0>(S), liftDeviationFace<DIM, N, 1>(S)), liftDeviationFace<DIM, N, 2>(S));
    if constexpr (DIM == 3) w = std::max(w, liftDeviationFace<DIM, N, 3>(S));
    return w;
}
template <int N>
inline double liftDeviation(const Setup& S) { return S.dim == 3 ? liftDeviationDim<3, N>(S) : liftDeviationDim<2, N>(S); }

}  // namespace bb
}  // namespace dgb
