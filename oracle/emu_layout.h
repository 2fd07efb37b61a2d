// Host copy of the engine's device layout for the CUDA-emulation harnesses — TEST INFRASTRUCTURE ONLY.
// Follows dgfem-acoustic_b200/csrc/dgb_api.cu (buildOperators, createImpl) for a single-GPU handle: the reference-element
// operators DwT / nLiftT, the inverse Jacobians, per-face geometry, neighbour ids, flags and the de-duplicated face-node maps.
#pragma once
#include <cmath>
#include <map>
#include <stdexcept>
#include <vector>

#include "../dgfem-acoustic_b200/csrc/dgb_internal.h"

namespace emu {

typedef long double real;

inline void invertDense(std::vector<real>& A, int n) {
    std::vector<real> B((size_t)n * n, 0);
    for (int i = 0; i < n; ++i) B[(size_t)i * n + i] = 1;
    for (int c = 0; c < n; ++c) {
        int piv = c;
        for (int r = c + 1; r < n; ++r) if (fabsl(A[(size_t)r * n + c]) > fabsl(A[(size_t)piv * n + c])) piv = r;
        if (A[(size_t)piv * n + c] == 0) throw std::runtime_error("singular matrix");
        if (piv != c)
            for (int k = 0; k < n; ++k) { std::swap(A[(size_t)c * n + k], A[(size_t)piv * n + k]); std::swap(B[(size_t)c * n + k], B[(size_t)piv * n + k]); }
        const real d = 1 / A[(size_t)c * n + c];
        for (int k = 0; k < n; ++k) { A[(size_t)c * n + k] *= d; B[(size_t)c * n + k] *= d; }
        for (int r = 0; r < n; ++r) {
            if (r == c) continue;
            const real f = A[(size_t)r * n + c];
            if (f == 0) continue;
            for (int k = 0; k < n; ++k) { A[(size_t)r * n + k] -= f * A[(size_t)c * n + k]; B[(size_t)r * n + k] -= f * B[(size_t)c * n + k]; }
        }
    }
    A.swap(B);
}

struct Layout {
    dgb::DeviceMesh M{};
    std::vector<double> DwT, nLiftT, Ginv, fgeo;
    std::vector<int32_t> faceNodes, fnbr, fflags;
    std::vector<uint8_t> maps;
};

inline void build(const dgb_desc* d, Layout& L) {
    using namespace dgb;
    const int Np = d->Np, Nfp = d->Nfp, Nf = d->Nf, dim = d->dim, nG = d->nG, nGf = d->nGf, K = d->K, gE = d->nGeomEl, gF = d->nGeomF;
    // operators: Mref, K^u, Dw^u = Mref^-1 K^u (stored transposed), Mf, -LIFT (transposed)
    std::vector<real> Minv((size_t)Np * Np, 0), Mf((size_t)Nfp * Nfp, 0);
    for (int g = 0; g < nG; ++g)
        for (int i = 0; i < Np; ++i) {
            const real wi = (real)d->elWeight[g] * d->elBasisFct[(size_t)g * Np + i];
            for (int j = 0; j < Np; ++j) Minv[(size_t)i * Np + j] += wi * d->elBasisFct[(size_t)g * Np + j];
        }
    invertDense(Minv, Np);
    L.DwT.assign((size_t)dim * Np * Np, 0.0);
    for (int u = 0; u < dim; ++u) {
        std::vector<real> Ku((size_t)Np * Np, 0);
        for (int g = 0; g < nG; ++g)
            for (int i = 0; i < Np; ++i) {
                const real wi = (real)d->elWeight[g] * d->elUGradBasisFct[((size_t)g * Np + i) * 3 + u];
                for (int j = 0; j < Np; ++j) Ku[(size_t)i * Np + j] += wi * d->elBasisFct[(size_t)g * Np + j];
            }
        for (int i = 0; i < Np; ++i)
            for (int j = 0; j < Np; ++j) {
                real s = 0;
                for (int k = 0; k < Np; ++k) s += Minv[(size_t)i * Np + k] * Ku[(size_t)k * Np + j];
                L.DwT[((size_t)u * Np + j) * Np + i] = (double)s;
            }
    }
    for (int g = 0; g < nGf; ++g)
        for (int n = 0; n < Nfp; ++n) {
            const real wn = (real)d->fWeight[g] * d->fBasisFct[(size_t)g * Nfp + n];
            for (int m = 0; m < Nfp; ++m) Mf[(size_t)n * Nfp + m] += wn * d->fBasisFct[(size_t)g * Nfp + m];
        }
    L.faceNodes.resize((size_t)Nf * Nfp);
    for (int lf = 0; lf < Nf; ++lf) {
        const int f = d->elFId[lf];
        if (d->fNbrElId[2 * (size_t)f] != 0) throw std::runtime_error("element 0 must be the first owner of its faces");
        for (int m = 0; m < Nfp; ++m) L.faceNodes[lf * Nfp + m] = d->fNToElNId[((size_t)f * Nfp + m) * 2];
    }
    const int NFL = Nf * Nfp;
    L.nLiftT.assign((size_t)NFL * Np, 0.0);
    for (int i = 0; i < Np; ++i)
        for (int lf = 0; lf < Nf; ++lf)
            for (int m = 0; m < Nfp; ++m) {
                real s = 0;
                for (int n = 0; n < Nfp; ++n) s += Minv[(size_t)i * Np + L.faceNodes[lf * Nfp + n]] * Mf[(size_t)n * Nfp + m];
                L.nLiftT[((size_t)lf * Nfp + m) * Np + i] = (double)(-s);
            }
    // per-element / per-face geometry and connectivity
    L.Ginv.resize((size_t)K * dim * dim); L.fgeo.resize((size_t)K * Nf * 4); L.fnbr.resize((size_t)K * Nf); L.fflags.resize((size_t)K * Nf);
    std::map<std::vector<uint8_t>, int> mapIds;
    std::vector<int> pos(Np, -1);
    for (int el = 0; el < K; ++el) {
        const double* J = &d->elJacobian[(size_t)el * gE * 9];
        real A[3][3], B[3][3];
        for (int r = 0; r < dim; ++r) for (int c = 0; c < dim; ++c) A[r][c] = J[r * 3 + c];
        if (dim == 1) B[0][0] = 1 / A[0][0];
        else if (dim == 2) {
            const real det = A[0][0] * A[1][1] - A[0][1] * A[1][0];
            B[0][0] = A[1][1] / det; B[0][1] = -A[0][1] / det; B[1][0] = -A[1][0] / det; B[1][1] = A[0][0] / det;
        } else {
            const real det = A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) +
                             A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) {
                    const int r1 = (c + 1) % 3, r2 = (c + 2) % 3, c1 = (r + 1) % 3, c2 = (r + 2) % 3;
                    B[r][c] = (A[r1][c1] * A[r2][c2] - A[r1][c2] * A[r2][c1]) / det;
                }
        }
        for (int x = 0; x < dim; ++x) for (int u = 0; u < dim; ++u) L.Ginv[(size_t)el * dim * dim + x * dim + u] = (double)B[x][u];
        const double detE = d->elJacobianDet[(size_t)el * gE];
        for (int lf = 0; lf < Nf; ++lf) {
            const int f = d->elFId[(size_t)el * Nf + lf];
            const int side = d->fNbrElId[2 * (size_t)f] == el ? 0 : 1;
            const int o = d->elFOrientation[(size_t)el * Nf + lf];
            double* fg = &L.fgeo[((size_t)el * Nf + lf) * 4];
            for (int x = 0; x < 3; ++x) fg[x] = o * d->fNormal[(size_t)f * gF * 3 + x];
            fg[3] = d->fJacobianDet[(size_t)f * gF] / detE;
            int flags;
            if (d->fIsBoundary[f]) {
                flags = d->fBC[f] == 1 ? FACE_REFLECTING : FACE_ABSORBING;
                L.fnbr[(size_t)el * Nf + lf] = -1;
            } else {
                L.fnbr[(size_t)el * Nf + lf] = d->fNbrElId[2 * (size_t)f + (1 - side)];
                const int tau = d->fc * o * (side == 0 ? 1 : -1);
                flags = FACE_INTERIOR | (tau < 0 ? FLAG_TAU_NEG : 0);
            }
            std::fill(pos.begin(), pos.end(), -1);
            for (int m = 0; m < Nfp; ++m) pos[L.faceNodes[lf * Nfp + m]] = m;
            std::vector<uint8_t> mp(Nfp, 0);
            for (int n = 0; n < Nfp; ++n) {
                const int own = d->fNToElNId[((size_t)f * Nfp + n) * 2 + side];
                const int nb = d->fIsBoundary[f] ? own : d->fNToElNId[((size_t)f * Nfp + n) * 2 + (1 - side)];
                mp[pos[own]] = (uint8_t)nb;
            }
            auto it = mapIds.find(mp);
            if (it == mapIds.end()) {
                it = mapIds.emplace(mp, (int)mapIds.size()).first;
                L.maps.insert(L.maps.end(), mp.begin(), mp.end());
            }
            L.fflags[(size_t)el * Nf + lf] = flags | (it->second << FLAG_MAP_SHIFT);
        }
    }
    DeviceMesh& M = L.M;
    M.dim = dim; M.order = d->order; M.Np = Np; M.Nfp = Nfp; M.Nf = Nf; M.L = dim * Np + Nf * Nfp;
    M.Kown = M.Ktot = K;
    M.stride = (int64_t)K * Np;
    M.DwT = L.DwT.data(); M.nLiftT = L.nLiftT.data(); M.tiledOps = nullptr;
    M.faceNodes = L.faceNodes.data(); M.nbrMaps = L.maps.data(); M.nMaps = (int)mapIds.size();
    M.Ginv = L.Ginv.data(); M.fgeo = L.fgeo.data(); M.fnbr = L.fnbr.data(); M.fflags = L.fflags.data();
    M.c0 = d->c0; M.rho0 = d->rho0; M.v0[0] = d->v0[0]; M.v0[1] = d->v0[1]; M.v0[2] = d->v0[2];
}

}  // namespace emu
