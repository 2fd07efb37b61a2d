// Runs the Bernstein-Bezier CUDA kernels (dgfem-acoustic_b200/csrc/stage_bb.cu, the file itself) on the CPU through
// cuda_emu.h — TEST INFRASTRUCTURE ONLY. The engine's device layout (what dgb_create uploads: Ginv, per-face geometry,
// neighbour ids, flags, de-duplicated face-node maps) is rebuilt here for one GPU from the desc, following
// csrc/dgb_api.cu (createImpl), so that the kernels see on the host exactly what they see on the device.
#define DGB_EMULATE 1
#include "cuda_emu.h"

#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../dgfem-acoustic_b200/csrc/bb_setup.h"
#include "../dgfem-acoustic_b200/csrc/stage_bb.cu"

using namespace dgb;

namespace {
thread_local std::string g_err;

struct Emu {
    DeviceMesh M{};
    bb::Setup S;
    StageKernel kernel;
    std::vector<double> Ginv, fgeo;
    std::vector<int32_t> fnbr, fflags, faceNodes;
    std::vector<uint8_t> maps;
    int Np = 0, K = 0;
    double dt = 0;
};

Emu* build(const dgb_desc* d, int variant) {
    auto* E = new Emu;
    try {
        E->S = bb::buildSetup(d);
        const int Np = d->Np, Nfp = d->Nfp, Nf = d->Nf, K = d->K, gE = d->nGeomEl, gF = d->nGeomF;
        E->Np = Np; E->K = K; E->dt = d->dt;
        E->faceNodes = E->S.faceNodes;
        E->Ginv.resize((size_t)K * 9); E->fgeo.resize((size_t)K * Nf * 4); E->fnbr.resize((size_t)K * Nf); E->fflags.resize((size_t)K * Nf);
        std::map<std::vector<uint8_t>, int> mapIds;
        std::vector<int> pos(Np, -1);
        for (int el = 0; el < K; ++el) {
            const double* J = &d->elJacobian[(size_t)el * gE * 9];
            double A[3][3], B[3][3];
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) A[r][c] = J[r * 3 + c];
            const double det = A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) +
                               A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) {
                    const int r1 = (c + 1) % 3, r2 = (c + 2) % 3, c1 = (r + 1) % 3, c2 = (r + 2) % 3;
                    B[r][c] = (A[r1][c1] * A[r2][c2] - A[r1][c2] * A[r2][c1]) / det;
                }
            for (int x = 0; x < 3; ++x) for (int u = 0; u < 3; ++u) E->Ginv[(size_t)el * 9 + x * 3 + u] = B[x][u];
            const double detE = d->elJacobianDet[(size_t)el * gE];
            for (int lf = 0; lf < Nf; ++lf) {
                const int f = d->elFId[(size_t)el * Nf + lf];
                const int side = d->fNbrElId[2 * (size_t)f] == el ? 0 : 1;
                const int o = d->elFOrientation[(size_t)el * Nf + lf];
                double* fg = &E->fgeo[((size_t)el * Nf + lf) * 4];
                for (int x = 0; x < 3; ++x) fg[x] = o * d->fNormal[(size_t)f * gF * 3 + x];
                fg[3] = d->fJacobianDet[(size_t)f * gF] / detE;
                int flags;
                if (d->fIsBoundary[f]) {
                    flags = d->fBC[f] == 1 ? FACE_REFLECTING : FACE_ABSORBING;
                    E->fnbr[(size_t)el * Nf + lf] = -1;
                } else {
                    E->fnbr[(size_t)el * Nf + lf] = d->fNbrElId[2 * (size_t)f + (1 - side)];
                    const int tau = d->fc * o * (side == 0 ? 1 : -1);
                    flags = FACE_INTERIOR | (tau < 0 ? FLAG_TAU_NEG : 0);
                }
                std::fill(pos.begin(), pos.end(), -1);
                for (int m = 0; m < Nfp; ++m) pos[E->faceNodes[lf * Nfp + m]] = m;
                std::vector<uint8_t> mp(Nfp, 0);
                for (int n = 0; n < Nfp; ++n) {
                    const int own = d->fNToElNId[((size_t)f * Nfp + n) * 2 + side];
                    const int nb = d->fIsBoundary[f] ? own : d->fNToElNId[((size_t)f * Nfp + n) * 2 + (1 - side)];
                    mp[pos[own]] = (uint8_t)nb;
                }
                auto it = mapIds.find(mp);
                if (it == mapIds.end()) {
                    it = mapIds.emplace(mp, (int)mapIds.size()).first;
                    E->maps.insert(E->maps.end(), mp.begin(), mp.end());
                }
                E->fflags[(size_t)el * Nf + lf] = flags | (it->second << FLAG_MAP_SHIFT);
            }
        }
        DeviceMesh& M = E->M;
        M.dim = 3; M.order = d->order; M.Np = Np; M.Nfp = Nfp; M.Nf = Nf; M.L = 3 * Np + Nf * Nfp;
        M.Kown = M.Ktot = K;
        M.stride = (int64_t)K * Np;
        M.faceNodes = E->faceNodes.data(); M.nbrMaps = E->maps.data(); M.nMaps = (int)mapIds.size();
        M.Ginv = E->Ginv.data(); M.fgeo = E->fgeo.data(); M.fnbr = E->fnbr.data(); M.fflags = E->fflags.data();
        M.c0 = d->c0; M.rho0 = d->rho0; M.v0[0] = d->v0[0]; M.v0[1] = d->v0[1]; M.v0[2] = d->v0[2];
        setBBTables(d->order, E->S.T);
        E->kernel = selectBBKernel(3, d->order, variant);
        if (!E->kernel.launch) throw std::runtime_error("no Bernstein kernel for this order");
        return E;
    } catch (...) {
        delete E;
        throw;
    }
}
}  // namespace

extern "C" {

const char* bbe_last_error(void) { return g_err.c_str(); }

// nodal u -> rhs = L(u) (MODE_RHS) with the emulated kernel `variant` (0: stage_bb, 1: stage_bb_seq)
int bbe_eval_rhs(const dgb_desc* d, int variant, const double* u, double* rhs) {
    try {
        Emu* E = build(d, variant);
        const size_t n = (size_t)4 * E->K * E->Np;
        std::vector<double> yin(n), yout(n, 0.0), U(n, 0.0), acc(n, 0.0);
        launchElementMatrix(u, yin.data(), E->M.stride, E->Np, E->K, E->S.Vinv.data(), nullptr);
        StageArgs A{};
        A.yin = yin.data(); A.u = U.data(); A.acc = acc.data(); A.yout = yout.data(); A.eBegin = 0; A.eEnd = E->K; A.mode = MODE_RHS; A.dt = 1.0;
        E->kernel.launch(E->M, A, nullptr);
        launchElementMatrix(yout.data(), rhs, E->M.stride, E->Np, E->K, E->S.V.data(), nullptr);
        delete E;
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

// nsteps of RK4 (integrator 1) or forward Euler (0) with the emulated kernel, nodal in / nodal out, with an optional hard source
// (node list in GLOBAL node ids, applied at the start of every step like dgb_run does), t accumulating from t0.
int bbe_run(const dgb_desc* d, int variant, int integrator, double* u, double t0, int nsteps, int nSrcNodes, const int32_t* srcNodes,
            double amp, double freq, double phase, double duration) {
    try {
        Emu* E = build(d, variant);
        const int Np = E->Np;
        const size_t n = (size_t)4 * E->K * Np;
        std::vector<double> U(n), ACC(n, 0.0), YA(n, 0.0), YB(n, 0.0);
        launchElementMatrix(u, U.data(), E->M.stride, Np, E->K, E->S.Vinv.data(), nullptr);
        // source nodes grouped by element, as dgb_set_sources does
        std::vector<int32_t> nodes(srcNodes, srcNodes + nSrcNodes), elList, nodeOff, nodeLocal;
        std::sort(nodes.begin(), nodes.end());
        nodes.erase(std::unique(nodes.begin(), nodes.end()), nodes.end());
        for (size_t k = 0; k < nodes.size(); ++k) {
            const int el = nodes[k] / Np;
            if (k == 0 || el != nodes[k - 1] / Np) { elList.push_back(el); nodeOff.push_back((int32_t)nodeLocal.size()); }
            nodeLocal.push_back(nodes[k] - el * Np);
        }
        elList.push_back(-1);
        nodeOff.push_back((int32_t)nodeLocal.size());
        double *pU = U.data(), *pYA = YA.data(), *pYB = YB.data();
        double t = t0;
        for (int step = 0; step < nsteps; ++step, t += E->dt) {
            if (nSrcNodes > 0 && t < duration)
                launchSetNodesBB(pU, Np, elList.data(), nodeOff.data(), nodeLocal.data(), (int)elList.size() - 1, amp * sin(2 * M_PI * freq * t + phase),
                                 E->S.V.data(), E->S.Vinv.data(), nullptr);
            StageArgs A{};
            A.u = pU; A.acc = ACC.data(); A.dt = E->dt; A.eBegin = 0; A.eEnd = E->K;
            if (integrator == 0) {
                A.yin = pU; A.yout = pYA; A.mode = MODE_EULER; E->kernel.launch(E->M, A, nullptr);
                std::swap(pU, pYA);
                continue;
            }
            A.yin = pU;  A.yout = pYA; A.mode = MODE_RK1; E->kernel.launch(E->M, A, nullptr);
            A.yin = pYA; A.yout = pYB; A.mode = MODE_RK2; E->kernel.launch(E->M, A, nullptr);
            A.yin = pYB; A.yout = pYA; A.mode = MODE_RK3; E->kernel.launch(E->M, A, nullptr);
            A.yin = pYA; A.yout = nullptr; A.mode = MODE_RK4; E->kernel.launch(E->M, A, nullptr);
        }
        launchElementMatrix(pU, u, E->M.stride, Np, E->K, E->S.V.data(), nullptr);
        delete E;
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}
}
