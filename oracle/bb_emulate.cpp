// Runs the Bernstein-Bezier CUDA kernels (dgfem-acoustic_b200/csrc/stage_bb.cu, the file itself) on the CPU through
// cuda_emu.h — TEST INFRASTRUCTURE ONLY. The engine's device layout (what dgb_create uploads: Ginv, per-face geometry,
// neighbour ids, flags, de-duplicated face-node maps) is rebuilt for one GPU from the desc by emu_layout.h, following
// csrc/dgb_api.cu (createImpl), so that the kernels see on the host exactly what they see on the device.
#define DGB_EMULATE 1
#include "cuda_emu.h"

#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../dgfem-acoustic_b200/csrc/bb_setup.h"
#include "../dgfem-acoustic_b200/csrc/stage_bb.cu"
#include "emu_layout.h"

using namespace dgb;

namespace {
thread_local std::string g_err;
int g_tile = 32;  // elements per CTA (bbe_set_tile)

struct Emu {
    emu::Layout L;
    DeviceMesh& M = L.M;
    bb::Setup S;
    StageKernel kernel;
    int Np = 0, K = 0;
    double dt = 0;
};

Emu* build(const dgb_desc* d, int variant) {
    auto* E = new Emu;
    try {
        E->S = bb::buildSetup(d);
        E->Np = d->Np; E->K = d->K; E->dt = d->dt;
        emu::build(d, E->L);  // what dgb_create uploads
        if (E->L.faceNodes != E->S.faceNodes) throw std::runtime_error("face-node tables disagree");
        setBBTables(d->order, E->S.T);
        E->kernel = selectBBKernel(3, d->order, variant, g_tile);
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
void bbe_set_tile(int tile) { g_tile = tile; }

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
