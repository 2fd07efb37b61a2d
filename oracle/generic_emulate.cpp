// Runs the generic CUDA stage kernel and the small kernels (dgfem-acoustic_b200/csrc/stage_generic.cu, the file itself) on the
// CPU through cuda_emu.h — TEST INFRASTRUCTURE ONLY. This kernel is verified on B200 hardware; running it here as well gives
// the CPU suite a regression net for every (dimension, order) and shows what the emulation is worth on a known-good kernel.
#define DGB_EMULATE 1
#include "cuda_emu.h"

#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../dgfem-acoustic_b200/csrc/stage_generic.cu"
#include "emu_layout.h"

using namespace dgb;

namespace {
thread_local std::string g_err;
}

extern "C" {
const char* gne_last_error(void) { return g_err.c_str(); }

// integrator 1: nsteps of RK4, 0: forward Euler, 2: u <- L(u). Optional hard source (node list, global ids) and receivers
// (el, weights) recorded at the start of every step into rec[nsteps][nrecv][4].
int gne_run(const dgb_desc* d, int integrator, double* u, double t0, int nsteps, int nSrcNodes, const int32_t* srcNodes, double amp,
            double freq, double phase, double duration, int nrecv, const int32_t* recvEl, const double* recvW, double* rec) {
    try {
        emu::Layout L;
        emu::build(d, L);
        const StageKernel k = selectGenericKernel(d->dim, d->order);
        if (!k.launch) throw std::runtime_error("no generic kernel for this dimension / order");
        const size_t n = (size_t)4 * d->K * d->Np;
        std::vector<double> U(u, u + n), ACC(n, 0.0), YA(n, 0.0), YB(n, 0.0);
        double *pU = U.data(), *pYA = YA.data(), *pYB = YB.data();
        StageArgs A{};
        A.acc = ACC.data(); A.dt = d->dt; A.eBegin = 0; A.eEnd = d->K;
        if (integrator == 2) {
            A.yin = pU; A.u = pU; A.yout = pYA; A.mode = MODE_RHS; A.dt = 1.0;
            k.launch(L.M, A, nullptr);
            std::copy(YA.begin(), YA.end(), u);
            return 0;
        }
        double t = t0;
        for (int step = 0; step < nsteps; ++step, t += d->dt) {
            if (nrecv > 0) launchGatherReceivers(pU, L.M.stride, d->Np, recvEl, recvW, nrecv, rec + (size_t)step * nrecv * 4, nullptr);
            if (nSrcNodes > 0 && t < duration) launchSetNodes(pU, srcNodes, nSrcNodes, amp * sin(2 * M_PI * freq * t + phase), nullptr);
            A.u = pU;
            if (integrator == 0) {
                A.yin = pU; A.yout = pYA; A.mode = MODE_EULER; k.launch(L.M, A, nullptr);
                std::swap(pU, pYA);
                continue;
            }
            A.yin = pU;  A.yout = pYA; A.mode = MODE_RK1; k.launch(L.M, A, nullptr);
            A.yin = pYA; A.yout = pYB; A.mode = MODE_RK2; k.launch(L.M, A, nullptr);
            A.yin = pYB; A.yout = pYA; A.mode = MODE_RK3; k.launch(L.M, A, nullptr);
            A.yin = pYA; A.yout = nullptr; A.mode = MODE_RK4; k.launch(L.M, A, nullptr);
        }
        std::copy(pU, pU + n, u);
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}
}
