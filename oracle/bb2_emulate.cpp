// Runs the product's Bernstein-Bezier stage kernels (dgfem-acoustic_b200/csrc/stage_bb2.cu and stage_bbe.cu, the files themselves) on the CPU
// through cuda_emu.h — TEST INFRASTRUCTURE ONLY. One OS thread per lane of the one-warp CTAs; warp barriers and shuffles are real exchanges
// between those threads; the asynchronous copies (TMA bulk copies on mbarriers, cp.async) are synchronous host stand-ins (csrc/dgb_async.cuh under
// DGB_EMULATE). This checks what a host can check — tile walking, partial tiles, trace gathers through the padded neighbour table, face order,
// the two-pass epilogue, every index — against the oracle; what only the hardware shows (asynchrony, proxies, occupancy) is covered by the
// GPU tests. The device tables are rebuilt here the way dgb_create builds them (csrc/dgb_api.cu).
#define DGB_EMULATE 1
#include "cuda_emu.h"

#include <cstring>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "../dgfem-acoustic_b200/csrc/bb_setup.h"
#include "../dgfem-acoustic_b200/csrc/stage_bb2.cu"
#include "../dgfem-acoustic_b200/csrc/stage_bbe.cu"
#include "emu_layout.h"

using namespace dgb;

namespace {
thread_local std::string g_err;

struct Emu2 {
    emu::Layout L;
    bb::Setup S;
    std::vector<uint4> tabStore;        // 16-byte aligned storage of the byte tables
    std::vector<double> VC, VinvC;      // conversion matrices, canonical coefficient order
    StageKernel kernel;
};

void build2(const dgb_desc* d, int kernelId, Emu2& E) {
    E.S = bb::buildSetup(d);
    emu::build(d, E.L);
    if (E.L.faceNodes != E.S.faceNodes) throw std::runtime_error("face-node tables disagree");
    DeviceMesh& M = E.L.M;
    const int Np = d->Np, Nfp = d->Nfp, Nf = d->Nf;
    std::vector<uint8_t> permG2C(Np, 0);
    for (int i = 0; i < Np; ++i) permG2C[E.S.T.permC2G[i]] = (uint8_t)i;
    const size_t RS = (size_t)(Nfp + 15) / 16 * 16;
    std::vector<uint8_t> tab((size_t)M.nMaps * 4 * RS, 0);
    for (int J = 0; J < Nf; ++J)
        for (int b = 0; b < Nfp; ++b) {
            M.bbOwn[J][b] = (uint8_t)E.S.ownIdx[(size_t)J * Nfp + b];
            for (int mp = 0; mp < M.nMaps; ++mp)
                tab[((size_t)mp * 4 + J) * RS + b] = permG2C[E.L.maps[(size_t)mp * Nfp + E.S.T.facePos[J][b]]];
        }
    E.tabStore.assign((tab.size() + 15) / 16 + 1, uint4{0, 0, 0, 0});
    memcpy(E.tabStore.data(), tab.data(), tab.size());
    M.bbTab = nullptr;
    M.bbNbr16 = reinterpret_cast<const uint8_t*>(E.tabStore.data());
    for (int J = 0; J < Nf; ++J) M.bbFaceLf[J] = E.S.T.faceLf[J];
    E.VC.resize((size_t)Np * Np);
    E.VinvC.resize((size_t)Np * Np);
    for (int n = 0; n < Np; ++n)
        for (int i = 0; i < Np; ++i) {
            E.VC[(size_t)n * Np + i] = E.S.V[(size_t)n * Np + E.S.T.permC2G[i]];
            E.VinvC[(size_t)i * Np + n] = E.S.Vinv[(size_t)E.S.T.permC2G[i] * Np + n];
        }
    E.kernel = kernelId == 7 ? selectBBEKernel(d->dim, d->order) : selectBB2Kernel(d->dim, d->order);
    if (!E.kernel.launch) throw std::runtime_error("no such Bernstein kernel for this dimension / order");
}
}  // namespace

extern "C" {
const char* bb2e_last_error(void) { return g_err.c_str(); }

// kernel 6: stage_bb2, 7: stage_bbe. integrator 1: nsteps of RK4, 0: forward Euler, 2: u <- L(u). Nodal in, nodal out (the conversions to and from
// the interleaved Bernstein layout are done here on the host with the set-up's matrices).
int bb2e_run(const dgb_desc* d, int kernelId, int integrator, double* u, int nsteps) {
    try {
        Emu2 E;
        build2(d, kernelId, E);
        const int Np = d->Np, K = d->K;
        const size_t N = (size_t)K * Np, n = 4 * N;
        std::vector<double> U(n), ACC(n, 0.0), YA(n, 0.0), YB(n, 0.0);
        for (int el = 0; el < K; ++el)
            for (int i = 0; i < Np; ++i)
                for (int q = 0; q < 4; ++q) {
                    double s = 0;
                    for (int m = 0; m < Np; ++m) s += E.VinvC[(size_t)i * Np + m] * u[q * N + (size_t)el * Np + m];
                    U[((size_t)el * Np + i) * 4 + q] = s;
                }
        double *pU = U.data(), *pYA = YA.data(), *pYB = YB.data();
        StageArgs A{};
        A.acc = ACC.data(); A.dt = d->dt; A.eBegin = 0; A.eEnd = K; A.fx = nullptr;
        const DeviceMesh& M = E.L.M;
        if (integrator == 2) {
            A.yin = pU; A.u = pU; A.yout = pYA; A.mode = MODE_RHS; A.dt = 1.0;
            E.kernel.launch(M, A, nullptr);
            pU = pYA;
        } else {
            for (int step = 0; step < nsteps; ++step) {
                A.u = pU;
                if (integrator == 0) {
                    A.yin = pU; A.yout = pYA; A.mode = MODE_EULER; E.kernel.launch(M, A, nullptr);
                    std::swap(pU, pYA);
                    continue;
                }
                A.yin = pU;  A.yout = pYA; A.mode = MODE_RK1; E.kernel.launch(M, A, nullptr);
                A.yin = pYA; A.yout = pYB; A.mode = MODE_RK2; E.kernel.launch(M, A, nullptr);
                A.yin = pYB; A.yout = pYA; A.mode = MODE_RK3; E.kernel.launch(M, A, nullptr);
                A.yin = pYA; A.yout = nullptr; A.mode = MODE_RK4; E.kernel.launch(M, A, nullptr);
            }
        }
        for (int el = 0; el < K; ++el)
            for (int m = 0; m < Np; ++m)
                for (int q = 0; q < 4; ++q) {
                    double s = 0;
                    for (int i = 0; i < Np; ++i) s += E.VC[(size_t)m * Np + i] * pU[((size_t)el * Np + i) * 4 + q];
                    u[q * N + (size_t)el * Np + m] = s;
                }
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}
}
