// Runs the curved-element CUDA kernel (dgfem-acoustic_b200/csrc/stage_curved.cu, the file itself) on the CPU through
// cuda_emu.h — TEST INFRASTRUCTURE ONLY. The CurvedMesh is what dgb_create uploads for a curved mesh: the desc's own arrays
// plus the inverse element mass matrices of curved_setup.h.
#define DGB_EMULATE 1
#include "cuda_emu.h"

#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../dgfem-acoustic_b200/csrc/stage_curved.cu"
#include "../dgfem-acoustic_b200/csrc/stage_generic.cu"  // the collapsed kernel of the straight-sided prefix of a mixed handle
#include "emu_layout.h"

using namespace dgb;

namespace {
thread_local std::string g_err;

CurvedMesh fromDesc(const dgb_desc* d, const std::vector<double>& Minv, int first) {
    if (d->nGeomEl != d->nG || d->nGeomF != d->nGf) throw std::runtime_error("the desc must carry one Jacobian / normal per integration point");
    CurvedMesh C{};
    C.dim = d->dim; C.Np = d->Np; C.Nfp = d->Nfp; C.Nf = d->Nf; C.K = d->K; C.F = d->F; C.nG = d->nG; C.nGf = d->nGf; C.fc = d->fc;
    C.elBasis = d->elBasisFct; C.elUGrad = d->elUGradBasisFct; C.elWeight = d->elWeight; C.fBasis = d->fBasisFct; C.fWeight = d->fWeight;
    C.elJac = d->elJacobian; C.elDet = d->elJacobianDet; C.fNormal = d->fNormal; C.fDet = d->fJacobianDet;
    C.elFId = d->elFId; C.elFOrientation = d->elFOrientation; C.fNbrElId = d->fNbrElId; C.fNToElNId = d->fNToElNId;
    C.fIsBoundary = d->fIsBoundary; C.fBC = d->fBC; C.Minv = Minv.data(); C.firstCurved = first;
    C.c0 = d->c0; C.rho0 = d->rho0; C.v0[0] = d->v0[0]; C.v0[1] = d->v0[1]; C.v0[2] = d->v0[2];
    C.stride = (int64_t)d->K * d->Np;
    return C;
}
}  // namespace

extern "C" {
const char* cve_last_error(void) { return g_err.c_str(); }

int cve_is_curved(const dgb_desc* d) { return isCurved(d) ? 1 : 0; }

// first element of the curved suffix (dgb_create's decision): K if nothing is curved, 0 if everything goes through the curved kernel
int cve_first_curved(const dgb_desc* d) {
    const std::vector<uint8_t> flag = curvedElements(d);
    bool any = false;
    for (uint8_t f : flag) any = any || f;
    return any ? curvedSuffixStart(flag) : d->K;
}

// rhs = L(u) of the elements >= first only (the others keep u's values): the curved half of a mixed handle
int cve_rhs_suffix(const dgb_desc* d, int first, double* u) {
    try {
        const std::vector<double> Minv = curvedInverseMass(d, first);
        const CurvedMesh C = fromDesc(d, Minv, first);
        const size_t n = (size_t)4 * d->K * d->Np;
        std::vector<double> U(u, u + n), ACC(n, 0.0), Y(u, u + n);
        StageArgs A{};
        A.yin = U.data(); A.u = U.data(); A.acc = ACC.data(); A.yout = Y.data(); A.mode = MODE_RHS; A.dt = 1.0; A.eBegin = first; A.eEnd = d->K;
        launchCurved(C, A, nullptr);
        std::copy(Y.begin(), Y.end(), u);
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

// integrator 1: nsteps of RK4, 0: forward Euler, 2: rhs = L(u) written over u (MODE_RHS)
int cve_run(const dgb_desc* d, int integrator, double* u, int nsteps) {
    try {
        const std::vector<double> Minv = curvedInverseMass(d);
        const CurvedMesh C = fromDesc(d, Minv, 0);
        const size_t n = (size_t)4 * d->K * d->Np;
        std::vector<double> U(u, u + n), ACC(n, 0.0), YA(n, 0.0), YB(n, 0.0);
        double *pU = U.data(), *pYA = YA.data(), *pYB = YB.data();
        StageArgs A{};
        A.acc = ACC.data(); A.dt = d->dt; A.eBegin = 0; A.eEnd = d->K;
        if (integrator == 2) {
            A.yin = pU; A.u = pU; A.yout = pYA; A.mode = MODE_RHS; A.dt = 1.0;
            launchCurved(C, A, nullptr);
            std::copy(YA.begin(), YA.end(), u);
            return 0;
        }
        for (int step = 0; step < nsteps; ++step) {
            A.u = pU;
            if (integrator == 0) {
                A.yin = pU; A.yout = pYA; A.mode = MODE_EULER; launchCurved(C, A, nullptr);
                std::swap(pU, pYA);
                continue;
            }
            A.yin = pU;  A.yout = pYA; A.mode = MODE_RK1; launchCurved(C, A, nullptr);
            A.yin = pYA; A.yout = pYB; A.mode = MODE_RK2; launchCurved(C, A, nullptr);
            A.yin = pYB; A.yout = pYA; A.mode = MODE_RK3; launchCurved(C, A, nullptr);
            A.yin = pYA; A.yout = nullptr; A.mode = MODE_RK4; launchCurved(C, A, nullptr);
        }
        std::copy(pU, pU + n, u);
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

// A mixed handle as dgb_run drives it (launchStage in csrc/dgb_api.cu): per stage the collapsed generic kernel on the
// straight-sided prefix [0, first) and the curved kernel on the suffix [first, K), both updating the same RK registers.
int cve_run_mixed(const dgb_desc* d, int integrator, double* u, int nsteps) {
    try {
        const std::vector<uint8_t> flag = curvedElements(d);
        const int first = curvedSuffixStart(flag);
        const std::vector<double> Minv = curvedInverseMass(d, first);
        const CurvedMesh C = fromDesc(d, Minv, first);
        emu::Layout L;
        emu::build(d, L);
        const StageKernel k = selectGenericKernel(d->dim, d->order);
        const size_t n = (size_t)4 * d->K * d->Np;
        std::vector<double> U(u, u + n), ACC(n, 0.0), YA(n, 0.0), YB(n, 0.0);
        double *pU = U.data(), *pYA = YA.data(), *pYB = YB.data();
        StageArgs A{};
        A.acc = ACC.data(); A.dt = d->dt;
        auto stage = [&](const double* yin, double* yout, int mode) {
            A.yin = yin; A.yout = yout; A.mode = mode; A.u = pU;
            StageArgs B = A;
            B.eBegin = 0; B.eEnd = first;
            if (first > 0) k.launch(L.M, B, nullptr);
            B.eBegin = first; B.eEnd = d->K;
            launchCurved(C, B, nullptr);
        };
        for (int step = 0; step < nsteps; ++step) {
            if (integrator == 0) { stage(pU, pYA, MODE_EULER); std::swap(pU, pYA); continue; }
            stage(pU, pYA, MODE_RK1);
            stage(pYA, pYB, MODE_RK2);
            stage(pYB, pYA, MODE_RK3);
            stage(pYA, nullptr, MODE_RK4);
        }
        std::copy(pU, pU + n, u);
        return first;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}
}
