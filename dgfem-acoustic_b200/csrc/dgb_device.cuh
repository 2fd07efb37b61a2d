// Device helpers shared by the stage kernels.
#pragma once
#include "dgb_internal.h"

namespace dgb {

struct Phys {
    double c0, rho0, invRho, rc2, v0[3];
};

// Numerical flux through one face node, already multiplied by the outward orientation:
//   interior   : 1/2 n.(F(q-)+F(q+)) + 1/2 tau c0 (q- - q+)          Mesh.cpp:519-527 (+ getElFlux sign, :554)
//   absorbing  : RKR rows                                            Mesh.cpp:391-418, 652-667
//   reflecting : physical flux of the wall-tangent ghost state       Mesh.cpp:616-647
__device__ __forceinline__ void faceFlux(int bc, double tau, const double n[3], const Phys& ph, const double qm[4],
                                         const double qp[4], double fl[4]) {
    const double v0n = ph.v0[0] * n[0] + ph.v0[1] * n[1] + ph.v0[2] * n[2];
    if (bc == FACE_INTERIOR) {
        const double ps = qm[0] + qp[0];
        const double vs0 = qm[1] + qp[1], vs1 = qm[2] + qp[2], vs2 = qm[3] + qp[3];
        const double vns = n[0] * vs0 + n[1] * vs1 + n[2] * vs2;
        const double pen = 0.5 * tau * ph.c0;
        const double psr = ps * ph.invRho;
        fl[0] = 0.5 * (v0n * ps + ph.rc2 * vns) + pen * (qm[0] - qp[0]);
        fl[1] = 0.5 * (v0n * vs0 + n[0] * psr) + pen * (qm[1] - qp[1]);
        fl[2] = 0.5 * (v0n * vs1 + n[1] * psr) + pen * (qm[2] - qp[2]);
        fl[3] = 0.5 * (v0n * vs2 + n[2] * psr) + pen * (qm[3] - qp[3]);
    } else if (bc == FACE_ABSORBING) {
        const double vn = n[0] * qm[1] + n[1] * qm[2] + n[2] * qm[3];
        const double a = 0.25 * ph.c0 * qm[0] + 0.25 * ph.c0 * ph.c0 * ph.rho0 * vn;
        const double b = 0.25 * qm[0] * ph.invRho + 0.25 * ph.c0 * vn;
        fl[0] = a;
        fl[1] = n[0] * b;
        fl[2] = n[1] * b;
        fl[3] = n[2] * b;
    } else {
        const double vn = n[0] * qm[1] + n[1] * qm[2] + n[2] * qm[3];
        const double g0 = qm[1] - vn * n[0], g1 = qm[2] - vn * n[1], g2 = qm[3] - vn * n[2];
        const double pr = qm[0] * ph.invRho;
        fl[0] = v0n * qm[0] + ph.rc2 * (n[0] * g0 + n[1] * g1 + n[2] * g2);
        fl[1] = v0n * g0 + n[0] * pr;
        fl[2] = v0n * g1 + n[1] * pr;
        fl[3] = v0n * g2 + n[2] * pr;
    }
}

__device__ __forceinline__ Phys makePhys(const DeviceMesh& M) {
    Phys ph;
    ph.c0 = M.c0; ph.rho0 = M.rho0; ph.invRho = 1.0 / M.rho0; ph.rc2 = M.rho0 * M.c0 * M.c0;
    ph.v0[0] = M.v0[0]; ph.v0[1] = M.v0[1]; ph.v0[2] = M.v0[2];
    return ph;
}

// k = dt * L(y) combined with the RK registers; `yown` is the stage input at this entry.
__device__ __forceinline__ void rkUpdate(const StageArgs& A, int64_t g, double rhs, double yown) {
    const double k = __dmul_rn(A.dt, rhs);  // rounded product first, like alpha*(A*x) in eigen::linEq
    switch (A.mode) {
        case MODE_RK1: A.acc[g] = k; A.yout[g] = yown + 0.5 * k; break;
        case MODE_RK2: A.acc[g] = A.acc[g] + 2 * k; A.yout[g] = A.u[g] + 0.5 * k; break;
        case MODE_RK3: A.acc[g] = A.acc[g] + 2 * k; A.yout[g] = A.u[g] + 1 * k; break;
        case MODE_RK4: A.u[g] = A.u[g] + (A.acc[g] + k) / 6.0; break;
        case MODE_EULER: A.yout[g] = 1.0 * yown + k; break;
        default: A.yout[g] = rhs; break;
    }
}

// Same update with k = dt*L(y) already formed (the tiled kernel folds dt into the operators) and the RK registers in
// hand (uval = u, or the stage input for Euler; accval = acc). The final combine multiplies by 1/6 instead of dividing:
// one FP64 instruction instead of a division sequence, <= 1 ulp away from the reference expression.
__device__ __forceinline__ void rkApplyK(const StageArgs& A, int64_t g, double k, double uval, double accval) {
    switch (A.mode) {
        case MODE_RK1: A.acc[g] = k; A.yout[g] = fma(0.5, k, uval); break;
        case MODE_RK2: A.acc[g] = fma(2.0, k, accval); A.yout[g] = fma(0.5, k, uval); break;
        case MODE_RK3: A.acc[g] = fma(2.0, k, accval); A.yout[g] = uval + k; break;
        case MODE_RK4: A.u[g] = fma(accval + k, 1.0 / 6.0, uval); break;
        case MODE_EULER: A.yout[g] = uval + k; break;
        default: A.yout[g] = k; break;
    }
}

// Same update with the RK registers already in hand (uval = u, or the stage input for Euler; accval = acc).
__device__ __forceinline__ void rkApply(const StageArgs& A, int64_t g, double rhs, double uval, double accval) {
    const double k = __dmul_rn(A.dt, rhs);
    switch (A.mode) {
        case MODE_RK1: A.acc[g] = k; A.yout[g] = uval + 0.5 * k; break;
        case MODE_RK2: A.acc[g] = accval + 2 * k; A.yout[g] = uval + 0.5 * k; break;
        case MODE_RK3: A.acc[g] = accval + 2 * k; A.yout[g] = uval + 1 * k; break;
        case MODE_RK4: A.u[g] = uval + (accval + k) / 6.0; break;
        case MODE_EULER: A.yout[g] = 1.0 * uval + k; break;
        default: A.yout[g] = rhs; break;
    }
}

}  // namespace dgb
