// Bernstein-Bezier form of the element operators on straight-sided tetrahedra ("sparse-operator variant").
//
// The nodal operators Dw^u (Np x Np dense) and LIFT (Np x Nf*Nfp dense) of the reference's scheme become sparse when the
// polynomial is carried by its Bernstein coefficients c_a, u = sum_|a|=N c_a B_a(lambda), B_a = N!/a! lambda^a, instead of
// its values at the equispaced nodes (u_nodal = V c, V_na = B_a(x_n), cond(V) = 15 at N = 4):
//   * derivative along a direction with barycentric weights w_j = d(lambda_j)/ds:
//         t_b = sum_j w_j c_{b+e_j}   (|b| = N-1, 4 terms),    (du/ds)_g = sum_k g_k t_{g-e_k}   (|g| = N, <= 4 terms)
//   * the numerical flux is linear with face-constant coefficients, so it acts on the coefficients directly, and the face
//     trace of u is carried by the coefficients with a_J = 0 (J = the vertex opposite the face), through the same
//     neighbour node maps as the nodal scheme (coefficient a <-> node a/N);
//   * lift of a face vector x (|b| = N, 2D): with the unnormalised 2D degree elevation (E_M x)_g = sum_k g_k x_{g-e_k},
//         layer 0 (a_J = 0):  y   = E_N^T E_N x
//         layer l (a_J = l):  z_l = -1/(l+1) * E_{N-l}^T z_{l-1},   z_0 = y
//     (Chan & Warburton's factorisation; the constants were fitted against Mref^-1 E_lf Mf of this code's reference
//     element and hold to 1e-13 for N = 2..5, tests/test_bb_ops.py).
// The strong form  rhs = -div F + LIFT(Fscale * (n.F(u-) - flux*))  equals the weak form the reference integrates
// (exact quadrature on affine elements): tests compare against the oracle to 1e-12.
//
// Everything here is index arithmetic that unrolls completely: arrays live in registers, coefficients are small
// integers. __host__ __device__ so that the CPU tests run the very same code the kernel runs (oracle/bb_check.cpp).
#pragma once

#if defined(__CUDACC__)
#define BB_HD __host__ __device__ __forceinline__
#define BB_UNROLL _Pragma("unroll")
#else
#define BB_HD inline
#define BB_UNROLL
#endif

#include <stdint.h>

namespace dgb {
namespace bb {

BB_HD constexpr int tri(int m) { return (m + 1) * (m + 2) / 2; }            // 2D indices of degree m
BB_HD constexpr int tet(int m) { return (m + 1) * (m + 2) * (m + 3) / 6; }  // 3D indices of degree m

// canonical order of the 3D indices a = (a0,a1,a2,a3), |a| = m: lexicographic in (a1, a2, a3), a0 = m - a1 - a2 - a3
BB_HD constexpr int vidx(int m, int a1, int a2, int a3) {
    // indices with a smaller a1: sum_{i<a1} tri(m-i) = tet(m) - tet(m-a1); with this a1 and a smaller a2: sum_{j<a2} (m-a1-j+1)
    return tet(m) - tet(m - a1) + a2 * (m - a1 + 1) - a2 * (a2 - 1) / 2 + a3;
}
// canonical order of the 2D indices b = (b0,b1,b2), |b| = m: lexicographic in (b1, b2)
BB_HD constexpr int fidx(int m, int b1, int b2) { return b1 * (m + 1) - b1 * (b1 - 1) / 2 + b2; }

// volume index of the coefficient in layer l of face J (a_J = l) whose other three entries, in increasing vertex order,
// are (b0, b1, b2), b0 = N - l - b1 - b2
template <int N, int J>
BB_HD constexpr int layerIdx(int l, int b1, int b2) {
    const int b0 = N - l - b1 - b2;
    // vertices other than J in increasing order: J=0 -> (1,2,3), J=1 -> (0,2,3), J=2 -> (0,1,3), J=3 -> (0,1,2)
    const int a1 = J == 0 ? b0 : J == 1 ? l : b1;
    const int a2 = J == 0 ? b1 : J == 1 ? b1 : J == 2 ? l : b2;
    const int a3 = J == 0 ? b2 : J == 1 ? b2 : J == 2 ? b2 : l;
    return vidx(N, a1, a2, a3);
}

// t_b (+)= sum_j w_j c_{b+e_j}  for |b| = N-1: the degree N-1 coefficients of the derivative of u along the direction whose
// barycentric rates are w (up to the factor N that the elevation back to degree N cancels)
template <int N, bool ACCUMULATE>
BB_HD void dirDeriv(const double (&c)[tet(N)], const double (&w)[4], double (&t)[tet(N - 1)]) {
    BB_UNROLL
    for (int b1 = 0; b1 <= N - 1; ++b1) {
        BB_UNROLL
        for (int b2 = 0; b2 <= N - 1 - b1; ++b2) {
            BB_UNROLL
            for (int b3 = 0; b3 <= N - 1 - b1 - b2; ++b3) {
                double s = w[0] * c[vidx(N, b1, b2, b3)];
                s = s + w[1] * c[vidx(N, b1 + 1, b2, b3)];
                s = s + w[2] * c[vidx(N, b1, b2 + 1, b3)];
                s = s + w[3] * c[vidx(N, b1, b2, b3 + 1)];
                const int i = vidx(N - 1, b1, b2, b3);
                t[i] = ACCUMULATE ? t[i] + s : s;
            }
        }
    }
}

// out_g += scale * sum_k g_k t_{g-e_k}  for |g| = N  (degree elevation N-1 -> N, unnormalised)
template <int N>
BB_HD void elevateAdd(const double (&t)[tet(N - 1)], double scale, double (&out)[tet(N)]) {
    BB_UNROLL
    for (int a1 = 0; a1 <= N; ++a1) {
        BB_UNROLL
        for (int a2 = 0; a2 <= N - a1; ++a2) {
            BB_UNROLL
            for (int a3 = 0; a3 <= N - a1 - a2; ++a3) {
                const int a0 = N - a1 - a2 - a3;
                double s = 0.0;
                if (a0 > 0) s = s + a0 * t[vidx(N - 1, a1, a2, a3)];
                if (a1 > 0) s = s + a1 * t[vidx(N - 1, a1 - 1, a2, a3)];
                if (a2 > 0) s = s + a2 * t[vidx(N - 1, a1, a2 - 1, a3)];
                if (a3 > 0) s = s + a3 * t[vidx(N - 1, a1, a2, a3 - 1)];
                const int i = vidx(N, a1, a2, a3);
                out[i] = out[i] + scale * s;
            }
        }
    }
}

// (E_M^T w)_b = sum_k (b_k + 1) w_{b+e_k} : degree M+1 -> M on a triangle, scaled
template <int M>
BB_HD void faceLower(const double (&w)[tri(M + 1)], double scale, double (&z)[tri(M)]) {
    BB_UNROLL
    for (int b1 = 0; b1 <= M; ++b1) {
        BB_UNROLL
        for (int b2 = 0; b2 <= M - b1; ++b2) {
            const int b0 = M - b1 - b2;
            const double s = (b0 + 1) * w[fidx(M + 1, b1, b2)] + (b1 + 1) * w[fidx(M + 1, b1 + 1, b2)] + (b2 + 1) * w[fidx(M + 1, b1, b2 + 1)];
            z[fidx(M, b1, b2)] = scale == 1.0 ? s : scale * s;
        }
    }
}

// (E_M x)_g = sum_k g_k x_{g-e_k} : degree M -> M+1 on a triangle
template <int M>
BB_HD void faceRaise(const double (&x)[tri(M)], double (&w)[tri(M + 1)]) {
    BB_UNROLL
    for (int g1 = 0; g1 <= M + 1; ++g1) {
        BB_UNROLL
        for (int g2 = 0; g2 <= M + 1 - g1; ++g2) {
            const int g0 = M + 1 - g1 - g2;
            double s = 0.0;
            if (g0 > 0) s = s + g0 * x[fidx(M, g1, g2)];
            if (g1 > 0) s = s + g1 * x[fidx(M, g1 - 1, g2)];
            if (g2 > 0) s = s + g2 * x[fidx(M, g1, g2 - 1)];
            w[fidx(M + 1, g1, g2)] = s;
        }
    }
}

// ---- lift of one face, split into a face-independent part and the scatter into the element's coefficients --------------
// face-local result: layer l (degree N-l, canonical 2D order) at offset layerOff(N, l); tet(N) values in all
BB_HD constexpr int layerOff(int N, int l) {
    int s = 0;
    for (int k = 0; k < l; ++k) s += tri(N - k);
    return s;
}

// cumulative layer factor: z_l = layerScale(l) * (E^T ... E^T y) with the unnormalised lowering operators
BB_HD constexpr double layerScale(int l) {
    double s = 1.0;
    for (int k = 1; k <= l; ++k) s *= -1.0 / (k + 1);
    return s;
}

template <int N, int L>
struct LiftLayers {
    // z = UNSCALED layer L (integer-coefficient sums only); stores it and descends with E_{N-L-1}^T. The factor
    // layerScale(L) is applied by the scatter (one fused multiply-add instead of a multiply here and an add there).
    static BB_HD void run(const double (&z)[tri(N - L)], double (&zl)[tet(N)]) {
        BB_UNROLL
        for (int b = 0; b < tri(N - L); ++b) zl[layerOff(N, L) + b] = z[b];
        if constexpr (L < N) {
            double zn[tri(N - L - 1)];
            faceLower<N - L - 1>(z, 1.0, zn);
            LiftLayers<N, L + 1>::run(zn, zl);
        }
    }
};

// zl = all layers of LIFT x in face-local order, layer l still to be multiplied by layerScale(l): x = Fscale * (n.F(u-) - flux*)
// as the Bernstein coefficients of the face polynomial (canonical 2D order). The same code for the four faces.
template <int N>
BB_HD void liftFaceLocal(const double (&x)[tri(N)], double (&zl)[tet(N)]) {
    double w[tri(N + 1)], y[tri(N)];
    faceRaise<N>(x, w);
    faceLower<N>(w, 1.0, y);
    LiftLayers<N, 0>::run(y, zl);
}

// out[coefficient of layer l, 2D index b of face J] += layerScale(l) * zl[layer l][b]
template <int N, int J>
BB_HD void scatterAddFace(const double (&zl)[tet(N)], double (&out)[tet(N)]) {
    BB_UNROLL
    for (int l = 0; l <= N; ++l) {
        BB_UNROLL
        for (int b1 = 0; b1 <= N - l; ++b1) {
            BB_UNROLL
            for (int b2 = 0; b2 <= N - l - b1; ++b2) {
                const int i = layerIdx<N, J>(l, b1, b2);
                const double v = zl[layerOff(N, l) + fidx(N - l, b1, b2)];
                out[i] = l == 0 ? out[i] + v : out[i] + layerScale(l) * v;
            }
        }
    }
}

// out += LIFT_J x
template <int N, int J>
BB_HD void liftFace(const double (&x)[tri(N)], double (&out)[tet(N)]) {
    double zl[tet(N)];
    liftFaceLocal<N>(x, zl);
    scatterAddFace<N, J>(zl, out);
}

// ---- triangles ----------------------------------------------------------------------------------------------------------
// The same operators one dimension down (configs 1 and 2 are 2D). Volume indices a = (a0,a1,a2), |a| = N, in the
// canonical order fidx(N, a1, a2); an edge carries the N+1 coefficients with a_J = 0 (J = the vertex opposite the edge),
// indexed by b1 where (b0, b1) are the other two entries in increasing vertex order. The lift has the very same closed form
// with the 1D elevation (E_M x)_g = g0 x_{g-e0} + g1 x_{g-e1}:  y = E_N^T E_N x,  z_l = -1/(l+1) E_{N-l}^T z_{l-1}
// (checked against Mref^-1 E_lf Mf at create time and in tests/test_bb_ops.py, N = 1..6).
namespace d2 {

template <int N, int J>
BB_HD constexpr int layerIdx(int l, int b1) {
    const int b0 = N - l - b1;
    const int a1 = J == 0 ? b0 : J == 1 ? l : b1;
    const int a2 = J == 0 ? b1 : J == 1 ? b1 : l;
    return fidx(N, a1, a2);
}

template <int N, bool ACCUMULATE>
BB_HD void dirDeriv(const double (&c)[tri(N)], const double (&w)[3], double (&t)[tri(N - 1)]) {
    BB_UNROLL
    for (int b1 = 0; b1 <= N - 1; ++b1) {
        BB_UNROLL
        for (int b2 = 0; b2 <= N - 1 - b1; ++b2) {
            double s = w[0] * c[fidx(N, b1, b2)];
            s = s + w[1] * c[fidx(N, b1 + 1, b2)];
            s = s + w[2] * c[fidx(N, b1, b2 + 1)];
            const int i = fidx(N - 1, b1, b2);
            t[i] = ACCUMULATE ? t[i] + s : s;
        }
    }
}

template <int N>
BB_HD void elevateAdd(const double (&t)[tri(N - 1)], double scale, double (&out)[tri(N)]) {
    BB_UNROLL
    for (int a1 = 0; a1 <= N; ++a1) {
        BB_UNROLL
        for (int a2 = 0; a2 <= N - a1; ++a2) {
            const int a0 = N - a1 - a2;
            double s = 0.0;
            if (a0 > 0) s = s + a0 * t[fidx(N - 1, a1, a2)];
            if (a1 > 0) s = s + a1 * t[fidx(N - 1, a1 - 1, a2)];
            if (a2 > 0) s = s + a2 * t[fidx(N - 1, a1, a2 - 1)];
            const int i = fidx(N, a1, a2);
            out[i] = out[i] + scale * s;
        }
    }
}

// (E_M^T w)_b = (b0 + 1) w_b + (b1 + 1) w_{b+e1} : degree M+1 -> M on an edge (index = b1)
template <int M>
BB_HD void edgeLower(const double (&w)[M + 2], double (&z)[M + 1]) {
    BB_UNROLL
    for (int b1 = 0; b1 <= M; ++b1) z[b1] = (M - b1 + 1) * w[b1] + (b1 + 1) * w[b1 + 1];
}

// (E_M x)_g = g0 x_g + g1 x_{g-e1} : degree M -> M+1 on an edge
template <int M>
BB_HD void edgeRaise(const double (&x)[M + 1], double (&w)[M + 2]) {
    BB_UNROLL
    for (int g1 = 0; g1 <= M + 1; ++g1) {
        const int g0 = M + 1 - g1;
        double s = 0.0;
        if (g0 > 0) s = s + g0 * x[g1];
        if (g1 > 0) s = s + g1 * x[g1 - 1];
        w[g1] = s;
    }
}

BB_HD constexpr int layerOff(int N, int l) {
    int s = 0;
    for (int k = 0; k < l; ++k) s += N - k + 1;
    return s;
}

template <int N, int L>
struct LiftLayers {
    static BB_HD void run(const double (&z)[N - L + 1], double (&zl)[tri(N)]) {
        BB_UNROLL
        for (int b = 0; b <= N - L; ++b) zl[layerOff(N, L) + b] = z[b];
        if constexpr (L < N) {
            double zn[N - L];
            edgeLower<N - L - 1>(z, zn);
            LiftLayers<N, L + 1>::run(zn, zl);
        }
    }
};

// all layers of LIFT x in edge-local order, layer l still to be multiplied by layerScale(l)
template <int N>
BB_HD void liftFaceLocal(const double (&x)[N + 1], double (&zl)[tri(N)]) {
    double w[N + 2], y[N + 1];
    edgeRaise<N>(x, w);
    edgeLower<N>(w, y);
    LiftLayers<N, 0>::run(y, zl);
}

template <int N, int J>
BB_HD void scatterAddFace(const double (&zl)[tri(N)], double (&out)[tri(N)]) {
    BB_UNROLL
    for (int l = 0; l <= N; ++l) {
        BB_UNROLL
        for (int b1 = 0; b1 <= N - l; ++b1) {
            const int i = layerIdx<N, J>(l, b1);
            const double v = zl[layerOff(N, l) + b1];
            out[i] = l == 0 ? out[i] + v : out[i] + layerScale(l) * v;
        }
    }
}

template <int N, int J>
BB_HD void liftFace(const double (&x)[N + 1], double (&out)[tri(N)]) {
    double zl[tri(N)];
    liftFaceLocal<N>(x, zl);
    scatterAddFace<N, J>(zl, out);
}

}  // namespace d2

// One interface over both simplices for the kernels that are written once (stage_bb2.cu): NV vertices / faces, NP volume and
// NFP face coefficients, ND coefficients of degree N-1. LIFT x = FACE_SCALE * scatter(liftLocal(x)): the closed forms are those
// of a reference face of measure 1/2 (triangle) resp. 1 (edge), Gmsh's reference edge [-1, 1] has measure 2 — the callers fold
// the factor into Fscale.
template <int DIM, int N>
struct Simplex;
template <int N>
struct Simplex<3, N> {
    static constexpr int NV = 4, NP = tet(N), NFP = tri(N), ND = tet(N - 1);
    static constexpr double FACE_SCALE = 1.0;
    static BB_HD void dirDerivAcc(const double (&c)[NP], const double (&w)[NV], double (&t)[ND]) { dirDeriv<N, true>(c, w, t); }
    static BB_HD void elevate(const double (&t)[ND], double scale, double (&out)[NP]) { elevateAdd<N>(t, scale, out); }
    static BB_HD void liftLocal(const double (&x)[NFP], double (&zl)[NP]) { liftFaceLocal<N>(x, zl); }
    template <int J>
    static BB_HD void scatterAdd(const double (&zl)[NP], double (&out)[NP]) { scatterAddFace<N, J>(zl, out); }
};
template <int N>
struct Simplex<2, N> {
    static constexpr int NV = 3, NP = tri(N), NFP = N + 1, ND = tri(N - 1);
    static constexpr double FACE_SCALE = 2.0;
    static BB_HD void dirDerivAcc(const double (&c)[NP], const double (&w)[NV], double (&t)[ND]) { d2::dirDeriv<N, true>(c, w, t); }
    static BB_HD void elevate(const double (&t)[ND], double scale, double (&out)[NP]) { d2::elevateAdd<N>(t, scale, out); }
    static BB_HD void liftLocal(const double (&x)[NFP], double (&zl)[NP]) { d2::liftFaceLocal<N>(x, zl); }
    template <int J>
    static BB_HD void scatterAdd(const double (&zl)[NP], double (&out)[NP]) {
        if constexpr (J < 3) d2::scatterAddFace<N, J>(zl, out);
    }
};

constexpr int MAX_ORDER = 6, MAX_NP = tet(MAX_ORDER), MAX_NFP = tri(MAX_ORDER);

// ---- face inputs of the lift ---------------------------------------------------------------------------------------------
// x = Fscale * (n.F(u-) - flux*) with the reference's numerical flux (interior: Mesh.cpp:519-527 with the penalty sign tau;
// absorbing: RKR rows, Mesh.cpp:391-418, 652-667; reflecting: Mesh.cpp:616-648). By linearity it only needs
//     a_q = q- - q+ (interior)   or   a_q = q- (boundary),     S = n . a_v :
//   interior    n.F(a) / 2 - tau c0 a_q / 2
//   reflecting  x_p = rho0 c0^2 S                          x_v = v0n S n
//   absorbing   x_p = (v0n - c0/4) a_p + 3/4 rho0 c0^2 S   x_v = v0n a_v + n (3/4 a_p / rho0 - c0/4 S)
// all of the form  x_p = app a_p + aps S,   x_vx = b a_vx + n_x (c a_p + d S)  with five face-constant coefficients.
constexpr int BC_INTERIOR = 0, BC_ABSORBING = 1, BC_REFLECTING = 2;  // == FACE_* of dgb_internal.h
struct FaceCoef {
    double app, aps, b, c, d;
};
BB_HD FaceCoef faceCoef(int bc, double tau, double fscale, double v0n, double c0, double rho0) {
    const double rc2 = rho0 * c0 * c0;
    FaceCoef k;
    if (bc == BC_INTERIOR) {
        k.app = 0.5 * fscale * (v0n - tau * c0);
        k.aps = 0.5 * fscale * rc2;
        k.b = k.app;
        k.c = 0.5 * fscale / rho0;
        k.d = 0.0;
    } else if (bc == BC_ABSORBING) {
        k.app = fscale * (v0n - 0.25 * c0);
        k.aps = 0.75 * fscale * rc2;
        k.b = fscale * v0n;
        k.c = 0.75 * fscale / rho0;
        k.d = -0.25 * fscale * c0;
    } else {
        k.app = 0.0;
        k.aps = fscale * rc2;
        k.b = 0.0;
        k.c = 0.0;
        k.d = fscale * v0n;
    }
    return k;
}
BB_HD void faceInput(const FaceCoef& k, const double (&n)[3], const double (&a)[4], double (&x)[4]) {
    const double S = n[0] * a[1] + n[1] * a[2] + n[2] * a[3];
    const double g = k.c * a[0] + k.d * S;
    x[0] = k.app * a[0] + k.aps * S;
    x[1] = k.b * a[1] + n[0] * g;
    x[2] = k.b * a[2] + n[1] * g;
    x[3] = k.b * a[3] + n[2] * g;
}

// Permutations between the mesh's element-local numbering and the canonical orders above (built by bb_setup.h, passed
// to the kernels by value)
struct Tables {
    uint8_t permC2G[MAX_NP];      // canonical volume index -> node of the mesh's element-local numbering
    uint8_t faceLf[4];            // canonical face J (opposite vertex J) -> local face of the mesh
    uint8_t facePos[4][MAX_NFP];  // canonical (J, 2D index) -> position m in the mesh's face-node list of that local face
};

template <int N, int J>
BB_HD void liftFaceFrom(const double* dphi, const Tables& T, double (&out)[tet(N)]) {
    double x[tri(N)];
    const int base = T.faceLf[J] * tri(N);
    BB_UNROLL
    for (int b = 0; b < tri(N); ++b) x[b] = dphi[base + T.facePos[J][b]];
    liftFace<N, J>(x, out);
}

// Right-hand side of ONE field of ONE element as Bernstein coefficients in canonical order:
//   out = -(v0.grad q + coupling) + sum_J LIFT_J dphi_J
//   q = 0 (p):   coupling = rho0 c0^2 div v          q = 1..3 (v_x): coupling = (1/rho0) dp/dx_x
// col0 + f*colStride points at the element's coefficients of field f (mesh node order), dphi at this field's face inputs
// Fscale * (n.F(u-) - flux*) in the mesh's (local face, face node) order, gl[j][x] = d lambda_j / d x.
// (No array is indexed with a run-time value: everything stays in registers.)
template <int N>
BB_HD void fieldVolume(int q, const double* col0, int colStride, const Tables& T, const double (&gl)[4][3], const double (&v0)[3],
                       bool flow, double rc2, double invRho, double (&out)[tet(N)]) {
    constexpr int NP = tet(N), ND = tet(N - 1);
    double t[ND];
    BB_UNROLL
    for (int i = 0; i < ND; ++i) t[i] = 0.0;
    const int nCoupling = q == 0 ? 3 : 1;
    const int nPass = nCoupling + (flow ? 1 : 0);
    for (int pass = 0; pass < nPass; ++pass) {  // run-time trip count: one copy of the body
        int field;
        double w[4];
        if (pass < nCoupling) {
            field = q == 0 ? 1 + pass : 0;
            const int x = q == 0 ? pass : q - 1;
            const double s = q == 0 ? rc2 : invRho;
            BB_UNROLL
            for (int j = 0; j < 4; ++j) w[j] = s * (x == 0 ? gl[j][0] : x == 1 ? gl[j][1] : gl[j][2]);
        } else {
            field = q;
            BB_UNROLL
            for (int j = 0; j < 4; ++j) w[j] = v0[0] * gl[j][0] + v0[1] * gl[j][1] + v0[2] * gl[j][2];
        }
        const double* col = col0 + field * colStride;
        double cc[NP];
        BB_UNROLL
        for (int i = 0; i < NP; ++i) cc[i] = col[T.permC2G[i]];
        dirDeriv<N, true>(cc, w, t);
    }
    BB_UNROLL
    for (int i = 0; i < NP; ++i) out[i] = 0.0;
    elevateAdd<N>(t, -1.0, out);
}

template <int N>
BB_HD void fieldRhs(int q, const double* col0, int colStride, const double* dphi, const Tables& T, const double (&gl)[4][3],
                    const double (&v0)[3], bool flow, double rc2, double invRho, double (&out)[tet(N)]) {
    fieldVolume<N>(q, col0, colStride, T, gl, v0, flow, rc2, invRho, out);
    liftFaceFrom<N, 0>(dphi, T, out);
    liftFaceFrom<N, 1>(dphi, T, out);
    liftFaceFrom<N, 2>(dphi, T, out);
    liftFaceFrom<N, 3>(dphi, T, out);
}

}  // namespace bb
}  // namespace dgb
