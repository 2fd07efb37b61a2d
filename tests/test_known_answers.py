"""Known answers that follow from the reference's formulas, checked on the CPU oracle."""
import numpy as np
import pytest

from conftest import rel_l2


def make(pkg, mesh_dir, name, order, reflecting=True, v0=(0, 0, 0), c0=343.0, rho0=1.225):
    model = pkg.Model.make_cube(int(name[5:]), -10.0, 10.0, order) if name.startswith("cube:") else pkg.Model.open_msh(mesh_dir / name, order)
    mesh = pkg.Mesh(model, pkg.Config())
    mesh.set_physics(c0=c0, rho0=rho0, v0=v0, dt=0.1 * mesh.h_min() / (c0 * (2 * order + 1)))
    mesh.fBC[:] = 1 if reflecting else 0
    return mesh


@pytest.mark.parametrize("name,order", [("disk.msh", 3), ("cube:2", 4), ("line.msh", 1)])
def test_free_stream(pkg, oracle_mod, mesh_dir, name, order):
    """Constant pressure, fluid at rest, rigid walls: L(u) = 0 to rounding (both oracle modes)."""
    mesh = make(pkg, mesh_dir, name, order)
    u = np.zeros((4, mesh.N))
    u[0] = 3.0
    orc = oracle_mod.Oracle(mesh, threads=2)
    for mode in (0, 1):
        r = orc.eval_rhs(mode, u)
        assert np.abs(r).max() < 1e-8 * 343.0 ** 2 * 1.225


def test_polynomial_exactness_interior(pkg, oracle_mod, mesh_dir):
    """p linear in x, v = 0: dp/dt = 0 and dv/dt = -grad p / rho0 exactly on elements without boundary faces."""
    mesh = make(pkg, mesh_dir, "square.msh", 2)
    x = mesh.node_coords
    u = np.zeros((4, mesh.N))
    u[0] = 2.0 * x[:, 0] - 0.5 * x[:, 1] + 1.0
    r = oracle_mod.Oracle(mesh, threads=2).eval_rhs(1, u)
    interior_el = ~(mesh.fIsBoundary[mesh.elFId] == 1).any(axis=1)
    sel = np.repeat(interior_el, mesh.Np)
    assert np.abs(r[0][sel]).max() < 1e-9
    np.testing.assert_allclose(r[1][sel], -2.0 / 1.225, atol=1e-9)
    np.testing.assert_allclose(r[2][sel], 0.5 / 1.225, atol=1e-9)
    assert np.abs(r[3]).max() == 0.0


def test_quadrature_independence(pkg, oracle_mod, mesh_dir):
    """Faithful mode (the reference's quadrature loops) == operator mode (collapsed operators) on random data."""
    mesh = make(pkg, mesh_dir, "cube:2", 3, reflecting=False, v0=(30.0, 10.0, 5.0))
    mesh.fBC[::2] = 1
    u = np.random.default_rng(2).standard_normal((4, mesh.N))
    orc = oracle_mod.Oracle(mesh, threads=2)
    a, b = orc.eval_rhs(0, u), orc.eval_rhs(1, u)
    for q in range(4):
        assert rel_l2(a[q], b[q]) < 1e-13


def test_energy_is_non_increasing_with_rigid_walls(pkg, oracle_mod, mesh_dir):
    """sigma = +1 (upwind) and reflecting walls: E = int p^2/(2 rho0 c0^2) + rho0 |v|^2 / 2 never grows."""
    mesh = make(pkg, mesh_dir, "square.msh", 2)
    x = mesh.node_coords
    u = np.zeros((4, mesh.N))
    u[0] = np.exp(-(x[:, 0] ** 2 + x[:, 1] ** 2))
    orc = oracle_mod.Oracle(mesh, threads=2)
    M = np.einsum("g,gi,gj->ij", mesh.elWeight, mesh.elBasisFct, mesh.elBasisFct)
    det = mesh.elJacobianDet[:, 0]

    def energy(v):
        e = 0.0
        for q, wq in ((0, 1.0 / (2 * 1.225 * 343.0 ** 2)), (1, 1.225 / 2), (2, 1.225 / 2)):
            w = v[q].reshape(mesh.K, mesh.Np)
            e += wq * np.einsum("k,ki,ij,kj->", det, w, M, w)
        return e

    e_prev = energy(u)
    for _ in range(6):
        orc.run(1, pkg.RUNGE_KUTTA, u, 0.0, 10)
        e = energy(u)
        assert e <= e_prev * (1 + 1e-12)
        e_prev = e


def test_operators_reproduce_derivatives(pkg, oracle_mod, mesh_dir):
    """Dw^u is the weak derivative: for nodal values of a polynomial of degree <= p, sum_u Dw^u applied to a constant
    vanishes after the lift of the same constant is subtracted (the free-stream identity in operator form)."""
    mesh = make(pkg, mesh_dir, "cube:1", 4)
    dw, lift = oracle_mod.Oracle(mesh).operators()
    ones = np.ones(mesh.Np)
    # M^-1 K^u 1 = M^-1 (boundary integral of phi_i n_u): equals sum over faces of LIFT_f (n_u on the reference faces)
    n_ref = np.array([[0, 0, -1], [0, -1, 0], [-1, 0, 0], [1, 1, 1]], dtype=float)  # faces {0,2,1},{0,1,3},{0,3,2},{3,1,2}
    area = np.array([1.0, 1.0, 1.0, 1.0])
    for u in range(3):
        lhs = dw[u] @ ones
        rhs = sum(lift[:, lf * mesh.Nfp:(lf + 1) * mesh.Nfp] @ (np.full(mesh.Nfp, n_ref[lf, u] * area[lf])) for lf in range(4))
        np.testing.assert_allclose(lhs, rhs, atol=1e-10)
