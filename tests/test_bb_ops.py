"""Bernstein-Bezier operator code (dgfem-acoustic_b200/csrc/bb_ops.h, bb_setup.h) on the CPU.

oracle/bb_check.cpp compiles the very templates the CUDA kernel stage_bb.cu instantiates for the host and evaluates
L(u) element by element as  V * rhs_Bernstein(V^-1 u); here it is compared with the oracle's operator mode
(tolerance 1e-12 relative L2 per field) on structured and unstructured tetrahedral meshes, all orders, both boundary
conditions, with and without mean flow, and both penalty signs. Also checked: the closed-form sparse lift equals
Mref^-1 E_lf Mf in the Bernstein basis, and the reference nodes recovered from the desc's tables are the equispaced ones."""
import ctypes as C

import numpy as np
import pytest

from conftest import ROOT, rel_l2

dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def bbc():
    lib = C.CDLL(str(ROOT / "oracle" / "libbbcheck.so"))
    lib.bbc_last_error.restype = C.c_char_p
    lib.bbc_eval_rhs.argtypes = [C.c_void_p, dp, dp, dp, C.POINTER(C.c_int32)]
    lib.bbc_eval_rhs_simplex.argtypes = [C.c_void_p, dp, dp, dp, C.POINTER(C.c_int32)]
    return lib


def _mesh(pkg, mesh_dir, name, order, v0):
    if name.startswith("cube:"):
        model = pkg.Model.make_cube(int(name.split(":")[1]), -10.0, 10.0, order)
    elif name.startswith("square:"):
        model = pkg.Model.make_square(int(name.split(":")[1]), -10.0, 10.0, order)
    else:
        model = pkg.Model.open_msh(mesh_dir / name, order)
    mesh = pkg.Mesh(model, pkg.Config())
    mesh.set_physics(c0=343.0, rho0=1.225, v0=v0, dt=1e-5)
    b = np.nonzero(mesh.fIsBoundary)[0]
    mesh.fBC[b[::2]] = 1
    return mesh


def _state(mesh, seed=0):
    rng = np.random.default_rng(seed)
    x = mesh.node_coords
    u = np.zeros((4, mesh.N))
    for q in range(4):
        k, ph = rng.uniform(0.5, 2, 3), rng.uniform(0, 6, 3)
        u[q] = np.cos(k[0] * x[:, 0] * 0.3 + ph[0]) * np.cos(k[1] * x[:, 1] * 0.3 + ph[1]) * np.cos(k[2] * x[:, 2] * 0.3 + ph[2])
    u[1:] *= 1e-3
    return u


@pytest.mark.parametrize("name,order,v0,flip_fc", [
    ("cube:2", 1, (0.0, 0.0, 0.0), False), ("cube:2", 2, (30.0, 10.0, -5.0), False), ("cube:3", 3, (0.0, 0.0, 0.0), False),
    ("cube:2", 4, (30.0, 10.0, -5.0), False), ("cube:3", 4, (0.0, 0.0, 0.0), True), ("cube:2", 5, (1.0, 2.0, 3.0), False),
    ("cube:2", 6, (0.0, 0.0, 0.0), False), ("cube.msh", 3, (3.0, 2.0, 1.0), False), ("sphere.msh", 4, (0.0, 0.0, 0.0), False)])
def test_bernstein_rhs_equals_the_oracle(pkg, oracle_mod, bbc, mesh_dir, name, order, v0, flip_fc):
    mesh = _mesh(pkg, mesh_dir, name, order, v0)
    if flip_fc:
        mesh.desc.fc = -mesh.desc.fc
    u = _state(mesh)
    ref = oracle_mod.Oracle(mesh).eval_rhs(oracle_mod.Oracle.OPERATOR, u)
    rhs = np.zeros_like(u)
    dev = C.c_double(-1.0)
    alpha = np.zeros((mesh.Np, 4), dtype=np.int32)
    rc = bbc.bbc_eval_rhs(C.cast(mesh.desc_p, C.c_void_p), u.ctypes.data_as(dp), rhs.ctypes.data_as(dp), C.byref(dev),
                          alpha.ctypes.data_as(C.POINTER(C.c_int32)))
    assert rc == 0, bbc.bbc_last_error()
    assert 0 <= dev.value < 1e-12  # closed-form lift == V^-1 (Mref^-1 E Mf) V_face
    for q in range(4):
        assert rel_l2(rhs[q], ref[q]) < 1e-12
    # the nodes recovered from the basis tables are where the mesh has them: x_n = sum_j (alpha_nj / N) * vertex_j
    x = mesh.node_coords.reshape(mesh.K, mesh.Np, 3)
    assert (alpha.sum(axis=1) == order).all()
    assert np.abs((alpha / order) @ x[5, :4] - x[5]).max() < 1e-11


@pytest.mark.parametrize("name,order,v0,flip_fc", [
    ("square:3", 1, (0.0, 0.0, 0.0), False), ("square.msh", 1, (30.0, 10.0, 0.0), False), ("square:3", 2, (30.0, 10.0, 0.0), False),
    ("disk.msh", 2, (0.0, 0.0, 0.0), False), ("square.msh", 3, (3.0, -2.0, 0.0), False), ("square_reflection.msh", 3, (0.0, 0.0, 0.0), True),
    ("square:2", 4, (30.0, 10.0, 0.0), False), ("square_reflection.msh", 4, (0.0, 0.0, 0.0), False), ("square:2", 5, (1.0, 2.0, 0.0), False),
    ("disk.msh", 6, (0.0, 0.0, 0.0), False), ("square:2", 6, (5.0, 0.0, 0.0), True),
    # tetrahedra once more through the Simplex<3, N> interface (what stage_bb2.cu calls)
    ("cube:2", 1, (30.0, 10.0, -5.0), False), ("cube:2", 3, (30.0, 10.0, -5.0), True), ("cube:2", 4, (0.0, 0.0, 0.0), False)])
def test_simplex_interface_equals_the_oracle(pkg, oracle_mod, bbc, mesh_dir, name, order, v0, flip_fc):
    """Triangles (orders 1..6, the reference's own 2D meshes and a refined square) and tetrahedra through bb::Simplex<DIM, N>."""
    mesh = _mesh(pkg, mesh_dir, name, order, v0)
    if flip_fc:
        mesh.desc.fc = -mesh.desc.fc
    u = _state(mesh)
    ref = oracle_mod.Oracle(mesh).eval_rhs(oracle_mod.Oracle.OPERATOR, u)
    rhs = np.zeros_like(u)
    dev = C.c_double(-1.0)
    alpha = np.zeros((mesh.Np, 4), dtype=np.int32)
    rc = bbc.bbc_eval_rhs_simplex(C.cast(mesh.desc_p, C.c_void_p), u.ctypes.data_as(dp), rhs.ctypes.data_as(dp), C.byref(dev),
                                  alpha.ctypes.data_as(C.POINTER(C.c_int32)))
    assert rc == 0, bbc.bbc_last_error()
    assert 0 <= dev.value < 1e-12
    for q in range(4):
        if np.abs(ref[q]).max() == 0:  # v_z on a 2D mesh without mean flow: continuous traces, n_z = 0 -> L(u)_vz == 0; the Bernstein path leaves rounding noise
            assert np.linalg.norm(rhs[q]) < 1e-10 * np.linalg.norm(u[q])
        else:
            assert rel_l2(rhs[q], ref[q]) < 1e-12
    assert (alpha.sum(axis=1) == order).all()
    nv = mesh.dim + 1
    x = mesh.node_coords.reshape(mesh.K, mesh.Np, 3)
    assert np.abs((alpha[:, :nv] / order) @ x[5, :nv] - x[5]).max() < 1e-11
