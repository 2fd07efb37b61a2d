"""Curved (non-affine) elements, SURVEY.md §8 f3.

CPU: the curved-element CUDA kernel (dgfem-acoustic_b200/csrc/stage_curved.cu, the file itself) runs through the CUDA
emulation of oracle/cuda_emu.h on warped isoparametric meshes and is compared with the oracle's FAITHFUL mode — the
reference's own loops, which tests/test_oracle_vs_reference.py pins against the reference binary on such meshes.
GPU (gated, the kernel was written after the round's GPU budget was spent): the engine itself through the C ABI."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT, rel_l2

dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def cve():
    lib = C.CDLL(str(ROOT / "oracle" / "libcurvedemu.so"))
    lib.cve_last_error.restype = C.c_char_p
    lib.cve_run.argtypes = [C.c_void_p, C.c_int, dp, C.c_int]
    lib.cve_is_curved.argtypes = [C.c_void_p]
    lib.cve_first_curved.argtypes = [C.c_void_p]
    lib.cve_rhs_suffix.argtypes = [C.c_void_p, C.c_int, dp]
    lib.cve_run_mixed.argtypes = [C.c_void_p, C.c_int, dp, C.c_int]
    return lib


def _case(pkg, mesh_dir, name, order, v0, warp):
    model = pkg.Model.make_cube(int(name.split(":")[1]), -10.0, 10.0, order) if name.startswith("cube:") else pkg.Model.open_msh(mesh_dir / name, order)
    if warp and len(warp) == 2:
        model.warp(*warp)
    elif warp:
        model.warp_local(warp[0], warp[1], warp[2], warp[3])  # a curved patch in a straight-sided mesh
    mesh = pkg.Mesh(model, pkg.Config())
    mesh.set_physics(c0=343.0, rho0=1.225, v0=v0, dt=0.05 * mesh.h_min() / (343.0 * (2 * order + 1)))
    b = np.nonzero(mesh.fIsBoundary)[0]
    mesh.fBC[b[::2]] = 1
    rng = np.random.default_rng(2)
    x = mesh.node_coords
    u = np.zeros((4, mesh.N))
    for q in range(4 if mesh.dim == 3 else 3):
        k, ph = rng.uniform(0.5, 2, 3), rng.uniform(0, 6, 3)
        u[q] = np.cos(k[0] * x[:, 0] * 0.3 + ph[0]) * np.cos(k[1] * x[:, 1] * 0.3 + ph[1]) * np.cos(k[2] * x[:, 2] * 0.3 + ph[2])
    u[1:] *= 1e-3
    return mesh, u


CASES = [("square.msh", 2, (8.0, -3.0, 0.0), (0.15, 0.9)), ("cube:3", 2, (8.0, -3.0, 2.0), (0.3, 0.4)), ("cube:2", 4, (0.0, 0.0, 0.0), (0.4, 0.3)),
         ("cube:2", 3, (1.0, 2.0, 3.0), (0.3, 0.5))]


@pytest.mark.parametrize("name,order,v0,warp", CASES)
def test_emulated_curved_kernel_equals_the_faithful_oracle(pkg, oracle_mod, cve, mesh_dir, name, order, v0, warp):
    mesh, u = _case(pkg, mesh_dir, name, order, v0, warp)
    d = C.cast(mesh.desc_p, C.c_void_p)
    assert cve.cve_is_curved(d) == 1
    orc = oracle_mod.Oracle(mesh)
    ref = orc.eval_rhs(oracle_mod.Oracle.FAITHFUL, u)
    got = u.copy()
    assert cve.cve_run(d, 2, got.ctypes.data_as(dp), 0) == 0, cve.cve_last_error()
    for q in range(4):
        if np.abs(ref[q]).max() > 0:
            assert rel_l2(got[q], ref[q]) < 1e-12
        else:
            assert np.abs(got[q]).max() == 0.0
    for integrator, ident in ((1, pkg.RUNGE_KUTTA), (0, pkg.EULER1)):
        got = u.copy()
        assert cve.cve_run(d, integrator, got.ctypes.data_as(dp), 2) == 0
        want = u.copy()
        oracle_mod.Oracle(mesh).run(oracle_mod.Oracle.FAITHFUL, ident, want, 0.0, 2)
        for q in range(4):
            if np.abs(want[q]).max() > 0:
                assert rel_l2(got[q], want[q]) < 1e-12


MIXED = [("cube:4", 3, (8.0, -3.0, 2.0), (0.4, 0.3, (0.0, 0.0, 0.0), 6.0)), ("square.msh", 3, (0.0, 0.0, 0.0), (0.1, 0.9, (1.0, 1.0, 0.0), 1.5))]


@pytest.mark.parametrize("name,order,v0,warp", MIXED)
def test_curved_patch_in_a_straight_mesh(pkg, oracle_mod, cve, mesh_dir, name, order, v0, warp):
    """The front end numbers the straight-sided elements first; dgb_create's classification (curved_setup.h) finds exactly that
    suffix; the curved kernel evaluated on the suffix alone agrees with the faithful oracle there, and on the straight-sided
    prefix the collapsed operator form agrees with the faithful oracle (so the two kernels of a mixed handle join up)."""
    mesh, u = _case(pkg, mesh_dir, name, order, v0, warp)
    d = C.cast(mesh.desc_p, C.c_void_p)
    det = mesh.elJacobianDet
    varies = np.abs(det.max(axis=1) - det.min(axis=1)) > 1e-10 * np.abs(det).mean()
    first = cve.cve_first_curved(d)
    assert 0 < first < mesh.K and varies[first:].all() and not varies[:first].any()
    orc = oracle_mod.Oracle(mesh)
    ref = orc.eval_rhs(oracle_mod.Oracle.FAITHFUL, u).reshape(4, mesh.K, mesh.Np)
    got = u.copy()
    assert cve.cve_rhs_suffix(d, first, got.ctypes.data_as(dp)) == 0, cve.cve_last_error()
    got = got.reshape(4, mesh.K, mesh.Np)
    for q in range(4 if mesh.dim == 3 else 3):
        assert rel_l2(got[q, first:], ref[q, first:]) < 1e-12
        assert np.array_equal(got[q, :first], u.reshape(4, mesh.K, mesh.Np)[q, :first])  # untouched
    # the collapsed operators are exact on the straight-sided elements, also next to curved neighbours (shared faces are flat):
    # checked with the host build of the Bernstein operator code, which evaluates every element from its first-point geometry
    if mesh.dim == 3:
        bbc = C.CDLL(str(ROOT / "oracle" / "libbbcheck.so"))
        bbc.bbc_eval_rhs.argtypes = [C.c_void_p, dp, dp, dp, C.POINTER(C.c_int32)]
        op = np.zeros_like(u)
        assert bbc.bbc_eval_rhs(d, u.ctypes.data_as(dp), op.ctypes.data_as(dp), None, None) == 0
        op = op.reshape(4, mesh.K, mesh.Np)
        for q in range(4):
            assert rel_l2(op[q, :first], ref[q, :first]) < 1e-12


@pytest.mark.parametrize("name,order,v0,warp", MIXED)
def test_emulated_mixed_handle_equals_the_faithful_oracle(pkg, oracle_mod, cve, mesh_dir, name, order, v0, warp):
    """Time marching of a mixed handle the way dgb_run drives it: two launches per stage on disjoint element ranges (the collapsed
    generic kernel on the straight-sided prefix, the curved kernel on the suffix) sharing the RK registers; RK4 and Euler."""
    mesh, u = _case(pkg, mesh_dir, name, order, v0, warp)
    d = C.cast(mesh.desc_p, C.c_void_p)
    for integrator, ident in ((1, pkg.RUNGE_KUTTA), (0, pkg.EULER1)):
        got = u.copy()
        first = cve.cve_run_mixed(d, integrator, got.ctypes.data_as(dp), 3)
        assert 0 < first < mesh.K, cve.cve_last_error()
        want = u.copy()
        oracle_mod.Oracle(mesh).run(oracle_mod.Oracle.FAITHFUL, ident, want, 0.0, 3)
        for q in range(4 if mesh.dim == 3 else 3):
            assert rel_l2(got[q], want[q]) < 1e-12


def test_straight_sided_meshes_are_not_flagged(pkg, cve, mesh_dir):
    mesh, _ = _case(pkg, mesh_dir, "cube:2", 3, (0.0, 0.0, 0.0), None)
    assert cve.cve_is_curved(C.cast(mesh.desc_p, C.c_void_p)) == 0
    assert mesh.desc.nGeomEl == 1 and mesh.desc.nGeomF == 1


@pytest.mark.gpu
@pytest.mark.parametrize("name,order,v0,warp", CASES + [("disk.msh", 3, (0.0, 0.0, 0.0), (0.05, 1.5))] + MIXED +
                         [("cube:6", 4, (0.0, 0.0, 0.0), (0.4, 0.3, (0.0, 0.0, 0.0), 6.0))])
def test_engine_on_curved_meshes(pkg, oracle_mod, mesh_dir, name, order, v0, warp):
    mesh, u = _case(pkg, mesh_dir, name, order, v0, warp)
    eng = pkg.Engine(mesh)
    assert eng.kernel_name.endswith("stage_curved") and (len(warp) == 2) == (eng.kernel_name == "stage_curved")
    orc = oracle_mod.Oracle(mesh)
    rhs = eng.eval_rhs(u)
    ref = orc.eval_rhs(oracle_mod.Oracle.FAITHFUL, u)
    for q in range(4):
        if np.abs(ref[q]).max() > 0:
            assert rel_l2(rhs[q], ref[q]) < 1e-10
    eng.set_state(u)
    eng.run(pkg.RUNGE_KUTTA, 0.0, 6)
    got = eng.get_state()
    want = u.copy()
    orc.run(oracle_mod.Oracle.FAITHFUL, pkg.RUNGE_KUTTA, want, 0.0, 6)
    for q in range(4):
        if np.abs(want[q]).max() > 0:
            assert rel_l2(got[q], want[q]) < 1e-10
    with pytest.raises(pkg.DgbError):
        eng.set_option("kernel", 4)  # no Bernstein representation next to curved elements
    eng.close()


def test_a_high_order_file_of_a_curved_mesh_is_recognised(pkg, mesh_dir, tmp_path):
    """What `gmsh -order p` writes for a curved geometry: high-order elements whose extra nodes are off the straight-sided
    positions. Written here from a warped model, read back: same geometry, curved layout; a straight-sided high-order file stays
    in the compressed affine layout."""
    model = pkg.Model.make_cube(2, -10.0, 10.0, 3).warp(0.4, 0.3)
    a = pkg.Mesh(model, pkg.Config())
    path = tmp_path / "curved_p3.msh"
    model.write_msh(path)
    b = pkg.Mesh(pkg.Model.open_msh(path, 1), pkg.Config())  # order <= 1: keep the file's order
    assert (b.order, b.K) == (3, a.K) and b.desc.nGeomEl == b.desc.nG and b.desc.nGeomF == b.desc.nGf
    assert np.array_equal(a.node_coords, b.node_coords) and np.array_equal(a.elJacobianDet, b.elJacobianDet)
    assert np.array_equal(a.fNormal, b.fNormal)
    straight = tmp_path / "straight_p3.msh"
    pkg.Model.make_cube(2, -10.0, 10.0, 3).write_msh(straight)
    c = pkg.Mesh(pkg.Model.open_msh(straight, 1), pkg.Config())
    assert c.order == 3 and c.desc.nGeomEl == 1 and c.desc.nGeomF == 1
