"""The Bernstein-Bezier CUDA kernels (dgfem-acoustic_b200/csrc/stage_bb.cu — the file itself, both schedules, plus the
source-update and conversion kernels) executed on the CPU through a small CUDA emulation (oracle/cuda_emu.h: one OS
thread per CUDA thread, a barrier for __syncthreads) on the engine's device layout, against the oracle: L(u), RK4 with a
hard source, forward Euler; full and partial tiles; orders 2..5; both boundary conditions; mean flow. Tolerance 1e-12.

This checks the kernels' index logic and arithmetic without a GPU. What it cannot check (cp.async, launch configuration,
occupancy, timing) is what the gated GPU tests of tests/test_zz_bb_gpu.py are for."""
import ctypes as C

import numpy as np
import pytest

from conftest import ROOT, rel_l2

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int32)


@pytest.fixture(scope="module")
def bbe():
    lib = C.CDLL(str(ROOT / "oracle" / "libbbemu.so"))
    lib.bbe_last_error.restype = C.c_char_p
    lib.bbe_eval_rhs.argtypes = [C.c_void_p, C.c_int, dp, dp]
    lib.bbe_run.argtypes = [C.c_void_p, C.c_int, C.c_int, dp, C.c_double, C.c_int, C.c_int, ip, C.c_double, C.c_double, C.c_double, C.c_double]
    return lib


def _case(pkg, mesh_dir, name, order, v0):
    model = pkg.Model.make_cube(int(name.split(":")[1]), -10.0, 10.0, order) if name.startswith("cube:") else pkg.Model.open_msh(mesh_dir / name, order)
    cfg = pkg.Config()
    cfg.add_source(2.0, 1.0, 0.0, 6.0, 10.0, 1500.0, 0.3, 1.0)
    mesh = pkg.Mesh(model, cfg)
    mesh.set_physics(c0=343.0, rho0=1.225, v0=v0, dt=0.1 * mesh.h_min() / (343.0 * (2 * order + 1)))
    b = np.nonzero(mesh.fIsBoundary)[0]
    mesh.fBC[b[::2]] = 1
    rng = np.random.default_rng(1)
    x = mesh.node_coords
    u = np.zeros((4, mesh.N))
    for q in range(4):
        k, ph = rng.uniform(0.5, 2, 3), rng.uniform(0, 6, 3)
        u[q] = np.cos(k[0] * x[:, 0] * 0.3 + ph[0]) * np.cos(k[1] * x[:, 1] * 0.3 + ph[1]) * np.cos(k[2] * x[:, 2] * 0.3 + ph[2])
    u[1:] *= 1e-3
    return mesh, u


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("name,order,v0,tile", [("cube:3", 4, (0.0, 0.0, 0.0), 32), ("cube:3", 3, (30.0, 10.0, -5.0), 16), ("cube:4", 2, (0.0, 0.0, 0.0), 8),
                                                ("cube:2", 5, (1.0, 2.0, 3.0), 32), ("cube.msh", 3, (0.0, 0.0, 0.0), 32), ("cube:3", 4, (3.0, 2.0, 1.0), 8)])
def test_emulated_kernels_equal_the_oracle(pkg, oracle_mod, bbe, mesh_dir, name, order, v0, tile, variant):
    mesh, u = _case(pkg, mesh_dir, name, order, v0)
    bbe.bbe_set_tile(tile)  # elements per CTA: 32, 16 or 8 (dgb_set_option("bb_tile", ...))
    d = C.cast(mesh.desc_p, C.c_void_p)
    orc = oracle_mod.Oracle(mesh)
    orc.set_sources_from_config()
    rhs = np.zeros_like(u)
    assert bbe.bbe_eval_rhs(d, variant, u.ctypes.data_as(dp), rhs.ctypes.data_as(dp)) == 0, bbe.bbe_last_error()
    ref = orc.eval_rhs(oracle_mod.Oracle.OPERATOR, u)
    for q in range(4):
        assert rel_l2(rhs[q], ref[q]) < 1e-12
    if name == "cube.msh":
        return  # 13 603 elements: L(u) is enough for the unstructured mesh
    _, idx = mesh.source_nodes()
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    assert len(idx) > 0
    got = u.copy()
    assert bbe.bbe_run(d, variant, 1, got.ctypes.data_as(dp), 0.0, 4, len(idx), idx.ctypes.data_as(ip), 10.0, 1500.0, 0.3, 1.0) == 0, bbe.bbe_last_error()
    want = u.copy()
    orc.run(oracle_mod.Oracle.OPERATOR, pkg.RUNGE_KUTTA, want, 0.0, 4)
    for q in range(4):
        assert rel_l2(got[q], want[q]) < 1e-12
    got = u.copy()
    assert bbe.bbe_run(d, variant, 0, got.ctypes.data_as(dp), 0.0, 3, 0, None, 0.0, 0.0, 0.0, 0.0) == 0
    want = u.copy()
    oracle_mod.Oracle(mesh).run(oracle_mod.Oracle.OPERATOR, pkg.EULER1, want, 0.0, 3)
    for q in range(4):
        assert rel_l2(got[q], want[q]) < 1e-12
