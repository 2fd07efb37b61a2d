"""The product's Bernstein-Bezier stage kernels (dgfem-acoustic_b200/csrc/stage_bb2.cu and stage_bbe.cu — the files themselves) executed on the CPU
through the CUDA emulation of oracle/cuda_emu.h, against the oracle: L(u), RK4, forward Euler; tetrahedra and triangles; full and partial tiles,
several tiles per persistent warp; both boundary conditions; mean flow. Tolerance 1e-12.

One OS thread per lane of the one-warp CTAs: `__syncwarp` is a real barrier, shuffles are real exchanges, the asynchronous copies (TMA bulk copies
completing on mbarriers, cp.async) are synchronous host stand-ins (csrc/dgb_async.cuh under DGB_EMULATE). This is the CPU-side net for the index
logic of the kernels the product runs by default; asynchrony, proxies and occupancy are what the GPU tests (tests/test_zz_bb_gpu.py) cover."""
import ctypes as C

import numpy as np
import pytest

from conftest import ROOT, rel_l2

dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def emu():
    lib = C.CDLL(str(ROOT / "oracle" / "libbb2emu.so"))
    lib.bb2e_last_error.restype = C.c_char_p
    lib.bb2e_run.argtypes = [C.c_void_p, C.c_int, C.c_int, dp, C.c_int]
    return lib


def _case(pkg, mesh_dir, name, order, v0):
    if name.startswith("cube:"):
        model = pkg.Model.make_cube(int(name.split(":")[1]), -10.0, 10.0, order)
    elif name.startswith("square:"):
        model = pkg.Model.make_square(int(name.split(":")[1]), -10.0, 10.0, order)
    else:
        model = pkg.Model.open_msh(mesh_dir / name, order)
    mesh = pkg.Mesh(model, pkg.Config())
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


def _close(got, want, u):
    for q in range(4):
        if np.abs(want[q]).max() == 0:  # v_z on triangles without mean flow
            assert np.linalg.norm(got[q]) < 1e-9 * max(np.linalg.norm(u[q]), 1e-300) * 1e4
        else:
            assert rel_l2(got[q], want[q]) < 1e-12


# kernel 6 = stage_bb2 (thread = (element, field), tiles of 8 elements), 7 = stage_bbe (thread = element, tiles of 32)
@pytest.mark.parametrize("kernel,name,order,v0,steps", [
    (7, "square.msh", 1, (30.0, 10.0, 0.0), 2), (7, "square:5", 2, (0.0, 0.0, 0.0), 2), (7, "cube:3", 1, (30.0, 10.0, -5.0), 2), (7, "disk.msh", 3, (0.0, 0.0, 0.0), 1),
    (6, "cube:2", 4, (0.0, 0.0, 0.0), 1), (6, "cube:2", 3, (30.0, 10.0, -5.0), 1), (6, "cube:2", 2, (0.0, 0.0, 0.0), 1), (6, "square:3", 5, (3.0, -2.0, 0.0), 1),
    (6, "square:4", 3, (0.0, 0.0, 0.0), 1), (6, "cube:1", 5, (1.0, 2.0, 3.0), 1), (6, "cube:3", 4, (30.0, 10.0, -5.0), 1), (6, "cube:3", 1, (0.0, 0.0, 0.0), 2),
    (6, "square:4", 4, (30.0, 10.0, 0.0), 1), (6, "square:3", 6, (0.0, 0.0, 0.0), 1), (6, "square:6", 2, (3.0, 2.0, 0.0), 1), (6, "square:5", 1, (0.0, 0.0, 0.0), 2),
    (7, "cube:2", 2, (0.0, 0.0, 0.0), 1), (7, "square_reflection.msh", 2, (30.0, 10.0, 0.0), 1), (6, "cube:1", 6, (3.0, 2.0, 1.0), 1), (6, "cube:2", 6, (0.0, 0.0, 0.0), 1)])
def test_emulated_kernels_equal_the_oracle(pkg, oracle_mod, emu, mesh_dir, kernel, name, order, v0, steps):
    mesh, u = _case(pkg, mesh_dir, name, order, v0)
    d = C.cast(mesh.desc_p, C.c_void_p)
    orc = oracle_mod.Oracle(mesh)
    rhs = u.copy()
    assert emu.bb2e_run(d, kernel, 2, rhs.ctypes.data_as(dp), 0) == 0, emu.bb2e_last_error()
    _close(rhs, orc.eval_rhs(oracle_mod.Oracle.OPERATOR, u), u)
    got = u.copy()
    assert emu.bb2e_run(d, kernel, 1, got.ctypes.data_as(dp), steps) == 0, emu.bb2e_last_error()
    want = u.copy()
    orc.run(oracle_mod.Oracle.OPERATOR, pkg.RUNGE_KUTTA, want, 0.0, steps)
    _close(got, want, u)
    got = u.copy()
    assert emu.bb2e_run(d, kernel, 0, got.ctypes.data_as(dp), 2) == 0, emu.bb2e_last_error()
    want = u.copy()
    oracle_mod.Oracle(mesh).run(oracle_mod.Oracle.OPERATOR, pkg.EULER1, want, 0.0, 2)
    _close(got, want, u)
