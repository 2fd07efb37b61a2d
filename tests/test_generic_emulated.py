"""The generic CUDA stage kernel and the small kernels (dgfem-acoustic_b200/csrc/stage_generic.cu — the file itself), executed
on the CPU through the CUDA emulation of oracle/cuda_emu.h, against the oracle. This kernel is verified on B200 hardware
(tests/test_gpu_parity.py); here it gives the CPU suite a regression net for every (dimension, order) — and shows that the
emulation reproduces a known-good kernel, which is what the emulated tests of the not-yet-run kernels (test_bb_emulated.py,
test_curved.py) rest on."""
import ctypes as C

import numpy as np
import pytest

from conftest import ROOT, rel_l2

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int32)


@pytest.fixture(scope="module")
def gne():
    lib = C.CDLL(str(ROOT / "oracle" / "libgenericemu.so"))
    lib.gne_last_error.restype = C.c_char_p
    lib.gne_run.argtypes = [C.c_void_p, C.c_int, dp, C.c_double, C.c_int, C.c_int, ip, C.c_double, C.c_double, C.c_double, C.c_double,
                            C.c_int, ip, dp, dp]
    return lib


def _case(pkg, mesh_dir, name, order, v0):
    model = pkg.Model.make_cube(int(name.split(":")[1]), -10.0, 10.0, order) if name.startswith("cube:") else pkg.Model.open_msh(mesh_dir / name, order)
    cfg = pkg.Config()
    cfg.add_source(2.0, 1.0, 0.0, 3.0, 10.0, 1500.0, 0.3, 1.0)
    mesh = pkg.Mesh(model, cfg)
    mesh.set_physics(c0=343.0, rho0=1.225, v0=v0, dt=0.1 * mesh.h_min() / (343.0 * (2 * order + 1)))
    b = np.nonzero(mesh.fIsBoundary)[0]
    mesh.fBC[b[::2]] = 1
    rng = np.random.default_rng(4)
    x = mesh.node_coords
    u = np.zeros((4, mesh.N))
    for q in range(1 + mesh.dim):
        k, ph = rng.uniform(0.5, 2, 3), rng.uniform(0, 6, 3)
        u[q] = np.cos(k[0] * x[:, 0] * 0.3 + ph[0]) * np.cos(k[1] * x[:, 1] * 0.3 + ph[1]) * np.cos(k[2] * x[:, 2] * 0.3 + ph[2])
    u[1:] *= 1e-3
    return mesh, u


CASES = [("line.msh", 1, (0.0, 0.0, 0.0)), ("square.msh", 1, (0.0, 0.0, 0.0)), ("square.msh", 2, (20.0, 5.0, 0.0)), ("square.msh", 4, (0.0, 0.0, 0.0)),
         ("cube:2", 1, (3.0, 2.0, 1.0)), ("cube:3", 2, (0.0, 0.0, 0.0)), ("cube:2", 3, (30.0, 10.0, 5.0)), ("cube:2", 4, (0.0, 0.0, 0.0)),
         ("cube:1", 5, (1.0, 2.0, 3.0)), ("cube:1", 6, (0.0, 0.0, 0.0))]


@pytest.mark.parametrize("name,order,v0", CASES)
def test_emulated_generic_kernel_equals_the_oracle(pkg, oracle_mod, gne, mesh_dir, name, order, v0):
    mesh, u = _case(pkg, mesh_dir, name, order, v0)
    d = C.cast(mesh.desc_p, C.c_void_p)
    orc = oracle_mod.Oracle(mesh)
    orc.set_sources_from_config()
    got = u.copy()
    assert gne.gne_run(d, 2, got.ctypes.data_as(dp), 0.0, 0, 0, None, 0.0, 0.0, 0.0, 0.0, 0, None, None, None) == 0, gne.gne_last_error()
    ref = orc.eval_rhs(oracle_mod.Oracle.OPERATOR, u)
    for q in range(1 + mesh.dim):
        assert rel_l2(got[q], ref[q]) < 1e-12
    # RK4 with the hard source and two receivers
    _, idx = mesh.source_nodes()
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    x = mesh.node_coords.reshape(mesh.K, mesh.Np, 3)
    pts = [tuple(x[mesh.K // 2, : mesh.dim + 1].mean(axis=0)), tuple(x[mesh.K // 3, 0])]
    el, w = mesh.locate_receivers(pts)
    w = np.ascontiguousarray(w)
    steps = 3
    rec = np.zeros((steps, 2, 4))
    got = u.copy()
    assert gne.gne_run(d, 1, got.ctypes.data_as(dp), 0.0, steps, len(idx), idx.ctypes.data_as(ip) if len(idx) else None, 10.0, 1500.0, 0.3, 1.0,
                       2, el.ctypes.data_as(ip), w.ctypes.data_as(dp), rec.ctypes.data_as(dp)) == 0, gne.gne_last_error()
    orc.set_receivers(el, w)
    want = u.copy()
    orc.run(oracle_mod.Oracle.OPERATOR, pkg.RUNGE_KUTTA, want, 0.0, steps)
    ref_rec = orc.get_receivers(steps)
    for q in range(1 + mesh.dim):
        assert rel_l2(got[q], want[q]) < 1e-12
        assert rel_l2(rec[:, :, q], ref_rec[:, :, q]) < 1e-12
    got = u.copy()
    assert gne.gne_run(d, 0, got.ctypes.data_as(dp), 0.0, 2, 0, None, 0.0, 0.0, 0.0, 0.0, 0, None, None, None) == 0
    want = u.copy()
    oracle_mod.Oracle(mesh).run(oracle_mod.Oracle.OPERATOR, pkg.EULER1, want, 0.0, 2)
    for q in range(1 + mesh.dim):
        assert rel_l2(got[q], want[q]) < 1e-12
