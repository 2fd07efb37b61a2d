"""Live comparison of the oracle with the reference's own sources (oracle/_ref/dgalerkin_ref, built by
oracle/Makefile where /root/reference exists). Skipped where the binary is absent."""
import os
import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import ROOT, rel_l2

REF_BIN = ROOT / "oracle" / "_ref" / "dgalerkin_ref"
pytestmark = pytest.mark.skipif(not REF_BIN.exists(), reason="oracle/_ref/dgalerkin_ref not built")

CONF = """timeStart=0
timeEnd=0.000405
timeStep=0.00002
timeRate=0.0001
elementType=Lagrange
timeIntMethod=Runge-Kutta
Reflecting = Reflecting
numThreads=2
v0_x = 8
v0_y = -3
v0_z = 0
rho0 = 1.225
c0 = 343
source1 = quadrupole, 0.5,0.2,0, 0.6, 10,1500,0,0.0002
initialCondtition1 = gaussian, -1,1,0,1,0.5
saveFile=out
"""


def read_view(path):
    b = Path(path).read_bytes()
    ns, ne, per = struct.unpack_from("<iii", b, 0)
    off, out = 12, []
    for _ in range(ns):
        st, t = struct.unpack_from("<id", b, off)
        off += 12
        out.append((st, t, np.frombuffer(b, dtype=np.float64, count=ne * per, offset=off).reshape(ne, per)))
        off += 8 * ne * per
    return out


def test_disk_p2_quadrupole_reflecting(pkg, oracle_mod, mesh_dir, tmp_path):
    conf = tmp_path / "case.conf"
    conf.write_text(CONF)
    env = dict(os.environ, GMSHLITE_QUIET="1", GMSHLITE_ORDER="2", OMP_NUM_THREADS="2")
    subprocess.run([str(REF_BIN), str(mesh_dir / "disk.msh"), str(conf)], cwd=tmp_path, env=env, check=True, timeout=600)
    P, V = read_view(tmp_path / "out.Pressure.bin"), read_view(tmp_path / "out.Velocity.bin")
    model = pkg.Model.open_msh(mesh_dir / "disk.msh", 2)
    cfg = model.parse_config(conf)
    assert cfg.c.nSources == 4  # quadrupole -> 4 monopoles (configParser.cpp:86-95)
    mesh = pkg.Mesh(model, cfg)
    assert (mesh.fBC[mesh.fIsBoundary == 1] == 1).all()
    steps, snaps = cfg.time_loop()
    assert [p[0] for p in P] == list(snaps)
    for mode in (0, 1):
        orc = oracle_mod.Oracle(mesh, threads=2)
        orc.set_sources_from_config()
        u = mesh.initial_condition()
        t, done = cfg.c.timeStart, 0
        for (st, tt, p), (_, _, v) in zip(P, V):
            t, _ = orc.run(mode, pkg.RUNGE_KUTTA, u, t, st - done)
            done = st
            assert t == tt
            v = v.reshape(mesh.K, mesh.Np, 3)
            ref = [p.reshape(-1), v[:, :, 0].reshape(-1), v[:, :, 1].reshape(-1)]
            for q in range(3):
                assert rel_l2(u[q], ref[q]) < 1e-12, (mode, st, q)
            assert np.abs(u[3]).max() == 0.0 and np.abs(v[:, :, 2]).max() == 0.0


CONF_CURVED = """timeStart=0
timeEnd=0.00021
timeStep=0.00002
timeRate=0.0001
elementType=Lagrange
timeIntMethod=Runge-Kutta
Absorbing = Reflecting
numThreads=2
v0_x = 8
v0_y = -3
v0_z = 0
rho0 = 1.225
c0 = 343
initialCondtition1 = gaussian, -1,1,0,1,0.5
saveFile=out
"""


@pytest.mark.parametrize("mesh_name,order,warp", [("square.msh", 3, (0.15, 0.9)), ("gen:cube4", 2, (0.3, 0.4)), ("gen:cube2", 3, (0.4, 0.3))])
def test_curved_elements_faithful_oracle_equals_the_reference(pkg, oracle_mod, mesh_dir, tmp_path, mesh_name, order, warp):
    """SURVEY §8 f3 groundwork: on a curved (warped isoparametric) mesh the reference keeps one Jacobian / normal per
    integration point; the front end produces that layout (nGeomEl = nG, nGeomF = nGf) and the oracle's faithful mode —
    the reference's own loops — reproduces the reference binary. (The collapsed operator form does not apply: the
    quadrature is no longer exact and the element mass matrices differ.)"""
    conf = tmp_path / "case.conf"
    text = CONF_CURVED if mesh_name == "square.msh" else CONF_CURVED.replace("v0_z = 0", "v0_z = 2")
    conf.write_text(text)
    if mesh_name.startswith("gen:cube"):  # a small generated cube, written as an order-1 MSH file for the reference
        msh = tmp_path / "cube.msh"
        pkg.Model.make_cube(int(mesh_name[8:]), -10.0, 10.0, 1).write_msh(msh)
    else:
        msh = mesh_dir / mesh_name
    env = dict(os.environ, GMSHLITE_QUIET="1", GMSHLITE_ORDER=str(order), GMSHLITE_WARP=f"{warp[0]},{warp[1]}", OMP_NUM_THREADS="2")
    subprocess.run([str(REF_BIN), str(msh), str(conf)], cwd=tmp_path, env=env, check=True, timeout=900)
    P, V = read_view(tmp_path / "out.Pressure.bin"), read_view(tmp_path / "out.Velocity.bin")
    model = pkg.Model.open_msh(msh, order).warp(*warp)
    cfg = model.parse_config(conf)
    mesh = pkg.Mesh(model, cfg)
    assert mesh.desc.nGeomEl == mesh.desc.nG and mesh.desc.nGeomF == mesh.desc.nGf
    det = mesh.elJacobianDet
    assert (np.abs(det.max(axis=1) - det.min(axis=1)) > 1e-6 * np.abs(det).mean()).mean() > 0.9  # the elements really are curved
    orc = oracle_mod.Oracle(mesh, threads=2)
    u = mesh.initial_condition()
    t, done = cfg.c.timeStart, 0
    for (st, tt, p), (_, _, v) in zip(P, V):
        t, _ = orc.run(oracle_mod.Oracle.FAITHFUL, pkg.RUNGE_KUTTA, u, t, st - done)
        done = st
        assert t == tt
        v = v.reshape(mesh.K, mesh.Np, 3)
        ref = [p.reshape(-1), v[:, :, 0].reshape(-1), v[:, :, 1].reshape(-1), v[:, :, 2].reshape(-1)]
        for q in range(4):
            if np.abs(ref[q]).max() == 0.0:
                assert np.abs(u[q]).max() == 0.0
            else:
                assert rel_l2(u[q], ref[q]) < 1e-12, (st, q)
