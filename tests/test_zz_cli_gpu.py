"""The product CLI `dgalerkin mesh.msh config.conf` (dgfem-acoustic_b200/lib/dgalerkin, the reference's command line,
src/dgalerkin.cpp:11-63) end to end on the GPU: the views it appends to saveFile and the receiver file against the oracle.

First run on hardware in round 2 (profiles/r02/cli_tests.log)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, rel_l2

pytestmark = [pytest.mark.gpu]

CONF = """timeStart=0
timeEnd=0.00205
timeStep=0.00005
timeRate=0.0005
elementType=Lagrange
timeIntMethod=Runge-Kutta
Absorbing = Absorbing
numThreads=4
v0_x = 0
v0_y = 0
v0_z = 0
rho0 = 1.225
c0 = 343.3
initialCondtition1 = gaussian, 0,0,0,1,1
source1 = monopole, 1.0,0.5,0, 0.5, 2.0,1000,0,0.001
receiverA = 0.37, -0.21, 0
receiverB = -2.2, 1.3, 0
receiverFile = rcv.txt
saveFile=data.msh
"""


def _blocks(path):
    text = path.read_text().split("\n")
    out, i = [], 0
    while i < len(text):
        if text[i] == "$ElementNodeData":
            name, t = text[i + 2].strip('"'), float(text[i + 4])
            step, ncomp, nel = int(text[i + 6]), int(text[i + 7]), int(text[i + 8])
            vals = np.array([[float(x) for x in text[i + 9 + k].split()[2:]] for k in range(nel)])
            out.append((name, t, step, ncomp, vals))
            i += 10 + nel
        else:
            i += 1
    return out


def test_cli_views_and_receivers(pkg, oracle_mod, mesh_dir, tmp_path):
    conf = tmp_path / "case.conf"
    conf.write_text(CONF)
    cli = ROOT / "dgfem-acoustic_b200" / "lib" / "dgalerkin"
    env = dict(os.environ, DGB_ORDER="2")
    subprocess.run([str(cli), str(mesh_dir / "square.msh"), str(conf)], cwd=tmp_path, env=env, check=True, timeout=600)
    model = pkg.Model.open_msh(mesh_dir / "square.msh", 2)
    cfg = model.parse_config(conf)
    mesh = pkg.Mesh(model, cfg)
    nsteps, snaps = cfg.time_loop()
    el, w = mesh.locate_receivers(cfg.receivers)
    orc = oracle_mod.Oracle(mesh)
    orc.set_sources_from_config()
    orc.set_receivers(el, w)
    u = mesh.initial_condition()
    blocks = _blocks(tmp_path / "data.msh")
    pressure = [b for b in blocks if b[0] == "Pressure"]
    velocity = [b for b in blocks if b[0] == "Velocity"]
    assert [b[2] for b in pressure] == list(snaps) and len(velocity) == len(pressure)
    t, done = cfg.c.timeStart, 0
    for (_, tt, step, _, p), (_, _, _, _, v) in zip(pressure, velocity):
        t, _ = orc.run(oracle_mod.Oracle.OPERATOR, pkg.RUNGE_KUTTA, u, t, step - done)
        done = step
        assert abs(t - tt) <= 1e-15 * max(1.0, abs(tt))  # %.16g in the file
        U = u.reshape(4, mesh.K, mesh.Np)
        assert rel_l2(p, U[0]) < 1e-10
        v = v.reshape(mesh.K, mesh.Np, 3)
        assert rel_l2(v[:, :, 0], U[1]) < 1e-10 and rel_l2(v[:, :, 1], U[2]) < 1e-10
    orc.run(oracle_mod.Oracle.OPERATOR, pkg.RUNGE_KUTTA, u, t, nsteps - done)
    ref = orc.get_receivers(nsteps)
    got = np.loadtxt(tmp_path / "rcv.txt")
    assert got.shape == (nsteps, 1 + 2 * 4)
    for j in range(2):
        for q in range(3):
            assert rel_l2(got[:, 1 + 4 * j + q], ref[:, j, q]) < 1e-10


def test_async_snapshot_equals_get_state(pkg, mesh_dir):
    """dgb_snapshot_begin / _end deliver the state of the moment of `begin`, while the next run is already under way."""
    import torch
    model = pkg.Model.make_cube(4, -10.0, 10.0, 4)
    cfg = pkg.Config()
    cfg.add_initial_condition(0.0, 0.0, 0.0, 20.0, 1.0)
    mesh = pkg.Mesh(model, cfg)
    mesh.set_physics(c0=343.0, rho0=1.225, v0=(0.0, 0.0, 0.0), dt=0.1 * mesh.h_min() / (343.0 * 9))
    u0 = mesh.initial_condition()
    for kernel in (0, 4):
        eng = pkg.Engine(mesh, options={"kernel": kernel})
        eng.set_state(u0)
        t = eng.run(pkg.RUNGE_KUTTA, 0.0, 5)
        want = eng.get_state().copy()
        snap = torch.empty((4, mesh.N), dtype=torch.float64, pin_memory=True).numpy()
        eng.snapshot_begin(snap)
        t = eng.run(pkg.RUNGE_KUTTA, t, 5)  # overlaps the copy
        eng.snapshot_end()
        later = eng.get_state()
        if kernel == 0:
            assert np.array_equal(snap, want)
        else:  # Bernstein mode converts on the device: same conversion kernel, same result
            assert np.array_equal(snap, want)
        assert not np.array_equal(later, want)
        with pytest.raises(pkg.DgbError):
            eng.snapshot_end()  # nothing in flight
        eng.close()
