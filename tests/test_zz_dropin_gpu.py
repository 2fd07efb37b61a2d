"""The literal drop-in: the reference's OWN main, config parser and Gmsh-based Mesh constructor,
unmodified, with src/solver.cpp replaced by integration/solver_dgb.cpp (the binding to the C ABI) and linked against libdgb.so
— oracle/_ref/dgalerkin_dgb — against the reference's own build of the same sources — oracle/_ref/dgalerkin_ref. Same mesh,
same config, same step count: the views both write must agree to 1e-10 relative L2 per field at every snapshot
(BASELINE.json north_star tolerance). Both binaries are built by oracle/Makefile where /root/reference exists and travel
prebuilt to the GPU box. The 3D cases are north_star configs 3 and 4 for their first snapshots."""
import os
import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import ROOT, rel_l2

REF_BIN = ROOT / "oracle" / "_ref" / "dgalerkin_ref"
DGB_BIN = ROOT / "oracle" / "_ref" / "dgalerkin_dgb"
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not (REF_BIN.exists() and DGB_BIN.exists()), reason="oracle/_ref binaries not built")]
TOL = 1e-10

CONF = """timeStart=0
timeEnd={tend}
timeStep={dt}
timeRate={rate}
elementType=Lagrange
timeIntMethod={method}
{bc}
numThreads=1
v0_x = {v0[0]}
v0_y = {v0[1]}
v0_z = {v0[2]}
rho0 = 1.225
c0 = 343
{source}
initialCondtition1 = gaussian, {ic}
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


CASES = {
    # 2D, reflecting walls, mean flow, quadrupole source (4 monopoles), CUDA-core kernel
    "disk_p2": dict(mesh="disk.msh", order=2, dt=2e-5, steps=20, rate=1e-4, method="Runge-Kutta", bc="Reflecting = Reflecting", v0=(8, -3, 0),
                    source="source1 = quadrupole, 0.5,0.2,0, 0.6, 10,1500,0,0.0002", ic="-1,1,0,1,0.5"),
    # forward Euler in 1D
    "line_p1_euler": dict(mesh="line.msh", order=1, dt=1e-5, steps=30, rate=1e-4, method="Euler1", bc="", v0=(0, 0, 0), source="", ic="0,0,0,1,1"),
    # north_star config 3: doc/3d/cube.msh, order 3, absorbing (default) boundary, linearised Euler with mean flow — Bernstein kernel
    "config3_cube_p3_flow": dict(mesh="cube.msh", order=3, dt=1.5e-5, steps=10, rate=7.5e-5, method="Runge-Kutta", bc="", v0=(30, 10, 5), source="",
                                 ic="0,0,0,1,1"),
    # north_star config 4 (auditorium mesh missing -> doc/3d/sphere.msh): order 4, Amphi_Pulse_3D.conf physics, monopole source
    "config4_sphere_p4_source": dict(mesh="sphere.msh", order=4, dt=1e-5, steps=5, rate=2e-5, method="Runge-Kutta", bc="", v0=(0, 0, 0),
                                     source="source1 = monopole, 0.2,0.1,0, 0.35, 10,2000,0,1", ic="0,0,0,1,0.5"),
}


@pytest.mark.parametrize("name", list(CASES))
def test_reference_front_end_on_the_engine_equals_the_reference(mesh_dir, tmp_path, name):
    c = CASES[name]
    text = CONF.format(tend=repr((c["steps"] - 0.5) * c["dt"]), dt=repr(c["dt"]), rate=repr(c["rate"]), method=c["method"], bc=c["bc"], v0=c["v0"],
                       source=c["source"], ic=c["ic"])
    views = {}
    for tag, exe in (("ref", REF_BIN), ("dgb", DGB_BIN)):
        wd = tmp_path / tag
        wd.mkdir()
        (wd / "case.conf").write_text(text)
        env = dict(os.environ, GMSHLITE_QUIET="1", GMSHLITE_ORDER=str(c["order"]), OMP_NUM_THREADS=str(os.cpu_count() or 1))
        subprocess.run([str(exe), str(mesh_dir / c["mesh"]), "case.conf"], cwd=wd, env=env, check=True, timeout=1500)
        views[tag] = {v: read_view(wd / f"out.{v}.bin") for v in ("Pressure", "Density", "Velocity")}
    nsnap = len(views["ref"]["Pressure"])
    assert nsnap >= 2 and len(views["dgb"]["Pressure"]) == nsnap
    for v in ("Pressure", "Density", "Velocity"):
        for (st_r, t_r, a), (st_d, t_d, b) in zip(views["ref"][v], views["dgb"][v]):
            assert st_r == st_d and t_r == t_d
            if v == "Velocity":
                a, b = a.reshape(a.shape[0], -1, 3), b.reshape(b.shape[0], -1, 3)
                comps = [(a[:, :, x], b[:, :, x]) for x in range(3)]
            else:
                comps = [(a, b)]
            for x, (ra, rb) in enumerate(comps):
                if np.abs(ra).max() == 0.0:
                    assert np.abs(rb).max() == 0.0, (v, x, st_r)
                else:
                    assert rel_l2(rb, ra) < TOL, (v, x, st_r, rel_l2(rb, ra))
    last = views["ref"]["Pressure"][-1]
    assert last[0] > 0 and np.abs(last[2]).max() > 0  # the run went somewhere
