"""Generates tests/golden/cases/*.npz by RUNNING THE REFERENCE ITSELF.

The reference's own translation units (/root/reference/src/*.cpp) are compiled unmodified into
oracle/_ref/dgalerkin_ref by oracle/Makefile (on top of the Gmsh/Eigen stand-ins, because the Gmsh SDK and Eigen
cannot be installed here). This script runs that binary on each case below, reads the views it "writes"
(oracle/shim/gmsh_shim.cpp: <saveFile>.<View>.bin) and stores selected snapshots. Run it in the build container:

    python tests/golden/make_golden.py

The GPU box has no /root/reference; it only consumes the committed .npz files.
"""
import gzip
import os
import shutil
import struct
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as graft  # noqa: E402

REF_BIN = ROOT / "oracle" / "_ref" / "dgalerkin_ref"

COMMON = """elementType=Lagrange
numThreads=2
rho0 = 1.225
saveFile=out
"""

# name: (mesh, order, config body, snapshots to keep)
CASES = {
    "line_p1_euler": ("line.msh", 1, """timeStart=0
timeEnd=0.00081
timeStep=0.00002
timeRate=0.000001
timeIntMethod=Euler1
Boundary = Absorbing
v0_x = 10
v0_y = 0
v0_z = 0
c0 = 343
initialCondtition1 = gaussian, 0,0,0,1,1
""", [10, 40]),
    "square_p1_rk4_config1": ("square.msh", 1, """timeStart=0
timeEnd=0.00301
timeStep=0.0001
timeRate=0.000001
timeIntMethod=Runge-Kutta
Absorbing = Absorbing
v0_x = 0
v0_y = 0
v0_z = 0
c0 = 343.3
initialCondtition1 = gaussian, 0,0,0,1,1
""", [30]),
    "square_p3_rk4_reflecting_flow": ("square.msh", 3, """timeStart=0
timeEnd=0.00021
timeStep=0.00002
timeRate=0.000001
timeIntMethod=Runge-Kutta
Absorbing = Reflecting
v0_x = 20
v0_y = 5
v0_z = 0
c0 = 343
initialCondtition1 = gaussian, 1,-1,0,2,1
""", [10]),
    "cube2_p3_rk4_flow_source": ("gen:cube2", 3, """timeStart=0
timeEnd=0.000101
timeStep=0.00001
timeRate=0.000001
timeIntMethod=Runge-Kutta
v0_x = 30
v0_y = 10
v0_z = 5
c0 = 343
initialCondtition1 = gaussian, 0,0,0,40,1
source1 = monopole, 2,1,0, 6, 10,1500,0,0.00005
""", [10]),
    "cube2_p4_rk4_reflecting": ("gen:cube2", 4, """timeStart=0
timeEnd=0.000051
timeStep=0.00001
timeRate=0.000001
timeIntMethod=Runge-Kutta
Boundary = Reflecting
v0_x = 0
v0_y = 0
v0_z = 0
c0 = 343
initialCondtition1 = gaussian, 3,-2,1,40,1
""", [5]),
    "cube2_p2_euler_absorbing": ("gen:cube2", 2, """timeStart=0
timeEnd=0.000081
timeStep=0.000004
timeRate=0.000001
timeIntMethod=Euler1
Boundary = Absorbing
v0_x = 5
v0_y = 0
v0_z = -5
c0 = 343
initialCondtition1 = gaussian, 0,0,0,40,1
""", [20]),
}


def read_view(path):
    b = Path(path).read_bytes()
    ns, ne, per = struct.unpack_from("<iii", b, 0)
    off = 12
    out = []
    for _ in range(ns):
        st, t = struct.unpack_from("<id", b, off)
        off += 12
        out.append((st, t, np.frombuffer(b, dtype=np.float64, count=ne * per, offset=off).reshape(ne, per).copy()))
        off += 8 * ne * per
    return out


def materialise_mesh(mesh, workdir):
    """Returns the path of an order-1 MSH 4.0 file for `mesh` (a shipped fixture or a generated cube)."""
    if mesh.startswith("gen:cube"):
        pkg = graft.load_package()
        path = Path(workdir) / (mesh[4:] + ".msh")
        pkg.Model.make_cube(int(mesh[8:]), -10.0, 10.0, 1).write_msh(path)
        return path
    path = Path(workdir) / mesh
    with gzip.open(HERE / "meshes" / (mesh + ".gz"), "rb") as src, open(path, "wb") as dst:
        shutil.copyfileobj(src, dst)
    return path


def run_reference(mesh_path, order, conf_text, workdir):
    conf = Path(workdir) / "case.conf"
    conf.write_text(conf_text)
    env = dict(os.environ, GMSHLITE_QUIET="1", GMSHLITE_ORDER=str(order), OMP_NUM_THREADS="2")
    subprocess.run([str(REF_BIN), str(mesh_path), str(conf)], cwd=workdir, env=env, check=True)
    P = read_view(Path(workdir) / "out.Pressure.bin")
    V = read_view(Path(workdir) / "out.Velocity.bin")
    return P, V


def main():
    if not REF_BIN.exists():
        raise SystemExit(f"{REF_BIN} missing: run `make -C oracle ref` where /root/reference exists")
    for name, (mesh, order, body, keep) in CASES.items():
        with tempfile.TemporaryDirectory() as wd:
            conf_text = body + COMMON
            P, V = run_reference(materialise_mesh(mesh, wd), order, conf_text, wd)
            steps = {p[0]: i for i, p in enumerate(P)}
            data = {"mesh": mesh, "order": order, "config": conf_text, "steps": np.array(keep)}
            for s in keep:
                i = steps[s]
                ne, npn = P[i][2].shape
                v = V[i][2].reshape(ne, npn, 3)
                data[f"t_{s}"] = P[i][1]
                data[f"u_{s}"] = np.stack([P[i][2].reshape(-1), v[:, :, 0].reshape(-1), v[:, :, 1].reshape(-1), v[:, :, 2].reshape(-1)])
            np.savez_compressed(HERE / "cases" / f"{name}.npz", **data)
            print(name, "snapshots", keep, "K", ne, "Np", npn, "max|p|", float(np.abs(data[f"u_{keep[-1]}"][0]).max()))


if __name__ == "__main__":
    main()
