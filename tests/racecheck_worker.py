"""Worker of tests/test_emulated_racecheck.py: runs emulated kernels from a ThreadSanitizer build of an emulation harness
(LD_PRELOAD=libtsan.so python racecheck_worker.py <kind> <library>); TSan reports go to stderr."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as graft  # noqa: E402

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int32)


def mesh_and_state(pkg, cells, order, v0, warp=None):
    model = pkg.Model.make_cube(cells, -10.0, 10.0, order)
    if warp:
        model.warp(*warp)
    cfg = pkg.Config()
    cfg.add_source(2.0, 1.0, 0.0, 6.0, 10.0, 1500.0, 0.3, 1.0)
    mesh = pkg.Mesh(model, cfg)
    mesh.set_physics(c0=343.0, rho0=1.225, v0=v0, dt=1e-5)
    b = np.nonzero(mesh.fIsBoundary)[0]
    mesh.fBC[b[::2]] = 1
    return mesh, np.random.default_rng(0).standard_normal((4, mesh.N))


def main():
    kind, path = sys.argv[1], sys.argv[2]
    pkg = graft.load_package()
    lib = C.CDLL(path)
    if kind == "bb":
        lib.bbe_eval_rhs.argtypes = [C.c_void_p, C.c_int, dp, dp]
        lib.bbe_run.argtypes = [C.c_void_p, C.c_int, C.c_int, dp, C.c_double, C.c_int, C.c_int, ip, C.c_double, C.c_double, C.c_double, C.c_double]
        lib.bbe_set_tile.argtypes = [C.c_int]
        mesh, u = mesh_and_state(pkg, 2, 3, (3.0, 2.0, 1.0))
        d = C.cast(mesh.desc_p, C.c_void_p)
        _, idx = mesh.source_nodes()
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        for tile in (32, 16, 8):
            lib.bbe_set_tile(tile)
            for variant in (0, 1):
                rhs = np.zeros_like(u)
                assert lib.bbe_eval_rhs(d, variant, u.ctypes.data_as(dp), rhs.ctypes.data_as(dp)) == 0
                w = u.copy()
                assert lib.bbe_run(d, variant, 1, w.ctypes.data_as(dp), 0.0, 1, len(idx), idx.ctypes.data_as(ip), 10.0, 1500.0, 0.3, 1.0) == 0
    elif kind == "bb2":  # the product's default kernels: stage_bb2 (kernel 6) and stage_bbe (kernel 7), tetrahedra and triangles
        lib.bb2e_run.argtypes = [C.c_void_p, C.c_int, C.c_int, dp, C.c_int]
        for cells, order, kernel in ((2, 4, 6), (2, 3, 6), (2, 2, 6), (3, 1, 7)):
            mesh, u = mesh_and_state(pkg, cells, order, (3.0, 2.0, 1.0))
            d = C.cast(mesh.desc_p, C.c_void_p)
            for integrator, steps in ((2, 0), (1, 1), (0, 1)):
                w = u.copy()
                assert lib.bb2e_run(d, kernel, integrator, w.ctypes.data_as(dp), steps) == 0
        for cells, order, kernel in ((3, 5, 6), (5, 1, 7), (4, 2, 7)):
            model = pkg.Model.make_square(cells, -10.0, 10.0, order)
            mesh = pkg.Mesh(model, pkg.Config())
            mesh.set_physics(c0=343.0, rho0=1.225, v0=(3.0, 2.0, 0.0), dt=1e-5)
            u = np.random.default_rng(0).standard_normal((4, mesh.N))
            assert lib.bb2e_run(C.cast(mesh.desc_p, C.c_void_p), kernel, 1, u.ctypes.data_as(dp), 1) == 0
    elif kind == "curved":
        lib.cve_run.argtypes = [C.c_void_p, C.c_int, dp, C.c_int]
        mesh, u = mesh_and_state(pkg, 2, 2, (3.0, 2.0, 1.0), warp=(0.3, 0.4))
        assert lib.cve_run(C.cast(mesh.desc_p, C.c_void_p), 1, u.ctypes.data_as(dp), 1) == 0
    elif kind == "generic":
        lib.gne_run.argtypes = [C.c_void_p, C.c_int, dp, C.c_double, C.c_int, C.c_int, ip, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, ip, dp, dp]
        mesh, u = mesh_and_state(pkg, 2, 2, (3.0, 2.0, 1.0))
        assert lib.gne_run(C.cast(mesh.desc_p, C.c_void_p), 1, u.ctypes.data_as(dp), 0.0, 1, 0, None, 0.0, 0.0, 0.0, 0.0, 0, None, None, None) == 0
    print("worker done", flush=True)


if __name__ == "__main__":
    main()
