"""Development aid: per-phase cycle breakdown of the tiled kernel (needs lib/libdgb_dbg.so built with -DDGB_TILED_PHASE_TIMERS)."""
import ctypes as C, sys, os, shutil, numpy as np
sys.path.insert(0, "/root/repo")
import __graft_entry__ as g
pkg = g.load_package()
import dgfem_acoustic_b200.capi as capi
dbg = capi.LIB_DIR / "libdgb_dbg.so"
orig = capi.LIB_DIR / "libdgb.so"
bak = capi.LIB_DIR / "libdgb_product.so"
shutil.copy(orig, bak); shutil.copy(dbg, orig)
try:
    lib = capi.load_dgb()
    cells = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    model = pkg.Model.make_cube(cells, -10, 10, 4); cfg = pkg.Config(); cfg.add_initial_condition(0, 0, 0, 1, 1)
    mesh = pkg.Mesh(model, cfg); mesh.set_physics(c0=343.0, rho0=1.225, v0=(0, 0, 0), dt=1e-6)
    eng = pkg.Engine(mesh); eng.set_state(mesh.initial_condition()); eng.run(1, 0.0, 2)
    buf = (C.c_ulonglong * 8)()
    lib.dgbTiledPhaseTimers(buf, 1)
    eng.run(1, 0.0, 5)
    lib.dgbTiledPhaseTimers(buf, 1)
    t = np.array(list(buf), dtype=np.float64)
    nunits = mesh.K / 4 * 20
    names = ["wait prefetch", "flux compute", "Bq/G + issue prefetch", "lift MMA", "volume MMA + combine", "RK epilogue"]
    print("cells", cells, "stage ms", eng.last_stage_kernel_ms, "kernel", eng.kernel_name)
    for n, v in zip(names, t[:6]): print(f"{n:26s} {v / nunits:10.0f} cycles per unit per warp ({100 * v / t[:6].sum():.1f}%)")
finally:
    shutil.copy(bak, orig); os.remove(bak)
