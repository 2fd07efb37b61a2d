"""Shared helpers for the golden-vector tests (tests/golden/cases/*.npz, produced by tests/golden/make_golden.py
by running the reference's own sources, see oracle/Makefile)."""
from pathlib import Path

import numpy as np

CASES_DIR = Path(__file__).resolve().parent / "golden" / "cases"
CASE_NAMES = sorted(p.stem for p in CASES_DIR.glob("*.npz"))


def load_case(pkg, mesh_dir, tmp_path, name):
    z = np.load(CASES_DIR / f"{name}.npz", allow_pickle=False)
    mesh_name, order = str(z["mesh"]), int(z["order"])
    if mesh_name.startswith("gen:cube"):
        model = pkg.Model.make_cube(int(mesh_name[8:]), -10.0, 10.0, order)
    else:
        model = pkg.Model.open_msh(mesh_dir / mesh_name, order)
    conf = tmp_path / f"{name}.conf"
    conf.write_text(str(z["config"]))
    cfg = model.parse_config(conf)
    mesh = pkg.Mesh(model, cfg)
    integ = pkg.RUNGE_KUTTA if cfg.c.timeIntMethod == b"Runge-Kutta" else pkg.EULER1
    snaps = [(int(s), float(z[f"t_{int(s)}"]), z[f"u_{int(s)}"]) for s in z["steps"]]
    return mesh, cfg, integ, snaps
