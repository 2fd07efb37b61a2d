"""Pins the CPU oracle (both modes) to golden vectors produced by the reference's own code."""
import numpy as np
import pytest

from conftest import rel_l2
from golden_util import CASE_NAMES, load_case


@pytest.mark.parametrize("name", CASE_NAMES)
@pytest.mark.parametrize("mode", [0, 1])
def test_oracle_reproduces_reference(pkg, oracle_mod, mesh_dir, tmp_path, name, mode):
    mesh, cfg, integ, snaps = load_case(pkg, mesh_dir, tmp_path, name)
    orc = oracle_mod.Oracle(mesh, threads=2)
    orc.set_sources_from_config()
    u = mesh.initial_condition()
    t, done = cfg.c.timeStart, 0
    for step, t_ref, u_ref in snaps:
        t, _ = orc.run(mode, integ, u, t, step - done)
        done = step
        assert t == t_ref  # the FP-accumulated time of the reference's loop header
        for q in range(4):
            if np.abs(u_ref[q]).max() == 0.0:
                assert np.abs(u[q]).max() == 0.0
            else:
                assert rel_l2(u[q], u_ref[q]) < 1e-12, (name, mode, step, q)
