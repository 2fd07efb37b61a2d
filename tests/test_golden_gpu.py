"""The CUDA engine (through the C ABI) against golden vectors produced by the reference's own code."""
import numpy as np
import pytest

from conftest import rel_l2
from golden_util import CASE_NAMES, load_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", CASE_NAMES)
def test_engine_reproduces_reference(pkg, mesh_dir, tmp_path, name):
    mesh, cfg, integ, snaps = load_case(pkg, mesh_dir, tmp_path, name)
    eng = pkg.Engine(mesh)
    eng.set_sources_from_config()
    eng.set_state(mesh.initial_condition())
    t, done = cfg.c.timeStart, 0
    for step, t_ref, u_ref in snaps:
        t = eng.run(integ, t, step - done)
        done = step
        assert t == t_ref
        u = eng.get_state()
        for q in range(4):
            if np.abs(u_ref[q]).max() == 0.0:
                assert np.abs(u[q]).max() == 0.0
            else:
                assert rel_l2(u[q], u_ref[q]) < 1e-10, (name, step, q)
