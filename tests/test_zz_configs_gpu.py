"""north_star configs 3 and 4 at scale: the shipped 3D meshes, the configured physics, hundreds of steps,
state AND time series against the oracle's operator mode to 1e-10 (their first steps are pinned against the reference's own
binary in test_zz_dropin_gpu.py; the oracle's operator mode equals its faithful mode and the reference binary to 1e-12,
test_oracle_vs_reference.py / test_known_answers.py)."""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-10


def test_config3_cube_order3_mean_flow_200_steps(pkg, oracle_mod, mesh_dir):
    """doc/3d/cube.msh elevated to order 3, physical group "Boundary" unmatched -> absorbing (Mesh.cpp:364, configParser.cpp:129-131),
    linearised Euler with mean flow v0 = (30, 10, 5), Gaussian pulse; 200 RK4 steps in chunks (snapshot cadence)."""
    model = pkg.Model.open_msh(mesh_dir / "cube.msh", 3)
    cfg = pkg.Config()
    cfg.add_initial_condition(0.0, 0.0, 0.0, 1.0, 1.0)
    mesh = pkg.Mesh(model, cfg)
    dt = 0.1 * mesh.h_min() / (343.0 * 7)
    mesh.set_physics(c0=343.0, rho0=1.225, v0=(30.0, 10.0, 5.0), dt=dt)
    assert mesh.K == 13603 and mesh.Np == 20
    probes = np.array([mesh.nearest_node(0, 0, 0), mesh.nearest_node(0.6, -0.4, 0.3), mesh.nearest_node(-0.8, 0.8, 0.5)], dtype=np.int32)
    u0 = mesh.initial_condition()
    eng = pkg.Engine(mesh)
    assert "bb2" in eng.kernel_name
    eng.set_probes(probes)
    eng.set_state(u0)
    t = 0.0
    for chunk in (1, 49, 50, 100):
        t = eng.run(pkg.RUNGE_KUTTA, t, chunk)
    got, rec = eng.get_state(), eng.get_probes(200)
    want = u0.copy()
    _, ref_rec = oracle_mod.Oracle(mesh).run(oracle_mod.Oracle.OPERATOR, pkg.RUNGE_KUTTA, want, 0.0, 200, probes)
    for q in range(4):
        assert rel_l2(got[q], want[q]) < TOL, q
        assert rel_l2(rec[:, :, q], ref_rec[:, :, q]) < TOL, q
    assert np.abs(want[3]).max() > 0  # v0_z and the 3D pulse make every field live
    eng.close()


def test_config4_sphere_order4_source_receivers_500_steps(pkg, oracle_mod, mesh_dir):
    """doc/3d/sphere.msh (stand-in for the missing auditorium mesh) at order 4 with the physics of doc/config/Amphi_Pulse_3D.conf
    (dt 1e-5, c0 343, rho0 1.225, absorbing walls), one monopole source (doc/config/Room_2D.conf:48 syntax), three interpolated
    receivers and three probes; 500 RK4 steps."""
    model = pkg.Model.open_msh(mesh_dir / "sphere.msh", 4)
    cfg = pkg.Config()
    cfg.add_source(0.2, 0.1, 0.0, 0.35, 10.0, 2000.0, 0.0, 0.004)  # x, y, z, size, amplitude, frequency, phase, duration (active for 400 steps)
    mesh = pkg.Mesh(model, cfg)
    mesh.set_physics(c0=343.0, rho0=1.225, v0=(0.0, 0.0, 0.0), dt=1e-5)
    assert mesh.K == 13905 and mesh.Np == 35
    offsets, idx = mesh.source_nodes()
    assert len(idx) > 0
    probes = np.array([mesh.nearest_node(0, 0, 0), mesh.nearest_node(0.5, 0.5, 0.5), int(idx[0])], dtype=np.int32)
    el, w = mesh.locate_receivers([(0.31, -0.22, 0.1), (-0.4, 0.4, 0.3), (0.0, 0.0, -0.7)])
    steps = 500
    eng = pkg.Engine(mesh)
    assert "bb2" in eng.kernel_name
    eng.set_sources_from_config()
    eng.set_probes(probes)
    eng.set_receivers(el, w)
    eng.set_state(np.zeros((4, mesh.N)))
    t = 0.0
    for chunk in (100, 150, 250):
        t = eng.run(pkg.RUNGE_KUTTA, t, chunk)
    got, rec_p, rec_r = eng.get_state(), eng.get_probes(steps), eng.get_receivers(steps)
    orc = oracle_mod.Oracle(mesh)
    orc.set_sources_from_config()
    orc.set_receivers(el, w)
    want = np.zeros((4, mesh.N))
    _, ref_p = orc.run(oracle_mod.Oracle.OPERATOR, pkg.RUNGE_KUTTA, want, 0.0, steps, probes)
    ref_r = orc.get_receivers(steps)
    assert np.abs(ref_r[:, :, 0]).max() > 1e-3  # the pulse reaches the receivers
    for q in range(4):
        assert rel_l2(got[q], want[q]) < TOL, q
        assert rel_l2(rec_p[:, :, q], ref_p[:, :, q]) < TOL, q
        assert rel_l2(rec_r[:, :, q], ref_r[:, :, q]) < TOL, q
    eng.close()
