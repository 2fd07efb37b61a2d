"""GPU parity: the CUDA engine, called through the C ABI (include/dgb.h), against the CPU oracle.

Tolerance (BASELINE.json north_star): relative L2 error per field <= 1e-10 in FP64.
"""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-10


def build_mesh(pkg, mesh_dir, name, order, v0=(0.0, 0.0, 0.0), c0=343.0, rho0=1.225, mixed_bc=True, cfl=0.1):
    if name.startswith("cube:"):
        model = pkg.Model.make_cube(int(name.split(":")[1]), -10.0, 10.0, order)
    elif name.startswith("square:"):
        model = pkg.Model.make_square(int(name.split(":")[1]), -10.0, 10.0, order)
    else:
        model = pkg.Model.open_msh(mesh_dir / name, order)
    cfg = pkg.Config()
    mesh = pkg.Mesh(model, cfg)
    dt = cfl * mesh.h_min() / (c0 * (2 * order + 1))
    mesh.set_physics(c0=c0, rho0=rho0, v0=v0, dt=dt)
    if mixed_bc:  # exercise both boundary conditions
        b = np.nonzero(mesh.fIsBoundary)[0]
        mesh.fBC[b[::2]] = 1
    return mesh


def smooth_state(mesh, seed=0):
    rng = np.random.default_rng(seed)
    x = mesh.node_coords
    L = np.abs(x).max() + 1e-12
    u = np.zeros((4, mesh.N))
    for q in range(4):
        k = rng.uniform(0.5, 2.0, size=3)
        ph = rng.uniform(0, 6.28, size=3)
        u[q] = np.cos(k[0] * x[:, 0] / L * 3 + ph[0]) * np.cos(k[1] * x[:, 1] / L * 3 + ph[1]) * np.cos(k[2] * x[:, 2] / L * 3 + ph[2])
    u[1:] *= 1e-3  # velocities ~ p/(rho c)
    return u


CASES = [
    ("line.msh", 1, (0, 0, 0)),
    ("square.msh", 1, (0, 0, 0)),
    ("square.msh", 2, (20.0, 5.0, 0)),
    ("square.msh", 3, (0, 0, 0)),
    ("disk.msh", 4, (10.0, -3.0, 0)),
    ("square.msh", 5, (0, 0, 0)),
    ("square.msh", 6, (5.0, 5.0, 0)),
    ("cube.msh", 1, (3.0, 2.0, 1.0)),
    ("cube.msh", 2, (0, 0, 0)),
    ("cube.msh", 3, (30.0, 10.0, 5.0)),
    ("sphere.msh", 4, (0, 0, 0)),
    ("cube:3", 4, (30.0, 10.0, 0.0)),
    ("cube:2", 5, (1.0, 2.0, 3.0)),
    ("cube:2", 6, (0, 0, 0)),
    ("square:7", 4, (15.0, -6.0, 0)),
]


@pytest.mark.parametrize("name,order,v0", CASES)
def test_rhs_matches_oracle(pkg, oracle_mod, mesh_dir, name, order, v0):
    """One operator evaluation L(u) on random (non-smooth) data: every term is exercised at full amplitude."""
    mesh = build_mesh(pkg, mesh_dir, name, order, v0)
    u = np.random.default_rng(1).standard_normal((4, mesh.N))
    eng = pkg.Engine(mesh)
    got = eng.eval_rhs(u)
    ref = oracle_mod.Oracle(mesh).eval_rhs(oracle_mod.Oracle.OPERATOR, u)
    for q in range(4):
        assert rel_l2(got[q], ref[q]) < 1e-12, (name, order, q, eng.kernel_name)


@pytest.mark.parametrize("name,order,v0,steps", [
    ("line.msh", 1, (0, 0, 0), 50),
    ("square.msh", 1, (0, 0, 0), 50),
    ("square.msh", 3, (20.0, 5.0, 0), 30),
    ("disk.msh", 6, (0, 0, 0), 10),
    ("cube.msh", 3, (30.0, 10.0, 5.0), 10),
    ("cube:3", 4, (30.0, 10.0, 0.0), 20),
])
def test_rk4_matches_oracle(pkg, oracle_mod, mesh_dir, name, order, v0, steps):
    mesh = build_mesh(pkg, mesh_dir, name, order, v0)
    u0 = smooth_state(mesh)
    eng = pkg.Engine(mesh)
    eng.set_state(u0)
    t_end = eng.run(pkg.RUNGE_KUTTA, 0.0, steps)
    got = eng.get_state()
    ref = u0.copy()
    t_ref, _ = oracle_mod.Oracle(mesh).run(oracle_mod.Oracle.OPERATOR, pkg.RUNGE_KUTTA, ref, 0.0, steps)
    assert t_end == t_ref
    for q in range(4):
        assert rel_l2(got[q], ref[q]) < TOL, (name, order, q)


def test_faithful_oracle_small(pkg, oracle_mod, mesh_dir):
    """Same comparison against the oracle's FAITHFUL mode (the reference's own loop nests)."""
    mesh = build_mesh(pkg, mesh_dir, "square.msh", 2, (20.0, 5.0, 0))
    u0 = smooth_state(mesh)
    eng = pkg.Engine(mesh)
    eng.set_state(u0)
    eng.run(pkg.RUNGE_KUTTA, 0.0, 20)
    got = eng.get_state()
    ref = u0.copy()
    oracle_mod.Oracle(mesh).run(oracle_mod.Oracle.FAITHFUL, pkg.RUNGE_KUTTA, ref, 0.0, 20)
    for q in range(4):
        assert rel_l2(got[q], ref[q]) < TOL


def test_euler_matches_oracle(pkg, oracle_mod, mesh_dir):
    mesh = build_mesh(pkg, mesh_dir, "square.msh", 2, (0, 0, 0), cfl=0.02)
    u0 = smooth_state(mesh)
    eng = pkg.Engine(mesh)
    eng.set_state(u0)
    eng.run(pkg.EULER1, 0.0, 25)
    got = eng.get_state()
    ref = u0.copy()
    oracle_mod.Oracle(mesh).run(oracle_mod.Oracle.OPERATOR, pkg.EULER1, ref, 0.0, 25)
    for q in range(4):
        assert rel_l2(got[q], ref[q]) < TOL


def test_sources_and_probes(pkg, oracle_mod, mesh_dir, config_dir):
    """Hard monopole source (solver.cpp:248-256) + probe time series, Room_2D.conf physics on the square mesh."""
    model = pkg.Model.open_msh(mesh_dir / "square.msh", 2)
    cfg = model.parse_config(config_dir / "room_source.conf")
    cfg.c.sources[0][4] = 0.6  # radius large enough to catch nodes of this coarse mesh
    mesh = pkg.Mesh(model, cfg)
    offsets, idx = mesh.source_nodes()
    assert len(idx) > 0
    probes = np.array([mesh.nearest_node(3.0, 1.2, 0.0), mesh.nearest_node(0.0, 0.0, 0.0), mesh.nearest_node(-4.0, 4.0, 0.0)], dtype=np.int32)
    steps = 60
    u0 = np.zeros((4, mesh.N))
    eng = pkg.Engine(mesh)
    eng.set_sources_from_config()
    eng.set_probes(probes)
    eng.set_state(u0)
    eng.run(pkg.RUNGE_KUTTA, cfg.c.timeStart, 25)
    t_mid = eng.run(pkg.RUNGE_KUTTA, 25 * 0.0, 0)  # zero-step call is a no-op
    # second chunk continues from the accumulated time of the first (the reference accumulates t += dt)
    t = cfg.c.timeStart
    for _ in range(25):
        t += cfg.c.timeStep
    eng.run(pkg.RUNGE_KUTTA, t, steps - 25)
    got = eng.get_state()
    rec = eng.get_probes(steps)
    orc = oracle_mod.Oracle(mesh)
    orc.set_sources_from_config()
    ref = u0.copy()
    _, rec_ref = orc.run(oracle_mod.Oracle.OPERATOR, pkg.RUNGE_KUTTA, ref, cfg.c.timeStart, steps, probes)
    assert rec.shape == rec_ref.shape == (steps, 3, 4)
    assert np.abs(ref[0]).max() > 0
    for q in range(4):
        assert rel_l2(got[q], ref[q]) < TOL
        assert rel_l2(rec[:, :, q], rec_ref[:, :, q]) < TOL


def test_downwind_sign_branch(pkg, oracle_mod, mesh_dir):
    """sigma = -1: arithmetic parity on the other penalty-sign branch, few steps (it is unstable)."""
    mesh = build_mesh(pkg, mesh_dir, "cube.msh", 2, (0, 0, 0))
    mesh.desc.fc = -mesh.desc.fc
    u0 = smooth_state(mesh)
    eng = pkg.Engine(mesh)
    eng.set_state(u0)
    eng.run(pkg.RUNGE_KUTTA, 0.0, 3)
    got = eng.get_state()
    ref = u0.copy()
    oracle_mod.Oracle(mesh).run(oracle_mod.Oracle.OPERATOR, pkg.RUNGE_KUTTA, ref, 0.0, 3)
    for q in range(4):
        assert rel_l2(got[q], ref[q]) < TOL


def test_reference_config_square(pkg, oracle_mod, mesh_dir, config_dir):
    """BASELINE config 1: square.msh + the reference's minimal config, the full 1000 steps."""
    model = pkg.Model.open_msh(mesh_dir / "square.msh", 1)
    cfg = model.parse_config(config_dir / "square_pulse.conf")
    mesh = pkg.Mesh(model, cfg)
    steps, snaps = cfg.time_loop()
    assert steps == 1000 and len(snaps) == 100
    u0 = mesh.initial_condition()
    eng = pkg.Engine(mesh)
    eng.set_state(u0)
    eng.run(pkg.RUNGE_KUTTA, cfg.c.timeStart, steps)
    got = eng.get_state()
    ref = u0.copy()
    oracle_mod.Oracle(mesh).run(oracle_mod.Oracle.OPERATOR, pkg.RUNGE_KUTTA, ref, cfg.c.timeStart, steps)
    for q in range(3):
        assert rel_l2(got[q], ref[q]) < TOL
    assert np.abs(got[3]).max() == 0.0


def test_large_mesh_properties(pkg, mesh_dir):
    """Size-independent properties at a size the oracle would not finish quickly: linearity of L and the
    free-stream / rigid-wall known answer (constant p, v = 0, reflecting walls => L(u) = 0)."""
    model = pkg.Model.make_cube(20, -10.0, 10.0, 4)  # 48 000 tets, 1.68 M nodes
    cfg = pkg.Config()
    mesh = pkg.Mesh(model, cfg)
    mesh.set_physics(c0=343.0, rho0=1.225, v0=(0, 0, 0), dt=1e-5)
    mesh.fBC[:] = 1
    eng = pkg.Engine(mesh)
    const = np.zeros((4, mesh.N))
    const[0] = 2.5
    r = eng.eval_rhs(const)
    assert np.abs(r).max() < 1e-9 * 343.0 ** 2
    rng = np.random.default_rng(3)
    a, b = rng.standard_normal((4, mesh.N)), rng.standard_normal((4, mesh.N))
    ra, rb, rab = eng.eval_rhs(a), eng.eval_rhs(b), eng.eval_rhs(2.0 * a - 0.5 * b)
    assert rel_l2(rab, 2.0 * ra - 0.5 * rb) < 1e-13


def test_error_behaviour(pkg, mesh_dir):
    mesh = build_mesh(pkg, mesh_dir, "line.msh", 1)
    eng = pkg.Engine(mesh)
    with pytest.raises(pkg.DgbError):
        eng.run(pkg.RUNGE_KUTTA, 0.0, 1)  # run before set_state
    eng.set_state(np.zeros((4, mesh.N)))
    with pytest.raises(pkg.DgbError):
        eng.run(7, 0.0, 1)  # unknown integrator
    with pytest.raises(pkg.DgbError):
        eng.set_option("no_such_option", 1)
    with pytest.raises(pkg.DgbError):
        eng.set_option("kernel", 3)  # no warp-specialised kernel for a 1D mesh
    with pytest.raises(pkg.DgbError):
        eng.set_option("overlap", 3)
    flow = build_mesh(pkg, mesh_dir, "cube:2", 4, (10.0, 0.0, 0.0))
    eng2 = pkg.Engine(flow)
    assert "bb2" in eng2.kernel_name  # the automatic choice for tetrahedra of order >= 3, with or without mean flow
    eng2.set_option("kernel", 2)
    assert "tiled" in eng2.kernel_name  # mean flow: the warp-specialised kernel is not offered
    with pytest.raises(pkg.DgbError):
        eng2.set_option("kernel", 3)


KERNEL_IDS = {"generic": 1, "tiled": 2, "ws": 3, "bb2": 6}


@pytest.mark.parametrize("name,order,v0", [("cube.msh", 3, (30.0, 10.0, 5.0)), ("sphere.msh", 4, (0, 0, 0)), ("cube:5", 4, (30.0, 10.0, 0.0)),
                                           ("cube:3", 3, (0.0, 0.0, 0.0)), ("cube.msh", 4, (0.0, 0.0, 0.0)), ("cube:1", 4, (0.0, 0.0, 0.0))])
def test_dmma_kernels_vs_generic_and_oracle(pkg, oracle_mod, mesh_dir, name, order, v0):
    """The FP64 tensor-core kernels (tiled; warp-specialised for zero mean flow), the Bernstein kernel and the CUDA-core kernel
    are independent implementations of the same operator. cube:1 (6 elements) and the shipped meshes give ragged last tiles."""
    mesh = build_mesh(pkg, mesh_dir, name, order, v0)
    u = np.random.default_rng(5).standard_normal((4, mesh.N))
    ref = oracle_mod.Oracle(mesh).eval_rhs(oracle_mod.Oracle.OPERATOR, u)
    eng = pkg.Engine(mesh)
    kernels = ["generic", "tiled"] + (["ws"] if all(v == 0 for v in v0) else []) + ["bb2"]
    assert "bb2" in eng.kernel_name  # the automatic choice for tetrahedra of order >= 3 (second-generation Bernstein kernel)
    for kern in kernels:
        eng.set_option("kernel", KERNEL_IDS[kern])
        assert kern in eng.kernel_name
        got = eng.eval_rhs(u)
        for q in range(4):
            assert rel_l2(got[q], ref[q]) < 1e-12, (kern, q)
    # and through a few RK4 / Euler steps with each kernel
    u0 = smooth_state(mesh)
    for integ, steps in ((pkg.RUNGE_KUTTA, 5), (pkg.EULER1, 3)):
        out = {}
        for kern in kernels:
            eng.set_option("kernel", KERNEL_IDS[kern])
            eng.set_state(u0)
            eng.run(integ, 0.0, steps)
            out[kern] = eng.get_state()
        for kern in kernels[1:]:
            for q in range(4):
                assert rel_l2(out[kern][q], out["generic"][q]) < 1e-12, (kern, q)


def test_ws_kernel_rk4_vs_oracle_many_tiles(pkg, oracle_mod, mesh_dir):
    """Warp-specialised kernel over many tiles per CTA (ring buffers wrap several times), against the oracle."""
    mesh = build_mesh(pkg, mesh_dir, "cube:9", 4, (0.0, 0.0, 0.0))  # 4374 tets = 547 tiles -> 3-4 tiles per CTA
    u0 = smooth_state(mesh)
    eng = pkg.Engine(mesh, options={"kernel": 3})
    assert "ws" in eng.kernel_name
    eng.set_state(u0)
    eng.run(pkg.RUNGE_KUTTA, 0.0, 10)
    got = eng.get_state()
    ref = u0.copy()
    oracle_mod.Oracle(mesh).run(oracle_mod.Oracle.OPERATOR, pkg.RUNGE_KUTTA, ref, 0.0, 10)
    for q in range(4):
        assert rel_l2(got[q], ref[q]) < TOL


def test_ws_kernel_deep_rings(pkg, mesh_dir):
    """48 000 tets = 6000 tiles (40 per CTA): the warp-specialised kernel against the CUDA-core kernel."""
    mesh = build_mesh(pkg, mesh_dir, "cube:20", 4, (0.0, 0.0, 0.0))
    u = np.random.default_rng(11).standard_normal((4, mesh.N))
    eng = pkg.Engine(mesh, options={"kernel": 3})
    assert "ws" in eng.kernel_name
    a = eng.eval_rhs(u)
    eng.set_option("kernel", 1)
    b = eng.eval_rhs(u)
    for q in range(4):
        assert rel_l2(a[q], b[q]) < 1e-12


@pytest.mark.parametrize("kernel,tag", [(3, "ws"), (6, "bb2")])
def test_ws_kernel_is_deterministic(pkg, mesh_dir, kernel, tag):
    """The warp-specialised kernel hands tiles between warp roles through shared-memory rings and mbarriers, the Bernstein
    kernel between the copy engines and its warp through mbarriers and async-proxy fences; a missed ordering would show up
    as run-to-run differences. Same input, same launch configuration -> bit-identical results."""
    mesh = build_mesh(pkg, mesh_dir, "cube:12", 4, (0.0, 0.0, 0.0))  # 10 368 tets = 1 296 tiles, ~9 per CTA
    u0 = np.random.default_rng(17).standard_normal((4, mesh.N))
    eng = pkg.Engine(mesh, options={"kernel": kernel})
    assert tag in eng.kernel_name
    runs = []
    for _ in range(3):
        eng.set_state(u0)
        eng.run(pkg.RUNGE_KUTTA, 0.0, 6)
        runs.append(eng.get_state().copy())
    assert np.array_equal(runs[0], runs[1]) and np.array_equal(runs[0], runs[2])
    assert np.isfinite(runs[0]).all()


@pytest.mark.parametrize("name,order", [("square.msh", 2), ("cube:3", 4)])
def test_cuda_graph_of_a_step_equals_eager_launches(pkg, mesh_dir, name, order):
    """Launch-bound runs replay one captured RK4 step (dgb_set_option("graph")): same kernels, same arguments -> bit-identical
    results and the same accumulated end time as the eager loop; the launch count still counts every stage."""
    mesh = build_mesh(pkg, mesh_dir, name, order, (0.0, 0.0, 0.0))
    u0 = smooth_state(mesh)
    out, tend, launches = {}, {}, {}
    for mode in (0, 1):
        eng = pkg.Engine(mesh)
        eng.set_option("graph", mode)
        eng.set_state(u0)
        tend[mode] = eng.run(pkg.RUNGE_KUTTA, 0.0, 40)
        tend[mode] = eng.run(pkg.RUNGE_KUTTA, tend[mode], 25)  # second call re-uses the instantiated graph
        out[mode] = eng.get_state().copy()
        launches[mode] = eng.launch_count
        conversions = 2 if "bb" in eng.kernel_name else 0  # Bernstein kernels: dgb_set_state / dgb_get_state each launch one conversion
        eng.close()
    assert np.array_equal(out[0], out[1])
    assert tend[0] == tend[1] and launches[0] == launches[1] == 4 * 65 + conversions
