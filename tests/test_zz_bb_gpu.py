"""Bernstein-Bezier stage kernel (csrc/stage_bb.cu, dgb_set_option("kernel", 4)) against the CPU oracle, through the C ABI.

The operator code of this kernel is also checked on the CPU (tests/test_bb_ops.py runs the same templates on the host,
tests/test_bb_emulated.py the kernel source through a CUDA emulation). Tolerance 1e-10 as everywhere."""

import numpy as np
import pytest

from conftest import rel_l2

pytestmark = [pytest.mark.gpu]
TOL = 1e-10


def _mesh(pkg, mesh_dir, name, order, v0):
    if name.startswith("cube:"):
        model = pkg.Model.make_cube(int(name.split(":")[1]), -10.0, 10.0, order)
    elif name.startswith("square:"):
        model = pkg.Model.make_square(int(name.split(":")[1]), -10.0, 10.0, order)
    else:
        model = pkg.Model.open_msh(mesh_dir / name, order)
    mesh = pkg.Mesh(model, pkg.Config())
    mesh.set_physics(c0=343.0, rho0=1.225, v0=v0, dt=0.1 * mesh.h_min() / (343.0 * (2 * order + 1)))
    b = np.nonzero(mesh.fIsBoundary)[0]
    mesh.fBC[b[::2]] = 1
    return mesh


def _state(mesh, seed=0):
    rng = np.random.default_rng(seed)
    x = mesh.node_coords
    u = np.zeros((4, mesh.N))
    for q in range(4):
        k, ph = rng.uniform(0.5, 2, 3), rng.uniform(0, 6, 3)
        u[q] = np.cos(k[0] * x[:, 0] * 0.3 + ph[0]) * np.cos(k[1] * x[:, 1] * 0.3 + ph[1]) * np.cos(k[2] * x[:, 2] * 0.3 + ph[2])
    u[1:] *= 1e-3
    return u


CASES = [("cube:3", 2, (0.0, 0.0, 0.0)), ("cube.msh", 3, (30.0, 10.0, 5.0)), ("cube:4", 4, (0.0, 0.0, 0.0)), ("sphere.msh", 4, (3.0, -2.0, 1.0)),
         ("cube:2", 5, (0.0, 0.0, 0.0))]


@pytest.mark.parametrize("variant", [4, 5, 6])
@pytest.mark.parametrize("name,order,v0", CASES)
def test_rhs_and_rk4(pkg, oracle_mod, mesh_dir, name, order, v0, variant):
    mesh = _mesh(pkg, mesh_dir, name, order, v0)
    u0 = _state(mesh)
    orc = oracle_mod.Oracle(mesh)
    tile = {2: 8, 3: 16}.get(order, 32)
    eng = pkg.Engine(mesh, options={"bb_tile": tile, "kernel": variant})
    assert eng.kernel_name == {4: f"stage_bb<3,{order}>/{tile}", 5: f"stage_bb_seq<3,{order}>/{tile}", 6: f"stage_bb2<3,{order}>"}[variant]
    rhs = eng.eval_rhs(u0)
    ref = orc.eval_rhs(oracle_mod.Oracle.OPERATOR, u0)
    for q in range(4):
        assert rel_l2(rhs[q], ref[q]) < TOL
    eng.set_state(u0)
    back = eng.get_state()
    for q in range(4):  # nodal -> Bernstein -> nodal
        assert rel_l2(back[q], u0[q]) < 1e-13
    t = eng.run(pkg.RUNGE_KUTTA, 0.0, 7)
    eng.run(pkg.RUNGE_KUTTA, t, 5)
    got = eng.get_state()
    want = u0.copy()
    orc.run(oracle_mod.Oracle.OPERATOR, pkg.RUNGE_KUTTA, want, 0.0, 12)
    for q in range(4):
        assert rel_l2(got[q], want[q]) < TOL
    eng.close()


# second-generation kernel only: triangles of orders 1..6 (the reference's 2D meshes, a refined square with partial tiles) and tetrahedra of order 1
CASES2 = [("square.msh", 1, (0.0, 0.0, 0.0)), ("square:5", 1, (30.0, 10.0, 0.0)), ("disk.msh", 2, (0.0, 0.0, 0.0)), ("square_reflection.msh", 3, (30.0, 10.0, 0.0)),
          ("square.msh", 4, (0.0, 0.0, 0.0)), ("square:3", 5, (3.0, -2.0, 0.0)), ("disk.msh", 6, (0.0, 0.0, 0.0)), ("cube:3", 1, (0.0, 0.0, 0.0)),
          ("sphere.msh", 1, (30.0, 10.0, 5.0)), ("cube:2", 6, (30.0, 10.0, 5.0)), ("cube:3", 6, (0.0, 0.0, 0.0))]


# element-per-thread kernel (stage_bbe.cu, kernel 7): triangles of orders 1..3, tetrahedra of orders 1 / 2
CASES_E = [(n, o, v, 7) for (n, o, v) in CASES2 if o <= 3 and not (o > 1 and n.startswith(("cube", "sphere")))] + [
    ("square:7", 2, (30.0, 10.0, 0.0), 7), ("cube:3", 2, (30.0, 10.0, 5.0), 7), ("sphere.msh", 2, (0.0, 0.0, 0.0), 7)]


@pytest.mark.parametrize("name,order,v0,variant", [(n, o, v, 6) for (n, o, v) in CASES2] + CASES_E)
def test_rhs_and_rk4_triangles_and_order_1(pkg, oracle_mod, mesh_dir, name, order, v0, variant):
    mesh = _mesh(pkg, mesh_dir, name, order, v0)
    u0 = _state(mesh)
    orc = oracle_mod.Oracle(mesh)
    eng = pkg.Engine(mesh, options={"kernel": variant})
    assert eng.kernel_name == (f"stage_bb2<{mesh.dim},{order}>" if variant == 6 else f"stage_bbe<{mesh.dim},{order}>")
    rhs = eng.eval_rhs(u0)
    ref = orc.eval_rhs(oracle_mod.Oracle.OPERATOR, u0)
    for q in range(4):
        if np.abs(ref[q]).max() == 0:  # v_z on a 2D mesh without mean flow: L(u)_vz == 0, the Bernstein path leaves rounding noise
            assert np.linalg.norm(rhs[q]) < 1e-10 * np.linalg.norm(u0[q]) * 343.0 / mesh.h_min()
        else:
            assert rel_l2(rhs[q], ref[q]) < TOL
    eng.set_state(u0)
    back = eng.get_state()
    for q in range(4):
        assert rel_l2(back[q], u0[q]) < 1e-13
    t = eng.run(pkg.RUNGE_KUTTA, 0.0, 7)
    eng.run(pkg.RUNGE_KUTTA, t, 5)
    got = eng.get_state()
    want = u0.copy()
    orc.run(oracle_mod.Oracle.OPERATOR, pkg.RUNGE_KUTTA, want, 0.0, 12)
    for q in range(4):
        assert rel_l2(got[q], want[q]) < TOL
    eng.run(pkg.EULER1, 0.0, 3)
    got = eng.get_state()
    orc.run(oracle_mod.Oracle.OPERATOR, pkg.EULER1, want, 0.0, 3)
    for q in range(4):
        assert rel_l2(got[q], want[q]) < TOL
    eng.close()


@pytest.mark.parametrize("variant", [6, 7])
def test_sources_probes_receivers_in_bernstein_mode_on_triangles(pkg, oracle_mod, mesh_dir, variant):
    model = pkg.Model.open_msh(mesh_dir / "square.msh", 3)
    cfg = pkg.Config()
    cfg.add_source(2.0, 1.0, 0.0, 1.0, 10.0, 1500.0, 0.3, 1.0)
    mesh = pkg.Mesh(model, cfg)
    mesh.set_physics(c0=343.0, rho0=1.225, v0=(0.0, 0.0, 0.0), dt=0.1 * mesh.h_min() / (343.0 * 7))
    offsets, idx = mesh.source_nodes()
    assert len(idx) > 3
    probes = np.array([mesh.nearest_node(0, 0, 0), mesh.nearest_node(3, 3, 0), int(idx[1])], dtype=np.int32)
    el, w = mesh.locate_receivers([(0.3, -0.2, 0.0), (2.0, 2.5, 0.0)])
    steps = 25
    u0 = _state(mesh, 5) * 1e-2
    eng = pkg.Engine(mesh, options={"kernel": variant})
    assert eng.kernel_name == ("stage_bb2<2,3>" if variant == 6 else "stage_bbe<2,3>")
    eng.set_sources_from_config()
    eng.set_probes(probes)
    eng.set_receivers(el, w)
    eng.set_state(u0)
    eng.run(pkg.RUNGE_KUTTA, 0.0, steps)
    got, rec_p, rec_r = eng.get_state(), eng.get_probes(steps), eng.get_receivers(steps)
    orc = oracle_mod.Oracle(mesh)
    orc.set_sources_from_config()
    orc.set_receivers(el, w)
    want = u0.copy()
    _, ref_p = orc.run(oracle_mod.Oracle.OPERATOR, pkg.RUNGE_KUTTA, want, 0.0, steps, probes)
    ref_r = orc.get_receivers(steps)
    for q in range(3):
        assert rel_l2(got[q], want[q]) < TOL
        assert rel_l2(rec_p[:, :, q], ref_p[:, :, q]) < TOL
        assert rel_l2(rec_r[:, :, q], ref_r[:, :, q]) < TOL
    eng.close()


@pytest.mark.parametrize("first,second", [(4, 3), (6, 3), (4, 6), (6, 5)])
def test_euler_and_switching_the_representation_with_a_resident_state(pkg, oracle_mod, mesh_dir, first, second):
    mesh = _mesh(pkg, mesh_dir, "cube:3", 4, (0.0, 0.0, 0.0))
    u0 = _state(mesh, 3)
    eng = pkg.Engine(mesh, options={"kernel": 3})
    eng.set_state(u0)
    eng.run(pkg.EULER1, 0.0, 5)           # warp-specialised kernel, nodal state
    eng.set_option("kernel", first)        # converts the resident state to Bernstein coefficients (layout of that kernel)
    t = 0.0
    for _ in range(5):
        t += mesh.desc.dt
    eng.run(pkg.EULER1, t, 6)
    eng.set_option("kernel", second)       # and back, or on to the other coefficient layout
    for _ in range(6):
        t += mesh.desc.dt
    eng.run(pkg.EULER1, t, 4)
    got = eng.get_state()
    want = u0.copy()
    oracle_mod.Oracle(mesh).run(oracle_mod.Oracle.OPERATOR, pkg.EULER1, want, 0.0, 15)
    for q in range(4):
        assert rel_l2(got[q], want[q]) < TOL
    eng.close()


@pytest.mark.parametrize("variant", [4, 6])
def test_sources_probes_receivers_in_bernstein_mode(pkg, oracle_mod, mesh_dir, variant):
    """Hard source (nodal overwrite expressed on the coefficients), probes and receivers (V-weighted gathers)."""
    model = pkg.Model.make_cube(4, -10.0, 10.0, 3)
    cfg = pkg.Config()
    cfg.add_source(2.0, 1.0, 0.0, 4.0, 10.0, 1500.0, 0.3, 1.0)
    cfg.add_source(-5.0, -5.0, 2.0, 3.0, 4.0, 900.0, 0.0, 1.0)
    mesh = pkg.Mesh(model, cfg)
    mesh.set_physics(c0=343.0, rho0=1.225, v0=(0.0, 0.0, 0.0), dt=0.1 * mesh.h_min() / (343.0 * 7))
    offsets, idx = mesh.source_nodes()
    assert len(idx) > 10
    probes = np.array([mesh.nearest_node(0, 0, 0), mesh.nearest_node(5, 5, 5), int(idx[3])], dtype=np.int32)
    el, w = mesh.locate_receivers([(0.3, -0.2, 0.9), (4.0, 4.5, -3.0)])
    steps = 20
    u0 = _state(mesh, 5) * 1e-2
    eng = pkg.Engine(mesh, options={"kernel": variant})
    eng.set_sources_from_config()
    eng.set_probes(probes)
    eng.set_receivers(el, w)
    eng.set_state(u0)
    eng.run(pkg.RUNGE_KUTTA, 0.0, steps)
    got, rec_p, rec_r = eng.get_state(), eng.get_probes(steps), eng.get_receivers(steps)
    orc = oracle_mod.Oracle(mesh)
    orc.set_sources_from_config()
    orc.set_receivers(el, w)
    want = u0.copy()
    _, ref_p = orc.run(oracle_mod.Oracle.OPERATOR, pkg.RUNGE_KUTTA, want, 0.0, steps, probes)
    ref_r = orc.get_receivers(steps)
    for q in range(4):
        assert rel_l2(got[q], want[q]) < TOL
        assert rel_l2(rec_p[:, :, q], ref_p[:, :, q]) < TOL
        assert rel_l2(rec_r[:, :, q], ref_r[:, :, q]) < TOL
    eng.close()


def test_switching_between_the_two_default_kernels_keeps_the_resident_state(pkg, oracle_mod, mesh_dir):
    """stage_bbe (kernel 7) and stage_bb2 (kernel 6) share the interleaved Bernstein representation: switching converts nothing."""
    mesh = _mesh(pkg, mesh_dir, "square.msh", 2, (30.0, 10.0, 0.0))
    u0 = _state(mesh, 2)
    eng = pkg.Engine(mesh)
    assert eng.kernel_name == "stage_bbe<2,2>" and eng.get_option("representation") == 2
    eng.set_state(u0)
    t = eng.run(pkg.RUNGE_KUTTA, 0.0, 4)
    eng.set_option("kernel", 6)
    assert eng.kernel_name == "stage_bb2<2,2>" and eng.get_option("representation") == 2
    t = eng.run(pkg.RUNGE_KUTTA, t, 3)
    eng.set_option("kernel", 1)            # nodal kernel: one conversion
    assert eng.get_option("representation") == 0
    t = eng.run(pkg.RUNGE_KUTTA, t, 2)
    eng.set_option("kernel", 0)            # back to the automatic choice
    assert eng.kernel_name == "stage_bbe<2,2>"
    eng.run(pkg.RUNGE_KUTTA, t, 3)
    got = eng.get_state()
    want = u0.copy()
    oracle_mod.Oracle(mesh).run(oracle_mod.Oracle.OPERATOR, pkg.RUNGE_KUTTA, want, 0.0, 12)
    for q in range(4):
        assert rel_l2(got[q], want[q]) < TOL
    eng.close()


def test_unsupported_combinations_fail_loudly(pkg, mesh_dir):
    mesh2 = pkg.Mesh(pkg.Model.open_msh(mesh_dir / "square.msh", 2), pkg.Config())
    eng2 = pkg.Engine(mesh2)
    with pytest.raises(pkg.DgbError):
        eng2.set_option("kernel", 4)  # triangles: no first-generation Bernstein kernel
    eng2.close()
    mesh1 = pkg.Mesh(pkg.Model.open_msh(mesh_dir / "line.msh", 1), pkg.Config())
    eng1 = pkg.Engine(mesh1)
    with pytest.raises(pkg.DgbError):
        eng1.set_option("kernel", 6)  # lines: no Bernstein kernel at all
    eng1.close()
    mesh6 = pkg.Mesh(pkg.Model.make_cube(2, -10.0, 10.0, 6), pkg.Config())
    eng6 = pkg.Engine(mesh6)
    with pytest.raises(pkg.DgbError):
        eng6.set_option("kernel", 4)  # order 6: no first-generation kernel
    with pytest.raises(pkg.DgbError):
        eng6.set_option("kernel", 7)  # nor an element-per-thread one
    eng6.close()
