"""Receivers: point location + Lagrange interpolation weights (front end), the oracle's receiver
record, the time-series writers, and — on the GPU — the engine's receiver record against the oracle's.

The reference has no receivers (its assets/ show receiver audio made outside the code), so the checks are
known answers: interpolation of polynomials of degree <= p is exact, a receiver on a DG node equals a probe."""
import struct

import numpy as np
import pytest

from conftest import rel_l2


def _mesh(pkg, mesh_dir, name, order):
    if name.startswith("cube:"):
        model = pkg.Model.make_cube(int(name.split(":")[1]), -10.0, 10.0, order)
    else:
        model = pkg.Model.open_msh(mesh_dir / name, order)
    cfg = pkg.Config()
    mesh = pkg.Mesh(model, cfg)
    mesh.set_physics(c0=343.0, rho0=1.225, v0=(0.0, 0.0, 0.0), dt=0.1 * mesh.h_min() / (343.0 * (2 * order + 1)))
    return mesh


def _poly(x, order, dim, seed):
    """a random polynomial of total degree <= order in the first dim coordinates"""
    rng = np.random.default_rng(seed)
    val = np.zeros(x.shape[:-1])
    for a in range(order + 1):
        for b in range(order + 1 - a if dim > 1 else 1):
            for c in range(order + 1 - a - b if dim > 2 else 1):
                val = val + rng.uniform(-1, 1) * (x[..., 0] / 10) ** a * (x[..., 1] / 10) ** b * (x[..., 2] / 10) ** c
    return val


@pytest.mark.parametrize("name,order", [("line.msh", 1), ("square.msh", 1), ("square.msh", 3), ("disk.msh", 5),
                                        ("cube.msh", 2), ("cube:3", 4), ("cube:2", 6)])
def test_interpolation_is_exact_for_polynomials(pkg, mesh_dir, name, order):
    mesh = _mesh(pkg, mesh_dir, name, order)
    rng = np.random.default_rng(7)
    x = mesh.node_coords.reshape(mesh.K, mesh.Np, 3)
    nv = mesh.dim + 1
    f_nodes = _poly(x, order, mesh.dim, 11)
    for _ in range(25):
        el = int(rng.integers(mesh.K))
        lam = rng.dirichlet(np.ones(nv))  # a point strictly inside element el (the first dim+1 nodes are its vertices)
        pt = lam @ x[el, :nv]
        got_el, w, uvw, outside = mesh.locate_point(*pt)
        assert not outside
        assert abs(w.sum() - 1.0) < 1e-12  # partition of unity
        # the located element contains the point: the interpolated coordinates give the point back
        assert np.allclose(w @ x[got_el], pt, atol=1e-11 * 10)
        exact = _poly(pt[None, :], order, mesh.dim, 11)[0]
        assert abs(w @ f_nodes[got_el] - exact) < 1e-11 * max(1.0, abs(exact))


def test_receiver_on_a_node_is_a_probe_and_shared_faces_pick_the_lowest_element(pkg, mesh_dir):
    mesh = _mesh(pkg, mesh_dir, "square.msh", 2)
    x = mesh.node_coords.reshape(mesh.K, mesh.Np, 3)
    el, n = 321, 4  # an edge node of some element
    got_el, w, _, outside = mesh.locate_point(*x[el, n])
    assert not outside
    # every element that carries a node at this position
    owners = np.nonzero((np.abs(x - x[el, n]).sum(axis=2) < 1e-12).any(axis=1))[0]
    assert got_el == owners.min()
    k = int(np.argmax(w))
    assert abs(w[k] - 1.0) < 1e-12 and np.abs(np.delete(w, k)).max() < 1e-12
    assert np.abs(x[got_el, k] - x[el, n]).max() < 1e-12


def test_points_outside_the_mesh_are_flagged(pkg, mesh_dir):
    mesh = _mesh(pkg, mesh_dir, "cube:2", 2)
    _, _, _, outside = mesh.locate_point(10.5, 0.0, 0.0)
    assert outside
    _, _, _, outside = mesh.locate_point(10.0, 10.0, 10.0)  # a corner of the cube is still inside
    assert not outside
    with pytest.raises(pkg.FrontError):
        mesh.locate_receivers([(0.0, 0.0, 0.0), (30.0, 0.0, 0.0)])


def test_oracle_receivers_against_probes_and_interpolated_state(pkg, oracle_mod, mesh_dir):
    mesh = _mesh(pkg, mesh_dir, "square.msh", 3)
    mesh.cfg.add_initial_condition(0.0, 0.0, 0.0, 1.0, 1.0)
    x = mesh.node_coords.reshape(mesh.K, mesh.Np, 3)
    node_pt = x[100, 0]
    pts = [tuple(node_pt), (0.37, -0.21, 0.0), (-2.2, 1.3, 0.0)]
    el, w = mesh.locate_receivers(pts)
    orc = oracle_mod.Oracle(mesh)
    orc.set_receivers(el, w)
    u = mesh.initial_condition()
    probe = np.array([el[0] * mesh.Np + int(np.argmax(w[0]))], dtype=np.int32)
    steps = 12
    u_start = u.copy()
    _, rec_probe = orc.run(oracle_mod.Oracle.OPERATOR, pkg.RUNGE_KUTTA, u, 0.0, steps, probe)
    rec = orc.get_receivers(steps)
    assert rec.shape == (steps, 3, 4)
    assert np.abs(rec[:, 0] - rec_probe[:, 0]).max() < 1e-13 * max(1.0, np.abs(rec_probe).max())
    # step 0 is the initial state interpolated at the points
    for j in range(3):
        for q in range(4):
            assert abs(rec[0, j, q] - w[j] @ u_start[q].reshape(mesh.K, mesh.Np)[el[j]]) < 1e-14
    assert orc.get_receivers(steps).shape[0] == 0  # the record is cleared by the read


def test_config_keys_and_writers(pkg, mesh_dir, config_dir, tmp_path):
    conf = tmp_path / "rcv.conf"
    conf.write_text((config_dir / "square_pulse.conf").read_text() +
                    "\nreceiverB = 1.0, -2.0, 0.0\nreceiverA = 0.5, 0.25, 0.0\nreceiverFile = out_rcv.txt\nreceiverWav = mic_\n")
    model = pkg.Model.open_msh(mesh_dir / "square.msh", 1)
    cfg = model.parse_config(conf)
    assert cfg.c.nReceivers == 2
    assert cfg.receivers == [[0.5, 0.25, 0.0], [1.0, -2.0, 0.0]]  # std::map order of the keys
    assert cfg.c.receiverFile == b"out_rcv.txt" and cfg.c.receiverWav == b"mic_"
    assert cfg.c.nSources == 0 and cfg.c.nInit == 1  # the reference's keys are parsed as before
    # the stock config has no receivers
    assert model.parse_config(config_dir / "square_pulse.conf").c.nReceivers == 0

    import ctypes as C
    lib = pkg.load_front()
    steps, dt = 50, 1.0 / 8000
    rec = np.zeros((steps, 2, 4))
    t = np.arange(steps) * dt
    rec[:, 0, 0] = np.sin(2 * np.pi * 440 * t)
    rec[:, 1, 0] = 0.25 * np.cos(2 * np.pi * 100 * t)
    rec[:, 1, 2] = 1e-3
    xyz = np.array(cfg.receivers)
    dp = C.POINTER(C.c_double)
    txt = tmp_path / "r.txt"
    assert lib.dgf_write_receivers(str(txt).encode(), 2, xyz.ctypes.data_as(dp), steps, 0.0, dt, rec.ctypes.data_as(dp)) == 0
    back = np.loadtxt(txt)
    assert back.shape == (steps, 1 + 8)
    assert np.array_equal(back[:, 1:].reshape(steps, 2, 4), rec)  # %.17g round-trips doubles
    acc = 0.0
    for k in range(steps):  # the time column accumulates like the loop header
        assert back[k, 0] == acc
        acc += dt
    wav = tmp_path / "r.wav"
    assert lib.dgf_write_wav(str(wav).encode(), 2, 0, 0, steps, dt, 0, rec.ctypes.data_as(dp)) == 0
    raw = wav.read_bytes()
    assert raw[:4] == b"RIFF" and raw[8:16] == b"WAVEfmt " and raw[36:40] == b"data"
    fmt, ch, rate = struct.unpack("<HHI", raw[20:28])
    assert (fmt, ch, rate) == (1, 1, 8000)
    pcm = np.frombuffer(raw[44:], dtype="<i2")
    assert len(pcm) == steps and np.abs(pcm / 32767.0 - rec[:, 0, 0] / np.abs(rec[:, 0, 0]).max()).max() < 1e-4
    assert lib.dgf_write_wav(str(wav).encode(), 2, 5, 0, steps, dt, 0, rec.ctypes.data_as(dp)) == -1  # bad receiver index


@pytest.mark.gpu
def test_engine_receivers_match_the_oracle(pkg, oracle_mod, mesh_dir, config_dir):
    """dgb_set_receivers / dgb_get_receivers against the oracle, with a source running, in two chunks; a receiver
    placed on a DG node equals the probe at that node bit for bit up to the summation of exact zeros."""
    model = pkg.Model.open_msh(mesh_dir / "square.msh", 3)
    cfg = model.parse_config(config_dir / "room_source.conf")
    cfg.c.sources[0][4] = 0.6
    mesh = pkg.Mesh(model, cfg)
    x = mesh.node_coords.reshape(mesh.K, mesh.Np, 3)
    pts = [tuple(x[700, 2]), (3.0, 1.2, 0.0), (0.1, -0.3, 0.0), (-4.0, 4.0, 0.0)]
    el, w = mesh.locate_receivers(pts)
    probe = np.array([el[0] * mesh.Np + int(np.argmax(w[0]))], dtype=np.int32)
    steps = 40
    eng = pkg.Engine(mesh)
    eng.set_sources_from_config()
    eng.set_receivers(el, w)
    eng.set_probes(probe)
    eng.set_state(np.zeros((4, mesh.N)))
    t = eng.run(pkg.RUNGE_KUTTA, cfg.c.timeStart, 15)
    eng.run(pkg.RUNGE_KUTTA, t, steps - 15)
    rec = eng.get_receivers(steps)
    rec_probe = eng.get_probes(steps)
    orc = oracle_mod.Oracle(mesh)
    orc.set_sources_from_config()
    orc.set_receivers(el, w)
    ref = np.zeros((4, mesh.N))
    orc.run(oracle_mod.Oracle.OPERATOR, pkg.RUNGE_KUTTA, ref, cfg.c.timeStart, steps)
    rec_ref = orc.get_receivers(steps)
    assert rec.shape == rec_ref.shape == (steps, 4, 4)
    assert np.abs(rec_ref[:, :, 0]).max() > 0
    for q in range(4):
        assert rel_l2(rec[:, :, q], rec_ref[:, :, q]) < 1e-10
    assert np.abs(rec[:, 0] - rec_probe[:, 0]).max() <= 1e-13 * max(1.0, np.abs(rec_probe).max())
    eng.close()


@pytest.mark.parametrize("name,order,warp", [("square.msh", 3, (0.15, 0.9)), ("cube:3", 2, (0.3, 0.4)), ("cube:2", 4, (0.4, 0.3))])
def test_point_location_on_curved_meshes(pkg, mesh_dir, name, order, warp):
    """Curved (warped isoparametric) elements: Newton on x(u) = sum_n phi_n(u) x_n — the located parametric point maps back to
    the physical point, the weights are a partition of unity and reproduce the (isoparametric) coordinate functions."""
    model = pkg.Model.make_cube(int(name.split(":")[1]), -10.0, 10.0, order) if name.startswith("cube:") else pkg.Model.open_msh(mesh_dir / name, order)
    model.warp(*warp)
    mesh = pkg.Mesh(model, pkg.Config())
    x = mesh.node_coords.reshape(mesh.K, mesh.Np, 3)
    rng = np.random.default_rng(3)
    nv = mesh.dim + 1
    for _ in range(20):
        el = int(rng.integers(mesh.K))
        lam = rng.dirichlet(np.ones(nv) * 2.0)
        pt = lam @ x[el, :nv]  # near element el (inside its straight-sided shadow); some element of the curved mesh contains it
        got_el, w, uvw, outside = mesh.locate_point(*pt)
        if outside:  # a point of the shadow that the warped boundary left outside the mesh
            continue
        assert abs(w.sum() - 1.0) < 1e-12
        assert np.abs(w @ x[got_el] - pt).max() < 1e-10 * 10
        assert uvw[: mesh.dim].min() > -1e-9 and uvw[: mesh.dim].sum() < 1 + 1e-9
