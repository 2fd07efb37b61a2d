"""Host-side domain decomposition (csrc/partition.cpp) exercised with a real 2-process exchange over gloo on CPU.

Each rank holds only its owned + halo elements (everything else is NaN), evaluates the operator with the CPU oracle,
exchanges the traces it owes its peers exactly the way the engine does (send list -> peer's halo slots, ordered by
global element id) and must reproduce the single-domain result on the elements it owns, stage after stage.
"""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _plan(pkg, mesh, part, rank, world):
    lib = pkg.load_front()
    lib.dgf_plan_create.restype = C.c_void_p
    lib.dgf_plan_create.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_int, C.c_int]
    lib.dgf_plan_sizes.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
    lib.dgf_plan_arrays.argtypes = [C.c_void_p] + [C.POINTER(C.c_int32)] * 5
    lib.dgf_plan_free.argtypes = [C.c_void_p]
    p32 = C.POINTER(C.c_int32)
    h = lib.dgf_plan_create(mesh.h, part.ctypes.data_as(p32), rank, world)
    assert h
    sizes = np.zeros(6, dtype=np.int32)
    lib.dgf_plan_sizes(h, sizes.ctypes.data_as(p32))
    kown, kint, khalo, npeers, nsend, _ = (int(x) for x in sizes)
    l2g = np.zeros(kown + khalo, dtype=np.int32)
    peers = np.zeros(max(npeers, 1), dtype=np.int32)
    roff = np.zeros(npeers + 1, dtype=np.int32)
    soff = np.zeros(npeers + 1, dtype=np.int32)
    send = np.zeros(max(nsend, 1), dtype=np.int32)
    lib.dgf_plan_arrays(h, *(a.ctypes.data_as(p32) for a in (l2g, peers, roff, soff, send)))
    lib.dgf_plan_free(h)
    return dict(kown=kown, kint=kint, khalo=khalo, l2g=l2g, peers=peers[:npeers], roff=roff, soff=soff, send=send[:nsend])


def _worker(rank, world, port, out_dir, partitioner):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, str(ROOT))
    import __graft_entry__ as graft
    from oracle.oracle_py import Oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = graft.load_package()
    mesh = pkg.Mesh(pkg.Model.make_cube(3, -10.0, 10.0, 2), pkg.Config())
    mesh.set_physics(c0=343.0, rho0=1.225, v0=(30.0, 10.0, 0.0), dt=1e-4)
    part = np.zeros(mesh.K, dtype=np.int32)
    if partitioner == "metis":
        assert pkg.load_front().dgf_partition_metis(mesh.h, world, part.ctypes.data_as(C.POINTER(C.c_int32)), None) == 0
    else:
        assert pkg.load_front().dgf_partition_rcb(mesh.h, world, part.ctypes.data_as(C.POINTER(C.c_int32))) == 0
    plan = _plan(pkg, mesh, part, rank, world)
    Np, N = mesh.Np, mesh.N
    l2g, kown = plan["l2g"], plan["kown"]
    assert (part[l2g[:kown]] == rank).all() and (part[l2g[kown:]] != rank).all()
    assert len(set(l2g.tolist())) == len(l2g)
    # interior elements come first and touch no foreign element
    nbr = np.where(mesh.fNbrElId[mesh.elFId, 0] == np.arange(mesh.K)[:, None], mesh.fNbrElId[mesh.elFId, 1], mesh.fNbrElId[mesh.elFId, 0])
    for l in range(kown):
        foreign = [(n >= 0 and part[n] != rank) for n in nbr[l2g[l]]]
        assert any(foreign) == (l >= plan["kint"])

    truth = np.random.default_rng(7).standard_normal((4, N))
    orc = Oracle(mesh, threads=1)
    y_true = truth.copy()
    node = lambda els: (np.asarray(els)[:, None] * Np + np.arange(Np)[None, :]).reshape(-1)
    own_nodes, halo_nodes = node(l2g[:kown]), node(l2g[kown:])
    local = np.full((4, N), np.nan)
    local[:, own_nodes] = truth[:, own_nodes]
    local[:, halo_nodes] = truth[:, halo_nodes]
    for stage in range(3):
        rhs_true = orc.eval_rhs(Oracle.OPERATOR, y_true)
        rhs_loc = orc.eval_rhs(Oracle.OPERATOR, local)
        assert np.isfinite(rhs_loc[:, own_nodes]).all()  # owned elements depend on owned + halo data only
        np.testing.assert_allclose(rhs_loc[:, own_nodes], rhs_true[:, own_nodes], rtol=0, atol=1e-9 * np.abs(rhs_true).max())
        # next stage input on the owned elements, then the halo exchange of the engine (dgb_api.cu: exchangeHalo)
        y_true = y_true + 1e-6 * rhs_true
        local[:, own_nodes] = y_true[:, own_nodes]
        local[:, halo_nodes] = np.nan
        reqs, bufs = [], []
        for i, peer in enumerate(plan["peers"]):
            s_el = l2g[plan["send"][plan["soff"][i]:plan["soff"][i + 1]]]
            sbuf = torch.from_numpy(np.ascontiguousarray(local[:, node(s_el)]))
            rbuf = torch.empty((4, (plan["roff"][i + 1] - plan["roff"][i]) * Np), dtype=torch.float64)
            reqs += [dist.isend(sbuf, int(peer)), dist.irecv(rbuf, int(peer))]
            bufs.append((i, sbuf, rbuf))
        for r in reqs:
            r.wait()
        for i, _, rbuf in bufs:
            h_el = l2g[kown + plan["roff"][i]:kown + plan["roff"][i + 1]]
            local[:, node(h_el)] = rbuf.numpy()
        np.testing.assert_array_equal(local[:, halo_nodes], y_true[:, halo_nodes])
    np.save(Path(out_dir) / f"ok{rank}.npy", np.array([kown, plan["khalo"]]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,partitioner", [(2, "rcb"), (2, "metis"), (3, "metis")])
def test_halo_plan_over_gloo(tmp_path, world, partitioner):
    import torch.multiprocessing as mp

    mp.spawn(_worker, args=(world, 29731 + world + (7 if partitioner == "metis" else 0), str(tmp_path), partitioner), nprocs=world, join=True)
    sizes = [np.load(tmp_path / f"ok{r}.npy") for r in range(world)]
    assert sum(int(s[0]) for s in sizes) == 3 ** 3 * 6
    assert all(int(s[1]) > 0 for s in sizes)
