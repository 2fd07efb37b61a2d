"""Worker of tests/test_multi_gpu.py: one process per GPU (launched with torch.distributed.run).

Every rank builds the same mesh, partitions it with RCB, runs RK4 through dgb_create_partitioned / dgb_run and
writes the elements it owns; rank 0 also runs the single-GPU engine as the reference result.
"""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as graft  # noqa: E402


def main():
    import torch
    import torch.distributed as dist

    out_dir, cells, order, steps, overlap = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    flow = int(sys.argv[6]) if len(sys.argv) > 6 else 1  # 0: zero mean flow (warp-specialised kernel at orders 3, 4)
    exchange = int(sys.argv[7]) if len(sys.argv) > 7 else 0  # 1: direct peer-to-peer stores instead of ncclSend/ncclRecv
    kernel = int(sys.argv[8]) if len(sys.argv) > 8 else 0      # 4: Bernstein-Bezier kernel (every rank switches)
    dim = int(sys.argv[9]) if len(sys.argv) > 9 else 3          # 2: refined square of triangles instead of the cube of tetrahedra
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = graft.load_package()
    model = pkg.Model.make_cube(cells, -10.0, 10.0, order) if dim == 3 else pkg.Model.make_square(cells, -10.0, 10.0, order)
    zs = 1.0 if dim == 3 else 0.0  # the square lies in z = 0
    cfg = pkg.Config()
    cfg.add_initial_condition(1.0, -2.0, 0.5 * zs, 30.0, 1.0)
    cfg.add_source(2.0, 1.0, 0.0, 6.0, 10.0, 1500.0, 0.0, 1.0)
    mesh = pkg.Mesh(model, cfg)
    mesh.set_physics(c0=343.0, rho0=1.225, v0=(30.0, 10.0, 0.0) if flow else (0.0, 0.0, 0.0), dt=0.1 * mesh.h_min() / (343.0 * (2 * order + 1)))
    b = np.nonzero(mesh.fIsBoundary)[0]
    mesh.fBC[b[::3]] = 1
    part = np.zeros(mesh.K, dtype=np.int32)
    assert pkg.load_front().dgf_partition_rcb(mesh.h, world, part.ctypes.data_as(C.POINTER(C.c_int32))) == 0
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt = torch.tensor(list(pkg.nccl_unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(idt, 0)
    probes = np.array([mesh.nearest_node(0, 0, 0), mesh.nearest_node(5, 5, 5 * zs), mesh.nearest_node(-7, 3, -2 * zs)], dtype=np.int32)
    u0 = mesh.initial_condition()
    eng = pkg.Engine(mesh, el_part=part, rank=rank, nranks=world, nccl_id=idt.cpu().numpy().tobytes(), options={"overlap": overlap})
    if kernel:
        eng.set_option("kernel", kernel)
    if exchange:
        eng.set_option("p2p_timeout_ms", 5000)
        eng.set_option("exchange", exchange)  # collective
    eng.set_sources_from_config()
    eng.set_probes(probes)
    # interpolated receivers in different parts of the cube: with 2 ranks at least one lies in an element of rank 1
    r_el, r_w = mesh.locate_receivers([(-6.3, 2.2, 1.1 * zs), (6.1, -3.3, 0.4 * zs), (0.2, 7.7, -5.1 * zs), (1.3, -8.2, 6.6 * zs)])
    assert len(set(int(part[e]) for e in r_el)) > 1
    eng.set_receivers(r_el, r_w)
    pinned = cells % 2 == 1  # odd sizes: page-locked caller buffers (the GPU gathers / scatters them over PCIe), even: pageable (staged)
    if pinned:
        hu, hg = torch.empty((4, mesh.N), dtype=torch.float64, pin_memory=True), torch.empty((4, mesh.N), dtype=torch.float64, pin_memory=True)
        hu.numpy()[...] = u0
        eng.set_state(hu.numpy())
    else:
        eng.set_state(u0)
    t_half = eng.run(pkg.RUNGE_KUTTA, 0.0, steps // 2)
    eng.run(pkg.RUNGE_KUTTA, t_half, steps - steps // 2)
    got = hg.numpy() if pinned else np.empty((4, mesh.N))
    got[...] = np.nan
    eng.get_state(got)
    got = np.array(got)
    rec = eng.get_probes(steps)
    rcv = eng.get_receivers(steps)
    owned = np.repeat(part == rank, mesh.Np)
    np.savez(Path(out_dir) / f"rank{rank}.npz", u=got, owned=owned, probes=rec, receivers=rcv, launches=eng.launch_count)
    dist.barrier()
    eng.close()
    if rank == 0:
        single = pkg.Engine(mesh, options={"kernel": kernel} if kernel else None)
        single.set_sources_from_config()
        single.set_probes(probes)
        single.set_receivers(r_el, r_w)
        single.set_state(u0)
        t_half = single.run(pkg.RUNGE_KUTTA, 0.0, steps // 2)
        single.run(pkg.RUNGE_KUTTA, t_half, steps - steps // 2)
        np.savez(Path(out_dir) / "single.npz", u=single.get_state(), probes=single.get_probes(steps), receivers=single.get_receivers(steps))
        single.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
