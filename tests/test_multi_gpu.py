"""Domain decomposition across GPUs: partitioned run == single-GPU run, for every halo-exchange overlap
mode (0 none, 1 same stage, 2 next stage, -1 automatic), with and without mean flow, for every stage kernel. Needs >= 2 CUDA
devices; one process per GPU, NCCL halo exchange or direct peer-to-peer stores. (Driver-side evidence on a 1-GPU box: the
`parity` object of every multi-GPU bench.py line, and profiles/r02/multi_gpu_tests.log.)"""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _device_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


# kernel 0 = the automatic choice (the second-generation Bernstein kernel for tetrahedra of order >= 3: the halo carries
# interleaved coefficients; the CUDA-core kernel at order 2), 2 / 3 = the DMMA kernels, 4 = the first-generation Bernstein kernel
@pytest.mark.parametrize("world,cells,order,overlap,flow,kernel", [
    (2, 6, 4, 1, 1, 0), (2, 6, 4, 0, 1, 0), (2, 5, 2, 1, 1, 0), (2, 6, 4, 1, 0, 0), (2, 7, 4, 0, 0, 0), (2, 6, 4, 2, 0, 0), (2, 6, 4, 2, 1, 0),
    (2, 5, 2, 2, 1, 0), (2, 6, 4, -1, 0, 0), (2, 6, 3, 0, 1, 0),
    (2, 6, 4, 1, 1, 2), (2, 6, 4, 0, 0, 3), (2, 6, 4, 1, 0, 3), (2, 6, 4, 2, 0, 3), (2, 6, 4, 2, 1, 2), (2, 6, 4, 0, 0, 4), (2, 6, 3, 1, 1, 4)])
def test_partitioned_equals_single(tmp_path, world, cells, order, overlap, flow, kernel):
    _run_and_compare(tmp_path, world, cells, order, overlap, flow, 0, kernel)


@pytest.mark.parametrize("world,cells,order,overlap,flow,kernel", [
    (2, 6, 4, 0, 0, 0), (2, 6, 4, 1, 1, 0), (2, 5, 2, 1, 1, 0), (2, 7, 4, 0, 1, 0), (2, 6, 4, 0, 0, 3), (2, 6, 4, 1, 1, 2), (2, 6, 4, 0, 1, 4)])
def test_partitioned_equals_single_direct_exchange(tmp_path, world, cells, order, overlap, flow, kernel):
    """dgb_set_option("exchange", 1): stores into the peers' halo slots over NVLink + epoch flags (csrc/halo_p2p.cu)."""
    _run_and_compare(tmp_path, world, cells, order, overlap, flow, 1, kernel)


# Low orders and triangles: the automatic choice is the element-per-thread Bernstein kernel (triangles of orders 1 / 2, tetrahedra of order 1:
# direct stores or NCCL, never the fused exchange) or stage_bb2 on triangles (fused exchange); dim 2 = refined square.
@pytest.mark.parametrize("world,cells,order,overlap,flow,exchange,kernel,dim", [
    (2, 7, 1, 0, 1, 0, 0, 3), (2, 7, 1, 1, 0, 2, 0, 3), (2, 24, 1, 0, 1, 2, 0, 2), (2, 24, 2, 1, 0, 0, 0, 2), (2, 24, 2, 0, 1, 1, 7, 2),
    (2, 20, 3, 0, 1, 2, 0, 2), (2, 16, 4, 1, 0, 0, 0, 2), (2, 12, 6, 0, 1, 2, 0, 2), (2, 7, 1, 0, 1, 2, 6, 3)])
def test_partitioned_equals_single_low_orders_and_triangles(tmp_path, world, cells, order, overlap, flow, exchange, kernel, dim):
    _run_and_compare(tmp_path, world, cells, order, overlap, flow, exchange, kernel, dim=dim)


def _run_and_compare(tmp_path, world, cells, order, overlap, flow, exchange, kernel=0, tol=1e-12, dim=3):
    """The single-GPU run on rank 0 uses the same kernel as the partitioned run: the two must agree to rounding."""
    if _device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    worker = Path(__file__).resolve().parent / "multi_gpu_worker.py"
    steps = 12
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29000 + overlap * 7 + order + 20 * flow + cells + 100 * exchange + 200 * kernel + 1500 * (dim == 2)), str(worker), str(tmp_path),
           str(cells), str(order), str(steps), str(overlap), str(flow), str(exchange), str(kernel), str(dim)]
    subprocess.run(cmd, check=True, timeout=600)
    single = np.load(tmp_path / "single.npz")
    merged = np.zeros_like(single["u"])
    covered = np.zeros(merged.shape[1], dtype=int)
    probes = np.zeros_like(single["probes"])
    receivers = np.zeros_like(single["receivers"])
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        merged[:, z["owned"]] = z["u"][:, z["owned"]]
        covered += z["owned"]
        probes += z["probes"]  # probes owned by another rank are returned as zeros
        receivers += z["receivers"]  # likewise the receivers (at least one lies on a non-zero rank)
        assert z["launches"] > 0
    assert (covered == 1).all()
    for q in range(4):
        assert rel_l2(merged[q], single["u"][q]) < tol
        assert rel_l2(probes[:, :, q], single["probes"][:, :, q]) < tol
        assert rel_l2(receivers[:, :, q], single["receivers"][:, :, q]) < tol
