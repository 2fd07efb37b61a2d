#!/usr/bin/env python
"""bench.py — headline benchmark of the RK4 stage evaluation (BASELINE.json: DOF-updates/s per RK4 stage).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--cells n] [--order p]

* workload (BASELINE config 5): synthetic cube of n^3 x 6 Kuhn tetrahedra (n = 62 -> 1 429 968 tets), order 4
  (35 nodes / element, 50.05 M DG nodes, 200.2 M unknowns), absorbing boundary, c0 = 343, rho0 = 1.225,
  Gaussian pressure pulse, dt = 0.1 h_min / (c0 (2p+1)), RK4. One bench "step" = one RK4 time step = 4 stage
  launches over the whole mesh. DOF-updates/s = 4*K*Np unknowns x 4 stages x steps / time.
* `value`  : state resident in HBM, timed with CUDA events inside the engine (dgb_last_run_ms), max over ranks.
* `e2e`    : the user-facing call sequence through the C ABI from HOST buffers: dgb_set_state (pinned host ->
             device), dgb_run(K steps), dgb_get_state (device -> pinned host), timed with a host clock around
             the three synchronous calls, max over ranks.
* `roofline`: the stage kernel against the measured HBM roof (algorithmic bytes; `traffic` = DRAM bytes of the
             committed ncu capture of the same kernel) and, for comparison with the dense-operator kernels of round 1, against
             the FP64 roof (dense flops; the Bernstein kernels of the default path execute ~0.3x as many, see the `note`).
* `cpu_baseline` / `--impl reference`: the reference's OWN sources (oracle/_ref/dgalerkin_ref) on the host cores
             on a bounded sample of the same workload (falls back to the oracle's faithful mode, kind "port").
N > 1 (launched by torchrun): the same mesh is partitioned over the ranks (recursive coordinate bisection or METIS); the halo
exchange is fused into the stage kernel (peer-to-peer bulk stores over NVLink, `--exchange 0` = NCCL send/recv); "strong"
scaling; the line carries a `parity` object (a fixed sub-case run partitioned and on one GPU in the same process).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as graft  # noqa: E402

METRIC = "DOF-updates/s per RK4 stage"
UNIT = "DOF-updates/s"
FP64_PEAK_TFLOPS = 37.1  # profiles/microbench/r01_fp64_peaks_b200.txt (DMMA m8n8k4 sustained on this pool's B200)
HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return json.loads(p.read_text()), "measured"
        except Exception:
            pass
    return {"hbm_gbs": HBM_FALLBACK_GBS}, "fallback"


class ClockSampler:
    """Samples SM clocks / throttle reasons of one GPU while the timed region runs: NVML polled every few milliseconds from a
    thread (nvidia-ml-py), `nvidia-smi -lms` as the fallback."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.thread = index, [], None, None
        self.stop = threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll(self):
        n = self.nvml
        reasons = {"hw_slowdown": getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                   "hw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                   "sw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                   "sw_power_cap": getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self.stop.is_set():
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.rows.append((mhz, self.max_mhz, [k for k, bit in reasons.items() if mask & bit]))
            except Exception:
                pass
            time.sleep(0.003)

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            try:
                self.rows.append((float(r[0]), float(r[1]), [nm for k, nm in enumerate(names) if r[3 + k].lower().startswith("active")]))
            except Exception:
                continue

    def __enter__(self):
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        self.stop.set()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        if self.thread:
            self.thread.join(timeout=2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = [r[0] for r in self.rows]
        reasons = sorted({x for r in self.rows for x in r[2]})
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(r[1] for r in self.rows)), "reasons": reasons,
                "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def build_workload(pkg, cells, order, v0, dim=3):
    model = pkg.Model.make_cube(cells, -10.0, 10.0, order) if dim == 3 else pkg.Model.make_square(cells, -10.0, 10.0, order)
    cfg = pkg.Config()
    cfg.add_initial_condition(0.0, 0.0, 0.0, 1.0, 1.0)  # config.conf:57
    mesh = pkg.Mesh(model, cfg)
    if dim == 2:  # BASELINE config 2: refined square, reflecting walls
        mesh.fBC[np.nonzero(mesh.fIsBoundary)[0]] = 1
    c0, rho0 = 343.0, 1.225
    dt = 0.1 * mesh.h_min() / (c0 * (2 * order + 1))
    mesh.set_physics(c0=c0, rho0=rho0, v0=v0, dt=dt)
    return model, cfg, mesh


def measured_traffic(kernel_name, K):
    """DRAM bytes per launch of the stage kernel from the committed `ncu --set full` capture (profiles/r01_traffic.json:
    dram__bytes_read.sum + dram__bytes_write.sum of one launch, per element of the profiled mesh), scaled to K elements."""
    for name in ("r02_traffic.json", "r01_traffic.json"):  # newest capture that knows this kernel
        try:
            rec = json.loads((ROOT / "profiles" / name).read_text())[kernel_name]
            return rec["dram_bytes_per_element"] * K, dict(rec, file="profiles/" + name)
        except Exception:
            continue
    return None, None


def alg_counts(mesh, v0_zero):
    """Algorithmic bytes / flops of ONE stage launch over the whole mesh."""
    K, Np, Nfp = mesh.K, mesh.Np, mesh.Nfp
    d, Nf = mesh.desc.dim, mesh.desc.Nf
    # per element: d*d inverse-Jacobian entries + per face (normal, Fscale) + 2 int32 = 232 B for a tetrahedron
    bytes_stage = 34.0 * 4 * K * Np + (8.0 * d * d + 40.0 * Nf) * K
    # volume: 2d matvecs (div v, grad p), 4d with mean flow; lift: 4 fields x Nf faces; flux + pointwise
    flops_el = (4.0 if v0_zero else 8.0) * d * Np * Np + 8.0 * Nf * Np * Nfp + 40.0 * Nf * Nfp + 40.0 * Np
    return bytes_stage, flops_el * K


# ------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own sources on the host cores, bounded sample
# ------------------------------------------------------------------------------------------------------------
REF_CONF = """timeStart=0
timeEnd={tend}
timeStep={dt}
timeRate=1000
elementType=Lagrange
timeIntMethod=Runge-Kutta
numThreads={threads}
v0_x = {v0x}
v0_y = {v0y}
v0_z = {v0z}
rho0 = 1.225
c0 = 343
initialCondtition1 = gaussian, 0,0,0,1,1
saveFile=out
"""


def time_reference(pkg, order, v0, steps, sample_cells=8):
    """DOF-updates/s of the reference on a cube of sample_cells^3 x 6 tets at `order` (same generator, same physics).
    Set-up (the reference's O(F^2) connectivity) is removed by differencing a 1-step and a (1+steps)-step run."""
    ref_bin = ROOT / "oracle" / "_ref" / "dgalerkin_ref"
    cores = os.cpu_count() or 1
    model = pkg.Model.make_cube(sample_cells, -10.0, 10.0, 1)
    cfg = pkg.Config()
    mesh1 = pkg.Mesh(pkg.Model.make_cube(sample_cells, -10.0, 10.0, order), cfg)
    K, Np = mesh1.K, mesh1.Np
    dt = 0.1 * mesh1.h_min() / (343.0 * (2 * order + 1))
    sample = f"cube {sample_cells}^3x6 = {K} tets, order {order}, {steps} RK4 steps"
    if ref_bin.exists():
        with tempfile.TemporaryDirectory() as wd:
            msh = Path(wd) / "cube.msh"
            model.write_msh(msh)
            env = dict(os.environ, GMSHLITE_QUIET="1", GMSHLITE_ORDER=str(order), OMP_NUM_THREADS=str(cores))

            def run(nsteps):
                # the loop runs while t <= timeEnd: (nsteps - 0.5) * dt gives exactly nsteps iterations
                conf = Path(wd) / f"c{nsteps}.conf"
                conf.write_text(REF_CONF.format(tend=repr((nsteps - 0.5) * dt), dt=repr(dt), threads=cores, v0x=v0[0], v0y=v0[1], v0z=v0[2]))
                t0 = time.perf_counter()
                subprocess.run([str(ref_bin), str(msh), str(conf)], cwd=wd, env=env, check=True)
                return time.perf_counter() - t0

            t1 = run(1)
            if t1 > 45.0:  # slow host: keep the whole measurement within a few minutes
                steps = 1
            t2 = run(1 + steps)
        sec = max(t2 - t1, 1e-9)
        sample = f"cube {sample_cells}^3x6 = {K} tets, order {order}, {steps} RK4 steps"
        return {"value": 4.0 * K * Np * 4 * steps / sec, "unit": UNIT, "cores": cores, "kind": "reference", "steps": steps,
                "sample": sample + " (reference sources on Gmsh/Eigen stand-ins; set-up removed by differencing)", "seconds": sec}
    from oracle.oracle_py import Oracle
    mesh1.set_physics(c0=343.0, rho0=1.225, v0=v0, dt=dt)
    orc = Oracle(mesh1, threads=cores)
    u = mesh1.initial_condition() if cfg.c.nInit else np.zeros((4, mesh1.N))
    orc.run(Oracle.FAITHFUL, pkg.RUNGE_KUTTA, u, 0.0, 1)  # builds the per-element matrices
    t0 = time.perf_counter()
    orc.run(Oracle.FAITHFUL, pkg.RUNGE_KUTTA, u, 0.0, steps)
    sec = time.perf_counter() - t0
    return {"value": 4.0 * K * Np * 4 * steps / sec, "unit": UNIT, "cores": cores, "kind": "port", "steps": steps,
            "sample": sample + " (oracle, faithful mode)", "seconds": sec}


def multi_gpu_parity(pkg, torch, dist, args, rank, world):
    """Small fixed sub-case run partitioned over all ranks (same kernel, same exchange as the timed run) and, on rank 0, on one
    GPU: relative L2 difference per field of the merged state after 12 RK4 steps with a source running. Collective."""
    import ctypes as C
    n, steps = args.parity_cells, 12
    model = pkg.Model.make_cube(n, -10.0, 10.0, args.order)
    cfg = pkg.Config()
    cfg.add_initial_condition(1.0, -2.0, 0.5, 30.0, 1.0)
    cfg.add_source(2.0, 1.0, 0.0, 6.0, 10.0, 1500.0, 0.0, 1.0)
    mesh = pkg.Mesh(model, cfg)
    mesh.set_physics(c0=343.0, rho0=1.225, v0=args.v0, dt=0.1 * mesh.h_min() / (343.0 * (2 * args.order + 1)))
    b = np.nonzero(mesh.fIsBoundary)[0]
    mesh.fBC[b[::3]] = 1  # mixed absorbing / reflecting walls
    part = np.zeros(mesh.K, dtype=np.int32)
    fn = pkg.load_front().dgf_partition_metis if args.partitioner == "metis" else pkg.load_front().dgf_partition_rcb
    rc = fn(mesh.h, world, part.ctypes.data_as(C.POINTER(C.c_int32)), None) if args.partitioner == "metis" else fn(mesh.h, world, part.ctypes.data_as(C.POINTER(C.c_int32)))
    assert rc == 0
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt = torch.tensor(list(pkg.nccl_unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(idt, 0)
    eng = pkg.Engine(mesh, el_part=part, rank=rank, nranks=world, nccl_id=bytes(idt.cpu().numpy().tobytes()))
    if args.exchange is not None:
        eng.set_option("exchange", args.exchange)
    if args.kernel:
        eng.set_option("kernel", args.kernel)
    eng.set_sources_from_config()
    u0 = mesh.initial_condition()
    eng.set_state(u0)
    t_half = eng.run(pkg.RUNGE_KUTTA, 0.0, steps // 2)
    eng.run(pkg.RUNGE_KUTTA, t_half, steps - steps // 2)
    got = np.zeros((4, mesh.N))
    eng.get_state(got)
    owned = np.repeat(part == rank, mesh.Np)
    got[:, ~owned] = 0.0
    exchange, kernel = eng.get_option("exchange"), eng.kernel_name
    dist.barrier()
    eng.close()
    merged = torch.from_numpy(got).cuda()
    dist.all_reduce(merged)  # every DG node is owned by exactly one rank
    out = None
    if rank == 0:
        single = pkg.Engine(mesh, options={"kernel": args.kernel} if args.kernel else None)
        single.set_sources_from_config()
        single.set_state(u0)
        t_half = single.run(pkg.RUNGE_KUTTA, 0.0, steps // 2)
        single.run(pkg.RUNGE_KUTTA, t_half, steps - steps // 2)
        ref = single.get_state()
        single.close()
        m = merged.cpu().numpy()
        rel = [float(np.linalg.norm(m[q] - ref[q]) / max(np.linalg.norm(ref[q]), 1e-300)) for q in range(4)]
        out = {"rel_l2_vs_single": max(rel), "per_field": rel, "case": f"cube n={n} ({mesh.K} tets) order {args.order}, {steps} RK4 steps, source + mixed walls, "
               f"{world} ranks ({args.partitioner}) vs 1 GPU", "kernel": kernel, "exchange": exchange, "tolerance": 1e-12}
    dist.barrier()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=62)
    ap.add_argument("--order", type=int, default=4)
    ap.add_argument("--dim", type=int, default=3, choices=[2, 3], help="3: cube of Kuhn tetrahedra (config 5); 2: refined square of triangles, reflecting walls (config 2)")
    ap.add_argument("--v0", type=float, nargs=3, default=[0.0, 0.0, 0.0])
    ap.add_argument("--kernel", type=int, default=0, help="0 auto, 1 generic, 2 tiled DMMA, 3 warp-specialised DMMA, 4 / 5 Bernstein-Bezier (sparse operators, CUDA cores; 5 = face-sequential schedule), 6 Bernstein-Bezier second generation (TMA pipeline, interleaved layout)")
    ap.add_argument("--bb-tile", type=int, default=0, help="elements per CTA of the Bernstein-Bezier kernels (32, 16, 8; 0: the engine's default)")
    ap.add_argument("--partitioner", default="rcb", choices=["rcb", "metis"], help="element partition for --gpus > 1")
    ap.add_argument("--no-overlap", action="store_true")
    ap.add_argument("--overlap", type=int, default=None, help="halo exchange overlap mode 0/1/2 (default: the engine's)")
    ap.add_argument("--sm-reserve", type=int, default=None, help="SMs left to the halo-exchange kernels during overlapped launches")
    ap.add_argument("--exchange", type=int, default=None, help="halo exchange: 0 ncclSend/ncclRecv, 1 direct peer-to-peer stores, 2 stores fused into the stage kernel (default: the engine's)")
    ap.add_argument("--parity-cells", type=int, default=12, help="--gpus > 1: size of the partitioned-vs-single-GPU parity case reported beside the timing (0: skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=3, help="RK4 steps of the reference's bounded sample (set-up removed by differencing against a 1-step run)")
    ap.add_argument("--ref-cells", type=int, default=12, help="cube size n (n^3 x 6 tetrahedra) of the reference's bounded sample")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: whatever libraries print (e.g. the NCCL version banner) is sent to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    pkg = graft.load_package()
    workload = (f"cube n={args.cells} ({args.cells ** 3 * 6} tets) order {args.order} RK4" if args.dim == 3 else
                f"square n={args.cells} ({args.cells ** 2 * 2} triangles, reflecting walls) order {args.order} RK4")
    npn = {3: (args.order + 1) * (args.order + 2) * (args.order + 3) // 6, 2: (args.order + 1) * (args.order + 2) // 2}[args.dim]
    nel = args.cells ** 3 * 6 if args.dim == 3 else args.cells ** 2 * 2
    config = {"workload": workload, "cells": args.cells, "order": args.order, "v0": args.v0, "boundary": "absorbing" if args.dim == 3 else "reflecting",
              "l2": f"inputs larger than L2 (4 state arrays of {4 * nel * npn * 8 / 1e6:.0f} MB each, 126 MB of L2)", "partition": args.partitioner if world > 1 else "none"}

    if args.impl == "reference":
        if rank != 0:
            return
        # The reference's own CPU implementation on a BOUNDED sample of the workload: a cube of --ref-cells^3 x 6 tetrahedra of
        # the same order, physics and generator (the full 1.43 M-tetrahedron mesh would take the reference ~50 min per step
        # and its O(F^2) set-up never finishes). `steps` / `warmup` report what actually ran.
        t0 = time.perf_counter()
        res = time_reference(pkg, args.order, args.v0, args.cpu_steps, sample_cells=args.ref_cells)
        wall = time.perf_counter() - t0
        config = dict(config, sample=res["sample"], sample_cells=args.ref_cells,
                      note="throughput (DOF-updates/s) of the reference is size-independent beyond cache; the Gmsh stand-in's tetrahedron rule has "
                           "(p+1)^3 = 125 points at order 4 where Gmsh's own degree-8 rule has ~43: the reference spends ~2.9x more time in "
                           "getElStiffVector here than with real Gmsh")
        line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": res["steps"], "warmup": 1,
                "ms_per_step": res["seconds"] / res["steps"] * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0, "wall_s": wall}
        emit(line)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    model, cfg, mesh = build_workload(pkg, args.cells, args.order, args.v0, args.dim)
    K, Np = mesh.K, mesh.Np
    unknowns = 4 * K * Np
    if world > 1:
        import ctypes as C
        part = np.zeros(K, dtype=np.int32)
        if args.partitioner == "metis":
            assert pkg.load_front().dgf_partition_metis(mesh.h, world, part.ctypes.data_as(C.POINTER(C.c_int32)), None) == 0
        else:
            assert pkg.load_front().dgf_partition_rcb(mesh.h, world, part.ctypes.data_as(C.POINTER(C.c_int32))) == 0
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.tensor(list(pkg.nccl_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(idt, 0)
        eng = pkg.Engine(mesh, el_part=part, rank=rank, nranks=world, nccl_id=bytes(idt.cpu().numpy().tobytes()))
        if args.no_overlap:
            eng.set_option("overlap", 0)
        if args.overlap is not None:
            eng.set_option("overlap", args.overlap)
        if args.sm_reserve is not None:
            eng.set_option("sm_reserve", args.sm_reserve)
        if args.exchange is not None:
            eng.set_option("exchange", args.exchange)
    else:
        eng = pkg.Engine(mesh)
    if args.bb_tile:
        eng.set_option("bb_tile", args.bb_tile)
    if args.kernel:
        eng.set_option("kernel", args.kernel)
    if world > 1:
        config["exchange"] = {0: "nccl send/recv", 1: "direct peer-to-peer stores (3 launches per stage)",
                              2: "direct peer-to-peer stores fused into the stage kernel"}[eng.get_option("exchange")]

    # pinned host buffers for the end-to-end leg
    host_u = torch.empty((4, mesh.N), dtype=torch.float64, pin_memory=True)
    u_np = host_u.numpy()
    u_np[...] = mesh.initial_condition()
    t_sim = 0.0
    eng.set_state(u_np)
    barrier()
    t_sim = eng.run(pkg.RUNGE_KUTTA, t_sim, args.warmup)
    barrier()

    launches0 = eng.launch_count
    with ClockSampler(local_rank) as clk:
        barrier()
        t_sim = eng.run(pkg.RUNGE_KUTTA, t_sim, args.steps)
        ms = eng.last_run_ms
        stage_ms = eng.last_stage_kernel_ms
        barrier()
    launches = eng.launch_count - launches0
    ms = max_over_ranks(ms)
    stage_ms = max_over_ranks(stage_ms)
    value = unknowns * 4.0 * args.steps / (ms * 1e-3)
    finite = bool(np.isfinite(eng.get_state(u_np)).all())

    # end to end: host -> device, K steps, device -> host, through the public C ABI
    barrier()
    t0 = time.perf_counter()
    eng.set_state(u_np)
    eng.run(pkg.RUNGE_KUTTA, t_sim, args.steps)
    eng.get_state(u_np)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = unknowns * 4.0 * args.steps / e2e_s
    state_bytes = unknowns * 8

    parity = None
    if world > 1 and args.parity_cells > 0:
        parity = multi_gpu_parity(pkg, torch, dist, args, rank, world)

    fp64_now = eng.measure_fp64_tflops() if rank == 0 else 0.0  # same box, same run (after the timed regions)

    if rank == 0:
        peaks, how = measured_peaks()
        v0_zero = all(v == 0.0 for v in args.v0)
        bytes_stage, flops_stage = alg_counts(mesh, v0_zero)
        frac_work = 1.0  # the timed stage launch covers all elements on 1 GPU
        if world > 1:
            frac_work = 1.0 / world
        gbs = bytes_stage * frac_work / (stage_ms * 1e-3) / 1e9 if stage_ms > 0 else 0.0
        tfl = flops_stage * frac_work / (stage_ms * 1e-3) / 1e12 if stage_ms > 0 else 0.0
        hbm_peak = float(peaks.get("hbm_gbs", HBM_FALLBACK_GBS))
        traffic, traffic_rec = measured_traffic(eng.kernel_name, K * frac_work)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config, "kernel": eng.kernel_name, "finite": finite,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": state_bytes / args.steps, "d2h_bytes_per_step": state_bytes / args.steps,
                    "what": "dgb_set_state(pinned host) + dgb_run(steps) + dgb_get_state(pinned host), host clock"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "traffic": traffic,
                         "traffic_source": (traffic_rec or {}).get("source"),
                         "peak_source": how + " (MEASURED_PEAKS.json hbm_gbs)" if how == "measured" else "fallback 6.65 TB/s",
                         "stage_kernel_ms": stage_ms, "alg_bytes_per_launch": bytes_stage * frac_work,
                         "fp64": {"achieved": tfl, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": tfl / FP64_PEAK_TFLOPS,
                                  "alg_flops_per_launch": flops_stage * frac_work,
                                  "peak_source": "profiles/microbench/r01_fp64_peaks_b200.txt (DMMA m8n8k4 sustained)",
                                  "peak_measured_this_run": fp64_now, "peak_measured_how": "dgb_measure_fp64_tflops: DFMA, 16 chains x 1024 threads per SM, best of 3"}},
            "clocks": clk.summary(),
        }
        if eng.kernel_name.startswith(("stage_bb", "stage_bbe")):
            # the Bernstein kernels execute ~0.3x the flops of the dense nodal operators this figure counts: it is kept for comparison with
            # the DMMA kernels of round 1 and can exceed 1; HBM is the roof that binds these kernels
            line["roofline"]["fp64"]["note"] = ("alg_flops count the DENSE nodal operators (SURVEY.md §8 d3); this kernel applies sparse Bernstein-Bezier operators "
                                                "(~0.3x as many flops executed), so the fraction is a comparison with the dense kernels, not a pipe utilisation")
        if parity is not None:
            line["parity"] = parity
        if world == 1 and not args.no_cpu_baseline:
            res = time_reference(pkg, args.order, args.v0, args.cpu_steps, sample_cells=args.ref_cells)
            line["cpu_baseline"] = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
        emit(line)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
