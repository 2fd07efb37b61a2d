"""Shared-memory race check (and out-of-bounds check) of the emulated CUDA kernels on the CPU: the emulation of oracle/cuda_emu.h runs one OS thread per
CUDA thread with a pthread barrier for __syncthreads, so a ThreadSanitizer build of an emulation harness sees a missing barrier
as a data race on the shared-memory buffer — what compute-sanitizer's racecheck reports on the GPU. The check is itself
checked: a copy of stage_bb.cu with one barrier removed must be reported.

Needs g++ with libtsan (skipped otherwise)."""
import os
import shutil
import subprocess
import sys
from pathlib import Path

import pytest

from conftest import ROOT

WORKER = Path(__file__).resolve().parent / "racecheck_worker.py"


def _libtsan():
    try:
        p = subprocess.run(["gcc", "-print-file-name=libtsan.so"], capture_output=True, text=True, check=True).stdout.strip()
        return p if os.path.isabs(p) and Path(p).exists() else None
    except Exception:
        return None


def _libasan():
    try:
        p = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True, check=True).stdout.strip()
        return p if os.path.isabs(p) and Path(p).exists() else None
    except Exception:
        return None


TSAN = _libtsan()
ASAN = _libasan()
pytestmark = pytest.mark.skipif(TSAN is None, reason="libtsan not available")


def _cuda_include():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    return Path(nvcc).resolve().parent.parent / "include"


def _build(src_root, harness, out, sanitizer="thread"):
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-fopenmp", "-fPIC", "-shared", "-pthread", f"-fsanitize={sanitizer}", "-Wno-unknown-pragmas",
           f"-I{src_root / 'include'}", f"-I{_cuda_include()}", str(src_root / "oracle" / harness), "-o", str(out)]
    subprocess.run(cmd, check=True, timeout=900)


def _run(kind, lib):
    env = dict(os.environ, LD_PRELOAD=TSAN, TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 exitcode=0", OMP_NUM_THREADS="1")
    p = subprocess.run([sys.executable, str(WORKER), kind, str(lib)], env=env, capture_output=True, text=True, timeout=1200)
    assert "worker done" in p.stdout, p.stderr[-2000:]
    return p.stderr.count("WARNING: ThreadSanitizer: data race")


@pytest.mark.parametrize("kind,harness", [("bb", "bb_emulate.cpp"), ("curved", "curved_emulate.cpp"), ("generic", "generic_emulate.cpp"), ("bb2", "bb2_emulate.cpp")])
def test_no_shared_memory_race_in_the_emulated_kernels(tmp_path, kind, harness):
    lib = tmp_path / f"lib{kind}_tsan.so"
    _build(ROOT, harness, lib)
    assert _run(kind, lib) == 0


def test_the_race_check_sees_a_missing_barrier(tmp_path):
    """stage_bb.cu with the barrier between the face-input phase and the per-(element, field) phase removed."""
    root = tmp_path / "src"
    shutil.copytree(ROOT / "include", root / "include")
    shutil.copytree(ROOT / "dgfem-acoustic_b200" / "csrc", root / "dgfem-acoustic_b200" / "csrc")
    (root / "oracle").mkdir()
    for f in ("cuda_emu.h", "emu_layout.h", "bb_emulate.cpp"):
        shutil.copy(ROOT / "oracle" / f, root / "oracle" / f)
    cu = root / "dgfem-acoustic_b200" / "csrc" / "stage_bb.cu"
    text = cu.read_text()
    marker = "    __syncthreads();\n\n    // 3. one thread per (element, field): sparse Bernstein operators in registers"
    assert text.count(marker) == 1
    cu.write_text(text.replace(marker, "\n    // 3. one thread per (element, field): sparse Bernstein operators in registers"))
    lib = tmp_path / "libbb_broken_tsan.so"
    _build(root, "bb_emulate.cpp", lib)
    assert _run("bb", lib) > 0


def test_the_race_check_sees_a_missing_warp_barrier_in_stage_bb2(tmp_path):
    """stage_bb2.cu with the warp barrier between the wait for a face's traces (copied by OTHER lanes) and their use removed."""
    root = tmp_path / "src"
    shutil.copytree(ROOT / "include", root / "include")
    shutil.copytree(ROOT / "dgfem-acoustic_b200" / "csrc", root / "dgfem-acoustic_b200" / "csrc")
    (root / "oracle").mkdir()
    for f in ("cuda_emu.h", "emu_layout.h", "bb2_emulate.cpp"):
        shutil.copy(ROOT / "oracle" / f, root / "oracle" / f)
    cu = root / "dgfem-acoustic_b200" / "csrc" / "stage_bb2.cu"
    text = cu.read_text()
    marker = "may still travel)\n            __syncwarp();\n            double x[NFP];"
    assert text.count(marker) == 1
    cu.write_text(text.replace(marker, "may still travel)\n            double x[NFP];"))
    lib = tmp_path / "libbb2_broken_tsan.so"
    _build(root, "bb2_emulate.cpp", lib)
    assert _run("bb2", lib) > 0


@pytest.mark.skipif(ASAN is None, reason="libasan not available")
@pytest.mark.parametrize("kind,harness", [("bb", "bb_emulate.cpp"), ("curved", "curved_emulate.cpp"), ("generic", "generic_emulate.cpp"), ("bb2", "bb2_emulate.cpp")])
def test_no_out_of_bounds_access_in_the_emulated_kernels(tmp_path, kind, harness):
    """AddressSanitizer build: the emulation sizes the dynamic shared memory exactly and the global arrays are host vectors, so an
    index that runs off a tile, a state array or a table is reported (the CPU counterpart of compute-sanitizer memcheck)."""
    lib = tmp_path / f"lib{kind}_asan.so"
    _build(ROOT, harness, lib, sanitizer="address")
    env = dict(os.environ, LD_PRELOAD=ASAN, ASAN_OPTIONS="detect_leaks=0:halt_on_error=0", OMP_NUM_THREADS="1")
    p = subprocess.run([sys.executable, str(WORKER), kind, str(lib)], env=env, capture_output=True, text=True, timeout=1200)
    assert "worker done" in p.stdout, p.stderr[-2000:]
    assert "ERROR: AddressSanitizer" not in p.stderr, p.stderr[-3000:]
