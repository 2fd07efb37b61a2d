"""ctypes binding of oracle/liboracle.so — the CPU restatement of the reference's hot path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs. Nothing in the product package (dgfem-acoustic_b200/) imports this module.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

ORACLE_DIR = Path(__file__).resolve().parent

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
_oracle = None


def _as(ptr_type, arr):
    return arr.ctypes.data_as(ptr_type)


def load_oracle():
    global _oracle
    if _oracle is not None:
        return _oracle
    lib = C.CDLL(str(ORACLE_DIR / "liboracle.so"))
    lib.orc_last_error.restype = C.c_char_p
    lib.orc_create.restype = C.c_void_p
    lib.orc_create.argtypes = [C.c_void_p, C.c_int]
    lib.orc_destroy.argtypes = [C.c_void_p]
    lib.orc_set_sources.argtypes = [C.c_void_p, C.c_int, c_int32_p, c_int32_p, c_double_p, c_double_p, c_double_p, c_double_p]
    lib.orc_run.argtypes = [C.c_void_p, C.c_int, C.c_int, c_double_p, C.c_double, C.c_int, C.c_int, c_int32_p, c_double_p, c_double_p]
    lib.orc_create.restype = C.c_void_p
    lib.orc_eval_rhs.argtypes = [C.c_void_p, C.c_int, c_double_p, c_double_p]
    lib.orc_set_receivers.argtypes = [C.c_void_p, C.c_int, c_int32_p, c_double_p]
    lib.orc_get_receivers.argtypes = [C.c_void_p, c_double_p, C.c_int]
    lib.orc_get_operators.argtypes = [C.c_void_p, c_double_p, c_double_p]
    _oracle = lib
    return lib


class Oracle:
    """CPU restatement of the reference's hot path. mode 0 = faithful loops, 1 = operator form."""

    FAITHFUL, OPERATOR = 0, 1

    def __init__(self, mesh, threads: int = 0):
        self.lib = load_oracle()
        self.mesh = mesh
        self.N = mesh.N
        self.h = self.lib.orc_create(C.cast(mesh.desc_p, C.c_void_p), int(threads))
        if not self.h:
            raise RuntimeError("oracle: " + self.lib.orc_last_error().decode())
        self.nprobe = 0
        self.nrcv = 0

    def set_sources_from_config(self):
        src = self.mesh.cfg.sources
        if not src:
            return
        offsets, idx = self.mesh.source_nodes()
        s = np.array(src)
        a, f, p, d = (np.ascontiguousarray(s[:, k], dtype=np.float64) for k in (5, 6, 7, 8))
        self.lib.orc_set_sources(self.h, len(a), _as(c_int32_p, offsets), _as(c_int32_p, idx), _as(c_double_p, a),
                                 _as(c_double_p, f), _as(c_double_p, p), _as(c_double_p, d))

    def set_receivers(self, el, weights):
        """Receivers: value = sum_n weights[j][n] * u[q][el[j]*Np + n], recorded at the start of every step."""
        el = np.ascontiguousarray(el, dtype=np.int32)
        w = np.ascontiguousarray(weights, dtype=np.float64)
        assert w.shape == (len(el), self.mesh.Np)
        self.nrcv = len(el)
        self.lib.orc_set_receivers(self.h, len(el), _as(c_int32_p, el), _as(c_double_p, w))

    def get_receivers(self, capacity_steps):
        out = np.zeros((max(capacity_steps, 1), max(self.nrcv, 1), 4), dtype=np.float64)
        n = self.lib.orc_get_receivers(self.h, _as(c_double_p, out), int(capacity_steps))
        return out[:n, : self.nrcv]

    def run(self, mode, integrator, u, t_start, nsteps, probes=None):
        """Advances u in place; returns (t_end, probe record [nsteps][nprobe][4])."""
        assert u.dtype == np.float64 and u.flags.c_contiguous and u.size == 4 * self.N
        t_end = C.c_double(0.0)
        if probes is None:
            probes = np.zeros(0, dtype=np.int32)
        probes = np.ascontiguousarray(probes, dtype=np.int32)
        rec = np.zeros((max(nsteps, 1), max(len(probes), 1), 4), dtype=np.float64)
        rc = self.lib.orc_run(self.h, int(mode), int(integrator), _as(c_double_p, u), float(t_start), int(nsteps),
                              len(probes), _as(c_int32_p, probes), _as(c_double_p, rec), C.byref(t_end))
        if rc != 0:
            raise RuntimeError("oracle: " + self.lib.orc_last_error().decode())
        return t_end.value, rec[:nsteps, : len(probes)]

    def eval_rhs(self, mode, u):
        u = np.ascontiguousarray(u, dtype=np.float64)
        rhs = np.zeros((4, self.N), dtype=np.float64)
        rc = self.lib.orc_eval_rhs(self.h, int(mode), _as(c_double_p, u), _as(c_double_p, rhs))
        if rc != 0:
            raise RuntimeError("oracle: " + self.lib.orc_last_error().decode())
        return rhs

    def operators(self):
        """(Dw [dim][Np][Np], LIFT [Np][Nf*Nfp]) as the oracle's operator mode builds them."""
        m = self.mesh
        dw = np.zeros((m.dim, m.Np, m.Np))
        lift = np.zeros((m.Np, m.Nf * m.Nfp))
        self.lib.orc_get_operators(self.h, _as(c_double_p, dw), _as(c_double_p, lift))
        return dw, lift

    def __del__(self):
        try:
            if self.h:
                self.lib.orc_destroy(self.h)
                self.h = None
        except Exception:
            pass
