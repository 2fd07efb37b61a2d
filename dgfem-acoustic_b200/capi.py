"""ctypes bindings of the two product C-ABI libraries (used by tests/, bench.py and __graft_entry__.py).

* ``lib/libdgfront.so``  host front end (include/dgfront.h): MSH reader, config parser, Mesh set-up
* ``lib/libdgb.so``      the CUDA engine (include/dgb.h) — the product; fails loudly without a GPU

The CPU oracle is NOT reachable from here; its binding lives under ``oracle/`` (test infrastructure).

Python is plumbing here: it never computes anything on the hot path.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
REPO_ROOT = PKG_DIR.parent
LIB_DIR = PKG_DIR / "lib"

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_uint8_p = C.POINTER(C.c_uint8)


class DgbDesc(C.Structure):
    """struct dgb_desc (include/dgb.h)."""

    _fields_ = [
        ("dim", C.c_int32), ("order", C.c_int32), ("Np", C.c_int32), ("Nfp", C.c_int32), ("Nf", C.c_int32),
        ("K", C.c_int32), ("F", C.c_int32), ("nG", C.c_int32), ("nGf", C.c_int32),
        ("nGeomEl", C.c_int32), ("nGeomF", C.c_int32), ("fc", C.c_int32),
        ("elBasisFct", c_double_p), ("elUGradBasisFct", c_double_p), ("elWeight", c_double_p),
        ("fBasisFct", c_double_p), ("fWeight", c_double_p),
        ("elJacobian", c_double_p), ("elJacobianDet", c_double_p), ("fNormal", c_double_p), ("fJacobianDet", c_double_p),
        ("elFId", c_int32_p), ("elFOrientation", c_int32_p), ("fNbrElId", c_int32_p), ("fNToElNId", c_int32_p),
        ("fIsBoundary", c_uint8_p), ("fBC", c_int32_p),
        ("c0", C.c_double), ("rho0", C.c_double), ("v0", C.c_double * 3), ("dt", C.c_double),
    ]


class DgfConfig(C.Structure):
    """struct dgf_config (include/dgfront.h)."""

    _fields_ = [
        ("timeStart", C.c_double), ("timeEnd", C.c_double), ("timeStep", C.c_double), ("timeRate", C.c_double),
        ("elementType", C.c_char * 64), ("timeIntMethod", C.c_char * 64), ("saveFile", C.c_char * 512),
        ("numThreads", C.c_int32),
        ("v0", C.c_double * 3), ("rho0", C.c_double), ("c0", C.c_double),
        ("nSources", C.c_int32), ("sources", (C.c_double * 9) * 64),
        ("nInit", C.c_int32), ("initConditions", (C.c_double * 6) * 32),
        ("nPhysBC", C.c_int32), ("physBCTag", C.c_int32 * 64), ("physBCType", C.c_int32 * 64),
        ("nReceivers", C.c_int32), ("receivers", (C.c_double * 3) * 64), ("receiverFile", C.c_char * 512), ("receiverWav", C.c_char * 512),
    ]


EULER1, RUNGE_KUTTA = 0, 1

_front = None
_dgb = None


def _as(ptr_type, arr):
    return arr.ctypes.data_as(ptr_type)


# ------------------------------------------------------------------------------------------------
# front end
# ------------------------------------------------------------------------------------------------
def load_front():
    global _front
    if _front is not None:
        return _front
    lib = C.CDLL(str(LIB_DIR / "libdgfront.so"))
    lib.dgf_last_error.restype = C.c_char_p
    lib.dgf_open_msh.restype = C.c_void_p
    lib.dgf_open_msh.argtypes = [C.c_char_p, C.c_int]
    lib.dgf_make_cube.restype = C.c_void_p
    lib.dgf_make_cube.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int]
    lib.dgf_make_square.restype = C.c_void_p
    lib.dgf_make_square.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int]
    lib.dgf_model_free.argtypes = [C.c_void_p]
    lib.dgf_warp_model.argtypes = [C.c_void_p, C.c_double, C.c_double]
    lib.dgf_warp_model_local.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double]
    lib.dgf_write_msh.argtypes = [C.c_void_p, C.c_char_p]
    lib.dgf_model_dimension.argtypes = [C.c_void_p]
    lib.dgf_parse_config.argtypes = [C.c_char_p, C.c_void_p, C.POINTER(DgfConfig)]
    lib.dgf_default_config.argtypes = [C.POINTER(DgfConfig)]
    lib.dgf_mesh_build.restype = C.c_void_p
    lib.dgf_mesh_build.argtypes = [C.c_void_p, C.POINTER(DgfConfig)]
    lib.dgf_mesh_free.argtypes = [C.c_void_p]
    lib.dgf_mesh_desc.restype = C.POINTER(DgbDesc)
    lib.dgf_mesh_desc.argtypes = [C.c_void_p]
    lib.dgf_mesh_node_coords.restype = c_double_p
    lib.dgf_mesh_node_coords.argtypes = [C.c_void_p]
    lib.dgf_mesh_el_tags.restype = c_int32_p
    lib.dgf_mesh_el_tags.argtypes = [C.c_void_p]
    lib.dgf_mesh_el_node_tags.restype = c_int32_p
    lib.dgf_mesh_el_node_tags.argtypes = [C.c_void_p]
    lib.dgf_mesh_face_nodes.restype = c_int32_p
    lib.dgf_mesh_face_nodes.argtypes = [C.c_void_p]
    lib.dgf_initial_condition.argtypes = [C.c_void_p, C.POINTER(DgfConfig), c_double_p]
    lib.dgf_source_nodes.argtypes = [C.c_void_p, C.POINTER(DgfConfig), c_int32_p, c_int32_p]
    lib.dgf_time_loop.argtypes = [C.POINTER(DgfConfig), c_int32_p, C.c_int, C.POINTER(C.c_int)]
    lib.dgf_nearest_node.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
    lib.dgf_locate_point.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, c_double_p, c_double_p, C.POINTER(C.c_int)]
    lib.dgf_write_receivers.argtypes = [C.c_char_p, C.c_int, c_double_p, C.c_int, C.c_double, C.c_double, c_double_p]
    lib.dgf_write_wav.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, c_double_p]
    lib.dgf_write_views.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.POINTER(DgfConfig), C.c_int, c_int32_p,
                                    c_double_p, c_double_p]
    _front = lib
    return lib


class FrontError(RuntimeError):
    pass


class Config:
    """Config of the reference (include/configParser.h), parsed by the front end or built in code."""

    def __init__(self, c: DgfConfig | None = None):
        self.c = c if c is not None else DgfConfig()
        if c is None:
            load_front().dgf_default_config(C.byref(self.c))

    @property
    def sources(self):
        return [list(self.c.sources[i]) for i in range(self.c.nSources)]

    def add_source(self, x, y, z, size, amp, freq, phase, duration, pole=0.0):
        i = self.c.nSources
        for k, v in enumerate([pole, x, y, z, size, amp, freq, phase, duration]):
            self.c.sources[i][k] = v
        self.c.nSources = i + 1

    @property
    def receivers(self):
        return [list(self.c.receivers[i]) for i in range(self.c.nReceivers)]

    def add_receiver(self, x, y, z):
        i = self.c.nReceivers
        for k, v in enumerate([x, y, z]):
            self.c.receivers[i][k] = v
        self.c.nReceivers = i + 1

    def add_initial_condition(self, x, y, z, size, amp):
        i = self.c.nInit
        for k, v in enumerate([0.0, x, y, z, size, amp]):
            self.c.initConditions[i][k] = v
        self.c.nInit = i + 1

    def set_bc(self, phys_tag: int, reflecting: bool):
        i = self.c.nPhysBC
        self.c.physBCTag[i] = phys_tag
        self.c.physBCType[i] = 1 if reflecting else 0
        self.c.nPhysBC = i + 1

    def time_loop(self):
        """(number of steps, snapshot step indices) of the reference's FP-accumulating loop header."""
        lib = load_front()
        n = C.c_int(0)
        steps = lib.dgf_time_loop(C.byref(self.c), None, 0, C.byref(n))
        snaps = np.zeros(max(n.value, 1), dtype=np.int32)
        lib.dgf_time_loop(C.byref(self.c), _as(c_int32_p, snaps), n.value, C.byref(n))
        return steps, snaps[: n.value]


class Model:
    def __init__(self, handle):
        if not handle:
            raise FrontError(load_front().dgf_last_error().decode())
        self.h = handle

    @classmethod
    def open_msh(cls, path, order=1):
        return cls(load_front().dgf_open_msh(str(path).encode(), int(order)))

    @classmethod
    def make_cube(cls, n, lo=-10.0, hi=10.0, order=1):
        return cls(load_front().dgf_make_cube(int(n), float(lo), float(hi), int(order)))

    @classmethod
    def make_square(cls, n, lo=-10.0, hi=10.0, order=1):
        return cls(load_front().dgf_make_square(int(n), float(lo), float(hi), int(order)))

    def warp(self, amp, k):
        """Curved stand-in geometry: every node moves by a smooth field (dgf_warp_model)."""
        load_front().dgf_warp_model(self.h, float(amp), float(k))
        return self

    def warp_local(self, amp, k, center, radius):
        """A curved patch inside the ball (center, radius) of an otherwise straight-sided mesh (dgf_warp_model_local)."""
        load_front().dgf_warp_model_local(self.h, float(amp), float(k), float(center[0]), float(center[1]), float(center[2]), float(radius))
        return self

    @property
    def dimension(self):
        return load_front().dgf_model_dimension(self.h)

    def write_msh(self, path):
        if load_front().dgf_write_msh(self.h, str(path).encode()) != 0:
            raise FrontError(load_front().dgf_last_error().decode())

    def parse_config(self, path) -> Config:
        c = DgfConfig()
        if load_front().dgf_parse_config(str(path).encode(), self.h, C.byref(c)) != 0:
            raise FrontError(load_front().dgf_last_error().decode())
        return Config(c)

    def __del__(self):
        try:
            if self.h:
                load_front().dgf_model_free(self.h)
                self.h = None
        except Exception:
            pass


def _view(ptr, shape, dtype):
    n = int(np.prod(shape))
    if n == 0:
        return np.zeros(shape, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).reshape(shape)


class Mesh:
    """The reference's Mesh object as plain arrays (numpy views onto the front end's memory)."""

    def __init__(self, model: Model, cfg: Config):
        lib = load_front()
        self.model, self.cfg = model, cfg
        self.h = lib.dgf_mesh_build(model.h, C.byref(cfg.c))
        if not self.h:
            raise FrontError(lib.dgf_last_error().decode())
        self.desc_p = lib.dgf_mesh_desc(self.h)
        d = self.desc = self.desc_p.contents
        self.dim, self.order, self.Np, self.Nfp, self.Nf, self.K, self.F = d.dim, d.order, d.Np, d.Nfp, d.Nf, d.K, d.F
        self.N = self.K * self.Np
        self.node_coords = _view(lib.dgf_mesh_node_coords(self.h), (self.N, 3), np.float64)
        self.el_tags = _view(lib.dgf_mesh_el_tags(self.h), (self.K,), np.int32)
        self.el_node_tags = _view(lib.dgf_mesh_el_node_tags(self.h), (self.K, self.Np), np.int32)
        self.face_nodes = _view(lib.dgf_mesh_face_nodes(self.h), (self.Nf, self.Nfp), np.int32)
        self.elFId = _view(d.elFId, (self.K, self.Nf), np.int32)
        self.elFOrientation = _view(d.elFOrientation, (self.K, self.Nf), np.int32)
        self.fNbrElId = _view(d.fNbrElId, (self.F, 2), np.int32)
        self.fNToElNId = _view(d.fNToElNId, (self.F, self.Nfp, 2), np.int32)
        self.fIsBoundary = _view(d.fIsBoundary, (self.F,), np.uint8)
        self.fBC = _view(d.fBC, (self.F,), np.int32)
        self.fNormal = _view(d.fNormal, (self.F, d.nGeomF, 3), np.float64)
        self.fJacobianDet = _view(d.fJacobianDet, (self.F, d.nGeomF), np.float64)
        self.elJacobian = _view(d.elJacobian, (self.K, d.nGeomEl, 9), np.float64)
        self.elJacobianDet = _view(d.elJacobianDet, (self.K, d.nGeomEl), np.float64)
        self.elBasisFct = _view(d.elBasisFct, (d.nG, self.Np), np.float64)
        self.elUGradBasisFct = _view(d.elUGradBasisFct, (d.nG, self.Np, 3), np.float64)
        self.elWeight = _view(d.elWeight, (d.nG,), np.float64)
        self.fBasisFct = _view(d.fBasisFct, (d.nGf, self.Nfp), np.float64)
        self.fWeight = _view(d.fWeight, (d.nGf,), np.float64)

    def set_physics(self, c0=None, rho0=None, v0=None, dt=None):
        d = self.desc
        if c0 is not None:
            d.c0 = self.cfg.c.c0 = float(c0)
        if rho0 is not None:
            d.rho0 = self.cfg.c.rho0 = float(rho0)
        if v0 is not None:
            for k in range(3):
                d.v0[k] = self.cfg.c.v0[k] = float(v0[k])
        if dt is not None:
            d.dt = self.cfg.c.timeStep = float(dt)

    def initial_condition(self) -> np.ndarray:
        u = np.zeros((4, self.N), dtype=np.float64)
        load_front().dgf_initial_condition(self.h, C.byref(self.cfg.c), _as(c_double_p, u))
        return u

    def source_nodes(self):
        lib = load_front()
        ns = self.cfg.c.nSources
        offsets = np.zeros(ns + 1, dtype=np.int32)
        total = lib.dgf_source_nodes(self.h, C.byref(self.cfg.c), _as(c_int32_p, offsets), None)
        idx = np.zeros(max(total, 1), dtype=np.int32)
        lib.dgf_source_nodes(self.h, C.byref(self.cfg.c), _as(c_int32_p, offsets), _as(c_int32_p, idx))
        return offsets, idx[:total]

    def nearest_node(self, x, y, z) -> int:
        return load_front().dgf_nearest_node(self.h, float(x), float(y), float(z))

    def locate_point(self, x, y, z):
        """(element id, Lagrange weights [Np], parametric coordinates [3], outside flag) of a point (receivers)."""
        w = np.zeros(self.Np, dtype=np.float64)
        uvw = np.zeros(3, dtype=np.float64)
        outside = C.c_int(0)
        el = load_front().dgf_locate_point(self.h, float(x), float(y), float(z), _as(c_double_p, w), _as(c_double_p, uvw), C.byref(outside))
        if el < 0:
            raise FrontError(load_front().dgf_last_error().decode())
        return el, w, uvw, bool(outside.value)

    def locate_receivers(self, points):
        """(el [n], weights [n][Np]) for dgb_set_receivers; points outside the mesh raise."""
        els, ws = [], []
        for pt in points:
            el, w, _, outside = self.locate_point(*pt)
            if outside:
                raise FrontError(f"receiver {tuple(pt)} lies outside the mesh")
            els.append(el)
            ws.append(w)
        return np.array(els, dtype=np.int32), np.array(ws, dtype=np.float64).reshape(len(els), self.Np)

    def h_min(self) -> float:
        """Smallest inscribed-sphere-like length d*|el|/|faces| (used to pick a CFL-stable dt in tests/bench)."""
        det_e = np.abs(self.elJacobianDet[:, 0])
        det_f = self.fJacobianDet[:, 0]
        ratio = det_f[self.elFId] / det_e[:, None]  # Fscale
        return float(1.0 / ratio.max())

    def __del__(self):
        try:
            if self.h:
                load_front().dgf_mesh_free(self.h)
                self.h = None
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------
# CUDA engine
# ------------------------------------------------------------------------------------------------
class DgbError(RuntimeError):
    pass


def load_dgb():
    """Loads lib/libdgb.so. Raises if it has not been built — there is no fallback."""
    global _dgb
    if _dgb is not None:
        return _dgb
    path = Path(os.environ["DGB_LIB"]) if os.environ.get("DGB_LIB") else LIB_DIR / "libdgb.so"  # DGB_LIB: development builds (profiles/build_variant.sh)
    if not path.exists():
        raise DgbError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` first "
                       "(the CUDA engine has no CPU fallback)")
    lib = C.CDLL(str(path))
    lib.dgb_last_error.restype = C.c_char_p
    lib.dgb_version.restype = C.c_char_p
    lib.dgb_kernel_name.restype = C.c_char_p
    lib.dgb_kernel_name.argtypes = [C.c_void_p]
    lib.dgb_create.argtypes = [C.POINTER(DgbDesc), C.POINTER(C.c_void_p)]
    lib.dgb_create_partitioned.argtypes = [C.POINTER(DgbDesc), c_int32_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
    lib.dgb_nccl_unique_id.argtypes = [C.c_void_p]
    lib.dgb_destroy.argtypes = [C.c_void_p]
    lib.dgb_set_state.argtypes = [C.c_void_p, c_double_p]
    lib.dgb_get_state.argtypes = [C.c_void_p, c_double_p]
    lib.dgb_set_sources.argtypes = [C.c_void_p, C.c_int, c_int32_p, c_int32_p, c_double_p, c_double_p, c_double_p, c_double_p]
    lib.dgb_set_probes.argtypes = [C.c_void_p, C.c_int, c_int32_p]
    lib.dgb_get_probes.argtypes = [C.c_void_p, c_double_p, C.c_int, C.POINTER(C.c_int)]
    lib.dgb_set_receivers.argtypes = [C.c_void_p, C.c_int, c_int32_p, c_double_p]
    lib.dgb_get_receivers.argtypes = [C.c_void_p, c_double_p, C.c_int, C.POINTER(C.c_int)]
    lib.dgb_snapshot_begin.argtypes = [C.c_void_p, c_double_p]
    lib.dgb_snapshot_end.argtypes = [C.c_void_p]
    lib.dgb_host_alloc.restype = C.c_void_p
    lib.dgb_host_alloc.argtypes = [C.c_uint64]
    lib.dgb_host_free.argtypes = [C.c_void_p]
    lib.dgb_run.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, c_double_p]
    lib.dgb_eval_rhs.argtypes = [C.c_void_p, c_double_p, c_double_p]
    lib.dgb_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    lib.dgb_synchronize.argtypes = [C.c_void_p]
    lib.dgb_last_run_ms.restype = C.c_double
    lib.dgb_last_run_ms.argtypes = [C.c_void_p]
    lib.dgb_measure_fp64_tflops.restype = C.c_double
    lib.dgb_measure_fp64_tflops.argtypes = [C.c_void_p]
    lib.dgb_last_stage_kernel_ms.restype = C.c_double
    lib.dgb_last_stage_kernel_ms.argtypes = [C.c_void_p]
    lib.dgb_launch_count.restype = C.c_int64
    lib.dgb_launch_count.argtypes = [C.c_void_p]
    lib.dgb_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    lib.dgb_get_option.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int)]
    _dgb = lib
    return lib


class Engine:
    """Thin object wrapper over the dgb_* C ABI (one handle)."""

    def __init__(self, mesh: Mesh, el_part=None, rank=0, nranks=1, nccl_id=None, options=None):
        lib = self.lib = load_dgb()
        self.mesh = mesh
        self.N = mesh.N
        h = C.c_void_p()
        if nranks > 1:
            part = np.ascontiguousarray(el_part, dtype=np.int32)
            idbuf = (C.c_char * 128).from_buffer_copy(bytes(nccl_id))
            rc = lib.dgb_create_partitioned(mesh.desc_p, _as(c_int32_p, part), rank, nranks, idbuf, C.byref(h))
        else:
            rc = lib.dgb_create(mesh.desc_p, C.byref(h))
        self._check(rc)
        self.h = h
        for k, v in (options or {}).items():
            self.set_option(k, v)

    def _check(self, rc):
        if rc != 0:
            raise DgbError(f"dgb error {rc}: {self.lib.dgb_last_error().decode()}")

    def set_option(self, key, value):
        self._check(self.lib.dgb_set_option(self.h, key.encode(), int(value)))

    def measure_fp64_tflops(self):
        return self.lib.dgb_measure_fp64_tflops(self.h)

    def get_option(self, key):
        v = C.c_int(0)
        self._check(self.lib.dgb_get_option(self.h, key.encode(), C.byref(v)))
        return v.value

    @property
    def kernel_name(self):
        return self.lib.dgb_kernel_name(self.h).decode()

    def set_state(self, u):
        u = np.ascontiguousarray(u, dtype=np.float64)
        assert u.size == 4 * self.N
        self._check(self.lib.dgb_set_state(self.h, _as(c_double_p, u)))

    def get_state(self, out=None):
        u = out if out is not None else np.zeros((4, self.N), dtype=np.float64)
        self._check(self.lib.dgb_get_state(self.h, _as(c_double_p, u)))
        return u

    def snapshot_begin(self, out):
        """Asynchronous get_state into `out` (ideally pinned memory); snapshot_end() waits for it."""
        assert out.dtype == np.float64 and out.flags.c_contiguous and out.size == 4 * self.N
        self._check(self.lib.dgb_snapshot_begin(self.h, _as(c_double_p, out)))

    def snapshot_end(self):
        self._check(self.lib.dgb_snapshot_end(self.h))

    def set_sources(self, offsets, idx, amp, freq, phase, duration):
        offsets = np.ascontiguousarray(offsets, dtype=np.int32)
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        a, f, p, d = (np.ascontiguousarray(x, dtype=np.float64) for x in (amp, freq, phase, duration))
        self._check(self.lib.dgb_set_sources(self.h, len(a), _as(c_int32_p, offsets), _as(c_int32_p, idx),
                                             _as(c_double_p, a), _as(c_double_p, f), _as(c_double_p, p), _as(c_double_p, d)))

    def set_sources_from_config(self):
        src = self.mesh.cfg.sources
        if not src:
            return
        offsets, idx = self.mesh.source_nodes()
        s = np.array(src)
        self.set_sources(offsets, idx, s[:, 5], s[:, 6], s[:, 7], s[:, 8])

    def set_probes(self, idx):
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        self.nprobe = len(idx)
        self._check(self.lib.dgb_set_probes(self.h, len(idx), _as(c_int32_p, idx)))

    def get_probes(self, capacity_steps):
        out = np.zeros((capacity_steps, self.nprobe, 4), dtype=np.float64)
        n = C.c_int(0)
        self._check(self.lib.dgb_get_probes(self.h, _as(c_double_p, out), capacity_steps, C.byref(n)))
        return out[: n.value]

    def set_receivers(self, el, weights):
        el = np.ascontiguousarray(el, dtype=np.int32)
        w = np.ascontiguousarray(weights, dtype=np.float64)
        assert w.shape == (len(el), self.mesh.Np)
        self.nrecv = len(el)
        self._check(self.lib.dgb_set_receivers(self.h, len(el), _as(c_int32_p, el), _as(c_double_p, w)))

    def get_receivers(self, capacity_steps):
        out = np.zeros((capacity_steps, self.nrecv, 4), dtype=np.float64)
        n = C.c_int(0)
        self._check(self.lib.dgb_get_receivers(self.h, _as(c_double_p, out), capacity_steps, C.byref(n)))
        return out[: n.value]

    def run(self, integrator, t_start, nsteps):
        t_end = C.c_double(0.0)
        self._check(self.lib.dgb_run(self.h, int(integrator), float(t_start), int(nsteps), C.byref(t_end)))
        return t_end.value

    def eval_rhs(self, u):
        u = np.ascontiguousarray(u, dtype=np.float64)
        rhs = np.zeros((4, self.N), dtype=np.float64)
        self._check(self.lib.dgb_eval_rhs(self.h, _as(c_double_p, u), _as(c_double_p, rhs)))
        return rhs

    def set_stream(self, cuda_stream_ptr):
        self._check(self.lib.dgb_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def synchronize(self):
        self._check(self.lib.dgb_synchronize(self.h))

    @property
    def last_run_ms(self):
        return self.lib.dgb_last_run_ms(self.h)

    @property
    def last_stage_kernel_ms(self):
        return self.lib.dgb_last_stage_kernel_ms(self.h)

    @property
    def launch_count(self):
        return self.lib.dgb_launch_count(self.h)

    def close(self):
        if getattr(self, "h", None):
            self.lib.dgb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def nccl_unique_id() -> bytes:
    buf = (C.c_char * 128)()
    rc = load_dgb().dgb_nccl_unique_id(buf)
    if rc != 0:
        raise DgbError(f"dgb_nccl_unique_id failed: {load_dgb().dgb_last_error().decode()}")
    return bytes(buf)
