"""The C-ABI library loads without a GPU and exports every symbol include/dgb.h declares; without a CUDA device
the product path fails loudly (no CPU fallback)."""
import ctypes as C
import re

import numpy as np
import pytest

from conftest import ROOT


def declared(header):
    text = (ROOT / "include" / header).read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    prefix = header.split(".")[0][:3]
    return sorted(set(re.findall(r"\b(%s_[a-z0-9_]+)\s*\(" % prefix, text)))


def test_dgb_exports_every_declared_symbol(pkg):
    lib = pkg.load_dgb()
    names = declared("dgb.h")
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), n
    assert b"sm_100a" in lib.dgb_version()


def test_dgfront_exports_every_declared_symbol(pkg):
    lib = pkg.load_front()
    for n in declared("dgfront.h"):
        assert hasattr(lib, n), n


def test_desc_struct_layout_matches_header(pkg):
    """ctypes mirror of struct dgb_desc: 12 int32, 15 pointers, 6 doubles."""
    from dgfem_acoustic_b200.capi import DgbDesc
    assert C.sizeof(DgbDesc) == 12 * 4 + 15 * 8 + 6 * 8


def test_no_cpu_fallback(pkg, mesh_dir):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    mesh = pkg.Mesh(pkg.Model.open_msh(mesh_dir / "line.msh", 1), pkg.Config())
    with pytest.raises(pkg.DgbError) as e:
        pkg.Engine(mesh)
    assert "-3" in str(e.value) or "CUDA" in str(e.value)  # DGB_ERR_CUDA


def test_argument_validation_without_gpu(pkg):
    lib = pkg.load_dgb()
    h = C.c_void_p()
    assert lib.dgb_create(None, C.byref(h)) == -1  # DGB_ERR_ARG
    assert b"null" in lib.dgb_last_error()
    assert lib.dgb_run(None, 1, 0.0, 1, None) == -1
    assert lib.dgb_set_state(None, None) == -1
    assert lib.dgb_launch_count(None) == 0


def test_product_package_does_not_reference_the_oracle():
    pkg_dir = ROOT / "dgfem-acoustic_b200"
    for path in list(pkg_dir.rglob("*.py")) + list(pkg_dir.rglob("*.cu")) + list(pkg_dir.rglob("*.cpp")) + list(pkg_dir.rglob("*.h")) + list(pkg_dir.rglob("*.cuh")):
        text = path.read_text()
        assert "liboracle" not in text and "oracle_py" not in text and "orc_" not in text, path
