"""Direct peer-to-peer halo exchange (dgfem-acoustic_b200/csrc/halo_p2p.cu — the file itself) for 2..8 ranks in ONE process,
through the CUDA emulation of oracle/cuda_emu.h. On hardware the exchange ran with two ranks, i.e. one peer per rank
(tests/test_multi_gpu.py); here the multi-peer bookkeeping is checked — which peer a send element goes to, which halo slot of
that peer's array it lands in, which flag slot signals it — on RCB and METIS partitions: after one push + signal + wait every
halo entry of every rank holds its owner's value."""
import ctypes as C

import numpy as np
import pytest

from conftest import ROOT

ip = C.POINTER(C.c_int32)


@pytest.fixture(scope="module")
def p2e():
    lib = C.CDLL(str(ROOT / "oracle" / "libp2pemu.so"))
    lib.p2e_last_error.restype = C.c_char_p
    lib.p2e_check.argtypes = [C.c_void_p, ip, C.c_int, C.POINTER(C.c_int)]
    return lib


@pytest.mark.parametrize("partitioner", ["rcb", "metis"])
@pytest.mark.parametrize("nranks", [2, 3, 4, 8])
def test_every_halo_entry_arrives(pkg, p2e, mesh_dir, nranks, partitioner):
    front = pkg.load_front()
    mesh = pkg.Mesh(pkg.Model.open_msh(mesh_dir / "sphere.msh", 2) if partitioner == "metis" else pkg.Model.make_cube(6, -10.0, 10.0, 2), pkg.Config())
    part = np.zeros(mesh.K, dtype=np.int32)
    if partitioner == "metis":
        rc = front.dgf_partition_metis(mesh.h, nranks, part.ctypes.data_as(ip), None)
        if rc == -2:
            pytest.skip("front end built without METIS")
        assert rc == 0
    else:
        assert front.dgf_partition_rcb(mesh.h, nranks, part.ctypes.data_as(ip)) == 0
    max_peers = C.c_int(0)
    wrong = p2e.p2e_check(C.cast(mesh.desc_p, C.c_void_p), part.ctypes.data_as(ip), nranks, C.byref(max_peers))
    assert wrong == 0, p2e.p2e_last_error()
    assert max_peers.value >= 1 and (nranks < 4 or max_peers.value >= 2)  # several peers per rank are exercised
