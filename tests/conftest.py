import gzip
import shutil
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import __graft_entry__ as graft  # noqa: E402

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    return graft.load_package()


@pytest.fixture(scope="session")
def mesh_dir(tmp_path_factory):
    """Decompressed copies of the reference's sample meshes."""
    d = tmp_path_factory.mktemp("meshes")
    for gz in (GOLDEN / "meshes").glob("*.msh.gz"):
        with gzip.open(gz, "rb") as src, open(d / gz.name[:-3], "wb") as dst:
            shutil.copyfileobj(src, dst)
    return d


@pytest.fixture(scope="session")
def config_dir():
    return GOLDEN / "configs"


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle_py
    return oracle_py


def rel_l2(a, b):
    import numpy as np
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
