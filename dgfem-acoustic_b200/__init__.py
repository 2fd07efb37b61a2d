"""B200-native engine for the DGFEM-Acoustic time-marching hot path.

The directory name carries a hyphen (it mirrors the reference repository's name), so the package is
imported through ``__graft_entry__.load_package()`` / ``tests/conftest.py`` under the module name
``dgfem_acoustic_b200``.
"""
from .capi import (EULER1, RUNGE_KUTTA, Config, DgbError, Engine, FrontError, Mesh, Model, load_dgb, load_front,  # noqa: F401
                   nccl_unique_id)
