"""ctypes signatures + autograd wrappers for the compute kernels.  Filled in as kernels land;
``register`` marks an op available only if its C symbol exists in the built library."""
from __future__ import annotations

import ctypes
from typing import Dict


def register(lib, have: Dict[str, bool]) -> None:
    from . import gemm as _gemm
    _gemm.register(lib, have)
