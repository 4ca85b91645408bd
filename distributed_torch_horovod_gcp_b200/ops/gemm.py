"""tcgen05 GEMM binding (csrc/gemm_sm100.cu) — placeholder registration until the kernel lands."""
from __future__ import annotations


def register(lib, have):
    return
