"""NVTX ranges around the framework's phases: bucket launch, optimizer step, graph
replay, collectives.  Off by default (``B200DP_NVTX=1`` or ``hvd.start_timeline`` turn them on) so the
launch-bound LSTM step does not pay two driver calls per range; with Nsight Systems / ``ncu --nvtx`` the
ranges name what each kernel on the side stream belongs to (``bucket.3 FUSED_ALLREDUCE_NVLS 16.0MB``)."""
from __future__ import annotations

import contextlib
import os

_ON = os.environ.get("B200DP_NVTX", "0") == "1"
_nvtx = None


def enabled() -> bool:
    return _ON


def enable(on: bool = True) -> None:
    global _ON
    _ON = bool(on)


def _mod():
    global _nvtx
    if _nvtx is None:
        import torch
        _nvtx = torch.cuda.nvtx
    return _nvtx


def push(name: str) -> None:
    if _ON:
        _mod().range_push(name)


def pop() -> None:
    if _ON:
        _mod().range_pop()


def mark(name: str) -> None:
    if _ON:
        _mod().mark(name)


@contextlib.contextmanager
def range(name: str):       # noqa: A001 - mirrors torch.cuda.nvtx.range
    if not _ON:
        yield
        return
    _mod().range_push(name)
    try:
        yield
    finally:
        _mod().range_pop()
