"""Fused LSTM forward/backward (K5) + chained linear head (K6) for the reference model
(SURVEY.md §2.6 S3-S7).  Bound to csrc/lstm_kernels.cu when built; ``available`` gates use."""
from __future__ import annotations

import torch


def available(model, x: torch.Tensor) -> bool:
    try:
        from . import kernels
    except Exception:
        return False
    return kernels.enabled_for(x) and kernels.has("lstm_fused") and \
        model.h_size == 256 and x.dtype == torch.float32


def forward(model, x, hidden):
    from . import kernels
    return kernels.lstm_head_forward(model, x, hidden)
