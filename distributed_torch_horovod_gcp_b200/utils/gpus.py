"""GPU discovery — a dependency-free replacement for ``GPUtil.getGPUs()`` used by the
reference for its device listing (app/torch_train.py:215-216,236-239).  ``GPUtil`` shells
out to ``nvidia-smi``; here NVML (``pynvml``) is preferred, then ``torch.cuda``, so no
fork/exec happens in a process that already owns a CUDA context.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List


@dataclass
class GPU:
    id: int
    name: str
    memoryTotal: float = 0.0   # MiB
    memoryUsed: float = 0.0
    load: float = 0.0
    uuid: str = ""


def getGPUs() -> List[GPU]:
    try:
        import pynvml
        pynvml.nvmlInit()
        out = []
        for i in range(pynvml.nvmlDeviceGetCount()):
            h = pynvml.nvmlDeviceGetHandleByIndex(i)
            name = pynvml.nvmlDeviceGetName(h)
            if isinstance(name, bytes):
                name = name.decode()
            mem = pynvml.nvmlDeviceGetMemoryInfo(h)
            try:
                util = pynvml.nvmlDeviceGetUtilizationRates(h).gpu / 100.0
            except Exception:
                util = 0.0
            uuid = pynvml.nvmlDeviceGetUUID(h)
            out.append(GPU(i, name, mem.total / 2**20, mem.used / 2**20, util,
                           uuid.decode() if isinstance(uuid, bytes) else uuid))
        return out
    except Exception:
        pass
    try:
        import torch
        if torch.cuda.is_available():
            return [GPU(i, torch.cuda.get_device_name(i),
                        torch.cuda.get_device_properties(i).total_memory / 2**20)
                    for i in range(torch.cuda.device_count())]
    except Exception:
        pass
    return []
