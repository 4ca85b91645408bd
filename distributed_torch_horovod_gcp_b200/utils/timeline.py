"""Chrome-trace timeline (``HOROVOD_TIMELINE`` equivalent).

Horovod's timeline is written by its background thread with per-tensor phases
(NEGOTIATE / WAIT_FOR_DATA / MEMCPY_IN / NCCL_ALLREDUCE / MEMCPY_OUT).  This runtime has no
negotiation and no fusion memcpy, so the phases are: BUCKET_READY (hook side),
ALLREDUCE_BEGIN/END per bucket (CUDA-event timed on the side stream) and STEP.
Enable with ``HOROVOD_TIMELINE=/path.json`` (``{rank}`` is substituted; otherwise
``.rank<N>`` is appended for ranks > 0) or ``hvd.start_timeline(path)``.
"""
from __future__ import annotations

import json
import os
import threading
import time
from typing import List


class Timeline:
    def __init__(self, path: str, rank: int = 0):
        if "{rank}" in path:
            path = path.format(rank=rank)
        elif rank != 0:
            path = f"{path}.rank{rank}"
        self.path, self.rank = path, rank
        self._events: List[dict] = []
        self._cuda_pairs = []
        self._lock = threading.Lock()
        self._t0 = time.perf_counter()
        self._closed = False

    def _now_us(self) -> float:
        return (time.perf_counter() - self._t0) * 1e6

    def mark(self, name: str, phase: str, **args) -> None:
        with self._lock:
            self._events.append({"name": phase, "cat": name, "ph": "i", "s": "t",
                                 "ts": self._now_us(), "pid": self.rank, "tid": 0,
                                 "args": args})

    def begin(self, name: str, phase: str, tid: int = 0, **args) -> None:
        with self._lock:
            self._events.append({"name": phase, "cat": name, "ph": "B", "ts": self._now_us(),
                                 "pid": self.rank, "tid": tid, "args": args})

    def end(self, name: str, phase: str, tid: int = 0) -> None:
        with self._lock:
            self._events.append({"name": phase, "cat": name, "ph": "E", "ts": self._now_us(),
                                 "pid": self.rank, "tid": tid})

    def cuda_span(self, name: str, phase: str, start_event, end_event, **args) -> None:
        """Register a pair of recorded CUDA events; resolved to a device-timed span at close
        (never synchronises on the hot path)."""
        with self._lock:
            self._cuda_pairs.append((name, phase, self._now_us(), start_event, end_event, args))

    def close(self) -> None:
        if self._closed:
            return
        self._closed = True
        with self._lock:
            for name, phase, ts, s, e, args in self._cuda_pairs:
                try:
                    e.synchronize()
                    dur_us = s.elapsed_time(e) * 1e3
                except Exception:
                    dur_us = 0.0
                self._events.append({"name": phase, "cat": name, "ph": "X", "ts": ts,
                                     "dur": dur_us, "pid": self.rank, "tid": 1, "args": args})
            d = os.path.dirname(os.path.abspath(self.path))
            os.makedirs(d, exist_ok=True)
            with open(self.path, "w") as f:
                json.dump({"traceEvents": self._events, "displayTimeUnit": "ms"}, f)
