"""Process sets: collectives over a subset of ranks (Horovod ``hvd.ProcessSet`` shape).

Not used by the reference (SURVEY.md §2.3 "not used but part of the surface").  Subset
collectives run on ``torch.distributed`` sub-groups; the symmetric-memory fast path is
reserved for the global set.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch.distributed as dist

from .. import _state


class ProcessSet:
    def __init__(self, ranks: Optional[Sequence[int]] = None):
        self.ranks: Optional[List[int]] = sorted(ranks) if ranks is not None else None
        self.process_set_id: Optional[int] = None
        self.group = None

    def _materialise(self):
        rt = _state._require_init()
        if self.ranks is None:
            self.ranks = list(range(rt.size))
        if rt.size > 1 and self.group is None and len(self.ranks) != rt.size:
            self.group = dist.new_group(ranks=self.ranks, backend="gloo" if not
                                        _cuda_group_needed() else None)
        return self

    def size(self) -> int:
        self._materialise()
        return len(self.ranks)

    def rank(self) -> int:
        self._materialise()
        r = _state.rank()
        if r not in self.ranks:
            raise ValueError(f"rank {r} is not part of process set {self.ranks}")
        return self.ranks.index(r)

    def included(self) -> bool:
        self._materialise()
        return _state.rank() in self.ranks

    def __repr__(self):
        return f"ProcessSet(process_set_id={self.process_set_id}, ranks={self.ranks})"


def _cuda_group_needed() -> bool:
    import torch
    return torch.cuda.is_available()


global_process_set = ProcessSet()
global_process_set.process_set_id = 0


def add_process_set(ranks) -> ProcessSet:
    """Collective: every rank must call with the same ranks."""
    rt = _state._require_init()
    ps = ranks if isinstance(ranks, ProcessSet) else ProcessSet(ranks)
    ps._materialise()
    ps.process_set_id = max(rt.process_sets.keys(), default=0) + 1
    rt.process_sets[ps.process_set_id] = ps
    return ps


def remove_process_set(ps: ProcessSet) -> bool:
    rt = _state._require_init()
    if ps.process_set_id in rt.process_sets:
        del rt.process_sets[ps.process_set_id]
        if ps.group is not None:
            try:
                dist.destroy_process_group(ps.group)
            except Exception:
                pass
            ps.group = None
        return True
    return False
