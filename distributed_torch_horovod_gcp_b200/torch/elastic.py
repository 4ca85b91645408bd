"""Minimal elastic-state API (Horovod ``hvd.elastic`` shape): in-memory commit / restore /
sync of model + optimizer, and a ``run`` decorator that rolls back to the last commit when
a collective fails (``HorovodInternalError``).

The reference enables none of this; it is provided for surface parity and
as the failure-recovery hook for the kernel watchdog (a bounded spin-wait in the sm_100a
kernels sets an error flag -> ``HorovodInternalError``).  Dynamic host discovery /
re-rendezvous with a different world size is out of scope.
"""
from __future__ import annotations

import copy
import functools

import torch

from .mpi_ops import HorovodInternalError


class State:
    def __init__(self, **kwargs):
        self._attrs = dict(kwargs)
        for k, v in kwargs.items():
            setattr(self, k, v)
        self._saved = {}
        self._reset_callbacks = []
        self.save()

    def register_reset_callbacks(self, callbacks):
        self._reset_callbacks.extend(callbacks)

    def on_reset(self):
        for cb in self._reset_callbacks:
            cb()

    def save(self):
        self._saved = {k: copy.deepcopy(getattr(self, k)) for k in self._attrs}

    def restore(self):
        for k, v in self._saved.items():
            setattr(self, k, copy.deepcopy(v))

    def commit(self):
        self.save()

    def check_host_updates(self):
        return None

    def sync(self):
        from .functions import broadcast_object
        for k in self._attrs:
            setattr(self, k, broadcast_object(getattr(self, k), 0))


class TorchState(State):
    """State of a model + optimizer (+ arbitrary picklable kwargs such as epoch/batch)."""

    def __init__(self, model=None, optimizer=None, **kwargs):
        self.model, self.optimizer = model, optimizer
        self._model_sd, self._opt_sd = None, None
        super().__init__(**kwargs)

    def save(self):
        super().save()
        if self.model is not None:
            self._model_sd = {k: v.detach().clone() for k, v in self.model.state_dict().items()}
        if self.optimizer is not None:
            eng = getattr(self.optimizer, "fused_engine", None)
            if eng is not None:
                eng.export_state()
            self._opt_sd = copy.deepcopy(self.optimizer.state_dict())

    def restore(self):
        super().restore()
        if self.model is not None and self._model_sd is not None:
            with torch.no_grad():
                for k, v in self.model.state_dict().items():
                    v.copy_(self._model_sd[k])
        if self.optimizer is not None and self._opt_sd is not None:
            self.optimizer.load_state_dict(copy.deepcopy(self._opt_sd))
            eng = getattr(self.optimizer, "fused_engine", None)
            if eng is not None:
                eng.params_changed()
                eng.import_state()

    def sync(self):
        from .functions import broadcast_parameters, broadcast_optimizer_state
        if self.model is not None:
            broadcast_parameters(self.model.state_dict(), root_rank=0)
        if self.optimizer is not None:
            broadcast_optimizer_state(self.optimizer, root_rank=0)
        super().sync()


class ObjectState(State):
    """State of arbitrary picklable python objects (Horovod ``hvd.elastic.ObjectState``)."""


class ElasticSampler(torch.utils.data.Sampler):
    """Sharded sampler that remembers which indices were already processed in the current epoch,
    so that after a reset (rollback / world-size change) the remaining samples are re-partitioned
    over the new set of ranks (Horovod ``hvd.elastic.ElasticSampler`` shape)."""

    def __init__(self, dataset, shuffle: bool = True, seed: int = 0):
        self.dataset, self.shuffle, self.seed = dataset, shuffle, seed
        self.epoch = 0
        self.processed_indices = set()
        self.reset()

    def set_epoch(self, epoch: int):
        self.epoch = epoch
        self.processed_indices = set()
        self.reset()

    def record_batch(self, batch_idx: int, batch_size: int):
        lo = batch_idx * batch_size
        self.processed_indices.update(self.indices[lo: lo + batch_size])

    def load_state_dict(self, sd):
        self.epoch = sd["epoch"]
        self.processed_indices = set(sd["processed_indices"])
        self.reset()

    def state_dict(self):
        return {"epoch": self.epoch, "processed_indices": sorted(self.processed_indices)}

    def reset(self):
        from .. import _state
        self.num_replicas = _state.size() if _state.is_initialized() else 1
        self.rank = _state.rank() if _state.is_initialized() else 0
        remaining = [i for i in range(len(self.dataset)) if i not in self.processed_indices]
        if self.shuffle:
            g = torch.Generator().manual_seed(self.seed + self.epoch)
            perm = torch.randperm(len(remaining), generator=g).tolist()
            remaining = [remaining[i] for i in perm]
        self.num_samples = -(-len(remaining) // self.num_replicas) if remaining else 0
        total = self.num_samples * self.num_replicas
        remaining += remaining[: total - len(remaining)]
        self.indices = remaining[self.rank: total: self.num_replicas]

    def __iter__(self):
        return iter(self.indices)

    def __len__(self):
        return len(self.indices)


def run(func):
    """Decorator: ``func(state, *args)`` is retried from the last commit after a collective
    failure.  ``B200DP_ELASTIC_MAX_RETRIES`` bounds the retries (default 3)."""
    import os

    @functools.wraps(func)
    def wrapper(state, *args, **kwargs):
        retries = int(os.environ.get("B200DP_ELASTIC_MAX_RETRIES", "3"))
        state.sync()
        while True:
            try:
                return func(state, *args, **kwargs)
            except HorovodInternalError:
                if retries <= 0:
                    raise
                retries -= 1
                state.restore()
                state.on_reset()
                state.sync()
    return wrapper
