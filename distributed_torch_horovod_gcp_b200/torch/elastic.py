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
