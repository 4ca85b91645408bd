"""CPU tests of the autograd hand-off protocols in ops/grad_sink.py (no kernels involved): the GradBox
producer/consumer ordering rules and the deferred-cast queue's no-op behaviour outside a backward pass."""
import torch

from distributed_torch_horovod_gcp_b200.ops import grad_sink
from distributed_torch_horovod_gcp_b200.ops.grad_sink import GradBox


def test_gradbox_unarmed_producer_keeps_its_gradient():
    box = GradBox()
    g = torch.ones(3)
    assert not box.park(g)              # no consumer registered in forward -> return the gradient to autograd
    assert box.take() == (None, None)


def test_gradbox_park_then_take():
    box = GradBox()
    box.armed = True                    # consumer's forward
    g, m = torch.ones(3), torch.zeros(1, dtype=torch.uint8)
    assert box.park(g, m)
    assert not box.park(torch.zeros(3))     # a second producer must not overwrite a parked gradient
    got, mask = box.take()
    assert got is g and mask is m
    assert box.take() == (None, None)       # handed over exactly once
    assert box.consumed


def test_gradbox_consumer_first_means_no_parking():
    """Projection block: the downsample dgrad has no data dependency on conv1's dgrad; if conv1 ran first the
    producer must fall back to returning its gradient (autograd then adds the two)."""
    box = GradBox()
    box.armed = True
    assert box.take() == (None, None)       # consumer backward ran with nothing parked
    assert not box.park(torch.ones(2))


def test_gradbox_fresh_per_forward():
    a, b = GradBox(), GradBox()
    a.armed = True
    a.park(torch.ones(1))
    assert b.dres is None and not b.armed and not b.consumed and not b.want_mask and b.ext_mask is None


def test_deferred_cast_is_refused_outside_backward():
    """defer_cast registers an end-of-backward callback; outside a backward pass (or without the multi-tensor
    kernel) it must refuse so that the caller converts immediately."""
    class _NoLib:
        pass
    ws, dst = torch.zeros(8), torch.zeros(8, dtype=torch.bfloat16)
    assert not grad_sink.defer_cast(_NoLib(), ws, dst, 8, False)

    class _Lib:
        def b200dp_multi_cast_acc_zero(self, *a):
            raise AssertionError("must not be called")
    assert not grad_sink.defer_cast(_Lib(), ws, dst, 8, False)      # not inside backward -> RuntimeError path
    assert not grad_sink._pending_casts
    grad_sink.flush_casts()                                          # empty queue: no-op
