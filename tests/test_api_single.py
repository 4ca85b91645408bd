"""World-size-1 plumbing (SURVEY.md §4 tier 1; reference README.md:20-23 single-GPU form)."""
import pytest
import torch

from distributed_torch_horovod_gcp_b200.models import LSTM


def test_init_without_launcher(hvd_single):
    hvd = hvd_single
    assert (hvd.rank(), hvd.size(), hvd.local_rank(), hvd.local_size()) == (0, 1, 0, 1)
    assert (hvd.cross_rank(), hvd.cross_size()) == (0, 1)
    hvd.init()                                   # idempotent
    assert hvd.is_initialized() and hvd.gloo_built() and not hvd.mpi_built()


def test_uninitialised_raises():
    import distributed_torch_horovod_gcp_b200.torch as hvd
    hvd.shutdown()
    with pytest.raises(ValueError, match="has not been initialized"):
        hvd.rank()


def test_optimizer_is_a_and_noop(hvd_single):
    hvd = hvd_single
    m = LSTM(23, 10, 1, 32)
    base = torch.optim.Adam(m.parameters(), lr=1e-3)
    opt = hvd.DistributedOptimizer(base, named_parameters=m.named_parameters())
    assert isinstance(opt, torch.optim.Adam) and type(opt).__name__ == "Adam"
    assert opt.param_groups[0]["lr"] == 1e-3
    ref = LSTM(23, 10, 1, 32)
    ref.load_state_dict(m.state_dict())
    ropt = torch.optim.Adam(ref.parameters(), lr=1e-3)
    x, y = torch.randn(8, 10, 23), torch.randn(8, 1, 1)
    for mod, o in ((m, opt), (ref, ropt)):
        torch.manual_seed(0)
        torch.nn.functional.mse_loss(mod(x), y).backward()
        o.step()
        o.zero_grad()
    for a, b in zip(m.parameters(), ref.parameters()):
        torch.testing.assert_close(a, b)
    assert all(p.grad is None for p in m.parameters())     # plain optimizer semantics at size 1


def test_optimizer_argument_validation(hvd_single):
    hvd = hvd_single
    m = torch.nn.Linear(4, 4)
    with pytest.raises(ValueError, match="unique"):
        hvd.DistributedOptimizer(torch.optim.SGD(m.parameters(), lr=0.1),
                                 named_parameters=[("a", m.weight), ("a", m.bias)])
    with pytest.raises(ValueError, match="not named"):
        hvd.DistributedOptimizer(torch.optim.SGD(m.parameters(), lr=0.1),
                                 named_parameters=[("a", m.weight)])
    with pytest.raises(ValueError, match="tuples"):
        hvd.DistributedOptimizer(torch.optim.SGD(m.parameters(), lr=0.1),
                                 named_parameters=[m.weight, m.bias])
    with pytest.raises(NotImplementedError):
        hvd.DistributedOptimizer(torch.optim.SGD(m.parameters(), lr=0.1), op=hvd.Adasum)


def test_collectives_identity_at_size_one(hvd_single):
    hvd = hvd_single
    t = torch.arange(6.0)
    torch.testing.assert_close(hvd.allreduce(t), t)
    torch.testing.assert_close(hvd.allreduce(t, op=hvd.Sum, prescale_factor=2.0), 2 * t)
    torch.testing.assert_close(hvd.broadcast(t, 0), t)
    torch.testing.assert_close(hvd.allgather(t), t)
    torch.testing.assert_close(hvd.alltoall(t), t)
    torch.testing.assert_close(hvd.reducescatter(t), t)
    h = hvd.allreduce_async_(t.clone(), name="x")
    assert hvd.poll(h)
    hvd.synchronize(h)
    with pytest.raises(ValueError):
        hvd.synchronize(h)
    assert hvd.broadcast_object({"a": 1}) == {"a": 1}
    assert hvd.allgather_object(3) == [3]
    hvd.barrier()
    assert hvd.join() == 0
    hvd.broadcast_parameters(LSTM(23, 10, 1, 8).state_dict(), root_rank=0)
    with pytest.raises(ValueError):
        hvd.broadcast_parameters(5, root_rank=0)


def test_allreduce_autograd(hvd_single):
    hvd = hvd_single
    x = torch.ones(3, requires_grad=True)
    y = hvd.allreduce(x, op=hvd.Sum)
    y.sum().backward()
    torch.testing.assert_close(x.grad, torch.ones(3))


def test_compression_roundtrip(hvd_single):
    hvd = hvd_single
    t = torch.randn(8)
    c, ctx = hvd.Compression.fp16.compress(t)
    assert c.dtype == torch.float16
    assert hvd.Compression.fp16.decompress(c, ctx).dtype == torch.float32
    c, ctx = hvd.Compression.bf16.compress(t)
    assert c.dtype == torch.bfloat16
    assert hvd.Compression.none.compress(t)[0] is t


def test_timeline(hvd_single, tmp_path):
    import json
    hvd = hvd_single
    p = tmp_path / "tl.json"
    hvd.start_timeline(str(p))
    from distributed_torch_horovod_gcp_b200 import _state
    _state.runtime().timeline.mark("bucket.0", "BUCKET_READY", bytes=4)
    _state.runtime().timeline.begin("step", "STEP")
    _state.runtime().timeline.end("step", "STEP")
    hvd.stop_timeline()
    ev = json.load(open(p))["traceEvents"]
    assert [e["name"] for e in ev] == ["BUCKET_READY", "STEP", "STEP"]


def test_elastic_state_commit_restore(hvd_single):
    hvd = hvd_single
    m = torch.nn.Linear(2, 2)
    opt = torch.optim.SGD(m.parameters(), lr=0.1, momentum=0.9)
    st = hvd.elastic.TorchState(m, opt, epoch=0)
    w0 = m.weight.detach().clone()
    m(torch.ones(1, 2)).sum().backward()
    opt.step()
    st.epoch = 5
    assert not torch.equal(m.weight, w0)
    st.restore()
    assert torch.equal(m.weight, w0) and st.epoch == 0
    calls = []

    @hvd.elastic.run
    def train(state):
        calls.append(1)
        if len(calls) == 1:
            raise hvd.HorovodInternalError("boom")
        return "done"
    assert train(st) == "done" and len(calls) == 2


def test_sync_batch_norm_single(hvd_single):
    hvd = hvd_single
    bn = hvd.SyncBatchNorm(4)
    ref = torch.nn.BatchNorm2d(4)
    x = torch.randn(8, 4, 3, 3)
    torch.testing.assert_close(bn(x), ref(x))
