import torch

from distributed_torch_horovod_gcp_b200.models import LSTM, resnet50
from distributed_torch_horovod_gcp_b200.parallel.buckets import plan_buckets, plan_hash, arena_sizes


def _plan(model, **kw):
    named = [(n, p) for n, p in model.named_parameters()]
    return plan_buckets(named, {id(p): 0 for _, p in named}, **kw)


def test_lstm_single_bucket_alignment():
    bs = _plan(LSTM(23, 10, 1, 256))
    assert len(bs) == 1                         # 1.48 MB -> one bucket
    b = bs[0]
    assert [s.name for s in b.slots][0] == "linear3.bias"      # reverse registration order
    for s in b.slots:
        assert (s.offset * 4) % 16 == 0          # every tensor 16-byte aligned, incl. 4-byte one
    assert b.numel * 4 % 4096 == 0 and b.numel >= 370049
    ends = sorted((s.offset, s.offset + s.numel) for s in b.slots)
    assert all(a[1] <= c[0] for a, c in zip(ends, ends[1:]))   # no overlap


def test_resnet50_bucket_cap_and_hash():
    m = resnet50()
    bs = _plan(m, bucket_bytes=16 << 20)
    assert 6 <= len(bs) <= 9
    assert sum(len(b.slots) for b in bs) == 161
    sizes = arena_sizes(bs)
    assert len(sizes) == 1 and list(sizes.values())[0] >= 25557032
    h1 = plan_hash(bs)
    assert h1 == plan_hash(_plan(resnet50(), bucket_bytes=16 << 20))
    assert h1 != plan_hash(_plan(m, bucket_bytes=8 << 20))


def test_groups_and_param_groups_do_not_mix():
    m = LSTM(23, 10, 1, 256)
    named = list(m.named_parameters())
    gof = {id(p): (0 if p.dim() > 1 else 1) for _, p in named}
    bs = plan_buckets(named, gof)
    assert len(bs) == 2 and {b.group_index for b in bs} == {0, 1}
    bs = plan_buckets(named, {id(p): 0 for _, p in named}, num_groups=3)
    assert len(bs) >= 2
    grp = [[m.linear3.weight, m.linear3.bias]]
    bs = plan_buckets(named, {id(p): 0 for _, p in named}, explicit_groups=grp)
    assert {s.name for s in bs[0].slots} == {"linear3.weight", "linear3.bias"}
