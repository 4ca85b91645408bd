"""Worker bodies for the multi-process Gloo tests (each runs on every rank)."""
import copy

import torch
import torch.nn.functional as F

from distributed_torch_horovod_gcp_b200.models import LSTM


def _model(seed):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.Tanh(), torch.nn.Linear(16, 3))


def broadcast_parameters(hvd):
    m = LSTM(23, 10, 1, 16)
    torch.manual_seed(100 + hvd.rank())
    with torch.no_grad():
        for p in m.parameters():
            p.normal_()
    bn = torch.nn.BatchNorm1d(4)
    bn.running_mean.fill_(float(hvd.rank()))
    sd = dict(m.state_dict())
    sd.update({"bn." + k: v for k, v in bn.state_dict().items()})
    hvd.broadcast_parameters(sd, root_rank=0)
    flat = torch.cat([v.reshape(-1).float() for v in sd.values()])
    gathered = hvd.allgather(flat.view(1, -1))
    assert all(torch.equal(gathered[0], gathered[r]) for r in range(hvd.size()))
    assert float(bn.running_mean[0]) == 0.0
    # named_parameters() iterable form + non-zero root
    m2 = _model(hvd.rank())
    hvd.broadcast_parameters(m2.named_parameters(), root_rank=hvd.size() - 1)
    ref = _model(hvd.size() - 1)
    for a, b in zip(m2.parameters(), ref.parameters()):
        assert torch.equal(a, b)
    return True


def dp_equals_single(hvd, opt_name, passes):
    """N-rank DP step == 1-rank step on the concatenated batch (fp32 allclose)."""
    world, rank = hvd.size(), hvd.rank()
    torch.manual_seed(7)
    X = torch.randn(8 * world, 6)
    Y = torch.randn(8 * world, 3)
    m = _model(0)
    ref = copy.deepcopy(m)

    def mk(mod):
        if opt_name == "sgd":
            return torch.optim.SGD(mod.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-3)
        return torch.optim.Adam(mod.parameters(), lr=1e-2)
    opt = hvd.DistributedOptimizer(mk(m), named_parameters=m.named_parameters(),
                                   backward_passes_per_step=passes)
    assert isinstance(opt, type(mk(ref)).__mro__[0])
    ropt = mk(ref)
    hvd.broadcast_parameters(m.state_dict(), root_rank=0)
    for step in range(3):
        # reference: full batch, mean loss (== average of per-rank mean losses, equal shards)
        ropt.zero_grad()
        F.mse_loss(ref(X), Y).backward()
        ropt.step()
        xs, ys = X[rank * 8:(rank + 1) * 8], Y[rank * 8:(rank + 1) * 8]
        for k in range(passes):
            sl = slice(k * (8 // passes), (k + 1) * (8 // passes))
            (F.mse_loss(m(xs[sl]), ys[sl]) / passes).backward()
        opt.step()
        opt.zero_grad()
    for a, b in zip(m.parameters(), ref.parameters()):
        torch.testing.assert_close(a, b, rtol=2e-5, atol=2e-6)
    flat = torch.cat([p.detach().reshape(-1) for p in m.parameters()])
    g = hvd.allgather(flat.view(1, -1))
    assert all(torch.equal(g[0], g[r]) for r in range(world))     # replicas stay bit-identical
    return True


def optimizer_edge_cases(hvd):
    m = _model(0)
    opt = hvd.DistributedOptimizer(torch.optim.SGD(m.parameters(), lr=0.1),
                                   named_parameters=m.named_parameters())
    hvd.broadcast_parameters(m.state_dict(), 0)
    x = torch.randn(4, 6)
    # grads are views into one flat bucket
    ptrs = sorted(p.grad.data_ptr() for p in m.parameters())
    assert len(opt.bucket_plan()) == 1 and ptrs[0] != 0
    # step() without backward() still synchronises (launches the bucket with zero grads)
    before = [p.detach().clone() for p in m.parameters()]
    opt.step()
    for a, b in zip(m.parameters(), before):
        assert torch.equal(a, b)
    # zero_grad() with in-flight reductions raises
    m(x).sum().backward()
    try:
        opt.zero_grad()
        raise RuntimeError("zero_grad should have raised")
    except AssertionError as e:
        assert "race condition" in str(e)
    # a second backward before step() raises (backward_passes_per_step=1)
    try:
        m(x).sum().backward()
        raise RuntimeError("second backward should have raised")
    except AssertionError as e:
        assert "backward_passes_per_step" in str(e)
    opt.synchronize()
    with opt.skip_synchronize():
        opt.step()
    opt.zero_grad()
    # unused parameter: its hook never fires, step() must still reduce the bucket
    m2 = torch.nn.ModuleDict({"a": torch.nn.Linear(3, 3), "b": torch.nn.Linear(3, 3)})
    opt2 = hvd.DistributedOptimizer(torch.optim.SGD(m2.parameters(), lr=0.5),
                                    named_parameters=m2.named_parameters())
    hvd.broadcast_parameters(m2.state_dict(), 0)
    (m2["a"](torch.ones(1, 3)) * (hvd.rank() + 1)).sum().backward()
    w_before = m2["a"].weight.detach().clone()
    opt2.step()
    mean_scale = sum(r + 1 for r in range(hvd.size())) / hvd.size()
    torch.testing.assert_close(m2["a"].weight, w_before - 0.5 * mean_scale * torch.ones(3, 3))
    # plan mismatch across ranks is detected at construction
    m3 = torch.nn.Linear(3, 3 + (1 if hvd.rank() == 1 else 0))
    try:
        hvd.DistributedOptimizer(torch.optim.SGD(m3.parameters(), lr=0.1),
                                 named_parameters=m3.named_parameters())
        raise RuntimeError("plan mismatch should have raised")
    except RuntimeError as e:
        assert "differs across ranks" in str(e), str(e)
    return True


def collectives(hvd):
    r, n = hvd.rank(), hvd.size()
    t = torch.full((5,), float(r + 1))
    s = float(sum(range(1, n + 1)))
    torch.testing.assert_close(hvd.allreduce(t), torch.full((5,), s / n))
    torch.testing.assert_close(hvd.allreduce(t, op=hvd.Sum), torch.full((5,), s))
    torch.testing.assert_close(hvd.allreduce(t, op=hvd.Max), torch.full((5,), float(n)))
    torch.testing.assert_close(hvd.allreduce(t, op=hvd.Min), torch.full((5,), 1.0))
    torch.testing.assert_close(hvd.allreduce(t, op=hvd.Sum, prescale_factor=0.5,
                                             postscale_factor=4.0), torch.full((5,), 2 * s))
    u = t.clone()
    hvd.allreduce_(u, op=hvd.Sum)
    torch.testing.assert_close(u, torch.full((5,), s))
    ti = torch.tensor([r + 1], dtype=torch.int64)
    assert int(hvd.allreduce(ti, op=hvd.Sum)) == int(s)
    # async + poll + duplicate names
    h = hvd.allreduce_async(t, name="dup")
    try:
        hvd.allreduce_async(t, name="dup")
        raise RuntimeError("duplicate name should raise")
    except ValueError as e:
        assert "Duplicate" in str(e)
    torch.testing.assert_close(hvd.synchronize(h), torch.full((5,), s / n))
    # compression
    torch.testing.assert_close(hvd.allreduce(t, compression=hvd.Compression.fp16),
                               torch.full((5,), s / n))
    # grouped
    outs = hvd.grouped_allreduce([t, torch.ones(2, 2) * r, ti.float()], op=hvd.Sum)
    torch.testing.assert_close(outs[0], torch.full((5,), s))
    torch.testing.assert_close(outs[1], torch.ones(2, 2) * sum(range(n)))
    # broadcast
    b = torch.arange(4.0) + 10 * r
    torch.testing.assert_close(hvd.broadcast(b, root_rank=n - 1), torch.arange(4.0) + 10 * (n - 1))
    assert torch.equal(b, torch.arange(4.0) + 10 * r)
    hvd.broadcast_(b, 0)
    torch.testing.assert_close(b, torch.arange(4.0))
    # allgather with uneven first dims
    g = hvd.allgather(torch.full((r + 1, 2), float(r)))
    assert g.shape == (sum(range(1, n + 1)), 2)
    off = 0
    for q in range(n):
        assert torch.all(g[off:off + q + 1] == q)
        off += q + 1
    # alltoall
    a = torch.arange(n * 2, dtype=torch.float32) + 100 * r
    out = hvd.alltoall(a)
    want = torch.cat([torch.arange(2 * r, 2 * r + 2, dtype=torch.float32) + 100 * q for q in range(n)])
    torch.testing.assert_close(out, want)
    splits = torch.tensor([q + 1 for q in range(n)])
    a2 = torch.cat([torch.full((q + 1,), float(r * 10 + q)) for q in range(n)])
    out2, rs = hvd.alltoall(a2, splits=splits)
    assert rs.tolist() == [r + 1] * n and out2.numel() == n * (r + 1)
    # reducescatter
    rsin = torch.arange(n * 3, dtype=torch.float32).view(n, 3) * (r + 1)
    got = hvd.reducescatter(rsin, op=hvd.Sum)
    torch.testing.assert_close(got, (torch.arange(n * 3, dtype=torch.float32).view(n, 3) * s)[r:r + 1])
    # objects / barrier / join
    assert hvd.broadcast_object({"k": r}, root_rank=1 % n) == {"k": 1 % n}
    assert hvd.allgather_object(r * r) == [q * q for q in range(n)]
    hvd.barrier()
    assert 0 <= hvd.join() < n
    # process sets
    if n >= 2:
        ps = hvd.add_process_set([0, 1])
        if ps.included():
            v = hvd.allreduce(torch.tensor([float(r)]), op=hvd.Sum, process_set=ps)
            assert float(v) == 1.0
        hvd.remove_process_set(ps)
    return True


def broadcast_optimizer_state(hvd):
    m = _model(hvd.rank())
    opt = torch.optim.Adam(m.parameters(), lr=1e-3 * (hvd.rank() + 1))
    if hvd.rank() == 0:        # only root has stepped (resume-from-checkpoint convention)
        m(torch.randn(2, 6)).sum().backward()
        opt.step()
    hvd.broadcast_parameters(m.state_dict(), 0)
    hvd.broadcast_optimizer_state(opt, 0)
    assert opt.param_groups[0]["lr"] == 1e-3
    st = opt.state_dict()["state"]
    assert len(st) == 4
    flat = torch.cat([st[i]["exp_avg"].reshape(-1) for i in sorted(st)])
    g = hvd.allgather(flat.view(1, -1))
    assert all(torch.equal(g[0], g[q]) for q in range(hvd.size())) and float(g[0].abs().sum()) > 0
    assert float(st[0]["step"]) == 1.0
    return True


def sync_batch_norm(hvd):
    r, n = hvd.rank(), hvd.size()
    torch.manual_seed(3)
    X = torch.randn(4 * n, 5, 3, 3)
    bn = hvd.SyncBatchNorm(5)
    ref = torch.nn.BatchNorm2d(5)
    with torch.no_grad():
        for mod in (bn, ref):
            mod.weight.copy_(torch.linspace(0.5, 1.5, 5))
            mod.bias.copy_(torch.linspace(-1, 1, 5))
    xs = X[r * 4:(r + 1) * 4].clone().requires_grad_(True)
    xr = X.clone().requires_grad_(True)
    y = bn(xs)
    yr = ref(xr)
    torch.testing.assert_close(y, yr[r * 4:(r + 1) * 4], rtol=1e-4, atol=1e-5)
    w = torch.linspace(-1, 1, yr.numel()).view_as(yr)
    (yr * w).sum().backward()
    (y * w[r * 4:(r + 1) * 4]).sum().backward()
    torch.testing.assert_close(xs.grad, xr.grad[r * 4:(r + 1) * 4], rtol=1e-3, atol=1e-5)
    torch.testing.assert_close(bn.running_mean, ref.running_mean, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(bn.running_var, ref.running_var, rtol=1e-4, atol=1e-6)
    return True


def timeline_and_elastic(hvd, tmpdir):
    import json, os
    path = os.path.join(tmpdir, "tl.json")
    hvd.start_timeline(path)
    m = _model(0)
    opt = hvd.DistributedOptimizer(torch.optim.SGD(m.parameters(), lr=0.1),
                                   named_parameters=m.named_parameters())
    hvd.broadcast_parameters(m.state_dict(), 0)
    m(torch.randn(4, 6)).sum().backward()
    opt.step()
    opt.zero_grad()
    hvd.stop_timeline()
    f = path if hvd.rank() == 0 else f"{path}.rank{hvd.rank()}"
    names = [e["name"] for e in json.load(open(f))["traceEvents"]]
    assert "BUCKET_READY" in names and "ALLREDUCE_BEGIN" in names and names.count("STEP") == 2
    # elastic sampler partitions the not-yet-processed samples over the ranks
    s = hvd.elastic.ElasticSampler(list(range(20)), shuffle=False)
    assert len(s) == 20 // hvd.size() and list(s)[0] == hvd.rank()
    s.record_batch(0, 2)
    s.reset()
    assert len(s) == (20 - 2 + hvd.size() - 1) // hvd.size()
    st = hvd.elastic.ObjectState(epoch=3)
    st.epoch = 7 if hvd.rank() == 0 else 1
    st.sync()
    assert st.epoch == 7
    return True


def optimizer_options(hvd):
    """groups / num_groups / compression / gradient_predivide_factor / op=Sum through the bucket path."""
    world, rank = hvd.size(), hvd.rank()
    torch.manual_seed(11)
    X, Y = torch.randn(4 * world, 6), torch.randn(4 * world, 3)
    xs, ys = X[rank * 4:(rank + 1) * 4], Y[rank * 4:(rank + 1) * 4]

    def run(**kw):
        m = _model(0)
        lr = kw.pop("lr", 0.1)
        opt = hvd.DistributedOptimizer(torch.optim.SGD(m.parameters(), lr=lr),
                                       named_parameters=m.named_parameters(), **kw)
        hvd.broadcast_parameters(m.state_dict(), 0)
        F.mse_loss(m(xs), ys).backward()
        opt.step()
        opt.zero_grad()
        return m, opt

    ref = _model(0)
    ropt = torch.optim.SGD(ref.parameters(), lr=0.1)
    F.mse_loss(ref(X), Y).backward()
    ropt.step()
    want = [p.detach().clone() for p in ref.parameters()]

    m, opt = run(num_groups=2)
    assert len(opt.bucket_plan()) >= 2
    for a, b in zip(m.parameters(), want):
        torch.testing.assert_close(a, b, rtol=2e-5, atol=2e-6)
    mg = _model(0)
    ps = list(mg.parameters())
    m, opt = run(gradient_predivide_factor=2.0)
    for a, b in zip(m.parameters(), want):
        torch.testing.assert_close(a, b, rtol=2e-5, atol=2e-6)
    m, opt = run(compression=hvd.Compression.fp16)
    for a, b in zip(m.parameters(), want):
        torch.testing.assert_close(a, b, rtol=2e-3, atol=2e-3)
    # op=Sum with lr/world == Average with lr
    m, opt = run(op=hvd.Sum, lr=0.1 / world)
    for a, b in zip(m.parameters(), want):
        torch.testing.assert_close(a, b, rtol=2e-5, atol=2e-6)
    # explicit groups: first two params share a bucket, the rest are planned automatically
    m = _model(0)
    ps = list(m.parameters())
    opt = hvd.DistributedOptimizer(torch.optim.SGD(m.parameters(), lr=0.1),
                                   named_parameters=m.named_parameters(), groups=[[ps[0], ps[1]]])
    plan = opt.bucket_plan()
    assert {s.name for s in plan[0].slots} == {"0.weight", "0.bias"} and len(plan) == 2
    try:
        hvd.DistributedOptimizer(torch.optim.SGD(m.parameters(), lr=0.1), num_groups=2, groups=[[ps[0]]])
        raise RuntimeError("num_groups+groups should raise")
    except ValueError:
        pass
    return True
