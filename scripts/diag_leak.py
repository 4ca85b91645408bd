"""torchrun --nproc-per-node 2 scripts/diag_leak.py : free device memory around init/shutdown cycles."""
import os, sys, time, torch
sys.path.insert(0, "/root/repo")
import distributed_torch_horovod_gcp_b200.torch as hvd
from distributed_torch_horovod_gcp_b200 import _state
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
def free():
    torch.cuda.synchronize()
    return torch.cuda.mem_get_info(dev)[0] >> 20
print(f"[{os.environ.get('RANK')}] start free={free()}", flush=True)
for cycle in range(4):
    hvd.init()
    r = hvd.rank()
    f1 = free()
    m = torch.nn.Linear(512, 512).to(dev)
    opt = hvd.DistributedOptimizer(torch.optim.SGD(m.parameters(), lr=0.1), named_parameters=m.named_parameters())
    for _ in range(2):
        m(torch.randn(4, 512, device=dev)).sum().backward(); opt.step(); opt.zero_grad()
    t = torch.full((1024,), float(r), device=dev)
    hvd.allreduce(t, op=hvd.Sum)
    f2 = free()
    s = _state._RT.symm
    nb = len(s.buffers) if s is not None else -1
    sizes = [b.padded >> 20 for b in s.buffers] if s is not None else []
    opt.remove_hooks()
    hvd.shutdown()
    f3 = free()
    time.sleep(1.0)
    f4 = free()
    print(f"[{r}] cycle {cycle}: after init {f1} MiB, after train {f2}, buffers {nb} {sizes}, after shutdown {f3}, +1s {f4}, torch reserved {torch.cuda.memory_reserved(dev) >> 20}", flush=True)
