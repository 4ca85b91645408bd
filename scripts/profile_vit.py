"""One ViT-B/16 training step between cudaProfilerStart/Stop (for ncu launch lists)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

impl = sys.argv[1] if len(sys.argv) > 1 else "ours"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 128
if impl != "ours":
    os.environ["B200DP_REFERENCE_OPS"] = "1"
os.environ.setdefault("B200DP_FUSED_SINGLE", "1")
import distributed_torch_horovod_gcp_b200.torch as hvd
from distributed_torch_horovod_gcp_b200.models import vit_b_16

hvd.init()
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
m = vit_b_16().to(dev).to(torch.bfloat16)
opt = torch.optim.SGD(m.parameters(), lr=0.01, momentum=0.9)
if impl == "ours":
    opt = hvd.DistributedOptimizer(opt, named_parameters=m.named_parameters())
x = torch.randn(batch, 3, 224, 224, device=dev).to(torch.bfloat16)
y = torch.randint(0, 1000, (batch,), device=dev)


def step():
    loss = F.cross_entropy(m(x).float(), y)
    loss.backward()
    opt.step()
    opt.zero_grad()


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one ViT step", impl)
