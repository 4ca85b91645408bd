"""Run warm-up steps, then exactly ONE training step between cudaProfilerStart/Stop so that
`ncu --profile-from-start off` captures one step's kernels (never a bench number)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

impl = sys.argv[1] if len(sys.argv) > 1 else "ours"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
if impl != "ours":
    os.environ["B200DP_REFERENCE_OPS"] = "1"
os.environ.setdefault("B200DP_FUSED_SINGLE", "1")
import distributed_torch_horovod_gcp_b200.torch as hvd
from distributed_torch_horovod_gcp_b200.models import resnet50

hvd.init()
torch.cuda.set_device(0)
torch.backends.cudnn.benchmark = True
dev = torch.device("cuda", 0)
m = resnet50().to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last)
opt = torch.optim.SGD(m.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
if impl == "ours":
    opt = hvd.DistributedOptimizer(opt, named_parameters=m.named_parameters())
x = torch.randn(batch, 3, 224, 224, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
y = torch.randint(0, 1000, (batch,), device=dev)


def step():
    loss = F.cross_entropy(m(x).float(), y)
    loss.backward()
    opt.step()
    opt.zero_grad()


for _ in range(4):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step", impl)
