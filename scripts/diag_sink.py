import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["B200DP_FUSED_SINGLE"] = "1"
import torch, torch.nn.functional as F
import distributed_torch_horovod_gcp_b200.torch as hvd
from distributed_torch_horovod_gcp_b200.models import resnet18, resnet50
from distributed_torch_horovod_gcp_b200.ops import grad_sink
hvd.init()
res = []
mk = resnet50 if len(sys.argv) > 1 and sys.argv[1] == "50" else resnet18
for enabled in ((True, True) if os.environ.get('SAME') else (True, False)):
    grad_sink._ENABLED = enabled
    torch.manual_seed(0)
    model = mk(num_classes=10).cuda().to(torch.bfloat16).to(memory_format=torch.channels_last)
    opt = hvd.DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9),
                                   named_parameters=model.named_parameters(), backward_passes_per_step=2)
    x = torch.randn(8, 3, 64, 64, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 10, (8,), device="cuda")
    loss = F.cross_entropy(model(x).float(), y)
    loss.backward()
    torch.cuda.synchronize()
    res.append({n: p.grad.detach().float().clone() for n, p in model.named_parameters()})
    loss = F.cross_entropy(model(x).float(), y)
    loss.backward()
    opt.step(); opt.zero_grad()
    torch.cuda.synchronize()
    opt.remove_hooks()
a, b = res
for n in a:
    d = (a[n] - b[n]).norm() / b[n].norm().clamp_min(1e-9)
    if d > 1e-3:
        print(f"{n:40s} rel={d:.4f} |a|={a[n].norm():.4f} |b|={b[n].norm():.4f} shape={tuple(a[n].shape)}")
print("done")
