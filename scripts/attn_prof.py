import os, sys, torch
sys.path.insert(0, "/root/repo")
from distributed_torch_horovod_gcp_b200.ops import kernels
kernels.has("attention_fused")
B,H,S=128,12,197
mats=[(torch.randn(B*S,H*64,device="cuda")*0.5).to(torch.bfloat16).requires_grad_(True) for _ in range(3)]
q,k,v=[m.view(B,S,H,64).transpose(1,2) for m in mats]
g=torch.randn(B,S,H,64,device="cuda").to(torch.bfloat16).transpose(1,2)
for _ in range(3):
    o=kernels.attention_fused(q,k,v); o.backward(g)
torch.cuda.synchronize()
