"""One GEMM (optionally with the BN-statistics epilogue) for an ncu capture:
    python scripts/gemm_prof.py M N K [stats] [reps]
"""
import sys, torch
sys.path.insert(0, "/root/repo")
from distributed_torch_horovod_gcp_b200.ops import kernels, gemm as G
assert kernels.has("conv_implicit_gemm")
M, N, K = (int(v) for v in sys.argv[1:4])
want_stats = len(sys.argv) > 4 and sys.argv[4] == "stats"
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
b = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
stats = torch.zeros(2 * N, device="cuda") if want_stats else None
for _ in range(reps):
    G.gemm(a, b, out, M, N, K, stats=stats)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    G.gemm(a, b, out, M, N, K, stats=stats)
e1.record(); torch.cuda.synchronize()
print(f"M{M} N{N} K{K} stats={want_stats}: {e0.elapsed_time(e1) / 5 * 1e3:.1f} us/call")
