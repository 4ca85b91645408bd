#!/usr/bin/env bash
# Final check on a 2-GPU box: the driver's commands + watchdog test + NCCL-fallback path.
set -u
mkdir -p gpurun_out
N=$(python -c "import torch;print(torch.cuda.device_count())")
exec > >(tee gpurun_out/gpu_final_check.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1 || exit 1
J='import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(json.dumps({k:d.get(k) for k in ("impl","n_gpus","value","ms_per_step","gpu_launches")}), d["config"].get("cuda_graph"), d["config"].get("comm"))
    elif "rror" in l or "failed" in l or "FALL BACK" in l: print(l.strip()[:300])'
echo "== pytest -m gpu (all, world $N)"; timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -5
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== NCCL-fallback path (symmetric runtime disabled) N=$N"; B200DP_DISABLE_SYMM=1 timeout 400 $TR --master-port 29515 bench.py --gpus $N --steps 5 --warmup 3 --no-e2e 2>&1 | python -c "$J"
echo "== done"
