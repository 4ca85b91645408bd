#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
N=$(python -c "import torch;print(torch.cuda.device_count())")
exec > >(tee gpurun_out/gpu_final_$N.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1 || exit 1
J='import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(json.dumps({k:d.get(k) for k in ("impl","n_gpus","value","ms_per_step","gpu_launches","e2e")}), d["config"].get("cuda_graph"), d["config"].get("comm"))
    elif "rror" in l or "failed" in l: print(l.strip()[:300])'
echo "== pytest comm subset at world $N"; timeout 600 python -m pytest tests/test_gpu_comm.py -q -m gpu -x -k "allreduce_matches or (fused_optimizer and sgd-bf16-nvls) or (fused_optimizer and adam-fp32-oneshot)" 2>&1 | tail -5
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== bench ours N=$N"; timeout 400 $TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 2>&1 | python -c "$J"
echo "== done"
