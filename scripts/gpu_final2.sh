#!/usr/bin/env bash
# 2-GPU final validation: full comm test file + bench ours/stand-in at N=2
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
echo "== pytest comm (all) at world $N"; timeout 900 python -m pytest tests/test_gpu_comm.py -q -m gpu 2>&1 | tail -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== bench ours N=$N"; timeout 400 $TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 2>&1 | python -c "$J"
echo "== bench nccl_standin N=$N"; timeout 400 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --impl nccl_standin --no-e2e 2>&1 | python -c "$J"
echo "== sweep N=$N"; timeout 300 $TR --master-port 29513 benchmarks/allreduce_sweep.py --max-bytes $((256<<20)) --algos auto,nccl --out gpurun_out/allreduce_sweep_$N.json 2>&1 | grep -E '^\{|rror' | tail -10
echo "== done"
