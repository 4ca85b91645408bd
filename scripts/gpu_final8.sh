#!/usr/bin/env bash
# 8-GPU final numbers with the final kernels (ours only; the stand-in was measured in run_8gpu.log)
set -u
mkdir -p gpurun_out
N=$(python -c "import torch;print(torch.cuda.device_count())")
exec > >(tee gpurun_out/gpu_final_$N.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1 || exit 1
J='import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(json.dumps({k:d.get(k) for k in ("impl","n_gpus","value","ms_per_step","gpu_launches","e2e","clocks")}), d["config"].get("model"), d["config"].get("cuda_graph"), d["config"].get("comm"))
    elif "rror" in l or "failed" in l: print(l.strip()[:300])'
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== resnet50 ours N=$N"; timeout 300 $TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 2>&1 | python -c "$J"
echo "== resnet152 ours N=$N"; timeout 300 $TR --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --model resnet152 --batch 128 2>&1 | python -c "$J"
echo "== vit_b_16 ours N=$N"; timeout 300 $TR --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --model vit_b_16 --batch 128 2>&1 | python -c "$J"
echo "== done"
