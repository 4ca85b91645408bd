#!/usr/bin/env bash
# quick 1-GPU pass: numerics + bench variants + launch list
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_quick.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1 || exit 1
J='import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, d["config"].get("cuda_graph"), d.get("e2e") and d["e2e"]["value"])
    elif "rror" in l or "failed" in l: print(l.strip()[:300])'
echo "== pytest gemm/bn"; timeout 600 python -m pytest tests/test_gpu_gemm.py -q -m gpu 2>&1 | tail -15
echo "== bench ours (graph auto)"; timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 2>&1 | python -c "$J"
echo "== bench ours (graph off)"; timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-e2e --graph off 2>&1 | python -c "$J"
echo "== bench ours (cudnn 1x1+stem, graph auto)"; B200DP_STEM_GEMM=0 B200DP_CONV1X1_GEMM=0 timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-e2e 2>&1 | python -c "$J"
echo "== vit_b_16 ours / standin"; timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-e2e --model vit_b_16 --batch 128 2>&1 | python -c "$J"
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-e2e --model vit_b_16 --batch 128 --impl nccl_standin 2>&1 | python -c "$J"
echo "== launch list (ours)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_ours.csv python scripts/profile_step.py ours > gpurun_out/prof_ours.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_ours.csv 2>/dev/null | head -24
echo "== done"
