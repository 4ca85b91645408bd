#!/usr/bin/env bash
# quick 1-GPU pass: numerics + bench variants + launch list
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_quick.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1 || exit 1
echo "== pytest gemm/bn"; timeout 600 python -m pytest tests/test_gpu_gemm.py -q -m gpu 2>&1 | tail -15
echo "== bench ours (gemm 1x1)"; timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-e2e 2>&1 | grep -E '^\{' | python -c "import sys,json; [print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}) for d in map(json.loads, sys.stdin)]"
echo "== bench ours (cudnn 1x1)"; B200DP_CONV1X1_GEMM=0 timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-e2e 2>&1 | grep -E '^\{' | python -c "import sys,json; [print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}) for d in map(json.loads, sys.stdin)]"
echo "== launch list (ours, ${LL_ENV:-default})"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_ours.csv python scripts/profile_step.py ours > gpurun_out/prof_ours.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_ours.csv 2>/dev/null | head -28
echo "== done"
