#!/usr/bin/env bash
# 1-GPU pass: numerics, bench, launch list, ncu --set full captures of the top kernels.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_kernels2.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
echo "== pytest gemm/bn"; timeout 600 python -m pytest tests/test_gpu_gemm.py -q -m gpu 2>&1 | tail -25
echo "== pytest single"; timeout 600 python -m pytest tests/test_gpu_single.py -q -m gpu 2>&1 | tail -6
echo "== bench ours"; timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 2>&1 | grep -E '^\{' | tail -2
echo "== bench ours (no stats fusion)"; B200DP_FUSE_BN_STATS=0 timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-e2e 2>&1 | grep -E '^\{' | tail -1
echo "== bench ours (cudnn 1x1)"; B200DP_CONV1X1_GEMM=0 timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-e2e 2>&1 | grep -E '^\{' | tail -1
echo "== launch list (ours)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_ours.csv python scripts/profile_step.py ours > gpurun_out/prof_ours.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_ours.csv 2>/dev/null | head -40
echo "== ncu full: gemm"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_bf16_kernel -c 6 -f -o gpurun_out/ncu_gemm python scripts/profile_step.py ours > gpurun_out/ncu_gemm.log 2>&1; tail -2 gpurun_out/ncu_gemm.log
echo "== ncu full: bn"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:bn_ -c 8 -f -o gpurun_out/ncu_bn python scripts/profile_step.py ours > gpurun_out/ncu_bn.log 2>&1; tail -2 gpurun_out/ncu_bn.log
ls -la gpurun_out/*.ncu-rep
echo "== done"
