#!/usr/bin/env bash
# 1-GPU pass: tcgen05 GEMM + fused BN numerics, single-GPU tests, bench (ours vs NCCL stand-in).
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_kernels.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
echo "== pytest gemm/bn"; timeout 600 python -m pytest tests/test_gpu_gemm.py -q -m gpu -x 2>&1 | tail -40
echo "== pytest single"; timeout 600 python -m pytest tests/test_gpu_single.py -q -m gpu 2>&1 | tail -15
echo "== bench ours"; timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 2>&1 | tail -4
echo "== bench ours (cudnn 1x1)"; B200DP_CONV1X1_GEMM=0 timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-e2e 2>&1 | tail -2
echo "== done"
