#!/usr/bin/env bash
# 1-GPU validation pass: env probe, single-GPU tests, short bench (ours + NCCL stand-in), launch list.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_single.log) 2>&1
echo "== env"; nvidia-smi -L; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total --format=csv
python -c "import torch;print(torch.__version__, torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))"
echo "== build"; python -c "import __graft_entry__ as g; g.build()" || exit 1
echo "== pytest gpu (single)"; timeout 900 python -m pytest tests/test_gpu_single.py -x -q -m gpu 2>&1 | tail -30
echo "== bench ours"; timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 2>&1 | tail -5
echo "== bench nccl_standin"; timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --impl nccl_standin 2>&1 | tail -5
echo "== done"
