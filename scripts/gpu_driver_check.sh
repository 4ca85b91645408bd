#!/usr/bin/env bash
# What the driver runs at round end, on 1 GPU: pytest -m gpu, smoke(), bench.py (both arms).
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_driver_check.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1 || exit 1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -6
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench reference arm"; timeout 120 python bench.py --impl reference --gpus 1 --steps 5 --warmup 3
echo "== bench default"; timeout 600 python bench.py 2>&1 | grep -E '^\{|rror'
echo "== loaded libs during bench"; python - <<'PY'
import os, subprocess, sys
os.environ["B200DP_FUSED_SINGLE"]="1"
import torch
import distributed_torch_horovod_gcp_b200.torch as hvd
from distributed_torch_horovod_gcp_b200.ops import kernels
hvd.init(); print("kernels:", kernels.has("gemm"), kernels.has("bn_act"), kernels.has("layer_norm"), kernels.has("lstm_fused"))
print([l.split()[-1] for l in open("/proc/self/maps") if "libb200dp" in l and "r-xp" in l])
PY
echo "== done"
