#!/usr/bin/env bash
# N-GPU pass (N = visible GPUs): comm kernel tests vs NCCL, bench ours vs stand-in at N, sweep, launch list.
set -u
mkdir -p gpurun_out
N=$(python -c "import torch;print(torch.cuda.device_count())")
exec > >(tee gpurun_out/gpu_multi_$N.log) 2>&1
echo "GPUs: $N"; nvidia-smi topo -m 2>/dev/null | head -12
python -c "import __graft_entry__ as g; g.build()" || exit 1
echo "== pytest comm"; timeout 900 python -m pytest tests/test_gpu_comm.py -q -m gpu -x 2>&1 | tail -25
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== bench ours N=$N"; timeout 600 $TR --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 2>&1 | grep -E '^\{|Error|error' | tail -3
echo "== bench nccl_standin N=$N"; timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --impl nccl_standin --no-e2e 2>&1 | grep -E '^\{|Error|error' | tail -3
echo "== sweep"; timeout 600 $TR --master-port 29513 benchmarks/allreduce_sweep.py --max-bytes $((256<<20)) --out gpurun_out/allreduce_sweep_$N.json 2>&1 | grep -E '^\{|Error|error' | tail -12
if [ "${PROFILE:-1}" = "1" ]; then
echo "== launch list (1 GPU, ours)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_ours.csv python scripts/profile_step.py ours > gpurun_out/prof_ours.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_ours.csv | head -45
echo "== launch list (1 GPU, standin)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_standin.csv python scripts/profile_step.py standin > gpurun_out/prof_standin.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_standin.csv | head -30
fi
echo "== done"
