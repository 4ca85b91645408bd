#!/usr/bin/env bash
# N-GPU pass (N = visible GPUs): comm tests, bench ours vs stand-in, allreduce sweep, LSTM app via the launcher.
set -u
mkdir -p gpurun_out
N=$(python -c "import torch;print(torch.cuda.device_count())")
exec > >(tee gpurun_out/gpu_multi_$N.log) 2>&1
echo "GPUs: $N"
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1 || exit 1
J='import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(json.dumps({k:d.get(k) for k in ("impl","n_gpus","value","ms_per_step","gpu_launches","e2e","clocks")}), d["config"].get("cuda_graph"), d["config"].get("comm"))
    elif "rror" in l or "failed" in l: print(l.strip()[:300])'
echo "== pytest comm (subset)"; timeout 900 python -m pytest tests/test_gpu_comm.py -q -m gpu -x -k "runtime_setup or allreduce_matches or broadcast or (fused_optimizer and (sgd-bf16-nvls or adam-fp32-oneshot or sgd_nesterov)) or lstm" 2>&1 | tail -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== bench ours N=$N"; timeout 600 $TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 2>&1 | python -c "$J"
echo "== bench nccl_standin N=$N"; timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --impl nccl_standin --no-e2e 2>&1 | python -c "$J"
echo "== sweep (default blocks)"; timeout 600 $TR --master-port 29513 benchmarks/allreduce_sweep.py --out gpurun_out/allreduce_sweep_$N.json 2>&1 | grep -E '^\{|rror' | tail -14
echo "== sweep (96 blocks, >=4MB)"; B200DP_COMM_BLOCKS=96 timeout 600 $TR --master-port 29514 benchmarks/allreduce_sweep.py --min-bytes 4194304 --algos twoshot,nvls --out gpurun_out/allreduce_sweep_${N}_b96.json 2>&1 | grep -E '^\{|rror' | tail -8
echo "== LSTM reference config via launcher (-np $N)"
( cd gpurun_out && B200DP_OFFLINE=1 B200DP_SYNTH_ROWS=20000 timeout 600 ../bin/horovodrun -np $N -H localhost:$N python ../app/torch_train.py --epochs $((N*3)) 2>&1 | grep -E "epoch: (0|2),|avg_time|total training|rror" | head -24 )
echo "== LSTM reference config, CUDA graph"
( cd gpurun_out && B200DP_OFFLINE=1 B200DP_SYNTH_ROWS=20000 timeout 600 ../bin/horovodrun -np $N -H localhost:$N python ../app/torch_train.py --epochs $((N*3)) --cuda-graph 2>&1 | grep -E "epoch: (0|2),|avg_time|total training|rror" | head -24 )
echo "== done"
