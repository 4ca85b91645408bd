#!/usr/bin/env bash
# Round-2 multi-GPU pass (N = visible GPUs): new comm/engine tests, bench (ours vs NCCL stand-in measured in the
# same run) for ResNet-50 / the reference LSTM / ViT-B/16, all-reduce sweep, reference app through the launcher.
set -u
mkdir -p gpurun_out
N=$(python -c "import torch;print(torch.cuda.device_count())")
exec > >(tee gpurun_out/r2_multi_$N.log) 2>&1
echo "GPUs: $N"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== pytest comm (round-2 additions)"
timeout 600 python -m pytest tests/test_gpu_comm.py -q -m gpu -k "native_collectives or sync_bn or fused_engine_cuda_graph or compressed or allreduce_large or init_shutdown or model_to" 2>&1 | tail -6
echo "== bench resnet50 N=$N"; timeout 400 $TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 2>&1 | tail -1
echo "== bench lstm N=$N"; timeout 300 $TR --master-port 29512 bench.py --gpus $N --model lstm --steps 100 --warmup 10 2>&1 | tail -1
echo "== bench vit_b_16 N=$N"; timeout 300 $TR --master-port 29513 bench.py --gpus $N --model vit_b_16 --batch 128 --steps 10 --warmup 3 --no-baseline 2>&1 | tail -1
echo "== sweep"; timeout 400 $TR --master-port 29514 benchmarks/allreduce_sweep.py --algos auto,nvls,nccl --out gpurun_out/r2_allreduce_sweep_$N.json 2>&1 | grep -E '^\{|rror' | tail -14
echo "== LSTM reference config via launcher (-np $N), CUDA graph"
( cd gpurun_out && B200DP_OFFLINE=1 B200DP_SYNTH_ROWS=20000 timeout 300 ../bin/horovodrun -np $N -H localhost:$N python ../app/torch_train.py --epochs $((N*3)) --cuda-graph 2>&1 | grep -E "epoch: (0|2),|avg_time|total training|rror" | head -12 )
echo "== done"
