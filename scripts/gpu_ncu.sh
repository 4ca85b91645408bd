#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_ncu.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1 || exit 1
echo "== ncu full (vit step): gemm 2cta"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gemm_bf16" -s 10 -c 9 -f -o gpurun_out/ncu_vit python scripts/profile_vit.py ours > gpurun_out/ncu_vit.log 2>&1; tail -1 gpurun_out/ncu_vit.log
echo "== ncu full (resnet step)"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gemm_bf16|bn_bwd|allreduce_oneshot|maxpool|stem_im2col" -s 3 -c 14 -f -o gpurun_out/ncu_resnet python scripts/profile_step.py ours > gpurun_out/ncu_resnet.log 2>&1; tail -1 gpurun_out/ncu_resnet.log
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
