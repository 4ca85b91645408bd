#!/bin/bash
# Round-2 evidence on ONE GPU: ncu --set full captures of the new kernels (read on the CPU box with
# scripts/ncu_summary.py) and compute-sanitizer racecheck / synccheck / memcheck logs of tiny instances.
set -x
mkdir -p gpurun_out
NCU="ncu --set full --import-source on --clock-control none"
# GEMM, K = 64 output-bound shape with the BN-statistics epilogue (layer1 conv3 of ResNet-50 at batch 256)
$NCU -k regex:gemm_bf16 -s 2 -c 1 -o gpurun_out/r2_gemm_k64_stats -f python scripts/gemm_prof.py 802816 256 64 stats 3 > /dev/null 2>&1
# 3x3 64-channel: halo fprop / dgrad + halo wgrad (conv_bench shape 0); 3x3 256-channel generic kernel (shape 4)
$NCU -k regex:conv_ -s 4 -c 1 -o gpurun_out/r2_conv_halo64_fprop -f python benchmarks/conv_bench.py --only 0 --ours-only --iters 1 > /dev/null 2>&1
$NCU -k regex:conv_halo -s 8 -c 1 -o gpurun_out/r2_conv_halo64_dgrad -f python benchmarks/conv_bench.py --only 0 --ours-only --iters 1 > /dev/null 2>&1
$NCU -k regex:conv_wgrad -s 3 -c 1 -o gpurun_out/r2_conv_halo64_wgrad -f python benchmarks/conv_bench.py --only 0 --ours-only --iters 1 > /dev/null 2>&1
$NCU -k regex:conv_bf16 -s 4 -c 1 -o gpurun_out/r2_conv256_fprop -f python benchmarks/conv_bench.py --only 4 --ours-only --iters 1 > /dev/null 2>&1
# attention forward + backward (ViT-B/16 shape), LSTM recurrence (reference config)
$NCU -k regex:attn_ -s 6 -c 3 -o gpurun_out/r2_attn -f python scripts/attn_prof.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
# sanitizers (tiny shapes)
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_kernels.py > gpurun_out/sanitizer_$tool.log 2>&1
  tail -4 gpurun_out/sanitizer_$tool.log
done
