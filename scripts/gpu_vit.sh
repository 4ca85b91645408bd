#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_vit.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1 || exit 1
J='import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print({k:d.get(k) for k in ("impl","value","ms_per_step","gpu_launches")}, d["config"].get("cuda_graph"))
    elif "rror" in l or "failed" in l: print(l.strip()[:300])'
echo "== pytest gemm/bn/ln"; timeout 600 python -m pytest tests/test_gpu_gemm.py -q -m gpu 2>&1 | tail -8
echo "== pytest single"; timeout 600 python -m pytest tests/test_gpu_single.py -q -m gpu 2>&1 | tail -8
echo "== vit ours"; timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-e2e --model vit_b_16 --batch 128 2>&1 | python -c "$J"
echo "== resnet50 ours"; timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-e2e 2>&1 | python -c "$J"
echo "== ncu full (resnet step)"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gemm_bf16_kernel|bn_|allreduce_oneshot|maxpool|stem_im2col" -s 6 -c 22 -f -o gpurun_out/ncu_resnet python scripts/profile_step.py ours > gpurun_out/ncu_resnet.log 2>&1; tail -1 gpurun_out/ncu_resnet.log
echo "== ncu full (vit step): gemm + ln"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gemm_bf16_kernel|ln_" -s 8 -c 10 -f -o gpurun_out/ncu_vit python scripts/profile_vit.py ours > gpurun_out/ncu_vit.log 2>&1; tail -1 gpurun_out/ncu_vit.log
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
echo "== done"
