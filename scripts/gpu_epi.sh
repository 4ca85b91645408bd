#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_epi.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1 || exit 1
J='import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print({k:d.get(k) for k in ("impl","value","ms_per_step","gpu_launches")}, d["config"].get("cuda_graph"))
    elif "rror" in l or "failed" in l: print(l.strip()[:300])'
echo "== pytest gemm all"; timeout 240 python -m pytest tests/test_gpu_gemm.py -q -m gpu -x 2>&1 | tail -6
echo "== vit ours (split qkv)"; timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-e2e --model vit_b_16 --batch 128 2>&1 | python -c "$J"
echo "== vit ours (packed qkv)"; B200DP_SPLIT_QKV=0 timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-e2e --model vit_b_16 --batch 128 2>&1 | python -c "$J"
echo "== resnet50 ours"; timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-e2e 2>&1 | python -c "$J"
echo "== done"
