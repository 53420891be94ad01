#!/bin/bash
# iteration script (rewritten per experiment)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_path_gpu.py -x -q -k "vit or dino" 2>&1 | tail -2
for sub in 32 16 8 32 16; do
WVN_VIT_SUB_ATTN=$sub timeout 600 python bench.py --steps 8 --warmup 3 --cpu-frames 0 2> gpurun_out/bench_iter.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('sub_attn $sub fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'attn TF', round(r['achieved'],1), 'attn share', round(r['share_of_step'],3), 'gemm share', round(r['gemm_share_of_step'],3))
"
done
