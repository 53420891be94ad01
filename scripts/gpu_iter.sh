#!/bin/bash
# iteration script (rewritten per experiment)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for pg in "" "--profile-gemm" "" "--profile-gemm"; do
timeout 600 python bench.py --steps 10 --warmup 3 --cpu-frames 0 $pg 2> gpurun_out/bench_iter.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('[$pg] fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'attn TF', round(r['achieved'],1), 'attn share', round(r['share_of_step'],3), 'gemm share', r['gemm_share_of_step'])
"
done
