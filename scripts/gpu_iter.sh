#!/bin/bash
# iteration script (rewritten per experiment)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for ing in f32 u8; do
timeout 600 python bench.py --steps 8 --warmup 3 --cpu-frames 0 --ingest $ing 2> gpurun_out/bench_iter.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('ingest $ing fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'h2d', d['e2e']['h2d_bytes_per_step'], 'attn TF', round(r['achieved'],1), 'traffic', r['traffic'])
"
done
tail -3 gpurun_out/bench_iter.err
