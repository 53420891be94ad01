#!/bin/bash
# iteration script (rewritten per experiment)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 8 --warmup 3 --cpu-frames 0 2> gpurun_out/bench_iter.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'attn TF', round(r['achieved'],1))
"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:segment_ -c 6 python bench.py --steps 1 --warmup 0 --profile-only 2>&1 | grep -E "segment_[a-z_]*kernel|gpu__time_duration" | sed 's/(const.*//' | head -12
