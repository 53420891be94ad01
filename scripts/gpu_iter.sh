#!/bin/bash
# iteration script (rewritten per experiment)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for lib in libwvn_b200.so libwvn_b200_spin.so libwvn_b200.so libwvn_b200_spin.so; do
  echo "== $lib"
  REPS=20 WVN_B200_LIB=$PWD/wild_visual_navigation_b200/$lib timeout 300 python scripts/gemm_timing.py 2>&1 | grep TFLOP
done
for lib in libwvn_b200.so libwvn_b200_spin.so; do
WVN_B200_LIB=$PWD/wild_visual_navigation_b200/$lib timeout 600 python bench.py --steps 8 --warmup 3 --cpu-frames 0 2> gpurun_out/bench_iter.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('$lib fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'attn TF', round(r['achieved'],1))
"
done
for lib in libwvn_b200.so libwvn_b200_spin.so; do
B=32 WVN_B200_LIB=$PWD/wild_visual_navigation_b200/$lib timeout 300 python scripts/bench_attention.py 2>&1 | tail -1 | sed "s/^/$lib /"
done
