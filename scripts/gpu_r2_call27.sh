#!/bin/bash
cd "$(dirname "$0")/.."
out=gpurun_out/r2c27.txt
: > $out
L=wild_visual_navigation_b200
echo "== correctness impl 7" >> $out
WVN_ATTN_IMPL=7 timeout 300 python -m pytest tests/test_kernels_gpu.py -q --tb=short -k "attention" 2>&1 | tail -3 >> $out
WVN_ATTN_IMPL=7 timeout 300 python -m pytest tests/test_path_gpu.py -q --tb=short -k "vit" 2>&1 | tail -3 >> $out
echo "== microbench" >> $out
for impl in 5 7; do
  for std in 1.2 1.8; do
  QK_STD=$std B=32 WVN_ATTN_IMPL=$impl timeout 200 python scripts/bench_attention.py 2>&1 | grep -E "poly=" | tail -1 | sed "s/^/impl=$impl std=$std /" >> $out
  done
done
for poly in 1 3; do
  B=32 WVN_ATTN_IMPL=7 WVN_ATTN_POLY=$poly timeout 200 python scripts/bench_attention.py 2>&1 | grep -E "poly=" | tail -1 | sed "s/^/impl=7 poly=$poly /" >> $out
done
echo "== timing impl 7" >> $out
B=32 WVN_B200_LIB=$L/libwvn_b200_timing.so WVN_ATTN_TIMING=1 WVN_ATTN_IMPL=7 timeout 200 python scripts/bench_attention.py 2>&1 | grep -E "timing|poly=" | tail -2 >> $out
echo "== bench impl 7 / 5" >> $out
for impl in 7 5 7 5; do
WVN_ATTN_IMPL=$impl timeout 600 python bench.py --steps 8 --warmup 3 --cpu-frames 0 --no-extras > gpurun_out/bench_c27_$impl.json 2> gpurun_out/bench_c27_$impl.err
python - <<PY >> $out
import json
try:
    d = json.loads(open("gpurun_out/bench_c27_$impl.json").read().strip().split("\n")[-1])
    print("impl $impl fps", round(d["value"],1), "ms", round(d["ms_per_step"],3), "attn_ms", round(d["roofline"]["avg_launch_ms"],4), "frac", round(d["roofline"]["frac"],3), "clocks", d["clocks"]["sm_mhz"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_c27_$impl.err").read()[-1500:])
PY
done
cat $out
