#!/bin/bash
cd "$(dirname "$0")/.."
echo "== gemm tests"
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm" 2>&1 | tail -3
echo "== gemm timing (timing lib)"
REPS=1 WVN_B200_LIB=wild_visual_navigation_b200/libwvn_b200_timing.so timeout 300 python scripts/gemm_timing.py 2>&1 | grep "^\[gemm" | awk 'NR%3==0' 
echo "== gemm speed (shipped lib, 20 reps)"
REPS=20 timeout 300 python scripts/gemm_timing.py 2>&1 | tail -4
echo "== bench"
timeout 600 python bench.py --steps 8 --warmup 3 --cpu-frames 0 --no-extras > gpurun_out/bench_c19.json 2> gpurun_out/bench_c19.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_c19.json").read().strip().split("\n")[-1])
print("fps", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "attn_ms", round(d["roofline"]["avg_launch_ms"],4), "frac", round(d["roofline"]["frac"],3), "clocks", d["clocks"])
PY
