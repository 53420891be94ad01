#!/bin/bash
cd "$(dirname "$0")/.."
for sub in 0 16 8 0 16; do
WVN_VIT_SUBCHUNK=$sub timeout 600 python bench.py --steps 8 --warmup 3 --cpu-frames 0 --no-extras > gpurun_out/bench_c31.json 2> gpurun_out/bench_c31.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_c31.json").read().strip().split("\n")[-1])
print("sub $sub fps", round(d["value"],1), "ms", round(d["ms_per_step"],3), "attn_ms", round(d["roofline"]["avg_launch_ms"],4), "clocks", d["clocks"]["sm_mhz"])
PY
done
