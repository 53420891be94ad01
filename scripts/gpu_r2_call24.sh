#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_path_gpu.py -q -x -k "attention or vit" 2>&1 | tail -3
for i in 1 2; do
timeout 600 python bench.py --steps 8 --warmup 3 --cpu-frames 0 --no-extras > gpurun_out/bench_c24_$i.json 2> gpurun_out/bench_c24.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_c24_$i.json").read().strip().split("\n")[-1])
print("fps", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "attn_ms", round(d["roofline"]["avg_launch_ms"],4), "frac", round(d["roofline"]["frac"],3), "clocks", d["clocks"]["sm_mhz"])
PY
done
