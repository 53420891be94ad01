#!/bin/bash
cd "$(dirname "$0")/.."
echo "== full GPU suite"
timeout 2400 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -6
echo "== bench"
timeout 600 python bench.py --steps 8 --warmup 3 --cpu-frames 0 --no-extras > gpurun_out/bench_c21.json 2> gpurun_out/bench_c21.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_c21.json").read().strip().split("\n")[-1])
print("fps", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "attn_ms", round(d["roofline"]["avg_launch_ms"],4), "frac", round(d["roofline"]["frac"],3), "launches", d["gpu_launches_per_step"], "clocks", d["clocks"])
PY
