#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_path_gpu.py tests/test_kernels_gpu.py -q -x -k "vit or gemm or dino or ingest or u8" 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gemm_bf16_kernel -c 3 python bench.py --steps 1 --warmup 1 --profile-only 2>&1 | grep -E "gpu__time_duration|gemm_bf16" | head -6
timeout 600 python bench.py --steps 8 --warmup 3 --cpu-frames 0 --no-extras > gpurun_out/bench_c29.json 2> gpurun_out/bench_c29.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_c29.json").read().strip().split("\n")[-1])
print("fps", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "attn_ms", round(d["roofline"]["avg_launch_ms"],4), "frac", round(d["roofline"]["frac"],3), "clocks", d["clocks"]["sm_mhz"])
PY
