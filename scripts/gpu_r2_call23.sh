#!/bin/bash
cd "$(dirname "$0")/.."
for std in 1.2 1.8 2.5; do
  for thr in 8 16 32; do
    QK_STD=$std B=32 WVN_ATTN_RESCALE=$thr timeout 200 python scripts/bench_attention.py 2>&1 | grep -E "poly=" | tail -1 | sed "s/^/std=$std thr=$thr /"
  done
done
for thr in 8 16 32; do
WVN_ATTN_RESCALE=$thr timeout 600 python bench.py --steps 8 --warmup 3 --cpu-frames 0 --no-extras > gpurun_out/bench_c23_$thr.json 2> gpurun_out/bench_c23.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_c23_$thr.json").read().strip().split("\n")[-1])
print("thr $thr fps", round(d["value"],1), "ms", round(d["ms_per_step"],3), "attn_ms", round(d["roofline"]["avg_launch_ms"],4), "frac", round(d["roofline"]["frac"],3), "clocks", d["clocks"]["sm_mhz"])
PY
done
