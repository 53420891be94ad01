#!/bin/bash
cd "$(dirname "$0")/.."
echo "== kmeans tests"
timeout 900 python -m pytest tests/test_path_gpu.py tests/test_bench_path_gpu.py -q -x -k "kmeans or bench_step" 2>&1 | tail -3
echo "== kmeans timing"
WVN_B200_LIB=wild_visual_navigation_b200/libwvn_b200_timing.so timeout 300 python - > gpurun_out/km_timing.log 2>&1 <<'PY'
import torch, bench
from wild_visual_navigation_b200.feature_extractor import FeatureExtractor
cfg, sd, hd = bench.make_weights()
fe = FeatureExtractor("cuda", segmentation_type="stego", feature_type="dino", input_size=448, state_dict=sd, head_state_dict=hd,
                      flip_tta=False, max_batch=32, chunk=32)
img = torch.rand(32, 3, 448, 448).cuda()
for _ in range(2):
    r = fe.extract_batch(img)
torch.cuda.synchronize()
PY
grep -E "kmeans timing|Error|error" gpurun_out/km_timing.log | tail -2
echo "== bench"
timeout 600 python bench.py --steps 8 --warmup 3 --cpu-frames 0 --no-extras > gpurun_out/bench_c22.json 2> gpurun_out/bench_c22.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_c22.json").read().strip().split("\n")[-1])
print("fps", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "attn_ms", round(d["roofline"]["avg_launch_ms"],4), "frac", round(d["roofline"]["frac"],3), "launches", d["gpu_launches_per_step"], "clocks", d["clocks"])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:stego_kmeans -c 2 python bench.py --steps 1 --warmup 1 --profile-only 2>&1 | grep -E "gpu__time_duration|stego_kmeans" | head -6
