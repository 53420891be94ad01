#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/r2c10.txt
: > $out
echo "== kmeans phase timing" >> $out
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
grep -E "kmeans timing|Error|error" gpurun_out/km_timing.log | tail -3 >> $out
echo "== tests" >> $out
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -8 >> $out
echo "== bench (8 steps, no extras)" >> $out
timeout 600 python bench.py --steps 8 --warmup 3 --cpu-frames 0 --no-extras --profile-gemm > gpurun_out/bench_c10.json 2> gpurun_out/bench_c10.err
python - <<PY >> $out
import json
try:
    d=json.load(open('gpurun_out/bench_c10.json'))
    r=d['roofline']
    print('fps', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'attn_ms', round(r['avg_launch_ms'],4), 'frac', round(r['frac'],3), 'attn share', round(r['share_of_step'],3), 'gemm share', r['gemm_share_of_step'], 'clocks', d['clocks'])
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/bench_c10.err').read()[-1500:])
PY
echo "== launch list" >> $out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r02d.csv \
  python bench.py --steps 1 --warmup 1 --profile-only > gpurun_out/launch_r02d.log 2>&1
python scripts/summarize_ncu.py --launches gpurun_out/launches_r02d.csv 2>&1 | head -40 >> $out
cat $out
