#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/r2c11.txt
: > $out
L=wild_visual_navigation_b200
echo "== v3 phase timing + kmeans timing" >> $out
B=2 WVN_B200_LIB=$L/libwvn_b200_timing.so WVN_ATTN_TIMING=1 WVN_ATTN_IMPL=3 timeout 200 python scripts/bench_attention.py 2>&1 | grep -E "timing|poly=" | tail -3 >> $out
B=32 WVN_B200_LIB=$L/libwvn_b200_timing.so WVN_ATTN_TIMING=1 WVN_ATTN_IMPL=3 timeout 200 python scripts/bench_attention.py 2>&1 | grep -E "timing|poly=" | tail -3 >> $out
WVN_B200_LIB=$L/libwvn_b200_timing.so timeout 300 python - > gpurun_out/km_timing.log 2>&1 <<'PY'
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
grep -E "kmeans timing|Error|error" gpurun_out/km_timing.log | tail -2 >> $out
grep -E "gemm timing" gpurun_out/km_timing.log | sort | uniq -c | sort -rn | head -8 >> $out
grep -E "pixel_head timing" gpurun_out/km_timing.log | tail -1 >> $out
echo "== kmeans tests" >> $out
timeout 600 python -m pytest tests/test_path_gpu.py -q --tb=short -k "kmeans" 2>&1 | tail -2 >> $out
echo "== bench (8 steps)" >> $out
timeout 600 python bench.py --steps 8 --warmup 3 --cpu-frames 0 --no-extras > gpurun_out/bench_c11.json 2> gpurun_out/bench_c11.err
python - <<PY >> $out
import json
try:
    d=json.load(open('gpurun_out/bench_c11.json'))
    r=d['roofline']
    print('fps', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'attn_ms', round(r['avg_launch_ms'],4), 'frac', round(r['frac'],3), 'clocks', d['clocks'])
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/bench_c11.err').read()[-1500:])
PY
cat $out
