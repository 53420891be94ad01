#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/r2c9b.txt
: > $out
echo "== kmeans phase timing" >> $out
WVN_B200_LIB=wild_visual_navigation_b200/libwvn_b200_timing.so timeout 300 python - >> $out 2>&1 <<'PY'
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
grep -E "kmeans timing" $out | tail -2 > /tmp/km.txt; grep -vE "gemm timing|pixel_head timing|attn" $out | tail -5 > /tmp/rest.txt; cat /tmp/km.txt /tmp/rest.txt > $out
echo "== bench-path test" >> $out
timeout 900 python -m pytest tests/test_bench_path_gpu.py -q --tb=short -x 2>&1 | tail -4 >> $out
echo "== N=1 default bench (all legs)" >> $out
( time timeout 1200 python bench.py > gpurun_out/bench_n1_full.json 2> gpurun_out/bench_n1_full.err ) 2>> $out
tail -c 3800 gpurun_out/bench_n1_full.json >> $out
tail -4 gpurun_out/bench_n1_full.err >> $out
echo "== reference arm, default threads" >> $out
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>> $out
python -c "
import json; d=json.load(open('gpurun_out/bench_ref.json')); print('reference arm', d['value'], d['cpu_baseline'])" >> $out 2>&1
cat $out
