#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/r2c7.txt
: > $out
L=wild_visual_navigation_b200
echo "== attention correctness impl 3" >> $out
WVN_ATTN_IMPL=3 timeout 300 python -m pytest tests/test_kernels_gpu.py -q --tb=short -x -k "attention" 2>&1 | tail -3 >> $out
echo "== attention microbench (B=32)" >> $out
for v in "1 2 1" "3 2 1" "3 1 1" "3 3 1" "3 0 1" "3 9 1"; do
  set -- $v
  B=32 WVN_ATTN_IMPL=$1 WVN_ATTN_POLY=$2 WVN_ATTN_TOKEN=$3 timeout 200 python scripts/bench_attention.py 2>&1 | tail -1 | sed "s/^/impl=$1 token=$3 /" >> $out
done
echo "== kmeans + bench-path tests" >> $out
timeout 900 python -m pytest tests/test_path_gpu.py tests/test_bench_path_gpu.py -q --tb=short -s -k "kmeans or bench_step or d90 or vit_base_logit" 2>&1 | grep -E "passed|failed|agreement|rel_l2|bench step|vit-b|D=90|Error|assert" | head -40 >> $out
echo "== full gpu test suite" >> $out
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -25 >> $out
echo "== bench (5 steps, no extras)" >> $out
for v in "1" "3"; do
  WVN_ATTN_IMPL=$v timeout 600 python bench.py --steps 5 --warmup 3 --cpu-frames 0 --no-extras --profile-gemm > gpurun_out/bench_c7_$v.json 2> gpurun_out/bench_c7_$v.err
  python - <<PY >> $out
import json
try:
    d=json.load(open('gpurun_out/bench_c7_$v.json'))
    r=d['roofline']
    print('impl=$v fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'attn_ms', round(r['avg_launch_ms'],4), 'frac', round(r['frac'],3), 'attn share', round(r['share_of_step'],3), 'launches/step', d['gpu_launches_per_step'], 'clocks', d['clocks'])
except Exception as e:
    print('impl=$v bench failed', e); print(open('gpurun_out/bench_c7_$v.err').read()[-1500:])
PY
done
echo "== launch list (impl default)" >> $out
WVN_ATTN_IMPL=3 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r02c.csv \
  python bench.py --steps 1 --warmup 1 --profile-only > gpurun_out/launch_r02c.log 2>&1
tail -1 gpurun_out/launch_r02c.log >> $out
python scripts/summarize_ncu.py --launches gpurun_out/launches_r02c.csv >> $out 2>&1
cat $out
