#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/r2c5.txt
: > $out
L=wild_visual_navigation_b200
echo "== attention correctness (elect-based issue) impl 1 2 3" >> $out
for impl in 1 2 3; do
  WVN_ATTN_IMPL=$impl timeout 300 python -m pytest tests/test_kernels_gpu.py -q --tb=short -x -k "attention" 2>&1 | tail -2 >> $out
done
echo "== gemm correctness" >> $out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q --tb=short -x -k "not attention" 2>&1 | tail -2 >> $out
echo "== attention microbench (B=32)" >> $out
for v in "1 2 1" "2 2 1" "2 2 0" "3 2 1" "1 0 1" "2 0 1" "2 1 1" "2 3 1" "1 9 1" "2 9 1" "3 9 1"; do
  set -- $v
  B=32 WVN_ATTN_IMPL=$1 WVN_ATTN_POLY=$2 WVN_ATTN_TOKEN=$3 timeout 200 python scripts/bench_attention.py 2>&1 | tail -1 | sed "s/^/impl=$1 token=$3 /" >> $out
done
echo "== phase timing (timing lib, B=2)" >> $out
for v in "1 1" "2 1" "2 0"; do
  set -- $v
  B=2 WVN_B200_LIB=$L/libwvn_b200_timing.so WVN_ATTN_TIMING=1 WVN_ATTN_IMPL=$1 WVN_ATTN_TOKEN=$2 timeout 200 python scripts/bench_attention.py 2>&1 | grep -E "timing|poly=" | tail -2 | sed "s/^/impl=$1 token=$2 /" >> $out
done
echo "== full gpu test suite (default lib)" >> $out
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40 >> $out
echo "== bench (5 steps, no extras)" >> $out
for v in "1" "2"; do
  WVN_ATTN_IMPL=$v timeout 600 python bench.py --steps 5 --warmup 3 --cpu-frames 0 --no-extras --profile-gemm > gpurun_out/bench_c5_$v.json 2> gpurun_out/bench_c5_$v.err
  python - <<PY >> $out
import json
try:
    d=json.load(open('gpurun_out/bench_c5_$v.json'))
    r=d['roofline']
    print('impl=$v fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'attn_ms', round(r['avg_launch_ms'],4), 'frac', round(r['frac'],3), 'attn share', round(r['share_of_step'],3), 'gemm share', r['gemm_share_of_step'], 'launches/step', d['gpu_launches_per_step'], 'clocks', d['clocks'])
except Exception as e:
    print('impl=$v bench failed', e); print(open('gpurun_out/bench_c5_$v.err').read()[-1500:])
PY
done
echo "== launch list" >> $out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r02a.csv \
  python bench.py --steps 1 --warmup 1 --profile-only > gpurun_out/launch_r02a.log 2>&1
tail -1 gpurun_out/launch_r02a.log >> $out
python scripts/summarize_ncu.py --launches gpurun_out/launches_r02a.csv >> $out 2>&1
cat $out
