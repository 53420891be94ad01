#!/bin/bash
# round 2, call 1: attention v2 correctness + variant sweep + phase timing + short bench A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/r2c1.txt
: > $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv >> $out 2>&1
for impl in 2 1; do
  echo "== kernel tests attention IMPL=$impl" >> $out
  WVN_ATTN_IMPL=$impl timeout 300 python -m pytest tests/test_kernels_gpu.py -q --tb=short -x -k "attention" 2>&1 | tail -5 >> $out
done
echo "== gemm + others" >> $out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q --tb=short -x -k "not attention" 2>&1 | tail -3 >> $out
echo "== path tests" >> $out
timeout 900 python -m pytest tests/test_path_gpu.py -q --tb=short -x 2>&1 | tail -3 >> $out
echo "== attention microbench (B=32)" >> $out
for v in "1 2 1" "2 2 1" "2 2 0" "2 0 1" "2 1 1" "2 3 1" "2 4 1" "2 9 1" "1 9 1"; do
  set -- $v
  B=32 WVN_ATTN_IMPL=$1 WVN_ATTN_POLY=$2 WVN_ATTN_TOKEN=$3 timeout 200 python scripts/bench_attention.py 2>&1 | tail -1 | sed "s/^/impl=$1 token=$3 /" >> $out
done
echo "== phase timing (timing lib, B=2)" >> $out
for v in "1 1" "2 1" "2 0"; do
  set -- $v
  B=2 WVN_B200_LIB=wild_visual_navigation_b200/libwvn_b200_timing.so WVN_ATTN_TIMING=1 WVN_ATTN_IMPL=$1 WVN_ATTN_TOKEN=$2 timeout 200 python scripts/bench_attention.py 2>&1 | grep -E "timing|poly=" | tail -3 | sed "s/^/impl=$1 token=$2 /" >> $out
done
echo "== bench A/B (5 steps)" >> $out
for impl in 1 2; do
  WVN_ATTN_IMPL=$impl timeout 600 python bench.py --steps 5 --warmup 3 --cpu-frames 0 > gpurun_out/bench_impl$impl.json 2> gpurun_out/bench_impl$impl.err
  python - <<PY >> $out
import json
try:
    d=json.load(open('gpurun_out/bench_impl$impl.json'))
    print('impl=$impl fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'attn_ms', round(d['roofline']['avg_launch_ms'],4), 'frac', round(d['roofline']['frac'],3), 'clocks', d['clocks'])
except Exception as e:
    print('impl=$impl bench failed', e)
PY
done
cat $out
