#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/r2c4.txt
: > $out
L=wild_visual_navigation_b200
echo "== attention v3 correctness" >> $out
WVN_ATTN_IMPL=3 timeout 300 python -m pytest tests/test_kernels_gpu.py -q --tb=short -x -k "attention" 2>&1 | tail -3 >> $out
echo "== attention microbench (B=32)" >> $out
for lib in libwvn_b200.so libwvn_b200_auxfirst.so; do
  for v in "1 2" "2 2" "3 2" "3 0" "3 1" "3 3" "3 9"; do
    set -- $v
    B=32 WVN_B200_LIB=$L/$lib WVN_ATTN_IMPL=$1 WVN_ATTN_POLY=$2 timeout 200 python scripts/bench_attention.py 2>&1 | tail -1 | sed "s/^/$lib impl=$1 /" >> $out
  done
done
echo "== full gpu test suite (default lib)" >> $out
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -60 >> $out
echo "== bench (5 steps, no extras): attention impl 1/2/3, then gemm aux-first" >> $out
for v in "1 libwvn_b200.so" "2 libwvn_b200.so" "3 libwvn_b200.so" "1 libwvn_b200_gemmauxfirst.so"; do
  set -- $v
  WVN_B200_LIB=$L/$2 WVN_ATTN_IMPL=$1 timeout 600 python bench.py --steps 5 --warmup 3 --cpu-frames 0 --no-extras --profile-gemm > gpurun_out/bench_c4_$1_$2.json 2> gpurun_out/bench_c4_$1_$2.err
  python - <<PY >> $out
import json
try:
    d=json.load(open('gpurun_out/bench_c4_$1_$2.json'))
    r=d['roofline']
    print('impl=$1 lib=$2 fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'attn_ms', round(r['avg_launch_ms'],4), 'frac', round(r['frac'],3), 'attn share', round(r['share_of_step'],3), 'gemm share', r['gemm_share_of_step'], 'launches/step', d['gpu_launches_per_step'], 'clocks', d['clocks'])
except Exception as e:
    print('impl=$1 lib=$2 bench failed', e); print(open('gpurun_out/bench_c4_$1_$2.err').read()[-1500:])
PY
done
cat $out
