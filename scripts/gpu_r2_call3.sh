#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/r2c3.txt
: > $out
L=wild_visual_navigation_b200
echo "== correctness" >> $out
for impl in 2 1; do
  WVN_ATTN_IMPL=$impl timeout 300 python -m pytest tests/test_kernels_gpu.py -q --tb=short -x -k "attention" 2>&1 | tail -2 >> $out
done
echo "== attention microbench (B=32): aux warps last (default) vs first" >> $out
for lib in libwvn_b200.so libwvn_b200_auxfirst.so; do
  for v in "1 1" "2 1" "2 0"; do
    set -- $v
    B=32 WVN_B200_LIB=$L/$lib WVN_ATTN_IMPL=$1 WVN_ATTN_TOKEN=$2 timeout 200 python scripts/bench_attention.py 2>&1 | tail -1 | sed "s/^/$lib impl=$1 token=$2 /" >> $out
  done
done
echo "== phase timing (timing lib, B=2)" >> $out
for v in "1 1" "2 1" "2 0"; do
  set -- $v
  B=2 WVN_B200_LIB=$L/libwvn_b200_timing.so WVN_ATTN_TIMING=1 WVN_ATTN_IMPL=$1 WVN_ATTN_TOKEN=$2 timeout 200 python scripts/bench_attention.py 2>&1 | grep -E "timing|poly=" | tail -2 | sed "s/^/impl=$1 token=$2 /" >> $out
done
echo "== poly sweep, default lib" >> $out
for v in "1 0" "1 3" "1 9" "2 0" "2 1" "2 3" "2 4" "2 9"; do
  set -- $v
  B=32 WVN_ATTN_IMPL=$1 WVN_ATTN_POLY=$2 timeout 200 python scripts/bench_attention.py 2>&1 | tail -1 | sed "s/^/impl=$1 /" >> $out
done
cat $out
