#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/r2c2.txt
: > $out
L=wild_visual_navigation_b200
echo "== attention microbench (B=32): try_wait suspend hint sweep" >> $out
for lib in libwvn_b200.so libwvn_b200_h20.so libwvn_b200_h100.so libwvn_b200_h1000.so; do
  for v in "1 1" "2 1" "2 0"; do
    set -- $v
    B=32 WVN_B200_LIB=$L/$lib WVN_ATTN_IMPL=$1 WVN_ATTN_TOKEN=$2 timeout 200 python scripts/bench_attention.py 2>&1 | tail -1 | sed "s/^/$lib impl=$1 token=$2 /" >> $out
  done
done
echo "== phase timing (timing lib = hint 0, B=2)" >> $out
for v in "1 1" "2 1" "2 0"; do
  set -- $v
  B=2 WVN_B200_LIB=$L/libwvn_b200_timing.so WVN_ATTN_TIMING=1 WVN_ATTN_IMPL=$1 WVN_ATTN_TOKEN=$2 timeout 200 python scripts/bench_attention.py 2>&1 | grep -E "timing|poly=" | tail -2 | sed "s/^/impl=$1 token=$2 /" >> $out
done
echo "== v2 poly sweep, default lib" >> $out
for p in 0 1 3 9; do
  B=32 WVN_ATTN_IMPL=2 WVN_ATTN_POLY=$p timeout 200 python scripts/bench_attention.py 2>&1 | tail -1 >> $out
done
cat $out
