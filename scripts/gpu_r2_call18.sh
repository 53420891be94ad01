#!/bin/bash
cd "$(dirname "$0")/.."
echo "== gemm timing (timing lib)"
REPS=1 WVN_B200_LIB=wild_visual_navigation_b200/libwvn_b200_timing.so timeout 300 python scripts/gemm_timing.py 2>&1 | grep -v "^\[gemm" | tail -4
REPS=1 WVN_B200_LIB=wild_visual_navigation_b200/libwvn_b200_timing.so timeout 300 python scripts/gemm_timing.py 2>&1 | grep "^\[gemm" | awk 'NR%3==0' 
