#!/bin/bash
# iteration script (rewritten per experiment)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
WVN_GEMM_PAIR=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel --launch-skip 5 --launch-count 1 -f -o gpurun_out/gemm_fc1 python scripts/gemm_timing.py 2>&1 | tail -5
WVN_GEMM_PAIR=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel --launch-skip 11 --launch-count 1 -f -o gpurun_out/gemm_proj python scripts/gemm_timing.py 2>&1 | tail -3
ls -la gpurun_out/*.ncu-rep
