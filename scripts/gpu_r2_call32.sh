#!/bin/bash
cd "$(dirname "$0")/.."
echo "== gemm tests, pair kernel on for every epilogue"
WVN_GEMM_PAIR=15 timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm" 2>&1 | tail -3
WVN_GEMM_PAIR=15 timeout 900 python -m pytest tests/test_path_gpu.py -q -x -k "vit" 2>&1 | tail -3
echo "== gemm speed single"
REPS=20 WVN_GEMM_PAIR=0 timeout 300 python scripts/gemm_timing.py 2>&1 | tail -4
echo "== gemm speed pair"
REPS=20 WVN_GEMM_PAIR=15 timeout 300 python scripts/gemm_timing.py 2>&1 | tail -4
