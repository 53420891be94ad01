#!/bin/bash
# compute-sanitizer over the round-2 kernels (run on a GPU box: gpurun -- bash scripts/gpu_r2_sanitize.sh)
cd "$(dirname "$0")/.."
mem() { echo "== memcheck: $*"; timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 5 python -m pytest "$@" -x -q 2>&1 | grep -E "ERROR SUMMARY|Invalid|passed|failed" | tail -4; }
race() { echo "== racecheck: $*"; timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 6 python -m pytest "$@" -x -q 2>&1 | grep -E "RACECHECK SUMMARY|hazard|passed|failed" | tail -4; }
mem tests/test_slic_gpu.py -k "edge or bit_exact and not 448"
mem tests/test_footprint_gpu.py -k "golden or fused"
mem tests/test_confidence_methods_gpu.py -k "moving_average or running"
mem tests/test_path_gpu.py -k "kmeans"
mem tests/test_kernels_gpu.py -k "attention and not peaked"
race tests/test_slic_gpu.py -k "bit_exact and 96"
race tests/test_footprint_gpu.py -k "golden"
race tests/test_confidence_methods_gpu.py -k "running"
race tests/test_path_gpu.py -k "kmeans_kernel"
