#!/bin/bash
cd "$(dirname "$0")/.."
run() { echo "== racecheck: $*"; timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 6 python -m pytest "$@" -x -q 2>&1 | grep -E "RACECHECK SUMMARY|hazard|Race reported|passed|failed" | sort | uniq -c | sort -rn | head -8; }
run tests/test_slic_gpu.py -k "bit_exact and 96"
run tests/test_footprint_gpu.py -k "golden"
run tests/test_confidence_methods_gpu.py -k "running"
run tests/test_path_gpu.py -k "kmeans_kernel"
