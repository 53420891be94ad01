#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_bench_path_gpu.py -q -x -k "legs" 2>&1 | tail -3
echo "== bench full (N=1)"
timeout 900 python bench.py > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; tail -c 300 gpurun_out/bench_final_n1.json; tail -3 gpurun_out/bench_final_n1.err
