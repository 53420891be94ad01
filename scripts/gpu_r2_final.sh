#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash scripts/gpu_profile.sh r02f 2>&1 | tail -12
echo "== bench full (N=1)"
timeout 900 python bench.py > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; tail -c 300 gpurun_out/bench_final_n1.json
echo "== bench c5"
timeout 900 python bench.py --config c5 --steps 3 --warmup 3 --cpu-frames 0 --no-extras > gpurun_out/bench_final_c5.json 2> gpurun_out/bench_final_c5.err; tail -c 200 gpurun_out/bench_final_c5.json
