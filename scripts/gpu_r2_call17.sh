#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== full GPU suite (default attention impl 5)"
timeout 2400 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -8
echo "== launch list + attention capture"
tag=r02b
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_${tag}.csv \
  python bench.py --steps 1 --warmup 1 --profile-only > gpurun_out/launch_${tag}.log 2>&1
tail -1 gpurun_out/launch_${tag}.log
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:attention" --launch-skip 14 --launch-count 1 \
    -f -o gpurun_out/attn_${tag} python bench.py --steps 1 --warmup 1 --profile-only > gpurun_out/ncu_attn_${tag}.log 2>&1
tail -1 gpurun_out/ncu_attn_${tag}.log
echo "== bench full"
timeout 900 python bench.py > gpurun_out/bench_c17.json 2> gpurun_out/bench_c17.err; tail -c 600 gpurun_out/bench_c17.json
