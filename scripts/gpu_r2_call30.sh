#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 ncu --set full --clock-control none --import-source on -k regex:train_fwd_rows --launch-skip 1 --launch-count 1 -f -o gpurun_out/train_fwd_r02 python bench.py --steps 1 --warmup 1 --profile-only > gpurun_out/ncu_train_fwd.log 2>&1
tail -1 gpurun_out/ncu_train_fwd.log
