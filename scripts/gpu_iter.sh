#!/bin/bash
# iteration script (rewritten per experiment)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_path_gpu.py -q -s --durations=6 -k "vit_base or eager" 2>&1 | tail -22
