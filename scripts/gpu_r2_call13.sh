#!/bin/bash
cd "$(dirname "$0")/.."
echo "== footprint + slic tests"
timeout 900 python -m pytest tests/test_footprint_gpu.py tests/test_slic_gpu.py -q -s 2>&1 | tail -40
