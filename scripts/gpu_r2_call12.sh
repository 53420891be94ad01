#!/bin/bash
cd "$(dirname "$0")/.."
echo "== footprint tests"
timeout 600 python -m pytest tests/test_footprint_gpu.py -x -q -s 2>&1 | tail -25
echo "== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
