#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_confidence_methods_gpu.py -x -q -s 2>&1 | grep -E "step 6|step 2:|passed|failed|Error|assert" | tail -20
