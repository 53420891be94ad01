#!/bin/bash
cd "$(dirname "$0")/.."
bash scripts/gpu_profile.sh r02 2>&1 | tail -20
