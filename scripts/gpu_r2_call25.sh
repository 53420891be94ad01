#!/bin/bash
cd "$(dirname "$0")/.."
WVN_B200_LIB=wild_visual_navigation_b200/libwvn_b200_timing.so timeout 600 python bench.py --steps 1 --warmup 1 --profile-only 2>&1 | grep "pixel_head timing" | tail -4
