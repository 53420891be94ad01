#!/bin/bash
# Round-end evidence: launch list of one step + full ncu captures of the three heaviest kernels.
# Usage (here): gpurun --timeout 1800 -- bash scripts/gpu_profile.sh <tag>;  then scripts/summarize_ncu.py.
cd "$(dirname "$0")/.."
tag=${1:-r02}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_${tag}.csv \
  python bench.py --steps 1 --warmup 1 --profile-only > gpurun_out/launch_${tag}.log 2>&1
tail -1 gpurun_out/launch_${tag}.log
# kernel base name, launches to skip (the 2nd step's block 3), output tag.  GEMM launch order per step:
# patch-embed, then (QKV, proj, fc1, fc2) x 12, STEGO head x 3, per-pixel token GEMM x 4  = 56
for spec in "regex:attention 14 attn" "regex:gemm_bf16_kernel 69 gemm_qkv" "regex:gemm_bf16_kernel 70 gemm_proj" "regex:gemm_bf16_kernel 71 gemm_fc1" "regex:gemm_bf16_kernel 72 gemm_fc2" "regex:pixel_head_kernel 4 pixel_head" "regex:logits_argmax 1 logits_argmax" "regex:stego_kmeans 1 kmeans"; do
  set -- $spec
  timeout 900 ncu --set full --clock-control none --import-source on -k "$1" --launch-skip $2 --launch-count 1 \
    -f -o gpurun_out/$3_${tag} python bench.py --steps 1 --warmup 1 --profile-only > gpurun_out/ncu_$3_${tag}.log 2>&1
  tail -1 gpurun_out/ncu_$3_${tag}.log
done
ls -la gpurun_out/*_${tag}.ncu-rep
