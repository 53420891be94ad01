#!/bin/bash
# 2 GPUs: the in-library NCCL train-step collectives, then the default single-GPU bench line with all legs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/r2c8.txt
: > $out
nvidia-smi --query-gpu=index,name --format=csv >> $out 2>&1
echo "== N=2 bench" >> $out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "rc=$?" >> $out
tail -c 1500 gpurun_out/bench_n2.json >> $out
tail -5 gpurun_out/bench_n2.err >> $out
echo "== N=1 default bench (all legs)" >> $out
( time timeout 1200 python bench.py > gpurun_out/bench_n1_full.json 2> gpurun_out/bench_n1_full.err ) 2>> $out
tail -c 3500 gpurun_out/bench_n1_full.json >> $out
tail -5 gpurun_out/bench_n1_full.err >> $out
echo "== reference arm" >> $out
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>> $out
tail -c 600 gpurun_out/bench_ref.json >> $out
echo "== c5 (ViT-B/8, B=128) short" >> $out
timeout 900 python bench.py --config c5 --steps 3 --warmup 2 --no-extras --cpu-frames 0 > gpurun_out/bench_c5cfg.json 2> gpurun_out/bench_c5cfg.err
tail -c 1200 gpurun_out/bench_c5cfg.json >> $out
tail -3 gpurun_out/bench_c5cfg.err >> $out
cat $out
