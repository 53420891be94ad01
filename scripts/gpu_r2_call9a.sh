#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/r2c9a.txt
: > $out
nvidia-smi --query-gpu=index,name --format=csv >> $out 2>&1
echo "== N=2 bench" >> $out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "rc=$?" >> $out
tail -c 1800 gpurun_out/bench_n2.json >> $out
grep -v "^$" gpurun_out/bench_n2.err | tail -15 >> $out
echo "== N=2 reference arm (rank 0 only)" >> $out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 \
  bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err
echo "rc=$?" >> $out
tail -c 400 gpurun_out/bench_ref_n2.json >> $out
cat $out
