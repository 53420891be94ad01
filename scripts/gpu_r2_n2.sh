#!/bin/bash
cd "$(dirname "$0")/.."
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_final_n2.json 2> gpurun_out/bench_final_n2.err
tail -c 400 gpurun_out/bench_final_n2.json; tail -3 gpurun_out/bench_final_n2.err
python bench.py --steps 8 --warmup 3 --cpu-frames 0 --no-extras > gpurun_out/bench_final_n1b.json 2>/dev/null; python -c "
import json
for f in ['gpurun_out/bench_final_n2.json','gpurun_out/bench_final_n1b.json']:
    d=json.loads(open(f).read().strip().split('\n')[-1]); print(f, d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1))
"
