#!/bin/bash
cd "$(dirname "$0")/.."
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/bench_final_n$N.json 2> gpurun_out/bench_final_n$N.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_final_n$N.json').read().strip().split('\n')[-1]); print('N', d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['clocks'])
" || tail -5 gpurun_out/bench_final_n$N.err
