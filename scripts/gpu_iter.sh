timeout 300 python -m pytest tests/test_kernels_gpu.py -q --tb=short -k "attention" 2>&1 | tail -3
for p in 0 1 2 9; do WVN_ATTN_POLY=$p B=32 timeout 120 python scripts/bench_attention.py 2>&1 | tail -1; done
timeout 300 python bench.py --steps 5 --warmup 2 --cpu-frames 0 2>gpurun_out/bench_iter.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fps',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'attn TF',round(d['roofline']['achieved'],1),'attn share',round(d['roofline']['share_of_step'],3),'gemm share',round(d['roofline']['gemm_share_of_step'],3))"
tail -3 gpurun_out/bench_iter.err
