mkdir -p gpurun_out
for p in 0 9; do WVN_ATTN_POLY=$p B=32 timeout 120 python scripts/bench_attention.py 2>&1 | tail -1; done
for br in 0 1 2; do
  WVN_GEMM_BRES=$br timeout 300 python bench.py --steps 5 --warmup 2 --cpu-frames 0 2>gpurun_out/bench_iter.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bres',$br,'fps',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'attn TF',round(d['roofline']['achieved'],1),'attn share',round(d['roofline']['share_of_step'],3),'gemm share',round(d['roofline']['gemm_share_of_step'],3))"
  tail -3 gpurun_out/bench_iter.err
done
timeout 300 python -m pytest tests/test_kernels_gpu.py -q --tb=short -k "gemm" 2>&1 | tail -3
WVN_GEMM_BRES=2 timeout 300 python -m pytest tests/test_kernels_gpu.py -q --tb=short -k "gemm" 2>&1 | tail -3
