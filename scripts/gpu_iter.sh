timeout 300 python -m pytest tests/test_path_gpu.py -q --tb=short -s -k "vit_tokens" 2>&1 | grep -E "rel_l2|passed|failed|Error" | tail -3
for sb in 32 16 8 4; do
  WVN_VIT_SUBCHUNK=$sb timeout 300 python bench.py --steps 5 --warmup 2 --cpu-frames 0 2>gpurun_out/bench_iter.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('sub',$sb,'fps',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'attn TF',round(d['roofline']['achieved'],1),'attn share',round(d['roofline']['share_of_step'],3),'gemm share',round(d['roofline']['gemm_share_of_step'],3))"
  tail -3 gpurun_out/bench_iter.err
done
