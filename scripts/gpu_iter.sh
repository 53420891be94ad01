# quick iteration: re-verify changed kernels, attention variant sweep, bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q --tb=short -k "attention" 2>&1 | tail -4
timeout 400 python -m pytest tests/test_path_gpu.py -q --tb=short -s -k "pixel or vit_tokens" 2>&1 | grep -E "rel_l2|diff|passed|failed|Error|assert|wvn" | tail -12
for p in 0 1 2 3; do WVN_ATTN_POLY=$p timeout 120 python scripts/bench_attention.py 2>&1 | tail -1; done
for c in 32; do
  timeout 300 python bench.py --steps 5 --warmup 2 --chunk $c --cpu-frames 0 2>gpurun_out/bench_iter.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk',$c,'fps',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'attn TF',round(d['roofline']['achieved'],1),'attn share',round(d['roofline']['share_of_step'],3),'gemm share',round(d['roofline']['gemm_share_of_step'],3))"
  tail -3 gpurun_out/bench_iter.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_iter.csv \
   python bench.py --profile-only --steps 1 --warmup 1 --batch 8 --chunk 8 > gpurun_out/ncu_launch_iter.log 2>&1
