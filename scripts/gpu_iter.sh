timeout 300 python -m pytest tests/test_path_gpu.py -q --tb=short -s -k "pixel" 2>&1 | grep -E "diff|passed|failed|Error" | tail -4
timeout 300 python bench.py --steps 5 --warmup 2 --cpu-frames 0 2>gpurun_out/bench_iter.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fps',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'attn TF',round(d['roofline']['achieved'],1),'attn share',round(d['roofline']['share_of_step'],3),'gemm share',round(d['roofline']['gemm_share_of_step'],3))"
tail -3 gpurun_out/bench_iter.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pixel_head_kernel -s 1 -c 1 -o gpurun_out/pixel_head -f \
   python bench.py --profile-only --steps 1 --warmup 1 --batch 8 > gpurun_out/ncu_ph.log 2>&1
