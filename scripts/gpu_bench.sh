# bench + ncu evidence (1 GPU).  Usage: bash scripts/gpu_bench.sh [round-tag]
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 3000 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
for c in 4 16; do
  timeout 300 python bench.py --steps 5 --warmup 2 --chunk $c --cpu-frames 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk',$c,'fps',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'attn TF',round(d['roofline']['achieved'],1),'attn share',round(d['roofline']['share_of_step'],3),'gemm share',round(d['roofline']['gemm_share_of_step'],3))" >> gpurun_out/chunk_sweep_$TAG.txt 2>&1
done
cat gpurun_out/chunk_sweep_$TAG.txt
# launch list of one step (B=8, one ViT chunk): cold-cache serialised times, compare SHARES only
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$TAG.csv \
   python bench.py --profile-only --steps 1 --warmup 1 --batch 8 --chunk 8 > gpurun_out/ncu_launch_$TAG.log 2>&1
# full capture of the dominant kernels
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 14 -c 2 -o gpurun_out/attn_$TAG -f \
   python bench.py --profile-only --steps 1 --warmup 1 --batch 8 --chunk 8 > gpurun_out/ncu_attn_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 50 -c 5 -o gpurun_out/gemm_$TAG -f \
   python bench.py --profile-only --steps 1 --warmup 1 --batch 8 --chunk 8 > gpurun_out/ncu_gemm_$TAG.log 2>&1
ls -la gpurun_out | tail -12
