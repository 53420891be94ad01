mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for sel in "test_gemm" "test_attention" "test_layernorm or test_upsample or test_logits"; do
  name=$(echo $sel | tr ' ' '_')
  timeout 400 python -m pytest tests/test_kernels_gpu.py -q --tb=short --maxfail=6 -k "$sel" > gpurun_out/k_$name.log 2>&1
  echo "== $sel exit $?" >> gpurun_out/summary.txt
  tail -5 gpurun_out/k_$name.log >> gpurun_out/summary.txt
done
for sel in "vit_tokens" "dino_interface" "stego" "segment" "pixel_inference" "train_step or checkpoint"; do
  name=$(echo $sel | tr ' ' '_')
  timeout 400 python -m pytest tests/test_path_gpu.py -q --tb=short -s -k "$sel" > gpurun_out/p_$name.log 2>&1
  echo "== $sel exit $?" >> gpurun_out/summary.txt
  grep -E "rel_l2|agree|diff|step [0-9]|passed|failed|Error|error" gpurun_out/p_$name.log | tail -12 >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
