"""Phase timing of the tcgen05 GEMM (needs a library built with -DWVN_GEMM_TIMING)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wild_visual_navigation_b200 import ops
M = 102400
for (N, K, kind, act, name) in [(1152, 384, 0, 0, "qkv-like bf16"), (1536, 384, 0, 2, "fc1 gelu"), (384, 1536, 2, 0, "fc2 resid"), (384, 384, 2, 0, "proj resid")]:
    a = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda")
    out = torch.zeros(M, N, device="cuda", dtype=torch.float32 if kind else torch.bfloat16)
    for _ in range(2):
        ops.gemm_bf16(a, w, bias, out_kind=kind, act=act, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = int(os.environ.get("REPS", "1"))
    e0.record()
    for _ in range(reps):
        ops.gemm_bf16(a, w, bias, out_kind=kind, act=act, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name}: {ms*1000:.1f} us  {2*M*N*K/ms/1e9:.0f} TFLOP/s", flush=True)
