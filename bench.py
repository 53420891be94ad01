#!/usr/bin/env python
"""bench.py — frames/s of the WVN hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        (the reference algorithm on the host CPU cores)

One "step" = one batch of 32 synthetic 448x448 frames per GPU through the whole path:
  DINO ViT-S/8 forward -> STEGO head + per-pixel cluster argmax (segmentation) -> relabel ->
  per-segment feature pooling + centroids + adjacency -> per-pixel traversability MLP
  (trav + confidence maps) -> one online train step (fwd + loss + bwd + Adam) on the pooled rows,
with the MLP gradient all-reduced over NCCL when N > 1 (weak scaling: 32 frames per GPU).
STEGO's flip-TTA (a second backbone pass) is OFF in both arms, and said so in `config`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "frames/sec end-to-end (DINO feat + seg + MLP train-step) at 448x448"
IMG, PATCH, BATCH = 448, 8, 32
N_TOK = (IMG // PATCH) ** 2 + 1  # 3137
# algorithmic work: FLOPs of one fused-attention launch per (frame, head)
ATTN_FLOP_PER_FRAME_HEAD = 4.0 * N_TOK * N_TOK * 64


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for i, n in enumerate(names):
                if len(r) > 3 + i and r[3 + i].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_weights():
    from oracle.dino_vit import ViTConfig, synthetic_state_dict
    from oracle.stego_head import synthetic_head

    cfg = ViTConfig.from_name("vit_small", PATCH, IMG)
    return cfg, synthetic_state_dict(cfg, seed=1), synthetic_head(cfg.dim, 90, 32, 27, seed=3)


def synthetic_images(n_sets, batch, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.rand(batch, 3, IMG, IMG, generator=g) for _ in range(n_sets)]


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle (a port of the pure-Python reference path) on host cores
# ------------------------------------------------------------------------------------------------
def cpu_frames_per_s(frames_per_step, steps, warmup, cores):
    from oracle.pipeline import cpu_step
    from oracle.wvn_path import mlp_init

    torch.set_num_threads(cores)
    cfg, sd, hd = make_weights()
    mlp_sd, opt = mlp_init(cfg.dim, (256, 32), seed=42), None
    mean, std = torch.tensor([0.3]), torch.tensor([0.1])
    imgs = synthetic_images(1, frames_per_step, seed=0)[0]
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        _, _, mlp_sd, opt, _ = cpu_step(imgs, sd, cfg, hd, mlp_sd, opt, mean, std, supervision_seed=2 + i)
        times.append(time.perf_counter() - t0)
    t = sum(times[warmup:])
    return frames_per_step * steps / t, 1000.0 * t / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = min(os.cpu_count() or 1, args.cpu_threads)
    fps, ms = cpu_frames_per_s(args.cpu_frames, args.steps, args.warmup, cores)
    sample = f"{args.cpu_frames} frame(s) per step, {args.steps} steps, oracle port of the reference path (eager fp32 torch)"
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, flip_tta=False, frames_per_step=args.cpu_frames),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, flip_tta, frames_per_step):
    return {"workload": "C3+C2: ViT-S/8@448 -> STEGO head + per-pixel cluster argmax -> relabel -> segment pooling/"
                        "centroids/adjacency -> per-pixel MLP trav+confidence -> online train step (fwd+loss+bwd+Adam)",
            "frames_per_gpu_per_step": frames_per_step, "image": f"{IMG}x{IMG}", "backbone": "DINO ViT-S/8 (random-init, seeded)",
            "stego_flip_tta": flip_tta, "parallelism": f"dp{args.gpus} (independent frames; NCCL all-reduce of MLP stats+grads)",
            "l2": "inputs rotate over 3 image batches (3 x 77 MB > 126 MB L2); activations >> L2"}


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist

    from wild_visual_navigation_b200 import FeatureExtractor, TraversabilityEstimator, TraversabilityInference, _C

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
        pg = dist.group.WORLD
    _C.require_device()
    lib = _C.lib()

    cfg, sd, hd = make_weights()
    B = args.batch
    fe = FeatureExtractor(dev, segmentation_type="stego", feature_type="dino", input_size=IMG, state_dict=sd,
                          head_state_dict=hd, flip_tta=False, max_batch=B, chunk=args.chunk, backbone_type="vit_small",
                          patch_size=PATCH)
    smax = fe._stego._n_clusters
    te = TraversabilityEstimator(device=dev, process_group=pg, max_rows=B * smax)
    cg = te._traversability_loss._confidence_generator
    ti = TraversabilityInference(fe._dino, te._model, cg)
    # synthetic supervision for up to B*smax pooled rows
    g = torch.Generator().manual_seed(2 + rank)
    yv_all = (torch.rand(B * smax, generator=g) < 0.16).to(dev)
    y_all = torch.where(yv_all, torch.rand(B * smax, generator=g).clamp(min=0.001).to(dev), torch.zeros(B * smax, device=dev))
    row_ids = torch.arange(smax, device=dev)[None, :]

    frames = synthetic_images(3, B, seed=100 + rank)
    if args.ingest == "u8":
        frames = [(t * 255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous() for t in frames]
    host_imgs = [t.pin_memory() for t in frames]
    dev_imgs = [t.to(dev) for t in host_imgs]
    host_trav = torch.empty(B, IMG, IMG).pin_memory()
    host_conf = torch.empty(B, IMG, IMG).pin_memory()
    host_metrics = torch.empty(6).pin_memory()

    def step(img):
        r = fe.extract_batch(img)                                  # ViT + STEGO seg + pooling + graph
        trav, conf = ti.predict_from_tokens(r["tokens"], IMG)      # per-pixel MLP -> maps
        mask = row_ids < r["n_segments"][:, None]                  # pooled rows of existing segments
        x = r["feat"][mask]
        n = x.shape[0]
        te._trainer.step(x, y_all[:n], yv_all[:n])                 # fwd + loss + bwd + (all-reduce) + Adam
        ti.refresh_weights()                                       # inference sees the updated MLP
        return trav, conf

    # end-to-end leg: every step copies its frames host->device (pinned) and its maps + loss metrics
    # device->host, on copy streams so that step k's transfers overlap step k-1 / k+1's compute
    # (double-buffered, exactly what a camera loop around the public API would do).
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    img_buf = [torch.empty_like(host_imgs[0], device=dev) for _ in range(2)]
    out_buf = [(torch.empty(B, IMG, IMG, device=dev), torch.empty(B, IMG, IMG, device=dev), torch.empty(6, device=dev))
               for _ in range(2)]
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_used = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]
    state = {"primed": False}

    def h2d(k):
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_used[k % 2])                        # compute no longer reads this buffer
            img_buf[k % 2].copy_(host_imgs[k % 3], non_blocking=True)
            ev_in[k % 2].record(s_in)

    def step_e2e(k):
        cur = torch.cuda.current_stream()
        if not state["primed"]:
            h2d(k)
            state["primed"] = True
        cur.wait_event(ev_in[k % 2])
        h2d(k + 1)                                                 # prefetch the next step's frames
        trav, conf = step(img_buf[k % 2])
        ev_used[k % 2].record(cur)
        cur.wait_event(ev_out[k % 2])                              # previous D2H out of this slot finished
        out_buf[k % 2][0].copy_(trav)
        out_buf[k % 2][1].copy_(conf)
        out_buf[k % 2][2].copy_(te._trainer.metrics)
        ev_done[k % 2].record(cur)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_done[k % 2])
            host_trav.copy_(out_buf[k % 2][0], non_blocking=True)  # D2H of the per-pixel maps + metrics
            host_conf.copy_(out_buf[k % 2][1], non_blocking=True)
            host_metrics.copy_(out_buf[k % 2][2], non_blocking=True)
            ev_out[k % 2].record(s_out)

    def drain_e2e():
        torch.cuda.current_stream().wait_stream(s_out)
        torch.cuda.current_stream().wait_stream(s_in)
        state["primed"] = False

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, finish=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(steps):
            fn(k)
        if finish is not None:
            finish()                                               # the last D2H is inside the timed region
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for k in range(args.warmup):
        step(dev_imgs[k % 3])
    if args.profile_only:  # short, launch-list friendly run for ncu (no CPU leg, no e2e leg)
        for k in range(args.steps):
            step(dev_imgs[k % 3])
        torch.cuda.synchronize()
        print(json.dumps({"profile_only": True, "launches": int(lib.wvn_launch_count())}))
        return
    for k in range(min(2, args.warmup)):
        step_e2e(k)
    drain_e2e()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib.wvn_profile_enable(3 if args.profile_gemm else 1)   # CUDA events around the roofline kernel's launches
    l0 = lib.wvn_launch_count()
    total_ms = timed(lambda k: step(dev_imgs[k % 3]), args.steps)
    launches = lib.wvn_launch_count() - l0
    import ctypes
    prof_ms = (ctypes.c_float * 2)()
    prof_n = (ctypes.c_longlong * 2)()
    _C.check(lib.wvn_profile_collect(prof_ms, prof_n))
    lib.wvn_profile_enable(0)
    e2e_ms = timed(step_e2e, args.steps, finish=drain_e2e)
    clocks = sampler.stop() if rank == 0 else None

    ms_per_step = total_ms / args.steps
    value = world * B * args.steps / (total_ms / 1000.0)
    e2e_value = world * B * args.steps / (e2e_ms / 1000.0)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak_tf, peak_gbs, peak_src = read_peaks()
    chunk = fe._dino._model.max_batch if args.chunk <= 0 else args.chunk
    attn_launches = max(1, prof_n[0])
    frames_per_launch = B * 12 * args.steps / attn_launches      # 12 blocks per frame
    attn_avg_ms = prof_ms[0] / attn_launches
    flops_per_launch = ATTN_FLOP_PER_FRAME_HEAD * 6 * frames_per_launch
    achieved_tf = flops_per_launch / (attn_avg_ms * 1e-3) / 1e12 if attn_avg_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "attention_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")

    cores = min(os.cpu_count() or 1, args.cpu_threads)
    cpu_fps = cpu_frames_per_s(args.cpu_frames, 1, 0, cores)[0] if args.cpu_frames > 0 else None

    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16 operands / fp32 accumulate (ViT, STEGO head, per-pixel MLP); fp32 (train step)",
        "data": "synthetic", "config": workload_config(args, False, B) | {"vit_chunk_frames": args.chunk, "ingest": args.ingest},
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": host_imgs[0].numel() * host_imgs[0].element_size(),
                "d2h_bytes_per_step": 2 * B * IMG * IMG * 4 + 24, "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"kernel": "attention_kernel (fused QK^T-softmax-PV, tcgen05)", "bound": "tensor",
                     "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                     "peak_source": f"{peak_src} bf16_tflops_sustained (kernel timed inside the step)",
                     "traffic": traffic, "avg_launch_ms": attn_avg_ms, "launches": int(prof_n[0]),
                     "share_of_step": prof_ms[0] / total_ms if total_ms > 0 else None,
                     "gemm_share_of_step": prof_ms[1] / total_ms if (total_ms > 0 and prof_n[1] > 0) else None,
                     "gemm_launches": int(prof_n[1]) if prof_n[1] > 0 else None},
        "cpu_baseline": {"value": cpu_fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{args.cpu_frames} frame(s), one full step of the oracle port (eager fp32 torch, all host threads)"},
        "whole_path_tflops": value * (315.1 + 1.36 + 47.69) / 1000.0,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--chunk", type=int, default=32, help="frames per ViT activation chunk")
    ap.add_argument("--cpu-frames", type=int, default=1, help="frames in the bounded CPU sample (5.7 s/frame on 8 cores)")
    ap.add_argument("--cpu-threads", type=int, default=32,
                    help="torch CPU threads for the CPU legs (more than ~32 only adds sync overhead on these small ops)")
    ap.add_argument("--profile-only", action="store_true", help="setup + warmup + steps only (for ncu)")
    ap.add_argument("--profile-gemm", action="store_true",
                    help="also time every GEMM launch with CUDA events (gemm_share_of_step; ~280 extra event pairs per step)")
    ap.add_argument("--ingest", default="f32", choices=["f32", "u8"],
                    help="frame format at the boundary: f32 = (B,3,H,W) float in [0,1] (the reference's boundary, "
                         "BASELINE.json); u8 = camera frames (B,H,W,3) uint8, ingest fused into the patch loader (SURVEY.md §8f)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
