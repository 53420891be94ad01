"""Tensor-level wrappers over the C ABI (include/wvn_b200.h).

Everything here is plumbing: allocate outputs with torch, pass raw pointers + the current stream
to libwvn_b200.so.  The classes in ``feature_extractor/``, ``model/``, ``utils/`` and
``traversability_estimator/`` (the reference's API surface) are built on these.
"""
from __future__ import annotations

import ctypes
import math
from ctypes import byref, c_void_p

import torch
import torch.nn.functional as F

from . import _C
from ._C import TrainConfig, VitConfig, check, lib, ptr, stream

OUT_BF16, OUT_F32, OUT_RESID_F32 = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2


# --------------------------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------------------------
def gemm_bf16(a, w, bias=None, out_kind=OUT_BF16, act=ACT_NONE, out=None, block_n=0):
    """out = act(a @ w.T + bias).  a: [M,K] bf16, w: [N,K] bf16, bias: [N] f32."""
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=torch.bfloat16 if out_kind == OUT_BF16 else torch.float32)
    check(lib().wvn_gemm_bf16(ptr(a), a.stride(0), ptr(w), ptr(bias), ptr(out), out.stride(0), M, N, K, out_kind, act,
                              block_n, stream()))
    return out


def attention(q, k, vt, n_valid, scale=0.125):
    """q,k: [B,H,npad,64] bf16; vt: [B,H,64,npad] bf16 -> [B,npad,H*64] bf16."""
    B, H, npad, dh = q.shape
    assert dh == 64
    out = torch.empty(B, npad, H * 64, device=q.device, dtype=torch.bfloat16)
    check(lib().wvn_attention_bf16(ptr(q), ptr(k), ptr(vt), ptr(out), B, H, npad, n_valid, scale, stream()))
    return out


def layernorm(x, gamma, beta, eps=1e-6):
    rows, dim = x.shape
    out = torch.empty(rows, dim, device=x.device, dtype=torch.bfloat16)
    check(lib().wvn_layernorm(ptr(x), ptr(gamma), ptr(beta), ptr(out), rows, dim, eps, stream()))
    return out


def upsample_dense(tokens, gh, gw, out_h, out_w):
    """tokens [B, gh*gw, D] f32 -> (B, D, out_h, out_w) f32, bilinear align_corners=True."""
    B, P, D = tokens.shape
    out = torch.empty(B, D, out_h, out_w, device=tokens.device, dtype=torch.float32)
    check(lib().wvn_upsample_dense(ptr(tokens), ptr(out), B, D, gh, gw, out_h, out_w, stream()))
    return out


def logits_argmax(logits, col0, classes, batch, npad, gh, gw, out_h, out_w, col0_b=0, classes_b=0):
    """Per-pixel argmax of bilinearly (align_corners=False) upsampled class logits.  With a second
    column range the two segmentations come out of one pass: returns seg or (seg, seg_b)."""
    seg = torch.empty(batch, out_h, out_w, device=logits.device, dtype=torch.int64)
    seg_b = torch.empty_like(seg) if classes_b > 0 else None
    check(lib().wvn_logits_argmax(ptr(logits), logits.stride(0), col0, classes, col0_b, classes_b, batch, npad, gh, gw,
                                  out_h, out_w, ptr(seg), ptr(seg_b), stream()))
    return seg if seg_b is None else (seg, seg_b)


_WS_CACHE = {}


def _workspace(tag, nbytes, device):
    """Scratch that the kernels fully (re)initialise themselves: one buffer per (tag, device), grown on demand, so the
    per-frame calls do not allocate."""
    key = (tag, str(device))
    buf = _WS_CACHE.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1), device=device, dtype=torch.uint8)
        _WS_CACHE[key] = buf
    return buf


def segment_reduce(seg, smax, tokens=None, grid=None, want_centers=True, want_edges=True, max_edges=None):
    """seg: [B,H,W] int64.  Returns dict(feat [B,smax,D] | None, centers [B,smax,2] | None,
    edges [B,max_edges,2] | None, n_edges [B] int32 | None).  Only the first n_edges[b] rows of edges[b] are written
    (n_edges[b] < 0 flags an overflow of max_edges)."""
    B, H, W = seg.shape
    dev = seg.device
    if tokens is not None:
        gh, gw = grid
        D = tokens.shape[-1]
    else:
        gh = gw = 1
        D = 0
    ws_bytes = lib().wvn_segment_workspace_bytes(B, smax, gh, gw)
    ws = _workspace("segment_reduce", ws_bytes, dev)
    feat = torch.empty(B, smax, D, device=dev, dtype=torch.float32) if tokens is not None else None
    centers = torch.empty(B, smax, 2, device=dev, dtype=torch.float32) if want_centers else None
    if want_edges:
        if max_edges is None:
            # STEGO cluster labels are not connected regions: the directed label-pair count can reach smax*(smax-1),
            # so the planar-graph bound only applies to large segment counts (SLIC-like, <= 1024 here: 16 MB worst case)
            max_edges = smax * smax if smax <= 256 else min(smax * smax, 64 * smax)
        edges = torch.empty(B, max_edges, 2, device=dev, dtype=torch.int64)
        n_edges = torch.empty(B, device=dev, dtype=torch.int32)
    else:
        max_edges, edges, n_edges = 0, None, None
    check(lib().wvn_segment_reduce(ptr(seg), B, H, W, smax, ptr(tokens), gh, gw, D, ptr(feat), ptr(centers), ptr(edges),
                                   ptr(n_edges), max_edges, ptr(ws), stream()))
    return {"feat": feat, "centers": centers, "edges": edges, "n_edges": n_edges, "max_edges": max_edges}


def pool_supervision(seg: torch.Tensor, mask: torch.Tensor, smax: int):
    """Per-segment supervision labels (MissionNode.update_supervision_signal, nodes.py:400-440).
    seg (B,H,W) int64; mask (B,C,H,W) or (B,H,W) fp32 with NaN = unlabelled -> y (B,smax) f32, y_valid (B,smax) bool."""
    _C.require_device()
    if mask.dim() == 3:
        mask = mask[:, None]
    B, C, H, W = mask.shape
    assert seg.shape == (B, H, W) and seg.dtype == torch.int64 and mask.dtype == torch.float32
    seg, mask = seg.contiguous(), mask.contiguous()
    y = torch.empty(B, smax, device=seg.device, dtype=torch.float32)
    valid = torch.empty(B, smax, device=seg.device, dtype=torch.uint8)
    ws = torch.empty(B, smax, device=seg.device, dtype=torch.float32)
    check(lib().wvn_supervision_pool(ptr(seg), ptr(mask), B, C, H, W, smax, ptr(y), ptr(valid), ptr(ws), stream()))
    return y, valid.bool()


_slic_luts = {}


def slic_tables():
    """Host lookup tables of the library's 8-bit sRGB -> CIELAB conversion as int32 numpy arrays (g256, m9, f4096)."""
    import ctypes

    import numpy as np

    g, m, f = np.zeros(256, np.int32), np.zeros(9, np.int32), np.zeros(4096, np.int32)
    lib().wvn_slic_tables(g.ctypes.data_as(ctypes.c_void_p), m.ctypes.data_as(ctypes.c_void_p),
                          f.ctypes.data_as(ctypes.c_void_p))
    return g, m, f


def slic_geometry(h, w, num_components):
    """-> (grid interval S, nx, ny); the segmentation has nx * ny clusters."""
    import ctypes

    S, nx, ny = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    check(lib().wvn_slic_geometry(h, w, num_components, ctypes.byref(S), ctypes.byref(nx), ctypes.byref(ny)))
    return S.value, nx.value, ny.value


def slic(img, num_components=100, compactness=10.0, iters=10):
    """SLIC superpixels of img (B,3,H,W) fp32 in [0,1] -> labels (B,H,W) int64 in [0, nx*ny) (see include/wvn_b200.h)."""
    _C.require_device()
    B, C, H, W = img.shape
    assert C == 3 and img.dtype == torch.float32
    img = img.contiguous()
    key = str(img.device)
    if key not in _slic_luts:
        _slic_luts[key] = tuple(torch.from_numpy(t).to(img.device) for t in slic_tables())
    g, m, f = _slic_luts[key]
    ws = _workspace("slic", lib().wvn_slic_workspace_bytes(B, H, W, num_components), img.device)
    labels = torch.empty(B, H, W, device=img.device, dtype=torch.int64)
    check(lib().wvn_slic(ptr(img), B, H, W, num_components, float(compactness), iters, ptr(g), ptr(m), ptr(f), ptr(labels),
                         ptr(ws), stream()))
    return labels


def project_and_render(sK, pose_camera_in_world, points, colors, h, w, render=True, supervision=None, traversability=None):
    """Footprint projection + rasterisation (ImageProjector.project_and_render, image_projector.py:152-197).
    sK (B,4,4) scaled camera matrices, pose (B,4,4), points (B,N,3), colors (B,3) | (3,) | None.
    -> (masks (B,3,h,w) or None, projected (B,N,2), valid (B,N) bool); ``supervision`` (B,3,h,w) is updated in place
    with fmin(supervision, mask * traversability) (traversability_estimator.py:281-284)."""
    _C.require_device()
    B, N = points.shape[0], points.shape[1]
    dev = points.device
    sK, pose, points = sK.float().contiguous(), pose_camera_in_world.float().contiguous(), points.float().contiguous()
    assert sK.shape == (B, 4, 4) and pose.shape == (B, 4, 4) and points.shape == (B, N, 3)
    if colors is not None:
        colors = colors.to(dev, torch.float32).contiguous()
        assert colors.shape in ((3,), (B, 3))
    masks = torch.empty(B, 3, h, w, device=dev, dtype=torch.float32) if render else None
    proj = torch.empty(B, N, 2, device=dev, dtype=torch.float32)
    valid = torch.empty(B, N, device=dev, dtype=torch.uint8)
    if supervision is not None:
        assert supervision.shape == (B, 3, h, w) and supervision.dtype == torch.float32 and supervision.is_contiguous()
    if traversability is not None:
        traversability = traversability.to(dev, torch.float32).reshape(-1)[:1].contiguous()
    check(lib().wvn_project_and_render(ptr(sK), ptr(pose), ptr(points), ptr(colors),
                                       int(colors is not None and colors.dim() == 2), B, N, h, w, ptr(traversability),
                                       ptr(masks), ptr(proj), ptr(valid), ptr(supervision), stream()))
    return masks, proj, valid.bool()


def flip_average(head, batch, npad, grid):
    """STEGO flip-TTA merge of the head output [2*batch*npad, C] in place (see include/wvn_b200.h)."""
    assert head.is_contiguous() and head.shape[0] == 2 * batch * npad and head.dtype == torch.float32
    check(lib().wvn_flip_average(ptr(head), batch, npad, grid, head.stride(0), stream()))


def stego_kmeans(head, batch, npad, patches, code_col, code_dim, logit_col, k, iters, centroids_out=None):
    """Per-image k-means of the code columns of the head output ``head`` [batch*npad, ld]; leaves the per-patch
    nearest-centroid scores in columns [logit_col, logit_col + k) (see include/wvn_b200.h)."""
    ws = _workspace("stego_kmeans", lib().wvn_stego_kmeans_workspace_bytes(batch, k, code_dim), head.device)
    check(lib().wvn_stego_kmeans(ptr(head), head.stride(0), batch, npad, patches, code_col, code_dim, logit_col, k, iters,
                                 ptr(centroids_out), ptr(ws), stream()))


def relabel(seg, num_labels):
    """In-place relabel of each frame of seg [B,H,W] to 0..S-1; returns counts [B] int32."""
    B = seg.shape[0]
    scratch = _workspace("relabel", 4 * B * num_labels, seg.device).view(torch.int32)[: B * num_labels]
    counts = torch.empty(B, device=seg.device, dtype=torch.int32)
    check(lib().wvn_segment_relabel(ptr(seg), B, seg[0].numel(), num_labels, ptr(scratch), ptr(counts), stream()))
    return counts


# --------------------------------------------------------------------------------------------
# ViT backbone handle
# --------------------------------------------------------------------------------------------
def interpolate_pos_embed(pos_embed: torch.Tensor, grid: int) -> torch.Tensor:
    """DINO's interpolate_pos_encoding (bicubic, +0.1 trick), done once at weight-load time."""
    n_pre = pos_embed.shape[1] - 1
    if n_pre == grid * grid:
        return pos_embed
    dim = pos_embed.shape[-1]
    side = int(math.sqrt(n_pre))
    w0 = h0 = grid + 0.1
    patch = F.interpolate(pos_embed[:, 1:].reshape(1, side, side, dim).permute(0, 3, 1, 2).float(),
                          scale_factor=(h0 / side, w0 / side), mode="bicubic")
    patch = patch.permute(0, 2, 3, 1).reshape(1, -1, dim)
    return torch.cat((pos_embed[:, :1], patch), dim=1)


class ViTBackbone:
    """Owns a ``wvn_vit_t``: DINO ViT weights (bf16 on device) + L2-sized activation workspaces."""

    def __init__(self, image_size, patch_size, dim, depth, heads, mlp_dim, state_dict, max_batch=32, chunk=0,
                 head_weights=None, ln_eps=1e-6):
        _C.require_device()
        self.image_size, self.patch_size, self.dim = image_size, patch_size, dim
        self.grid = image_size // patch_size
        self.P = self.grid * self.grid
        self.max_batch = max_batch
        self.head_out = 0 if head_weights is None else int(head_weights["head_a.weight"].shape[0])
        cfg = VitConfig(image_size, patch_size, dim, depth, heads, mlp_dim, max_batch, chunk, ln_eps, self.head_out)
        h = c_void_p()
        check(lib().wvn_vit_create(byref(cfg), byref(h)))
        self._h = h
        self.npad = lib().wvn_vit_npad(h)
        self._load(state_dict, depth)
        if head_weights is not None:
            for k, v in head_weights.items():
                self._set("stego." + k, v)

    def _set(self, name, t):
        t = t.detach().to(dtype=torch.float32, device="cpu").contiguous()
        check(lib().wvn_vit_set_weight(self._h, name.encode(), c_void_p(t.data_ptr()), t.numel()))

    def _load(self, sd, depth):
        sd = {k.replace("module.", "").replace("backbone.", ""): v for k, v in sd.items()}
        self._set("cls_token", sd["cls_token"].reshape(-1))
        self._set("pos_embed", interpolate_pos_embed(sd["pos_embed"].float().cpu(), self.grid).reshape(-1))
        self._set("patch_embed.proj.weight", sd["patch_embed.proj.weight"].reshape(self.dim, -1))
        self._set("patch_embed.proj.bias", sd["patch_embed.proj.bias"])
        for i in range(depth):
            for leaf in ("norm1.weight", "norm1.bias", "attn.qkv.weight", "attn.qkv.bias", "attn.proj.weight",
                         "attn.proj.bias", "norm2.weight", "norm2.bias", "mlp.fc1.weight", "mlp.fc1.bias",
                         "mlp.fc2.weight", "mlp.fc2.bias"):
                self._set(f"blocks.{i}.{leaf}", sd[f"blocks.{i}.{leaf}"])
        self._set("norm.weight", sd["norm.weight"])
        self._set("norm.bias", sd["norm.bias"])

    def forward(self, img: torch.Tensor, resized_hw=None, flip_tta: bool = False) -> torch.Tensor:
        """img: (B,3,H,W) f32 CUDA in [0,1], or the camera frames themselves (B,H,W,3) uint8 RGB
        -> final-norm patch tokens (B, P, D) f32.  flip_tta (float input only): (2B, P, D), the second half being the
        pass over the horizontally flipped TRANSFORMED images (STEGO's get_code)."""
        img = img.contiguous()
        if flip_tta:
            B, C, H, W = img.shape
            assert img.dtype == torch.float32
            fn = lib().wvn_vit_forward_tta
        elif img.dtype == torch.uint8:
            B, H, W, C = img.shape
            fn = lib().wvn_vit_forward_u8
        else:
            B, C, H, W = img.shape
            assert img.dtype == torch.float32
            fn = lib().wvn_vit_forward
        assert C == 3
        if resized_hw is None:
            resized_hw = _resized_size(H, W, self.image_size)
        tokens = torch.empty(2 * B if flip_tta else B, self.P, self.dim, device=img.device, dtype=torch.float32)
        check(fn(self._h, ptr(img), B, H, W, resized_hw[0], resized_hw[1], ptr(tokens), stream()))
        self.last_tokens = tokens   # the handle's bf16 copy of exactly these tokens feeds the per-pixel head directly
        return tokens

    def stego_head(self, batch: int) -> torch.Tensor:
        out = torch.empty(batch * self.npad, self.head_out, device="cuda", dtype=torch.float32)
        check(lib().wvn_vit_stego_head(self._h, batch, ptr(out), stream()))
        return out

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().wvn_vit_destroy(self._h)
                self._h = None
        except Exception:
            pass


def _resized_size(h, w, size):
    # torchvision Resize(size:int): smaller edge -> size
    if h <= w:
        return size, int(size * w / h)
    return int(size * h / w), size


# --------------------------------------------------------------------------------------------
# traversability MLP: inference handle and trainer
# --------------------------------------------------------------------------------------------
class MlpInference:
    def __init__(self, dim=384, h1=256, h2=32, chunk_rows=0, tokens_per_frame=0):
        _C.require_device()
        self.dim, self.h1, self.h2 = dim, h1, h2
        h = c_void_p()
        check(lib().wvn_mlp_infer_create(dim, h1, h2, chunk_rows, byref(h)))
        self._h = h
        if tokens_per_frame > 0:  # all workspaces exist before the first frame arrives
            check(lib().wvn_mlp_infer_reserve(h, tokens_per_frame))

    def set_params(self, flat_params: torch.Tensor):
        assert flat_params.is_cuda and flat_params.dtype == torch.float32
        check(lib().wvn_mlp_infer_set_params(self._h, ptr(flat_params.contiguous()), stream()))

    def pixels(self, tokens, grid, out_hw, cg_mean, cg_std, std_factor):
        B = tokens.shape[0]
        trav = torch.empty(B, out_hw[0], out_hw[1], device=tokens.device, dtype=torch.float32)
        conf = torch.empty_like(trav)
        check(lib().wvn_mlp_infer_pixels(self._h, ptr(tokens), B, grid[0], grid[1], out_hw[0], out_hw[1], ptr(cg_mean),
                                         ptr(cg_std), float(std_factor), ptr(trav), ptr(conf), stream()))
        return trav, conf

    def pixels_from_vit(self, vit: "ViTBackbone", batch, out_hw, cg_mean, cg_std, std_factor):
        """The same maps from the backbone's own bf16 tokens of its last forward (frames [0, batch))."""
        trav = torch.empty(batch, out_hw[0], out_hw[1], device=cg_mean.device, dtype=torch.float32)
        conf = torch.empty_like(trav)
        check(lib().wvn_mlp_infer_pixels_vit(self._h, vit._h, batch, out_hw[0], out_hw[1], ptr(cg_mean), ptr(cg_std),
                                             float(std_factor), ptr(trav), ptr(conf), stream()))
        return trav, conf

    def rows(self, x, cg_mean, cg_std, std_factor):
        R = x.shape[0]
        trav = torch.empty(R, device=x.device, dtype=torch.float32)
        conf = torch.empty_like(trav)
        check(lib().wvn_mlp_infer_rows(self._h, ptr(x.contiguous()), R, ptr(cg_mean), ptr(cg_std), float(std_factor),
                                       ptr(trav), ptr(conf), stream()))
        return trav, conf

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().wvn_mlp_infer_destroy(self._h)
                self._h = None
        except Exception:
            pass


def mlp_forward_f32(flat_params, x, dim, h1, h2):
    """SimpleMLP.forward in fp32: returns (rows, 1+dim) with column 0 through the sigmoid."""
    R = x.shape[0]
    dev = x.device
    b1 = torch.empty(R, h1, device=dev, dtype=torch.float32)
    b2 = torch.empty(R, h2, device=dev, dtype=torch.float32)
    out = torch.empty(R, 1 + dim, device=dev, dtype=torch.float32)
    check(lib().wvn_mlp_forward_f32(dim, h1, h2, ptr(flat_params), ptr(x.contiguous()), R, ptr(b1), ptr(b2), ptr(out),
                                    stream()))
    return out


class MlpTrainer:
    """Fused online train step on a flat fp32 parameter buffer (csrc/mlp_train_fused.cu): four kernels, no host
    synchronisation, rows may arrive padded per frame (``step_padded``) exactly as the segment pooling leaves them.

    ``params`` is the tensor that SimpleMLP's layers view into, so the reference's state_dict layout stays intact.
    With ``process_group`` set the step is global-batch exact: the six statistic sums (incl. the row count) and the flat
    gradient are all-reduced between the kernels — by the library's own NCCL communicator when the group's backend is
    NCCL (one rank per GPU), otherwise (e.g. gloo in tests) by ``torch.distributed`` on the same buffers.
    ``legacy=True`` runs round 1's three-phase kernels (csrc/mlp_train.cu), kept for A/B parity tests."""

    def __init__(self, params, dim=384, h1=256, h2=32, max_rows=4096, w_trav=0.03, w_reco=0.5, std_factor=0.5,
                 anomaly_balanced=True, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, process_group=None, legacy=False):
        _C.require_device()
        self.dim, self.h1, self.h2 = dim, h1, h2
        self.n_params = lib().wvn_mlp_param_count(dim, h1, h2)
        assert params.numel() == self.n_params and params.is_cuda and params.dtype == torch.float32
        dev = params.device
        self.params = params
        self.grads = torch.zeros(self.n_params + 1, device=dev)
        self.exp_avg = torch.zeros(self.n_params, device=dev)
        self.exp_avg_sq = torch.zeros(self.n_params, device=dev)
        self.step_counter = torch.zeros(1, device=dev, dtype=torch.int64)
        self.cfg = TrainConfig(w_trav, w_reco, std_factor, int(anomaly_balanced), lr, betas[0], betas[1], eps)
        self.metrics = torch.zeros(6, device=dev)
        self.cg_mean = torch.zeros(1, device=dev)
        self.cg_std = torch.ones(1, device=dev)
        self.pg = process_group
        self.legacy = legacy
        self._conf = None     # (method id, var, running_n, running_sum, running_sum_of_squares, kf_proc_cov, kf_meas_cov)
        self._h = None
        self._lib_comm = False
        if legacy:
            self.scalars = torch.zeros(lib().wvn_mlp_train_scalars_bytes() // 8, device=dev, dtype=torch.float64)
            self._alloc_ws(max_rows)
            return
        self.scalars = torch.zeros((lib().wvn_mlp_trainer_scalars_bytes() + 7) // 8, device=dev, dtype=torch.float64)
        self._create(max_rows)

    # ---- fused path ---------------------------------------------------------------------------
    def _create(self, max_rows):
        if self._h is not None:
            lib().wvn_mlp_trainer_destroy(self._h)
            self._h = None
        self.max_rows = int(max_rows)
        h = c_void_p()
        check(lib().wvn_mlp_trainer_create(self.dim, self.h1, self.h2, self.max_rows, byref(self.cfg), ptr(self.scalars),
                                           ptr(self.grads), byref(h)))
        self._h = h
        if self._conf is not None:
            self.set_confidence(*self._conf)
        self.conf = torch.empty(self.max_rows + 32, device=self.params.device, dtype=torch.float32)
        self._lib_comm = False
        if self.pg is not None:
            import torch.distributed as dist

            if dist.get_backend(self.pg) == "nccl":
                # the library owns its communicator: rank 0 makes the id, torch.distributed only carries the 128 bytes
                rank, world = dist.get_rank(self.pg), dist.get_world_size(self.pg)
                buf = (ctypes.c_ubyte * 128)()
                if rank == 0:
                    check(lib().wvn_comm_unique_id(buf))
                idt = torch.tensor(list(buf), dtype=torch.uint8, device=self.params.device)
                dist.broadcast(idt, src=dist.get_global_rank(self.pg, 0), group=self.pg)
                raw = (ctypes.c_ubyte * 128)(*idt.cpu().tolist())
                check(lib().wvn_mlp_trainer_init_comm(self._h, raw, rank, world))
                self._lib_comm = True

    def set_confidence(self, method=0, var=None, running_n=None, running_sum=None, running_sum_of_squares=None,
                       kf_proc_cov=0.2, kf_meas_cov=1.0):
        """ConfidenceGenerator method of the step (0 latest_measurement, 1 running_mean, 2 kalman_filter, 3
        moving_average) and the device tensors holding its state (updated in place by the step; None = private)."""
        assert not self.legacy or method == 0, "the round-1 kernels implement latest_measurement only"
        self._conf = (int(method), var, running_n, running_sum, running_sum_of_squares, float(kf_proc_cov), float(kf_meas_cov))
        if self._h is not None:
            check(lib().wvn_mlp_trainer_set_confidence(self._h, int(method), ptr(var), ptr(running_n), ptr(running_sum),
                                                       ptr(running_sum_of_squares), float(kf_proc_cov), float(kf_meas_cov)))

    def _run(self, x, groups, rpg, n_rows, y, yv):
        if groups * rpg > self.max_rows:
            self._create(int(groups * rpg * 1.5))
        x = x.contiguous()
        y = y.contiguous().float()
        yv = yv.contiguous().to(torch.uint8)
        s = stream()

        def phase(mask):
            check(lib().wvn_mlp_train_step(self._h, ptr(self.params), ptr(self.exp_avg), ptr(self.exp_avg_sq),
                                           ptr(self.step_counter), ptr(x), groups, rpg, ptr(n_rows), ptr(y), ptr(yv),
                                           ptr(self.cg_mean), ptr(self.cg_std), ptr(self.conf), ptr(self.metrics), mask, s))

        if self.pg is None or self._lib_comm:
            phase(7)
        else:  # non-NCCL process group (tests): the same two exchanges through torch.distributed
            import torch.distributed as dist

            phase(1)
            dist.all_reduce(self.scalars[:6], group=self.pg)
            if self._conf is not None and self._conf[0] == 3:   # moving_average normalises by the global extrema
                dist.all_reduce(self.scalars[6:7], op=dist.ReduceOp.MIN, group=self.pg)
                dist.all_reduce(self.scalars[7:8], op=dist.ReduceOp.MAX, group=self.pg)
            phase(2)
            dist.all_reduce(self.grads, group=self.pg)
            phase(4)

    def step_padded(self, feat, n_rows, y, y_valid):
        """feat [G, S, D] f32 padded per group, n_rows [G] int32 (device): the first n_rows[g] rows of group g are live.
        y / y_valid are indexed by the compacted row number.  Returns the confidence buffer (compacted order; the live
        prefix has sum(n_rows) entries — no host sync happens here)."""
        assert not self.legacy
        G, S, D = feat.shape
        assert D == self.dim and n_rows.dtype == torch.int32 and y.numel() >= 1
        self._run(feat, G, S, n_rows, y, y_valid)
        return self.conf

    def step(self, x, y, y_valid, n_total=None):
        """x [R,D] f32, y [R] f32, y_valid [R] bool.  Returns the confidence vector [R]; metrics stay
        on the device in ``self.metrics`` (loss_total, loss_trav, loss_reco, loss_trav_conf, mean, std)."""
        if self.legacy:
            return self._step_legacy(x, y, y_valid, n_total)
        R = x.shape[0]
        self._run(x, 1, R, None, y, y_valid)
        return self.conf[:R]

    # ---- round-1 kernels (three phases, ~20 launches) --------------------------------------------
    def _alloc_ws(self, max_rows):
        self.max_rows = max_rows
        nbytes = lib().wvn_mlp_train_workspace_bytes(self.dim, self.h1, self.h2, max_rows)
        self.ws = torch.empty(nbytes // 4, device=self.params.device, dtype=torch.float32)
        self.conf = torch.empty(max_rows, device=self.params.device, dtype=torch.float32)

    def _step_legacy(self, x, y, y_valid, n_total=None):
        R = x.shape[0]
        if R > self.max_rows:
            self._alloc_ws(int(R * 1.5))
        x = x.contiguous()
        y = y.contiguous().float()
        yv = y_valid.contiguous().to(torch.uint8)
        d = (self.dim, self.h1, self.h2)
        s = stream()
        check(lib().wvn_mlp_train_forward_stats(*d, ptr(self.params), ptr(x), ptr(y), ptr(yv), R, self.max_rows,
                                                ptr(self.ws), ptr(self.scalars), s))
        if self.pg is not None:
            import torch.distributed as dist
            dist.all_reduce(self.scalars[:5], group=self.pg)
            if n_total is None:
                nt = torch.tensor([R], device=x.device, dtype=torch.int64)
                dist.all_reduce(nt, group=self.pg)
                n_total = int(nt.item())
        if n_total is None:
            n_total = R
        check(lib().wvn_mlp_train_backward(*d, ptr(self.params), ptr(x), ptr(y), ptr(yv), R, self.max_rows, n_total,
                                           byref(self.cfg), ptr(self.ws), ptr(self.scalars), ptr(self.cg_mean),
                                           ptr(self.cg_std), ptr(self.grads), ptr(self.conf), s))
        if self.pg is not None:
            import torch.distributed as dist
            dist.all_reduce(self.grads, group=self.pg)
        check(lib().wvn_mlp_train_apply(*d, ptr(self.params), ptr(self.grads), ptr(self.exp_avg), ptr(self.exp_avg_sq),
                                        ptr(self.step_counter), n_total, byref(self.cfg), ptr(self.scalars), s))
        check(lib().wvn_mlp_train_read_metrics(ptr(self.scalars), ptr(self.metrics), s))
        return self.conf[:R]

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().wvn_mlp_trainer_destroy(self._h)
                self._h = None
        except Exception:
            pass
