"""ORACLE (test infrastructure only) — the whole per-frame hot path on the CPU, in the reference's
own order of operations (eager fp32 PyTorch), used as the checker in tests and as the timed CPU
baseline / `--impl reference` arm of bench.py.

One call = what the two ROS nodes do for a batch of frames:
  feature node  (wvn_feature_extractor_node.py:306-370): transform -> ViT -> [STEGO head ->
      postprocess -> relabel -> adjacency / centers] -> bilinear(align_corners=True) dense features ->
      per-segment mean -> per-pixel MLP -> trav / reco-loss / confidence maps
  learning node (traversability_estimator.py:464-477): one train step on the pooled rows.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import stego_head, wvn_path
from .dino_vit import ViTConfig, vit_feature_map


@torch.no_grad()
def frame_features(img: torch.Tensor, sd: dict, cfg: ViTConfig, hd: dict | None, flip_tta: bool = False,
                   n_image_clusters: int = 0, kmeans_iters: int = 10):
    """img (1,3,H,W).  Returns dict(dense (1,D,H,H), seg (H,H) long | None, feat (S,D) | None,
    edges, centers)."""
    timg = wvn_path.wvn_transform(img, cfg.image_size)
    fmap = vit_feature_map(timg, sd, cfg)
    H = img.shape[2]
    dense = F.interpolate(fmap, (H, H), mode="bilinear", align_corners=True)
    out = {"dense": dense, "fmap": fmap, "seg": None, "feat": None, "edges": None, "centers": None}
    if hd is not None:
        fmap_f = vit_feature_map(timg.flip(dims=[3]), sd, cfg) if flip_tta else None
        _, cluster, _ = stego_head.stego_inference(fmap, fmap_f, hd, (cfg.image_size, cfg.image_size),
                                                   n_image_clusters=n_image_clusters, kmeans_iters=kmeans_iters)
        seg = wvn_path.relabel(cluster[0].long())
        out["seg"] = seg
        out["edges"] = wvn_path.adjacency_list(seg[None, None])
        out["centers"] = wvn_path.centers(seg[None, None])
        out["feat"] = wvn_path.sparsify_features(dense, seg)
    return out


def cpu_step(imgs: torch.Tensor, sd: dict, cfg: ViTConfig, hd: dict, mlp_sd: dict, opt_state, cg_mean, cg_std,
             std_factor: float = 0.5, supervision_seed: int = 2, flip_tta: bool = False, train: bool = True,
             n_image_clusters: int = 0, kmeans_iters: int = 10):
    """Full step over a batch of frames (frame by frame, as the reference's B=1 contract demands).
    Returns (trav [B,H,H], conf [B,H,H], new_mlp_sd, new_opt_state, metrics)."""
    travs, confs, rows = [], [], []
    for b in range(imgs.shape[0]):
        f = frame_features(imgs[b : b + 1], sd, cfg, hd, flip_tta, n_image_clusters, kmeans_iters)
        t, c = wvn_path.pixel_inference(f["dense"], mlp_sd, cg_mean, cg_std, std_factor)
        travs.append(t)
        confs.append(c)
        rows.append(f["feat"])
    metrics = None
    if train:
        x = torch.cat(rows, dim=0)
        y, yv = wvn_path.synthetic_supervision(x.shape[0], seed=supervision_seed)
        y, yv = y.to(x.device), yv.to(x.device)
        mlp_sd, opt_state, metrics = wvn_path.train_step(mlp_sd, opt_state, x, y, yv, lr=1e-3, std_factor=std_factor)
    return torch.stack(travs), torch.stack(confs), mlp_sd, opt_state, metrics
