"""Backbone / head weight sources.

The reference downloads DINO checkpoints from dl.fbaipublicfiles.com at construction
([EXTERNAL] stego.backbones.backbone.get_backbone) and STEGO Lightning checkpoints from disk
(stego_interface.py:23,43).  There is no network here, so:
  * ``pretrained_weights`` / ``$WVN_DINO_WEIGHTS`` may point at a DINO ``.pth`` state dict
    (keys cls_token, pos_embed, patch_embed.proj.*, blocks.N.*, norm.*), which is used as is;
  * otherwise seeded DINO-style random weights are generated — results are then
    numerically meaningful only for parity / throughput work, and a warning says so.
"""
from __future__ import annotations

import os
import warnings

import torch

VIT_SHAPES = {
    "vit_small": dict(dim=384, depth=12, heads=6, mlp_dim=1536),
    "vit_base": dict(dim=768, depth=12, heads=12, mlp_dim=3072),
}


def synthetic_dino_state_dict(dim, depth, mlp_dim, patch_size, pretrain_grid, seed=1, attn_std=0.09):
    g = torch.Generator().manual_seed(seed)

    def n(*shape, std):
        return torch.randn(*shape, generator=g) * std

    sd = {
        "cls_token": n(1, 1, dim, std=0.02),
        "pos_embed": n(1, 1 + pretrain_grid**2, dim, std=0.02),
        "patch_embed.proj.weight": n(dim, 3, patch_size, patch_size, std=0.05),
        "patch_embed.proj.bias": n(dim, std=0.02),
        "norm.weight": 1.0 + n(dim, std=0.05),
        "norm.bias": n(dim, std=0.02),
    }
    for i in range(depth):
        b = f"blocks.{i}."
        sd[b + "norm1.weight"] = 1.0 + n(dim, std=0.05)
        sd[b + "norm1.bias"] = n(dim, std=0.02)
        sd[b + "attn.qkv.weight"] = n(3 * dim, dim, std=attn_std)
        sd[b + "attn.qkv.bias"] = n(3 * dim, std=0.02)
        sd[b + "attn.proj.weight"] = n(dim, dim, std=0.02)
        sd[b + "attn.proj.bias"] = n(dim, std=0.02)
        sd[b + "norm2.weight"] = 1.0 + n(dim, std=0.05)
        sd[b + "norm2.bias"] = n(dim, std=0.02)
        sd[b + "mlp.fc1.weight"] = n(mlp_dim, dim, std=0.04)
        sd[b + "mlp.fc1.bias"] = n(mlp_dim, std=0.02)
        sd[b + "mlp.fc2.weight"] = n(dim, mlp_dim, std=0.02)
        sd[b + "mlp.fc2.bias"] = n(dim, std=0.02)
    return sd


def load_dino_state_dict(backbone_type: str, patch_size: int, pretrained_weights=None, seed: int = 1):
    shp = VIT_SHAPES[backbone_type]
    path = pretrained_weights or os.environ.get("WVN_DINO_WEIGHTS")
    if path:
        sd = torch.load(path, map_location="cpu", weights_only=False)  # DINO full checkpoints pickle argparse objects
        if "state_dict" in sd:
            sd = sd["state_dict"]
        if "teacher" in sd:
            sd = sd["teacher"]
        return {k.replace("module.", "").replace("backbone.", ""): v for k, v in sd.items()}
    warnings.warn("no DINO checkpoint given (pretrained_weights / $WVN_DINO_WEIGHTS): using seeded random "
                  "DINO-style weights — fine for parity and throughput work, not for real traversability")
    return synthetic_dino_state_dict(shp["dim"], shp["depth"], shp["mlp_dim"], patch_size, 224 // patch_size, seed)


def synthetic_stego_head(dim_in: int, code_dim: int = 90, n_clusters: int = 32, n_classes: int = 27, seed: int = 3):
    g = torch.Generator().manual_seed(seed)

    def n(*shape, std):
        return torch.randn(*shape, generator=g) * std

    return {
        "cluster1.0.weight": n(code_dim, dim_in, std=0.05), "cluster1.0.bias": n(code_dim, std=0.05),
        "cluster2.0.weight": n(dim_in, dim_in, std=0.05), "cluster2.0.bias": n(dim_in, std=0.05),
        "cluster2.2.weight": n(code_dim, dim_in, std=0.05), "cluster2.2.bias": n(code_dim, std=0.05),
        "cluster_probe.clusters": n(n_clusters, code_dim, std=1.0),
        "linear_probe.weight": n(n_classes, code_dim, std=0.3), "linear_probe.bias": n(n_classes, std=0.1),
    }


# Column layout of the stacked STEGO head output (see wvn_vit_stego_head in include/wvn_b200.h)
HEAD_CODE_COL, HEAD_CLUSTER_COL, HEAD_LINEAR_COL, HEAD_OUT = 0, 128, 192, 256


def fold_stego_head(hd: dict) -> dict:
    """Stack code / cluster-probe / linear-probe into the two affine maps the CUDA head applies:
        out = head_a(t) + head_b(relu(hidden(t)))
    rows [0,90): code = cluster1(t) + cluster2.2(relu(cluster2.0(t)));
    rows [128,128+K): normalize(clusters) @ code  (cluster-probe logits up to a positive per-pixel scale);
    rows [192,192+C): linear_probe(code).
    Exact algebra in fp32 on the host (init-time), so no bf16 rounding of the code sits between
    the head and its probes."""
    w1, b1 = hd["cluster1.0.weight"].float(), hd["cluster1.0.bias"].float()
    w2, b2 = hd["cluster2.2.weight"].float(), hd["cluster2.2.bias"].float()
    cn = torch.nn.functional.normalize(hd["cluster_probe.clusters"].float(), dim=1)
    wl, bl = hd["linear_probe.weight"].float(), hd["linear_probe.bias"].float()
    code_dim, dim_in = w1.shape
    K, C = cn.shape[0], wl.shape[0]
    assert code_dim <= 128 and K <= 64 and C <= 64, "head layout supports code<=128, clusters<=64, classes<=64"
    a_w = torch.zeros(HEAD_OUT, dim_in)
    b_w = torch.zeros(HEAD_OUT, dim_in)
    a_b = torch.zeros(HEAD_OUT)
    a_w[:code_dim], b_w[:code_dim], a_b[:code_dim] = w1, w2, b1 + b2
    a_w[HEAD_CLUSTER_COL : HEAD_CLUSTER_COL + K] = cn @ w1
    b_w[HEAD_CLUSTER_COL : HEAD_CLUSTER_COL + K] = cn @ w2
    a_b[HEAD_CLUSTER_COL : HEAD_CLUSTER_COL + K] = cn @ (b1 + b2)
    a_w[HEAD_LINEAR_COL : HEAD_LINEAR_COL + C] = wl @ w1
    b_w[HEAD_LINEAR_COL : HEAD_LINEAR_COL + C] = wl @ w2
    a_b[HEAD_LINEAR_COL : HEAD_LINEAR_COL + C] = wl @ (b1 + b2) + bl
    return {"head_a.weight": a_w, "head_a.bias": a_b, "hidden.weight": hd["cluster2.0.weight"].float(),
            "hidden.bias": hd["cluster2.0.bias"].float(), "head_b.weight": b_w}
