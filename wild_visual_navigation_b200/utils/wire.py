"""``ImageFeatures`` wire format without ROS.

The feature-extractor node publishes, per training camera frame (wvn_feature_extractor_node.py:373-393,
wild_visual_navigation_msgs/msg/ImageFeatures.msg):
    header            std_msgs/Header
    feature_segments  sensor_msgs/Image, ``seg.astype(int32)`` with "passthrough" encoding (one int32 per pixel)
    features          std_msgs/Float32MultiArray: data = feat.flatten() (row-major (S, D)), layout.dim =
                      [("n", S, S*D), ("feat", D, D)]
and the learning node decodes it back (wvn_learning_node.py:651-656) with
``np.array(data, dtype=float).reshape(dims).astype(np.float32)``.

These two functions produce / consume exactly that field layout as plain Python containers, so a thin ROS shim
(``msg.features.data = d["features"]["data"]`` ...) or any other transport can carry the product's outputs to an
unmodified learning node.  float32 -> Python float -> float32 is exact, so the round trip is bit-identical.
"""
from __future__ import annotations

import numpy as np
import torch


def encode_image_features(feat: torch.Tensor, seg: torch.Tensor, header: dict | None = None) -> dict:
    """feat (S, D) float, seg (H, W) integer -> dict mirroring ImageFeatures.msg."""
    feat_np = feat.detach().to("cpu", torch.float32).contiguous().numpy()
    seg_np = seg.detach().to("cpu").numpy().astype(np.int32)
    S, D = feat_np.shape
    H, W = seg_np.shape
    return {
        "header": dict(header or {}),
        "feature_segments": {"header": dict(header or {}), "height": H, "width": W, "encoding": "32SC1",
                             "is_bigendian": 0, "step": W * 4, "data": seg_np.tobytes()},
        "features": {"layout": {"dim": [{"label": "n", "size": S, "stride": S * D},
                                        {"label": "feat", "size": D, "stride": D}], "data_offset": 0},
                     "data": feat_np.flatten().tolist()},
    }


def decode_image_features(msg: dict, device="cpu"):
    """Inverse, with the learning node's own decoding arithmetic.  Returns (features (S, D) f32, segments (H, W) i32)."""
    ma = msg["features"]
    dims = tuple(d["size"] for d in ma["layout"]["dim"])
    features = torch.from_numpy(np.array(ma["data"], dtype=float).reshape(dims).astype(np.float32)).to(device)
    im = msg["feature_segments"]
    seg = np.frombuffer(im["data"], dtype=np.int32).reshape(im["height"], im["width"]).copy()
    return features, torch.from_numpy(seg).to(device)
