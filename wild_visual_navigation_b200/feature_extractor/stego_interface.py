"""StegoInterface (reference: wild_visual_navigation/feature_extractor/stego_interface.py:19-135).

``inference(img) -> (linear_pred, cluster_pred)`` with the ``features`` / ``cluster_segments`` /
``linear_segments`` properties.  The STEGO head runs as three tcgen05 GEMMs on the ViT tokens
(csrc/api.cu: wvn_vit_stego_head); the cluster / linear probes are folded into the head's output
columns (weights.fold_stego_head) and evaluated at patch resolution, then one kernel does the
bilinear(align_corners=False) upsampling + argmax per pixel — algebraically the upstream
``postprocess`` (upsample the 90-d code, then probe every pixel) without the 448x448x90 tensor.
``run_clustering=True`` (WVN's default through ``FeatureExtractor``): the cluster prediction comes from a per-image
k-means of the code with ``n_image_clusters`` clusters (csrc/stego_kmeans.cu — one launch per batch of frames; the
nearest-centroid scores replace the cluster-probe logits before the same upsample+argmax kernel).  CRF is out of scope
(``run_crf`` must be False; the reference's own default for WVN, feature_extractor.py:52).
"""
from __future__ import annotations

import torch

from .. import ops
from .dino_interface import DinoInterface, _Cfg
from .weights import (HEAD_CLUSTER_COL, HEAD_CODE_COL, HEAD_LINEAR_COL, fold_stego_head, synthetic_stego_head)


class StegoInterface:
    def __init__(self, device: str, input_size: int = 448, model_path: str = None, n_image_clusters: int = 40,
                 run_crf: bool = False, run_clustering: bool = False, cfg=None, backbone_type: str = "vit_small",
                 patch_size: int = 8, head_state_dict=None, backbone_state_dict=None, flip_tta: bool = True,
                 max_batch: int = 32, chunk: int = 0, code_dim: int = 90, kmeans_iters: int = 10):
        self._cfg = _Cfg(cfg) if cfg else _Cfg(model_path=model_path, input_size=input_size, run_crf=run_crf,
                                               run_clustering=run_clustering, n_image_clusters=n_image_clusters)
        if self._cfg.run_crf:
            raise ValueError("run_crf (pydensecrf on the CPU) is outside the B200 hot path; WVN itself runs with "
                             "run_crf=False (feature_extractor.py:52)")
        self._kmeans_iters = kmeans_iters
        self._device = device
        self._flip_tta = flip_tta
        if head_state_dict is None:
            if model_path:
                # a STEGO Lightning checkpoint pickles OmegaConf hyper-parameters: needs the full unpickler (trusted file)
                ck = torch.load(model_path, map_location="cpu", weights_only=False)
                sd = ck.get("state_dict", ck)
                head_state_dict = {k.replace("segmentation_head.", ""): v.squeeze(-1).squeeze(-1) if v.dim() == 4 else v
                                   for k, v in sd.items() if k.startswith(("segmentation_head.", "cluster_probe.",
                                                                           "linear_probe."))}
                if backbone_state_dict is None:
                    backbone_state_dict = {k.split("backbone.model.", 1)[1]: v for k, v in sd.items()
                                           if "backbone.model." in k} or None
            else:
                from .weights import VIT_SHAPES
                head_state_dict = synthetic_stego_head(VIT_SHAPES[backbone_type]["dim"], code_dim)
        self._head = head_state_dict
        self._code_dim = int(head_state_dict["cluster1.0.weight"].shape[0])
        self._n_clusters = int(head_state_dict["cluster_probe.clusters"].shape[0])
        if self._cfg.run_clustering and not (1 <= self._cfg.n_image_clusters <= 64):
            raise ValueError("n_image_clusters must be in [1, 64]")
        self._n_classes = int(head_state_dict["linear_probe.weight"].shape[0])
        self._dino = DinoInterface(device, input_size=input_size, backbone_type=backbone_type, patch_size=patch_size,
                                   max_batch=max_batch * (2 if flip_tta else 1), chunk=chunk,
                                   state_dict=backbone_state_dict, head_weights=fold_stego_head(head_state_dict))
        self._code = self._cluster_pred = self._linear_pred = None
        self._tokens = None

    def change_device(self, device):
        self._dino.change_device(device)

    @torch.no_grad()
    def inference(self, img: torch.Tensor):
        """img (B,3,H,W) -> (linear_pred, cluster_pred), each (1,B,H,H) int32 like the reference.
        Camera frames (B,H0,W0,3) uint8 are accepted too: they are resized (NEAREST) / center-cropped to
        ``input_size`` inside the patch loader, i.e. H = W = input_size as after ``ImageProjector.resize_image``."""
        if img.dtype == torch.uint8:
            img = img.to(self._device)
            B, H, W = img.shape[0], self._cfg.input_size, self._cfg.input_size
            assert not self._flip_tta, "flip TTA needs the float image (flip is applied to the transformed image)"
        else:
            img = img.to(self._device, dtype=torch.float32)
            B, _, H, W = img.shape
        vit = self._dino._model
        g, npad = vit.grid, vit.npad
        if self._flip_tta:
            # Stego.get_code: average with the pass over the horizontally flipped TRANSFORMED image (the flip happens
            # inside the patch loader, after resize + crop); flip back at patch level
            tokens = vit.forward(img, flip_tta=True)
            out = vit.stego_head(2 * B).view(2, B, npad, -1)
            a = out[0, :, 1 : 1 + g * g].reshape(B, g, g, -1)
            b = out[1, :, 1 : 1 + g * g].reshape(B, g, g, -1).flip(dims=[2])
            head = torch.zeros(B, npad, out.shape[-1], device=img.device)
            head[:, 1 : 1 + g * g] = ((a + b) * 0.5).reshape(B, g * g, -1)
            head = head.view(B * npad, -1)
            self._tokens = tokens[:B]
        else:
            self._tokens = vit.forward(img)
            head = vit.stego_head(B)
        S = self._cfg.input_size
        n_cluster_logits = self._n_clusters
        if self._cfg.run_clustering:
            # per-image k-means of the code: the nearest-centroid scores overwrite the cluster-probe logit columns
            n_cluster_logits = self._cfg.n_image_clusters
            ops.check(ops.lib().wvn_stego_kmeans(ops.ptr(head), head.stride(0), B, npad, g * g, HEAD_CODE_COL,
                                                 self._code_dim, HEAD_CLUSTER_COL, n_cluster_logits, self._kmeans_iters,
                                                 None, ops.stream()))
        cl, li = ops.logits_argmax(head, HEAD_CLUSTER_COL, n_cluster_logits, B, npad, g, g, S, S,
                                   col0_b=HEAD_LINEAR_COL, classes_b=self._n_classes)
        code = head.view(B, npad, -1)[:, 1 : 1 + g * g, HEAD_CODE_COL : HEAD_CODE_COL + self._code_dim].contiguous()
        self._code_tokens = code  # (B, P, 90) at patch resolution — what the fused consumers use
        self._code = None         # dense (B,90,H,H) only on demand (property `features`)
        self._img_hw = (H, W)
        if (S, S) != (H, H):
            cl = torch.nn.functional.interpolate(cl[None].float(), (H, H), mode="nearest")[0].long()
            li = torch.nn.functional.interpolate(li[None].float(), (H, H), mode="nearest")[0].long()
        self._cluster_pred = cl[None].int()
        self._linear_pred = li[None].int()
        return self._linear_pred, self._cluster_pred

    @property
    def model(self):
        return self._dino._model

    @property
    def max_segments(self):
        """Upper bound of the cluster ids ``cluster_segments`` can hold (+1)."""
        return self._cfg.n_image_clusters if self._cfg.run_clustering else self._n_clusters

    @property
    def input_size(self):
        return self._cfg.input_size

    @property
    def linear_segments(self):
        return self._linear_pred

    @property
    def cluster_segments(self):
        return self._cluster_pred

    @property
    def code_tokens(self):
        return self._code_tokens

    @property
    def backbone_tokens(self):
        return self._tokens

    @property
    def features(self):
        """Dense (B, 90, H, H) code, bilinear align_corners=True (stego_interface.py:107) — materialised lazily."""
        if self._code is None:
            g = self._dino.grid
            H = self._img_hw[0]
            self._code = ops.upsample_dense(self._code_tokens, g, g, H, H)
        return self._code
