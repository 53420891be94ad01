"""ConfidenceGenerator (reference: wild_visual_navigation/utils/confidence_generator.py:13-212).

State (``mean (1,)``, ``var (1,1)``, ``std (1,)``) lives in non-trainable parameters so it rides in
``state_dict`` exactly as upstream; the CUDA kernels read ``mean`` / ``std`` through device
pointers (no host sync on the per-frame path).  Only the default method ``latest_measurement``
(cfg/experiment_params.py:54) is on the hot path; its update is fused into the train step
(csrc/mlp_train.cu), and ``inference_without_update`` into the per-pixel MLP epilogue.  The
elementwise fallbacks below are tensor plumbing for direct calls outside the fused paths.
"""
from __future__ import annotations

import torch


class ConfidenceGenerator(torch.nn.Module):
    def __init__(self, std_factor, method, log_enabled: bool = False, log_folder: str = "/tmp"):
        super().__init__()
        if method not in ("latest_measurement",):
            raise ValueError(f"method '{method}' is not on the B200 hot path (only 'latest_measurement')")
        self.std_factor = std_factor
        self.method = method
        self.log_enabled = log_enabled
        self.log_folder = log_folder
        self.mean = torch.nn.Parameter(torch.zeros(1, dtype=torch.float32), requires_grad=False)
        self.var = torch.nn.Parameter(torch.ones((1, 1), dtype=torch.float32), requires_grad=False)
        self.std = torch.nn.Parameter(torch.ones(1, dtype=torch.float32), requires_grad=False)

    @torch.no_grad()
    def update(self, x: torch.Tensor, x_positive: torch.Tensor, step: int = 0, log_step: bool = False):
        self.mean[0] = x_positive.mean()
        self.std[0] = x_positive.std()
        return self.inference_without_update(x)

    @torch.no_grad()
    def inference_without_update(self, x: torch.Tensor):
        if x.device != self.mean.device:
            return torch.zeros_like(x)
        shifted_mean = self.mean + self.std * self.std_factor
        interval_min = torch.maximum(shifted_mean - self.std, torch.zeros_like(self.std))
        interval_max = shifted_mean + self.std
        x = torch.clip(x, interval_min, interval_max)
        return (1 - ((x - interval_min) / (interval_max - interval_min))).type(torch.float32)

    def reset(self):
        with torch.no_grad():
            self.mean[0] = 0
            self.var[0] = 1
            self.std[0] = 1

    def get_dict(self):
        return {"mean": self.mean, "var": self.var, "std": self.std}
