"""ConfidenceGenerator (reference: wild_visual_navigation/utils/confidence_generator.py:13-212).

State (``mean (1,)``, ``var (1,1)``, ``std (1,)``, and per method ``running_n / running_sum /
running_sum_of_squares`` (fp64) or the ``_kalman_filter`` sub-module) lives in non-trainable parameters so it rides in
``state_dict`` exactly as upstream; the CUDA kernels read and update it through device pointers (no host sync on the
per-frame path).  All four methods — ``latest_measurement`` (the configured default, cfg/experiment_params.py:54),
``running_mean``, ``moving_average``, ``kalman_filter`` — are evaluated inside the fused train step
(csrc/mlp_train_fused.cu: train_conf_kernel / row_confidence) and ``inference_without_update`` inside the per-pixel MLP
epilogue.  The methods below are the same updates as tensor plumbing for direct calls outside the fused paths.
"""
from __future__ import annotations

from collections import deque

import torch

from .kalman_filter import KalmanFilter

METHODS = {"latest_measurement": 0, "running_mean": 1, "kalman_filter": 2, "moving_average": 3}


class ConfidenceGenerator(torch.nn.Module):
    def __init__(self, std_factor, method, log_enabled: bool = False, log_folder: str = "/tmp"):
        super().__init__()
        if method not in METHODS:
            raise ValueError("Unknown method")
        self.std_factor = std_factor
        self.method = method
        self.log_enabled = log_enabled
        self.log_folder = log_folder
        self.mean = torch.nn.Parameter(torch.zeros(1, dtype=torch.float32), requires_grad=False)
        self.var = torch.nn.Parameter(torch.ones((1, 1), dtype=torch.float32), requires_grad=False)
        self.std = torch.nn.Parameter(torch.ones(1, dtype=torch.float32), requires_grad=False)
        if method == "kalman_filter":
            self._kalman_filter = KalmanFilter(dim_state=1, dim_control=1, dim_meas=1)
            self._kalman_filter.init_process_model(proc_model=torch.eye(1) * 1, proc_cov=torch.eye(1) * 0.2)
            self._kalman_filter.init_meas_model(meas_model=torch.eye(1), meas_cov=torch.eye(1) * 1.0)
        elif method == "running_mean":
            for name in ("running_n", "running_sum", "running_sum_of_squares"):
                setattr(self, name, torch.nn.Parameter(torch.zeros(1, dtype=torch.float64), requires_grad=False))
        elif method == "moving_average":
            self.data_window = deque(maxlen=5)

    @property
    def method_id(self) -> int:
        return METHODS[self.method]

    @torch.no_grad()
    def update(self, x: torch.Tensor, x_positive: torch.Tensor, step: int = 0, log_step: bool = False):
        if self.method == "latest_measurement":
            self.mean[0] = x_positive.mean()
            self.std[0] = x_positive.std()
            return self.inference_without_update(x)
        if self.method == "running_mean":
            self.running_n += x_positive.numel()
            self.running_sum += x_positive.sum()
            self.running_sum_of_squares += (x_positive**2).sum()
            self.mean[0] = self.running_sum[0] / self.running_n
            self.var[0] = self.running_sum_of_squares / self.running_n - self.mean**2
            self.std[0] = torch.sqrt(self.var)
            return self.inference_without_update(x)
        if self.method == "moving_average":
            self.data_window.append(x_positive)
            window = torch.cat(list(self.data_window), dim=0)
            self.mean[0] = window.mean()
            self.std[0] = window.std()
            xc = torch.clip(x, self.mean - 2 * self.std, self.mean + 2 * self.std)
            return ((xc - xc.min()) / (xc.max() - xc.min())).type(torch.float32)
        # kalman_filter
        if x_positive.shape[0] != 0:
            mean, var = self._kalman_filter(self.mean, self.var, x_positive.mean())
            self.var[0, 0] = var[0, 0]
            self.mean[0] = mean[0]
        self.std[0] = torch.sqrt(self.var)[0, 0]
        confidence = torch.exp(-(((x - self.mean) / (self.std * self.std_factor)) ** 2) * 0.5)
        confidence[x < self.mean] = 1.0
        return confidence.type(torch.float32)

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
            if self.method == "running_mean":   # reset_running_mean (:199-202) clears only the sums
                self.running_n[0] = 0
                self.running_sum[0] = 0
                self.running_sum_of_squares[0] = 0
            else:
                self.mean[0] = 0
                self.var[0] = 1
                self.std[0] = 1

    def get_dict(self):
        return {"mean": self.mean, "var": self.var, "std": self.std}
