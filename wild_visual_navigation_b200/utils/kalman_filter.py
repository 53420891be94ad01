"""KalmanFilter (reference: wild_visual_navigation/utils/kalman_filter.py:10-117) — the linear filter behind
``ConfidenceGenerator(method="kalman_filter")``.

Same parameters (so ``state_dict`` carries ``proc_model, proc_cov, control_model, meas_model, meas_cov, eye``), same
``init_process_model / init_meas_model / prediction / correction / forward`` surface.  Inside the fused train step
(csrc/mlp_train_fused.cu: train_conf_kernel) the D = 1 filter of the ConfidenceGenerator is evaluated in-kernel; this
module is the general-dimension form for direct calls.
"""
from __future__ import annotations

import torch
from torch import nn


class KalmanFilter(nn.Module):
    def __init__(self, dim_state: int = 1, dim_control: int = 1, dim_meas: int = 1, outlier_rejection: str = "none",
                 outlier_delta: float = 1.0):
        super().__init__()
        self.dim_state, self.dim_control, self.dim_meas = dim_state, dim_control, dim_meas
        self.proc_model = nn.Parameter(torch.eye(dim_state))
        self.proc_cov = nn.Parameter(torch.eye(dim_state))
        self.control_model = nn.Parameter(torch.eye(dim_state, dim_control))
        self.meas_model = nn.Parameter(torch.eye(dim_meas, dim_state))
        self.meas_cov = nn.Parameter(torch.eye(dim_meas, dim_meas))
        self.eye = nn.Parameter(torch.eye(dim_state, dim_state), requires_grad=False)
        self.outlier_rejection, self.outlier_delta = outlier_rejection, outlier_delta

    def _set(self, name, value):
        if value is not None:
            cur = getattr(self, name)
            assert cur.shape == value.shape, f"{name}: {tuple(cur.shape)} (expected) != {tuple(value.shape)} (new)"
            setattr(self, name, nn.Parameter(value))

    def init_process_model(self, proc_model=None, proc_cov=None, control_model=None):
        self._set("proc_model", proc_model)
        self._set("proc_cov", proc_cov)
        self._set("control_model", control_model)

    def init_meas_model(self, meas_model=None, meas_cov=None):
        self._set("meas_model", meas_model)
        self._set("meas_cov", meas_cov)

    def prediction(self, state, state_cov, control=None):
        state = self.proc_model @ state if control is None else self.proc_model @ state + self.control_model @ control
        return state, self.proc_model @ state_cov @ self.proc_model.t() + self.proc_cov

    def get_outlier_weight(self, error, cov):
        if self.outlier_rejection == "none":
            return 1.0
        r = torch.sqrt(error.t() @ cov.inverse() @ error)
        if self.outlier_rejection == "hard":
            return torch.tensor([0.0]) if r.item() >= self.outlier_delta else torch.tensor([1.0])
        if self.outlier_rejection == "huber":
            return 1.0 if r.abs() <= self.outlier_delta else (self.outlier_delta / r.abs()).item()
        return 1.0

    def correction(self, state, state_cov, meas):
        innovation = meas - self.meas_model @ state
        weight = self.get_outlier_weight(innovation, self.meas_cov)
        innovation_cov = self.meas_model @ state_cov @ self.meas_model.t() + self.meas_cov
        gain = weight * state_cov @ self.meas_model.t() @ innovation_cov.inverse()
        return state + gain @ innovation, (self.eye - gain @ self.meas_model) @ state_cov

    def forward(self, state, state_cov, meas, control=None):
        state, state_cov = self.prediction(state, state_cov, control)
        return self.correction(state, state_cov, meas)
