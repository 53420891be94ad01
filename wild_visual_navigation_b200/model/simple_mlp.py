"""SimpleMLP (reference: wild_visual_navigation/model/simple_mlp.py:10-39).

``D -> 256 -> 32 -> (1 + D)`` with ReLU, sigmoid on column 0, reconstruction head on the rest.
The module structure (``layers.{0,2,4}.{weight,bias}``), the seeded init and the quirk of
mutating the caller's ``hidden_sizes`` list are kept; all parameters are views into one flat
fp32 buffer (``flat_params``) in state-dict order so the fused CUDA train step / inference
kernels work on the very storage that ``state_dict()`` exposes.
"""
from __future__ import annotations

import torch

from .. import ops


class SimpleMLP(torch.nn.Module):
    def __init__(self, input_size: int = 64, hidden_sizes=[255], reconstruction: bool = False):
        super().__init__()
        layers = []
        self.nr_sigmoid_layers = hidden_sizes[-1]
        self.input_size = input_size
        if reconstruction:
            hidden_sizes[-1] = hidden_sizes[-1] + input_size  # mutates the caller's list, as upstream
        inp = input_size
        for hs in hidden_sizes[:-1]:
            layers.append(torch.nn.Linear(inp, hs))
            layers.append(torch.nn.ReLU())
            inp = hs
        layers.append(torch.nn.Linear(inp, hidden_sizes[-1]))
        self.layers = torch.nn.Sequential(*layers)
        self.output_features = hidden_sizes[-1]
        self.hidden = [int(h) for h in hidden_sizes[:-1]]
        self.reconstruction = reconstruction
        self.flat_params = None
        self._flatten()

    # ---- flat storage ---------------------------------------------------------------------
    def _flatten(self):
        ps = list(self.layers.parameters())
        if self.flat_params is not None and ps and ps[0].device == self.flat_params.device:
            off, same = 0, True
            for p in ps:  # already views of the flat buffer (a second .to(same device) must not move the storage the
                same &= p.data_ptr() == self.flat_params.data_ptr() + 4 * off  # CUDA trainer / inference handles hold)
                off += p.numel()
            if same:
                return
        flat = torch.cat([p.detach().reshape(-1) for p in ps]).contiguous()
        off = 0
        for p in ps:
            n = p.numel()
            p.data = flat[off : off + n].view_as(p)
            off += n
        self.flat_params = flat

    def _apply(self, fn, *args, **kwargs):
        super()._apply(fn, *args, **kwargs)
        self._flatten()  # .to(device) re-allocates: rebuild the flat buffer and the views
        return self

    def fused_ok(self) -> bool:
        return (self.reconstruction and len(self.hidden) == 2 and self.nr_sigmoid_layers == 1
                and self.flat_params is not None and self.flat_params.is_cuda)

    # ---- forward --------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, data) -> torch.Tensor:
        """Returns (M, 1 + D) fp32: column 0 through the sigmoid (fp32 CUDA-core kernels).
        The per-pixel node path does not go through here — see ``TraversabilityInference``."""
        if not self.fused_ok():
            raise RuntimeError("SimpleMLP.forward: only the hot-path shape (2 hidden layers, reconstruction, "
                               "1 sigmoid output, parameters on a CUDA device) is implemented; no CPU fallback")
        x = data.x
        return ops.mlp_forward_f32(self.flat_params, x.float(), self.input_size, self.hidden[0], self.hidden[1])
