"""SegmentExtractor (reference: wild_visual_navigation/feature_extractor/segment_extractor.py:11-92).

Same two methods and return conventions; both run as one pass over the segmentation map plus a
tiny emit kernel (csrc/segment_kernels.cu) instead of 4 convolutions + float64 unique and a
Python loop with one host sync per segment.
"""
from __future__ import annotations

import torch

from .. import ops


class SegmentExtractor(torch.nn.Module):
    def __init__(self):
        super().__init__()

    @torch.no_grad()
    def adjacency_list(self, seg: torch.Tensor):
        """seg: (1,1,H,W) long -> (N,2) long, directed (left/top, right/bottom) pairs sorted like
        ``torch.unique(left + right*(max+1))`` in the reference."""
        assert seg.shape[0] == 1 and len(seg.shape) == 4, f"{seg.shape}"
        s = seg[0].long().contiguous()
        smax = int(s.max().item()) + 1
        r = ops.segment_reduce(s, smax, want_centers=False, want_edges=True, max_edges=smax * smax)
        n = int(r["n_edges"][0].item())
        return r["edges"][0, :n].clone()

    @torch.no_grad()
    def centers(self, seg: torch.Tensor):
        """seg: (1,1,H,W) long -> (S,2) float32 centroids in (x=col, y=row) order."""
        assert seg.shape[0] == 1 and len(seg.shape) == 4
        s = seg[0].long().contiguous()
        smax = int(s.max().item()) + 1
        r = ops.segment_reduce(s, smax, want_centers=True, want_edges=False)
        return r["centers"][0]
