"""Minimal ``Data`` / ``Batch`` containers (reference: wild_visual_navigation/utils/data.py:11-58).

``Data`` is an attribute bag; ``Batch.from_data_list`` concatenates every public tensor attribute
along dim 0, offsets ``edge_index`` by the running node count and records ``ptr`` / ``batch``.
(The reference mutates the class object itself; an instance is returned here so that two live
batches do not alias — the attribute names and values are the same.)
"""
from __future__ import annotations

from typing import List

import torch


class Data:
    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)


class Batch:
    @classmethod
    def from_data_list(cls, list_of_data: List[Data]):
        if len(list_of_data) == 0:
            return None
        out = cls()
        first = list_of_data[0]
        keys = ["x"] + [k for k in dir(first) if k[0] != "_" and getattr(first, k) is not None and k != "x"]
        running, ptrs, batches = 0, [0], []
        for j, d in enumerate(list_of_data):
            n = int(d.x.shape[0])
            running += n
            ptrs.append(running)
            batches += [j] * n
        out.ptr = torch.tensor(ptrs, dtype=torch.long)
        out.batch = torch.tensor(batches, dtype=torch.long)
        for k in keys:
            if k == "edge_index":
                out.edge_index = torch.cat([getattr(d, k) + out.ptr[j] for j, d in enumerate(list_of_data)], dim=-1)
            else:
                out.__dict__[k] = torch.cat([getattr(d, k) for d in list_of_data], dim=0)
        out.ba = out.x.shape[0]
        return out
