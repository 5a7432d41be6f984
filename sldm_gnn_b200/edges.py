"""Device-side construction of the vehicle-graph edges (reference: the double loops at src/gbuilder.py:88-112 and
:244-268, the latter run per sliding window by rcv.py:77).

    edge_index, edge_attr = build_proximity_edges(x, m_radius)

`x` is the `[V, T, F]` fp32 trajectory tensor of a pack BEFORE the heading is re-encoded (feature 0 = X, 1 = Y,
4 = presence flag, as at gbuilder.py:90-96); the result is what the reference stores in `Data.edge_index` (int64
`[2, E]`, ordered by (i, j)) and `Data.edge_attr` (`[E, 4]` = min / max / mean / mean-square distance).  One host
read of the edge count sizes the outputs (tensor shapes live on the host); everything else stays on the device.
CUDA only.
"""
from __future__ import annotations

import torch

from ._lib import lib, check
from .ops import _require_cuda, _stream


def build_proximity_edges(x: torch.Tensor, m_radius: float):
    if not isinstance(x, torch.Tensor) or x.dim() != 3:
        raise RuntimeError("build_proximity_edges: x must be a [vehicles, frames, features] tensor")
    _require_cuda(x, "x")
    if x.dtype != torch.float32:
        raise RuntimeError(f"build_proximity_edges: expected float32 trajectories, got {x.dtype}")
    x = x.contiguous()
    V, T, F = x.shape
    dev = x.device
    with torch.cuda.device(dev):
        counts = torch.empty((max(V, 1),), dtype=torch.int32, device=dev)
        offsets = torch.empty((V + 1,), dtype=torch.int32, device=dev)
        check(lib.sldm_edge_build_count(x.data_ptr() if V > 0 else None, V, T, F, float(m_radius),
                                        counts.data_ptr(), offsets.data_ptr(), _stream(dev)))
        E = int(offsets[V])                                  # the one host read: output shapes
        edge_index = torch.empty((2, E), dtype=torch.long, device=dev)
        edge_attr = torch.empty((E, 4), dtype=torch.float32, device=dev)
        check(lib.sldm_edge_build_fill(x.data_ptr() if V > 0 else None, V, T, F, float(m_radius), offsets.data_ptr(), E,
                                       edge_index.data_ptr() if E > 0 else None, edge_attr.data_ptr() if E > 0 else None,
                                       _stream(dev)))
    return edge_index, edge_attr
