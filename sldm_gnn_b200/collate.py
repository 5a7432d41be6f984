"""PyG-free mini-batch container and device-side collate.

The reference builds its batches with torch_geometric.loader.DataLoader (main.py:166-167; consumed at
src/utils.py:218-223 and src/models/grusage.py:153-173): graphs are loaded onto the device one at a time
(src/dataset.py:75-89) and PyG's Batch.from_data_list concatenates them.  `collate` does the same assembly with one
kernel launch per attribute (libsldm_sage.so: sldm_concat_chunks / sldm_collate_graph_index):

    batch = collate([GraphData(x=..., edge_index=..., xsttype=..., xdims=..., pos_raw=..., y=...), ...])
    batch.x, batch.edge_index, batch.batch, batch.ptr, batch.num_graphs, batch.y ...

PyG's rules, restated: a tensor attribute is concatenated along dim 0, except attributes whose name contains "index"
(edge_index), which are concatenated along the last dim after adding the graph's node offset; the number of nodes of a
graph is x.size(0); `batch` and `ptr` are added.  CUDA only (no CPU fallback): all tensors of all graphs must be on one
CUDA device.
"""
from __future__ import annotations

import torch

from ._lib import lib, check
from .ops import _require_cuda, _stream


class GraphData:
    """Attribute bag standing where torch_geometric.data.Data stands (src/gbuilder.py:133,300)."""

    def __init__(self, **fields):
        for k, v in fields.items():
            setattr(self, k, v)

    def keys(self):
        return [k for k in vars(self) if not k.startswith("_")]

    @property
    def num_nodes(self) -> int:
        return int(self.x.size(0))

    def to(self, device, non_blocking: bool = False):
        for k in self.keys():
            v = getattr(self, k)
            if isinstance(v, torch.Tensor):
                setattr(self, k, v.to(device, non_blocking=non_blocking))
        return self


class GraphBatch(GraphData):
    """What the training loop and GruSage.forward read from a PyG Batch: the concatenated attributes plus
    `batch`, `ptr`, `num_graphs` (src/utils.py:219-223, src/models/grusage.py:153)."""


def collate(data_list) -> GraphBatch:
    if len(data_list) == 0:
        raise ValueError("collate: empty list of graphs")
    first = data_list[0]
    keys = first.keys()
    if "x" not in keys or "edge_index" not in keys:
        raise ValueError("collate: every graph needs `x` and `edge_index`")
    G = len(data_list)
    dev = first.x.device
    _require_cuda(first.x, "x")
    out = GraphBatch()
    nodes = [int(d.x.size(0)) for d in data_list]
    ptr_h = [0] * (G + 1)
    for g, n in enumerate(nodes):
        ptr_h[g + 1] = ptr_h[g] + n
    N = ptr_h[G]
    # pass 1 (host only): one table for ALL attributes -- [ptr (G+1, padded to a multiple of 4) | G rows of 4 per attribute]
    # so that a single pinned upload feeds every launch
    flat = list(ptr_h) + [0] * ((-(G + 1)) % 4)
    plan = []
    with torch.cuda.device(dev):
        for g, d in enumerate(data_list):                   # every graph carries the same attributes (PyG's collate assumes it)
            if set(d.keys()) != set(keys):
                raise ValueError(f"collate: graph {g} has attributes {sorted(d.keys())}, graph 0 has {sorted(keys)}")
        for k in keys:
            vals = [getattr(d, k) for d in data_list]
            if k == "edge_index" and not all(isinstance(v, torch.Tensor) for v in vals):
                raise ValueError("collate: `edge_index` must be a tensor in every graph")
            if not all(isinstance(v, torch.Tensor) for v in vals):
                setattr(out, k, vals)                       # non-tensor attributes: a list, like PyG
                continue
            for v in vals:
                if v.device != dev:
                    raise RuntimeError(f"collate: attribute `{k}` is on {v.device}, expected {dev}")
            if "index" in k:                                # edge_index: cat along the last dim with node offsets
                vals = [v if v.stride(-1) == 1 else v.contiguous() for v in vals]
                for v in vals:
                    if v.dtype != torch.long or v.dim() != 2 or v.size(0) != 2:
                        raise ValueError(f"Expected '{k}' to be an int64 tensor of shape [2, num_edges]")
                e = [int(v.size(1)) for v in vals]
                res = torch.empty((2, sum(e)), dtype=torch.long, device=dev)
                row0, off = len(flat), 0
                for v, ei in zip(vals, e):
                    flat += [v.data_ptr(), off, ei, int(v.stride(0)) if ei > 0 else 0]
                    off += ei
                plan.append(("index", k, res, row0, off, max(e)))
            else:                                           # cat along dim 0
                vals = [v.contiguous() for v in vals]
                tail, dt = tuple(vals[0].shape[1:]), vals[0].dtype
                rows_total = 0
                for v in vals:
                    if tuple(v.shape[1:]) != tail or v.dtype != dt:
                        raise RuntimeError(f"collate: attribute `{k}` has inconsistent shapes / dtypes across graphs")
                    rows_total += int(v.size(0))
                res = torch.empty((rows_total,) + tail, dtype=dt, device=dev)
                row0, off, mx = len(flat), 0, 0
                for v in vals:
                    nb = v.numel() * v.element_size()
                    flat += [v.data_ptr() if nb > 0 else 0, off, nb, 0]
                    off += nb
                    mx = max(mx, nb)
                plan.append(("cat", k, res, row0, off, mx))
        table = torch.tensor(flat, dtype=torch.int64).pin_memory().to(dev, non_blocking=True)
        # pass 2: one launch per attribute
        stream = _stream(dev)
        tb = table.data_ptr()
        batch = torch.empty((N,), dtype=torch.long, device=dev)
        have_batch = False
        for kind, k, res, row0, total, mx in plan:
            if kind == "index":
                want_batch = (k == "edge_index")
                check(lib.sldm_collate_graph_index(tb + 8 * row0, tb, G, total, mx, max(nodes),
                                                   res.data_ptr() if total > 0 else None,
                                                   batch.data_ptr() if (want_batch and N > 0) else None, stream))
                have_batch |= want_batch
            else:
                check(lib.sldm_concat_chunks(tb + 8 * row0, G, mx, res.data_ptr() if total > 0 else None, stream))
            setattr(out, k, res)
        assert have_batch or N == 0, "collate: the edge_index launch fills `batch`"   # edge_index is always a tensor here
        out.batch = batch
        out.ptr = table[:G + 1]
        out.num_graphs = G
        # nothing else needs to be kept: the kernels are enqueued on the stream that owns every source tensor, so the
        # caching allocator cannot hand their memory out before the copies have run
    return out
