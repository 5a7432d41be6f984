"""Graph-sharded data parallelism for the batched-graph workload.

The reference is single-process (main.py:58; no torch.distributed anywhere).  A PyG
batch is block-diagonal -- independent graphs, no cross-graph edges -- so whole graphs
are the unit of sharding: every rank builds its own CSR and runs the full SageBlock on
its graphs with NO data-path collective.  Training needs exactly one exchange step: the
sum of the parameter gradients.  All parameters of the wrapped module live in ONE flat
fp32 buffer and so do their gradients, so that exchange is a single all-reduce of
~174 KB ([128,96,96]) over NCCL / NVLink, latency bound by construction.

LayerNorm is per node, so there are no cross-rank statistics.  A loss that is a mean
over the local graphs needs the per-rank gradient weighted by the local graph count to
equal single-process maths; the weight rides in the last slot of the same bucket, so it
costs no extra collective.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn as nn


def shard_graphs(num_graphs: int, rank: int, world_size: int) -> range:
    """Contiguous, balanced split of graph ids (whole graphs only)."""
    base, rem = divmod(num_graphs, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


class GraphDataParallel(nn.Module):
    def __init__(self, module: nn.Module, process_group=None, broadcast: bool = True):
        super().__init__()
        self.module = module
        self.process_group = process_group
        params = [p for p in module.parameters() if p.requires_grad]
        if not params:
            raise ValueError("GraphDataParallel: module has no trainable parameters")
        dev, dt = params[0].device, params[0].dtype
        if any(p.device != dev or p.dtype != dt for p in params):
            raise ValueError("GraphDataParallel: parameters must share one device and dtype")
        total = sum(p.numel() for p in params)
        # one flat buffer for values, one for gradients (+1 slot: the loss weight)
        self._flat = torch.empty(total, dtype=dt, device=dev)
        self._flat_grad = torch.zeros(total + 1, dtype=dt, device=dev)
        self._params, self._views = params, []
        off = 0
        with torch.no_grad():
            for p in params:
                n = p.numel()
                self._flat[off:off + n].copy_(p.reshape(-1))
                p.data = self._flat[off:off + n].view_as(p)          # same names, same shapes
                gview = self._flat_grad[off:off + n].view_as(p)
                p.grad = gview
                self._views.append(gview)
                off += n
        if broadcast and self._ready():
            dist.broadcast(self._flat, src=self._src_rank(), group=self.process_group)

    def _ready(self) -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.process_group) > 1

    def _src_rank(self) -> int:
        return dist.get_global_rank(self.process_group, 0) if self.process_group is not None else 0

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def zero_grad(self, set_to_none: bool = False) -> None:  # keeps the flat views attached
        self._flat_grad.zero_()
        for p, v in zip(self._params, self._views):
            p.grad = v

    @property
    def flat_grad(self) -> torch.Tensor:
        return self._flat_grad[:-1]

    def sync_gradients(self, local_weight: float | None = None) -> None:
        """One all-reduce.  local_weight=None: plain average over ranks.  Otherwise the
        gradients are combined as sum_r w_r g_r / sum_r w_r (w_r = e.g. local graph count)."""
        for p, v in zip(self._params, self._views):
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():  # someone replaced .grad (zero_grad(set_to_none=True))
                v.copy_(p.grad)
            p.grad = v
        if not self._ready():
            return
        w = 1.0 if local_weight is None else float(local_weight)
        if w != 1.0:
            self._flat_grad[:-1].mul_(w)
        self._flat_grad[-1:].fill_(w)          # device-side fill: no host copy, no sync
        dist.all_reduce(self._flat_grad, op=dist.ReduceOp.SUM, group=self.process_group)
        self._flat_grad[:-1].div_(self._flat_grad[-1])
