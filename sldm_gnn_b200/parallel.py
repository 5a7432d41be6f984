"""Graph-sharded data parallelism for the batched-graph workload.

The reference is single-process (main.py:58; no torch.distributed anywhere).  A PyG
batch is block-diagonal -- independent graphs, no cross-graph edges -- so whole graphs
are the unit of sharding: every rank builds its own CSR and runs the full SageBlock on
its graphs with NO data-path collective.  Training needs exactly one exchange step: the
sum of the parameter gradients (~130-175 KB for a SageBlock), latency bound by construction.

All parameters of the wrapped module live in ONE flat fp32 buffer and so do their
gradients.  The gradient buffer is cut into BUCKETS in the order the backward pass
finishes them (the last SageBlock layer first); each bucket carries one extra slot for
the loss weight.  Two ways a bucket gets on its way while backward is still running:
  * SageBlock layers (the whole block is ONE autograd node, so the engine would only
    hand over their gradients after the last layer's backward): the block's backward
    asks the armed wrapper for the bucket views of a layer's five parameters
    (`claim()`), lets the kernels write dW_l, db_l, dW_r, dgamma, dbeta STRAIGHT into
    them, records an event after the weight-gradient reduction -- before the layer's dx
    gather is launched -- and calls `bucket_ready()`: the all-reduce of layer l runs on a
    side stream under the gather of layer l and the backward kernels of layer l-1.
    Autograd sees `None` for those parameters (their .grad IS the bucket view).
  * every other parameter: a post-accumulate hook counts its bucket down and launches the
    all-reduce behind an event recorded when the last accumulation has been enqueued.
`sync_gradients()` only waits for what is still in flight.  The direct-write path needs
gradients that were zeroed since the last backward (`zero_grad()` arms it, `expect_sync()`
confirms it); otherwise everything falls back to autograd accumulation + hooks.

LayerNorm is per node, so there are no cross-rank statistics.  A loss that is a mean
over the local graphs needs the per-rank gradient weighted by the local graph count to
equal single-process maths: gradients are combined as sum_r w_r g_r / sum_r w_r, the
weight rides in the bucket's own last slot, so it costs no extra collective.
"""
from __future__ import annotations

import re

import torch
import torch.distributed as dist
import torch.nn as nn


_ACTIVE = None      # the wrapper armed by expect_sync() (one per process: one rank, one model replica)


def active_wrapper():
    return _ACTIVE


def shard_graphs(num_graphs: int, rank: int, world_size: int) -> range:
    """Contiguous, balanced split of graph ids (whole graphs only)."""
    base, rem = divmod(num_graphs, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def default_bucket_key(name: str):
    """Bucket of a parameter: SageBlock layers (`...convs.{i}.*`, `...posts.{i}.*`) get one bucket each, keyed so that
    sorting in DESCENDING order is the order in which backward finishes them; everything else shares bucket (-1,)."""
    m = re.search(r"(^|\.)(convs|posts)\.(\d+)\.", name)
    if m:
        return (name[: m.start(2)], int(m.group(3)))
    return ("", -1)


class GraphDataParallel(nn.Module):
    def __init__(self, module: nn.Module, process_group=None, broadcast: bool = True, bucket_key=default_bucket_key,
                 overlap: bool = True):
        super().__init__()
        self.module = module
        self.process_group = process_group
        named = [(n, p) for n, p in module.named_parameters() if p.requires_grad]
        if not named:
            raise ValueError("GraphDataParallel: module has no trainable parameters")
        dev, dt = named[0][1].device, named[0][1].dtype
        if any(p.device != dev or p.dtype != dt for _, p in named):
            raise ValueError("GraphDataParallel: parameters must share one device and dtype")
        # ---- buckets, in the order backward completes them (later layers first)
        keys = {}
        for n, p in named:
            keys.setdefault(bucket_key(n), []).append(p)
        order = sorted(keys, key=lambda k: (k[0], k[1]), reverse=True)
        self._buckets = []                       # dicts: params, views, lo, hi (the weight slot is flat_grad[hi])
        total = sum(p.numel() for _, p in named)
        self._flat = torch.empty(total, dtype=dt, device=dev)
        self._flat_grad = torch.zeros(total + len(order), dtype=dt, device=dev)
        self._params, self._views, self._bucket_of = [], [], {}
        off = goff = 0
        with torch.no_grad():
            for bi, k in enumerate(order):
                b = dict(params=keys[k], views=[], lo=goff, pending=len(keys[k]))
                for p in keys[k]:
                    n = p.numel()
                    self._flat[off:off + n].copy_(p.reshape(-1))
                    p.data = self._flat[off:off + n].view_as(p)          # same names, same shapes
                    gview = self._flat_grad[goff:goff + n].view_as(p)
                    p.grad = gview
                    b["views"].append(gview)
                    self._params.append(p); self._views.append(gview); self._bucket_of[id(p)] = (bi, gview)
                    off += n; goff += n
                b["hi"] = goff
                goff += 1                                                # the bucket's weight slot
                self._buckets.append(b)
        self._whole_bucket = {frozenset(id(p) for p in b["params"]): bi for bi, b in enumerate(self._buckets)}
        # zero_grad() is ONE copy of this template: zeros, and 1.0 in every bucket's weight slot (scaled with the bucket)
        self._template = torch.zeros_like(self._flat_grad)
        for b in self._buckets:
            self._template[b["hi"]] = 1.0
        self._flat_grad.copy_(self._template)
        self._grad_index = torch.cat([torch.arange(b["lo"], b["hi"], device=dev) for b in self._buckets]) \
            if len(self._buckets) > 1 else None
        if broadcast and self._ready():
            dist.broadcast(self._flat, src=self._src_rank(), group=self.process_group)
        # ---- overlap machinery (CUDA only): hooks + one side stream
        self._overlap = bool(overlap and self._flat.is_cuda)
        self._comm = torch.cuda.Stream(device=dev) if self._overlap else None
        self._weight = None                      # set by expect_sync(): buckets may leave during backward
        self._launched = set()
        self._zeroed = True                      # gradients are zero (fresh, or zero_grad() since the last backward)
        if self._overlap and self._ready():        # (single process: no exchange, no per-parameter Python hooks)
            for p in self._params:
                p.register_post_accumulate_grad_hook(self._on_grad)

    # ------------------------------------------------------------------ plumbing --
    def _ready(self) -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.process_group) > 1

    def _src_rank(self) -> int:
        return dist.get_global_rank(self.process_group, 0) if self.process_group is not None else 0

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def zero_grad(self, set_to_none: bool = False) -> None:  # keeps the flat views attached
        self._flat_grad.copy_(self._template)
        for p, v in zip(self._params, self._views):
            p.grad = v
        for b in self._buckets:
            b["pending"] = len(b["params"])
        self._launched.clear()
        self._zeroed = True

    @property
    def flat_grad(self) -> torch.Tensor:
        """All parameter gradients as one 1-D tensor (a view when there is a single bucket)."""
        if self._grad_index is None:
            return self._flat_grad[:-1]
        return self._flat_grad.index_select(0, self._grad_index)

    @property
    def num_buckets(self) -> int:
        return len(self._buckets)

    # ------------------------------------------------------------------ exchange --
    def expect_sync(self, local_weight: float | None = None) -> None:
        """Call after zero_grad() and before backward() to let every bucket's all-reduce start as soon as backward
        has produced it (needs the weight up front).  Without it sync_gradients() launches everything itself."""
        self._weight = 1.0 if local_weight is None else float(local_weight)
        global _ACTIVE
        _ACTIVE = self if (self._overlap and self._ready() and self._zeroed) else None

    # -- direct-write path, called from SageBlock's backward (sageblock._SageBlockFn) ------------------------------
    def claim(self, params):
        """The bucket views of `params` (in that order) if they are exactly one whole bucket of this armed wrapper
        whose gradients are still zero and un-exchanged -- the kernels may then write into them directly -- else None."""
        if _ACTIVE is not self or self._weight is None or not self._zeroed:
            return None
        bi = self._whole_bucket.get(frozenset(id(p) for p in params))
        if bi is None or bi in self._launched:
            return None
        views = []
        for p in params:
            b, v = self._bucket_of[id(p)]
            if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                return None
            views.append(v)
        return bi, views

    def bucket_ready(self, bi: int, event) -> None:
        """The gradients of bucket `bi` are final on the current stream at `event`: exchange it on the side stream."""
        self._comm.wait_event(event)
        with torch.cuda.stream(self._comm):
            self._reduce_bucket(bi, self._weight)

    def _reduce_bucket(self, bi: int, w: float) -> None:
        """sum_r w_r g_r / sum_r w_r for one bucket: the weight slot holds 1.0 (zero_grad's template), so scaling the
        whole bucket by w_r -- inside the collective with NCCL's pre-multiplied sum, one mul_ otherwise -- turns it
        into w_r; after the all-reduce it holds sum_r w_r and one div_ finishes the bucket."""
        b = self._buckets[bi]
        buf = self._flat_grad[b["lo"]:b["hi"] + 1]
        op = dist.ReduceOp.SUM
        if w != 1.0:
            premul = None
            if buf.is_cuda and dist.get_backend(self.process_group) == "nccl":
                try:
                    premul = dist._make_nccl_premul_sum(w)
                except Exception:
                    premul = None
            if premul is not None:
                op = premul
            else:
                buf.mul_(w)
        dist.all_reduce(buf, op=op, group=self.process_group)
        buf[:-1].div_(buf[-1])
        self._launched.add(bi)

    def _on_grad(self, p) -> None:
        if self._weight is None or not self._ready():
            return
        bi, view = self._bucket_of[id(p)]
        b = self._buckets[bi]
        if p.grad is None or p.grad.data_ptr() != view.data_ptr():
            return                                 # .grad was replaced: sync_gradients() repairs and launches it
        b["pending"] -= 1
        if b["pending"] != 0 or bi in self._launched:
            return
        main = torch.cuda.current_stream(self._flat.device)
        ev = torch.cuda.Event()
        ev.record(main)                            # every accumulation of this bucket has been enqueued before this point
        self._comm.wait_event(ev)
        with torch.cuda.stream(self._comm):
            self._reduce_bucket(bi, self._weight)

    def sync_gradients(self, local_weight: float | None = None) -> None:
        """Finishes the exchange: launches the buckets the hooks have not (all of them without expect_sync()) and
        makes the current stream wait for the side stream.  local_weight=None: plain average over ranks.  Otherwise
        the gradients are combined as sum_r w_r g_r / sum_r w_r (w_r = e.g. local graph count)."""
        w = 1.0 if local_weight is None else float(local_weight)
        if self._weight is not None and w != self._weight:
            raise ValueError("sync_gradients: local_weight differs from the one given to expect_sync()")
        for p, v in zip(self._params, self._views):
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():  # someone replaced .grad (zero_grad(set_to_none=True))
                assert self._bucket_of[id(p)][0] not in self._launched
                v.copy_(p.grad)
            p.grad = v
        if self._ready():
            if self._overlap:
                main = torch.cuda.current_stream(self._flat.device)
                rest = [bi for bi in range(len(self._buckets)) if bi not in self._launched]
                if rest:
                    self._comm.wait_stream(main)
                    with torch.cuda.stream(self._comm):
                        for bi in rest:
                            self._reduce_bucket(bi, w)
                main.wait_stream(self._comm)
            else:
                for bi in range(len(self._buckets)):
                    if bi not in self._launched:
                        self._reduce_bucket(bi, w)
        self._weight = None
        self._launched.clear()
        self._zeroed = False
        global _ACTIVE
        if _ACTIVE is self:
            _ACTIVE = None
        for b in self._buckets:
            b["pending"] = len(b["params"])
