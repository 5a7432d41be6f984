"""Drop-in SageBlock (reference: src/models/blocks/sageblock.py:4-20).

Same constructor, same forward(x, edge_index) (a third positional `batch` is accepted
and ignored), same module tree and therefore the same state-dict keys:

    convs.{i}.lin_l.weight [Fout,Fin]   convs.{i}.lin_l.bias [Fout]
    convs.{i}.lin_r.weight [Fout,Fin]   (no lin_r.bias)
    posts.{i}.0.weight [Fout]           posts.{i}.0.bias [Fout]

so snapshots written by the reference's src/utils.py:22-30 load strictly
(test.py:121-122, rcv.py:62-63).  The arithmetic is libsldm_sage.so: one CSR build
per edge_index, then per layer a deterministic segment-mean gather and a fused
projection + LayerNorm + activation kernel, with a hand-written backward.  Dropout
stays torch's own (posts[i][2]) so RNG consumption is identical to the reference.
CUDA only: CPU tensors raise.
"""
from __future__ import annotations

import math

import torch
from torch.autograd.function import once_differentiable
import torch.nn as nn

from . import ops


class _Linear(nn.Module):
    """Parameter holder with the names / shapes / init of torch_geometric.nn.dense.Linear."""

    def __init__(self, in_channels: int, out_channels: int, bias: bool):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self) -> None:
        # PyG: kaiming_uniform(a=sqrt(5), fan=in) == U(+-1/sqrt(in)); bias U(+-1/sqrt(in))
        bound = 1.0 / math.sqrt(self.in_channels) if self.in_channels > 0 else 0.0
        with torch.no_grad():
            self.weight.uniform_(-bound, bound)
            if self.bias is not None:
                self.bias.uniform_(-bound, bound)

    def extra_repr(self) -> str:
        return f"{self.in_channels}, {self.out_channels}, bias={self.bias is not None}"


class SageConvParams(nn.Module):
    """Stands where PyG's SAGEConv(in, out) stands in the module tree (lin_l, lin_r)."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin_l = _Linear(in_channels, out_channels, bias=True)
        self.lin_r = _Linear(in_channels, out_channels, bias=False)

    def reset_parameters(self) -> None:
        self.lin_l.reset_parameters()
        self.lin_r.reset_parameters()

    def extra_repr(self) -> str:
        return f"{self.in_channels}, {self.out_channels}, aggr=mean"


class _SageBlockFn(torch.autograd.Function):
    """The whole block as ONE autograd node: L x (CSR segment mean -> fused projection + LayerNorm + activation ->
    dropout).  Gradients between the layers never pass through the autograd engine (no per-layer node, no dtype
    casts in the bf16 feature mode), the host issues two C calls per layer and direction.

    Dropout (posts[i][2] of the reference, src/models/blocks/sageblock.py:13) is torch's own fused kernel,
    torch.native_dropout -- the one nn.Dropout dispatches to on CUDA -- so the global Philox stream is consumed
    exactly as by the reference; its mask is kept for the backward.
    """

    @staticmethod
    def forward(ctx, x, csr, eps, slopes, drops, *params):
        L = len(params) // 5
        save = any(ctx.needs_input_grad)
        h, keep = x, []
        for l in range(L):
            W_l, b_l, W_r, ln_w, ln_b = params[5 * l:5 * l + 5]
            h, out, agg, xhat, rstd = ops.layer_forward(h, csr, W_l, b_l, W_r, ln_w, ln_b, eps[l], slopes[l], save)
            mask = None
            if drops[l] is not None:                       # training mode and 0 < p < 1
                out, mask = torch.native_dropout(out, drops[l], True)
            if save:
                keep += [h, agg, xhat, rstd, mask]
            h = out
        if save:
            tensors = [t for t in keep if t is not None]
            ctx.layout = [t is not None for t in keep]
            ctx.save_for_backward(*tensors, *params)
            ctx.csr, ctx.slopes, ctx.drops, ctx.L = csr, slopes, drops, L
        return h

    @staticmethod
    @once_differentiable   # hand-written first-order gradients: a double backward raises instead of returning garbage
    def backward(ctx, dout):
        L = ctx.L
        saved = list(ctx.saved_tensors)
        params = saved[len(saved) - 5 * L:]
        it = iter(saved[:len(saved) - 5 * L])
        keep = [next(it) if present else None for present in ctx.layout]
        # a data-parallel wrapper armed by expect_sync() hands out its bucket views: the kernels write the parameter
        # gradients straight into them and the layer's exchange starts before its dx gather (parallel.py)
        sink = None
        if torch.distributed.is_available() and torch.distributed.is_initialized() and not torch.cuda.is_current_stream_capturing():
            from .parallel import active_wrapper
            sink = active_wrapper()
        grads = [None] * (5 * L)
        g = dout if dout.dtype == torch.float32 else dout.float()    # bf16 features: the gradient path stays fp32
        for l in range(L - 1, -1, -1):
            h, agg, xhat, rstd, mask = keep[5 * l:5 * l + 5]
            W_l, _, W_r, ln_w, ln_b = params[5 * l:5 * l + 5]
            if mask is not None:
                g = torch.ops.aten.native_dropout_backward(g, mask, 1.0 / (1.0 - ctx.drops[l]))
            g = g.contiguous()
            need_dx = l > 0 or ctx.needs_input_grad[0]
            claim = sink.claim(params[5 * l:5 * l + 5]) if sink is not None else None
            if claim is not None:
                bi, views = claim
                dx = ops.layer_backward(g, h, agg, xhat, rstd, ctx.csr, W_l, W_r, ln_w, ln_b, ctx.slopes[l], need_dx,
                                        grad_out=views, on_param_grads=lambda ev, bi=bi: sink.bucket_ready(bi, ev))[0]
                # grads stay None: the parameters' .grad ARE the bucket views the kernels just wrote
            else:
                dx, dW_l, db_l, dW_r, dln_w, dln_b = ops.layer_backward(
                    g, h, agg, xhat, rstd, ctx.csr, W_l, W_r, ln_w, ln_b, ctx.slopes[l], need_dx)
                grads[5 * l:5 * l + 5] = [dW_l, db_l, dW_r, dln_w, dln_b]
            g = dx
        return (g, None, None, None, None, *grads)


class SageBlock(nn.Module):
    def __init__(self, hdims: list[int], dropout: float | None = None, negative_slope: float | None = None):
        super().__init__()
        assert len(hdims) >= 1, "hdims must contain at least one element"
        self.convs = nn.ModuleList([SageConvParams(hdims[i], hdims[i + 1]) for i in range(len(hdims) - 1)])
        self.posts = nn.ModuleList([
            nn.Sequential(
                nn.LayerNorm(hdims[i + 1]),
                nn.LeakyReLU(negative_slope=negative_slope) if negative_slope is not None else nn.ReLU(),
                nn.Dropout(p=dropout) if dropout is not None else nn.Identity(),
            ) for i in range(len(hdims) - 1)
        ])
        self._csr_key = None   # (edge_index tensor kept alive, its version, num_nodes)
        self._csr = None

    # -- CSR cache: one entry, valid while the same tensor object is passed unmodified --------
    def _get_csr(self, edge_index: torch.Tensor, num_nodes: int) -> ops.Csr:
        try:
            version = edge_index._version
        except RuntimeError:  # inference tensors do not track versions: never cached
            version = None
        key = self._csr_key
        if (version is not None and key is not None and key[0] is edge_index and key[1] == version
                and key[2] == num_nodes):
            return self._csr
        csr = ops.build_csr(edge_index, num_nodes)
        if version is not None:
            self._csr_key, self._csr = (edge_index, version, num_nodes), csr
        return csr

    def clear_cache(self) -> None:
        self._csr_key = self._csr = None

    def graphed(self, max_nodes: int, max_edges: int, training: bool = False) -> "GraphedSageBlock":
        """The block through captured CUDA graphs for inputs of up to max_nodes - 1 nodes / max_edges edges
        (inference: one graph; training=True: a forward and a backward graph behind autograd)."""
        return GraphedSageBlock(self, max_nodes, max_edges, training=training)

    def forward(self, x, edge_index, batch=None):
        if len(self.convs) == 0:
            return x
        ops.check_edge_index(edge_index)
        if not isinstance(x, torch.Tensor) or x.dim() != 2:
            raise RuntimeError("SageBlock: x must be a 2-D [num_nodes, features] tensor")
        ops._require_cuda(x, "x")
        ops._require_cuda(edge_index, "edge_index")
        if edge_index.device != x.device:
            raise RuntimeError(f"SageBlock: x is on {x.device} but edge_index is on {edge_index.device}")
        if x.dtype not in (torch.float32, torch.bfloat16):
            raise RuntimeError(f"SageBlock: expected float32 (or bfloat16: bf16 feature storage) features, got {x.dtype}")
        if x.size(1) != self.convs[0].in_channels:
            raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied ({x.size(0)}x{x.size(1)} and "
                               f"{self.convs[0].in_channels}x{self.convs[0].out_channels})")
        x = x.contiguous()
        ops.index_checks.poll()            # a deferred out-of-range report of an earlier edge_index raises here
        csr = self._get_csr(edge_index, x.size(0))
        eps, slopes, drops, params = [], [], [], []
        for conv, post in zip(self.convs, self.posts):
            ln, act, drop = post[0], post[1], post[2]
            slope = float(act.negative_slope) if isinstance(act, nn.LeakyReLU) else 0.0
            for name, t in (("lin_l.weight", conv.lin_l.weight), ("lin_l.bias", conv.lin_l.bias),
                            ("lin_r.weight", conv.lin_r.weight), ("LayerNorm.weight", ln.weight), ("LayerNorm.bias", ln.bias)):
                # the kernels read raw fp32 pointers: model.double() / .half() must raise like torch's dtype check would
                if t.device != x.device:
                    raise RuntimeError(f"SageBlock: parameter {name} is on {t.device} but x is on {x.device}")
                if t.dtype != torch.float32:
                    raise RuntimeError(f"SageBlock: parameter {name} has dtype {t.dtype}; the kernels are float32 "
                                       "(mat1 and mat2 must have the same dtype)")
                if not t.is_contiguous():
                    raise RuntimeError(f"SageBlock: parameter {name} is not contiguous")
            p = float(drop.p) if isinstance(drop, nn.Dropout) else 0.0
            if p >= 1.0 and self.training:
                raise NotImplementedError("SageBlock: dropout p = 1 is not supported by the fused block")
            eps.append(float(ln.eps)); slopes.append(slope)
            drops.append(p if (self.training and p > 0.0) else None)
            params += [conv.lin_l.weight, conv.lin_l.bias, conv.lin_r.weight, ln.weight, ln.bias]
        return _SageBlockFn.apply(x, csr, eps, slopes, drops, *params)


class _UncachedBlock(nn.Module):
    """The block with its CSR cache switched off: inside a CUDA graph the CSR build must be part of every replay."""

    def __init__(self, block: "SageBlock"):
        super().__init__()
        self.block = block

    def forward(self, x, edge_index):
        self.block.clear_cache()
        y = self.block(x, edge_index)
        self.block.clear_cache()
        return y


class GraphedSageBlock:
    """The block through captured CUDA graphs, one per size bucket -- the reference's real operating point is
    launch-bound: training batches of 32 graphs (main.py:24), eval batches of 64 (test.py:58) and, online, one
    un-batched graph of tens to ~300 vehicles per call from a worker thread (rcv.py:77-84, :107), where the ~20
    (forward) / ~43 (forward + backward) kernel launches and their Python glue cost several times the kernels.

        g = blk.graphed(max_nodes=512, max_edges=4096)                 # inference: CSR build + all layers, one graph
        y = g(x, edge_index)                                           # copy-in, one graph launch, copy-out
        g = blk.graphed(max_nodes=8192, max_edges=40960, training=True)
        loss(g(x, edge_index)).backward()                              # forward graph + backward graph, autograd-aware

    The captured step works on static buffers of the bucket's size: x is copied into the first N rows, edge_index into
    the first E columns, the remaining columns point at the last padding node (p, p) -- edges among padding rows never
    touch a real node, and every kernel is row-wise independent, so rows [0, N) of the result (and of dx) are
    bit-identical to the un-captured module's; the parameter gradients see only zero upstream gradient from the
    padding rows.  Requires N < max_nodes (one padding node) and E <= max_edges.  Inference: eval-mode dropout, the
    result is a copy.  Training (torch.cuda.make_graphed_callables): the result is a view of the graph's static
    output, valid until the next call.  Thread-safe: calls serialise on a lock.
    """

    def __init__(self, block: "SageBlock", max_nodes: int, max_edges: int, training: bool = False, device=None):
        import threading
        if len(block.convs) == 0:
            raise ValueError("GraphedSageBlock: the block has no layers")
        p0 = block.convs[0].lin_l.weight
        dev = torch.device(device) if device is not None else p0.device
        if dev.type != "cuda":
            raise RuntimeError("GraphedSageBlock: CUDA only")
        self.block, self.dev, self.training = block, dev, bool(training)
        self.max_nodes, self.max_edges = int(max_nodes), int(max_edges)
        self.fin, self.fout = block.convs[0].in_channels, block.convs[-1].out_channels
        self._lock = threading.Lock()
        with torch.cuda.device(dev):
            self._x = torch.zeros((self.max_nodes, self.fin), dtype=torch.float32, device=dev)
            self._ei = torch.full((2, self.max_edges), self.max_nodes - 1, dtype=torch.long, device=dev)
            ops.index_checks.poll(block=True)                  # nothing may be pending when a capture starts
            if self.training:
                self._x.requires_grad_(True)
                self._call = torch.cuda.make_graphed_callables(_UncachedBlock(block), (self._x, self._ei))
                return
            with torch.inference_mode():
                was_training = block.training
                block.eval()
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):                  # warm-up outside the capture: opt-in attributes, tensor maps
                    for _ in range(2):
                        block.clear_cache()
                        block(self._x, self._ei)
                torch.cuda.current_stream(dev).wait_stream(side)
                torch.cuda.synchronize(dev)
                ops.index_checks.poll(block=True)
                self._graph = torch.cuda.CUDAGraph()
                block.clear_cache()
                with torch.cuda.graph(self._graph):
                    self._y = block(self._x, self._ei)
                block.clear_cache()
                block.train(was_training)

    def _fill(self, x, edge_index, N, E):
        with torch.no_grad():
            if x is not None:
                self._x[:N].copy_(x, non_blocking=True)
            self._ei[:, :E].copy_(edge_index, non_blocking=True)
            if E < self.max_edges:
                self._ei[:, E:].fill_(self.max_nodes - 1)
        # rows beyond N keep whatever an earlier, larger call left there: they are padding rows (no edge from a real
        # node reaches them), so their values never matter

    def __call__(self, x: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
        ops.check_edge_index(edge_index)
        N, E = int(x.size(0)), int(edge_index.size(1))
        if x.dim() != 2 or x.size(1) != self.fin or x.dtype != torch.float32:
            raise RuntimeError(f"GraphedSageBlock: expected float32 x of shape [N, {self.fin}]")
        if N >= self.max_nodes or E > self.max_edges:
            raise RuntimeError(f"GraphedSageBlock: N = {N}, E = {E} exceed the bucket ({self.max_nodes - 1} nodes, "
                               f"{self.max_edges} edges)")
        with self._lock, torch.cuda.device(self.dev):
            if self.training:
                self._fill(None, edge_index, N, E)
                # zero-padded copy of x through autograd (dx flows back to the caller's tensor); the graphed callable
                # copies it into its static input and replays the forward graph, its backward replays the other one
                xs = torch.nn.functional.pad(x, (0, 0, 0, self.max_nodes - N))
                return self._call(xs, self._ei)[:N]
            with torch.inference_mode():
                self._fill(x, edge_index, N, E)
                self._graph.replay()
                return self._y[:N].clone()
