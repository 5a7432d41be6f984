"""CPU oracle for the SageBlock hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product path (sldm_gnn_b200/) never does and has
no CPU fallback.

PARITY UNPINNED: the reference ships no tests, fixtures or golden vectors
(SURVEY F6), and the arithmetic of its hot path lives in a dependency that is not
vendored and not installable here: torch-geometric 2.7.0 (reference uv.lock:1406-1407,
pyproject.toml:15) on torch 2.8.0.  This file therefore RESTATES the published
algorithm of that pinned version, op for op, using the same ATen CPU operators PyG
itself dispatches to, and is anchored on the reference's own call sites:

  * src/models/blocks/sageblock.py:4-20   -- SageBlock: L x (SAGEConv -> LayerNorm ->
    LeakyReLU|ReLU -> Dropout|Identity); module names `convs`, `posts`.
  * src/models/grusage.py:109,182 and src/models/map/mapencoder.py:20-24,37 -- the two
    callers: forward(x, edge_index), edge_index int64 [2, E], row 0 = source j,
    row 1 = destination i (PyG flow 'source_to_target').

Restated upstream functions (torch_geometric 2.7.0):
  * nn/conv/sage_conv.py  SAGEConv.__init__/forward with the defaults the reference
    uses (aggr='mean', normalize=False, root_weight=True, project=False, bias=True):
        out = propagate(edge_index, x=(x, x)); out = lin_l(out); out = out + lin_r(x)
    lin_l = Linear(in, out, bias=True), lin_r = Linear(in, out, bias=False).
  * nn/conv/message_passing.py  propagate/_collect: x_j = x.index_select(0, edge_index[0]);
    aggregation index = edge_index[1]; dim_size = x.size(0); _check_input raises
    ValueError unless edge_index is an int64 tensor of shape [2, E].
  * nn/aggr/basic.py MeanAggregation -> utils/_scatter.py scatter(reduce='mean'):
        count = zeros(N).scatter_add_(0, index, ones(E)); count = count.clamp(min=1)
        out = zeros(N, F).scatter_add_(0, index[:, None].expand(E, F), x_j); out / count[:, None]
  * nn/dense/linear.py Linear.forward = F.linear; reset_parameters: weight
    kaiming_uniform(a=sqrt(5), fan=in) == U(-1/sqrt(in), 1/sqrt(in)); bias
    U(-1/sqrt(in), 1/sqrt(in)).
If torch_geometric is importable at run time, tests cross-check this file against
the real SAGEConv (tests/test_oracle.py::test_against_real_pyg_if_present).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def check_edge_index(edge_index: torch.Tensor) -> None:
    """MessagePassing._check_input of PyG 2.7.0 for a plain Tensor edge_index."""
    if not isinstance(edge_index, torch.Tensor):
        raise ValueError("`edge_index` must be a torch.Tensor of dtype torch.long and shape [2, num_edges]")
    if edge_index.dtype != torch.long:
        raise ValueError(f"Expected 'edge_index' to be of integer type (got '{edge_index.dtype}')")
    if edge_index.dim() != 2:
        raise ValueError(f"Expected 'edge_index' to be two-dimensional (got {edge_index.dim()} dimensions)")
    if edge_index.size(0) != 2:
        raise ValueError(f"Expected 'edge_index' to have size '2' in the first dimension (got '{edge_index.size(0)}')")


def scatter_mean(src: torch.Tensor, index: torch.Tensor, dim_size: int) -> torch.Tensor:
    """utils/_scatter.py::scatter(src, index, dim=0, dim_size, reduce='mean')."""
    count = src.new_zeros(dim_size)
    count.scatter_add_(0, index, src.new_ones(src.size(0)))
    count = count.clamp(min=1)
    idx = index.view(-1, 1).expand_as(src)
    out = src.new_zeros((dim_size,) + tuple(src.shape[1:])).scatter_add_(0, idx, src)
    return out / count.view(-1, 1)


class PygLinear(nn.Module):
    """torch_geometric.nn.dense.linear.Linear (weight [out, in], optional bias)."""

    def __init__(self, in_channels: int, out_channels: int, bias: bool = True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self) -> None:
        bound = 1.0 / math.sqrt(self.in_channels) if self.in_channels > 0 else 0.0
        with torch.no_grad():
            self.weight.uniform_(-bound, bound)
            if self.bias is not None:
                self.bias.uniform_(-bound, bound)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return F.linear(x, self.weight, self.bias)


class SAGEConvOracle(nn.Module):
    """SAGEConv(in, out) with PyG 2.7.0 defaults, restated (see module docstring)."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin_l = PygLinear(in_channels, out_channels, bias=True)
        self.lin_r = PygLinear(in_channels, out_channels, bias=False)
        self.reset_parameters()  # PyG initialises every tensor twice

    def reset_parameters(self) -> None:
        self.lin_l.reset_parameters()
        self.lin_r.reset_parameters()

    def aggregate(self, x: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
        x_j = x.index_select(0, edge_index[0])
        return scatter_mean(x_j, edge_index[1], x.size(0))

    def forward(self, x: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
        check_edge_index(edge_index)
        out = self.aggregate(x, edge_index)
        out = self.lin_l(out)
        out = out + self.lin_r(x)
        return out


class SageBlockOracle(nn.Module):
    """src/models/blocks/sageblock.py:4-20 with SAGEConv replaced by its restatement."""

    def __init__(self, hdims: list[int], dropout: float | None = None, negative_slope: float | None = None):
        super().__init__()
        assert len(hdims) >= 1, "hdims must contain at least one element"
        self.convs = nn.ModuleList([SAGEConvOracle(hdims[i], hdims[i + 1]) for i in range(len(hdims) - 1)])
        self.posts = nn.ModuleList([
            nn.Sequential(
                nn.LayerNorm(hdims[i + 1]),
                nn.LeakyReLU(negative_slope=negative_slope) if negative_slope is not None else nn.ReLU(),
                nn.Dropout(p=dropout) if dropout is not None else nn.Identity(),
            ) for i in range(len(hdims) - 1)
        ])

    def forward(self, x, edge_index):
        for conv, post in zip(self.convs, self.posts):
            x = conv(x, edge_index)
            x = post(x)
        return x


# ---- bf16 feature storage (an addition of the product, BASELINE configs[4]; the reference has no such mode) --------
class _RoundBf16(torch.autograd.Function):
    """Round to bfloat16 (nearest even) and back; the gradient passes straight through -- the product's backward
    works in fp32 on the stored (rounded) values and never differentiates the rounding."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


class SageBlockBf16Oracle(SageBlockOracle):
    """SageBlockOracle with the three stored feature matrices of every layer rounded to bf16 where the bf16 kernels
    round them: the layer input x (given already rounded), the aggregated rows, the layer output.  All arithmetic
    in between is the oracle's fp32 (or fp64) arithmetic."""

    def forward(self, x, edge_index):
        rnd = _RoundBf16.apply
        x = rnd(x)
        for conv, post in zip(self.convs, self.posts):
            check_edge_index(edge_index)
            agg = rnd(conv.aggregate(x, edge_index))
            x = rnd(post(conv.lin_l(agg) + conv.lin_r(x)))
        return x


# ---- one layer, forward + backward, for graphs too large to materialise [E, F] -----
def layer_fwd_bwd_chunked(x, ei, state, hdims, slope, w, dtype, chunk=1_000_000):
    """One SageBlock layer fwd+bwd on a graph too large to materialise [E, F] in fp64: the aggregation (a3/a4 of SURVEY
    8a) and its transpose are applied in edge chunks with index_add_ (edge order preserved), the dense part runs through
    torch autograd.  Same arithmetic as SageBlockOracle, restated; used as the fp64 adjudicator only."""
    N = x.size(0)
    xd = x.to(dtype)
    src, dst = ei[0], ei[1]
    cnt = torch.bincount(dst, minlength=N).clamp(min=1).to(dtype)
    agg = torch.zeros(N, hdims[0], dtype=dtype)
    for s in range(0, src.numel(), chunk):
        agg.index_add_(0, dst[s:s + chunk], xd.index_select(0, src[s:s + chunk]))
    agg = (agg / cnt[:, None]).requires_grad_(True)
    xroot = xd.clone().requires_grad_(True)
    p = {k: v.to(dtype).clone().requires_grad_(True) for k, v in state.items()}
    z = F.linear(agg, p["convs.0.lin_l.weight"], p["convs.0.lin_l.bias"]) + F.linear(xroot, p["convs.0.lin_r.weight"])
    y = F.leaky_relu(F.layer_norm(z, (hdims[1],), p["posts.0.0.weight"], p["posts.0.0.bias"], 1e-5), slope)
    y.backward(w.to(dtype))
    dmsg = agg.grad / cnt[:, None]
    dx = xroot.grad.clone()
    for s in range(0, src.numel(), chunk):
        dx.index_add_(0, src[s:s + chunk], dmsg.index_select(0, dst[s:s + chunk]))
    return y.detach(), dx, {k: v.grad for k, v in p.items()}


# ---- index oracle (bit-exact contract of the CSR build) ------------------------
def csr_oracle(edge_index: torch.Tensor, num_nodes: int):
    """rowptr/col pairs the device CSR must equal bit for bit.

    by destination: rowptr_dst = [0, cumsum(bincount(dst))], col_src = src[argsort(dst, stable)]
    by source     : rowptr_src = [0, cumsum(bincount(src))], col_dst = dst[argsort(src, stable)]
    Stability == edge order inside a segment == the order in which the reference's CPU
    scatter_add_ / index_add_ visit the edges.
    """
    src, dst = edge_index[0].cpu(), edge_index[1].cpu()

    def one(keys, vals):
        order = torch.sort(keys, stable=True).indices
        counts = torch.bincount(keys, minlength=num_nodes)
        rowptr = torch.zeros(num_nodes + 1, dtype=torch.int64)
        rowptr[1:] = torch.cumsum(counts, 0)
        return rowptr.to(torch.int32), vals[order].to(torch.int32)

    rp_d, col_s = one(dst, src)
    rp_s, col_d = one(src, dst)
    return rp_d, col_s, rp_s, col_d


# ---------------------------------------------------------------- graph readout --
# Restates torch_geometric 2.7.0 nn/pool/glob.py as called at src/models/grusage.py:113-120,185:
#   global_mean_pool(x, batch, size) = scatter(x, batch, dim=-2, dim_size=size, reduce='mean')
#   global_max_pool(x, batch, size)  = scatter(x, batch, dim=-2, dim_size=size, reduce='max')
#   batch is None -> x.mean(dim=-2, keepdim=True) / x.max(dim=-2, keepdim=True)[0]
# and utils/_scatter.py::scatter: 'mean' as above (count, clamp(min=1), scatter_add_, divide);
# 'max' = src.new_zeros(size).scatter_reduce_(dim, index, src, reduce='amax', include_self=False)
# (rows that receive nothing keep the 0 they were created with); dim_size = int(index.max()) + 1 when None.
def global_mean_pool_oracle(x, batch, size=None):
    if batch is None:
        return x.mean(dim=-2, keepdim=True)
    if size is None:
        size = int(batch.max()) + 1 if batch.numel() > 0 else 0
    return scatter_mean(x, batch, size)


def global_max_pool_oracle(x, batch, size=None):
    if batch is None:
        return x.max(dim=-2, keepdim=True)[0]
    if size is None:
        size = int(batch.max()) + 1 if batch.numel() > 0 else 0
    index = batch.view(-1, 1).expand(-1, x.size(1))
    return x.new_zeros((size, x.size(1))).scatter_reduce_(0, index, x, reduce="amax", include_self=False)


def global_double_pool_oracle(x, batch, size=None):
    """src/models/grusage.py:119."""
    return torch.cat([global_mean_pool_oracle(x, batch, size), global_max_pool_oracle(x, batch, size)], dim=1)
