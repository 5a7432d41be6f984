"""Seeded synthetic workloads of SURVEY 8(d) -- shared by tests, tools and bench.py (both arms).

Lives outside the package on purpose: the reference arm of bench.py imports it without loading libsldm_sage.so.

Plain torch on the CPU (inputs are generated on the host and copied, as the
reference's DataLoader would deliver them)."""
from __future__ import annotations

import torch


def unit_map_graphs(num_graphs: int, seed: int = 0, nodes_lo: int = 150, nodes_hi: int = 250, edges_per_node: int = 5):
    """Block-diagonal batch of unit map graphs: n_g ~ U{lo..hi}, e_g = 5 n_g, src != dst,
    edges sorted by (src, dst) like the reference builders (src/gbuilder.py:88-112).
    Returns (edge_index int64 [2,E], batch int64 [N], N)."""
    g = torch.Generator().manual_seed(seed)
    sizes = torch.randint(nodes_lo, nodes_hi + 1, (num_graphs,), generator=g)
    offs = torch.zeros(num_graphs + 1, dtype=torch.int64)
    offs[1:] = torch.cumsum(sizes, 0)
    N = int(offs[-1])
    gid = torch.repeat_interleave(torch.arange(num_graphs), sizes * edges_per_node)
    n_e = sizes[gid]
    E = gid.numel()
    src = (torch.rand(E, generator=g, dtype=torch.float64) * n_e).long()
    src = torch.minimum(src, n_e - 1)
    hop = 1 + (torch.rand(E, generator=g, dtype=torch.float64) * (n_e - 1)).long()
    hop = torch.minimum(hop, n_e - 1)
    dst = (src + hop) % n_e  # uniform over the other n_g - 1 nodes: src != dst
    src = src + offs[gid]
    dst = dst + offs[gid]
    key = src * N + dst
    order = torch.sort(key, stable=True).indices
    edge_index = torch.stack([src[order], dst[order]]).contiguous()
    batch = torch.repeat_interleave(torch.arange(num_graphs), sizes)
    return edge_index, batch, N


def skewed_graph(num_nodes: int, num_edges: int, seed: int = 0):
    """Config 4: src ~ U, dst = pi(floor(N u^3)): power-law in-degree, hottest node ~1% of E."""
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, num_nodes, (num_edges,), generator=g)
    u = torch.rand(num_edges, generator=g, dtype=torch.float64)
    raw = (num_nodes * u ** 3).long().clamp_(max=num_nodes - 1)
    perm = torch.randperm(num_nodes, generator=g)
    dst = perm[raw]
    return torch.stack([src, dst]).contiguous()
