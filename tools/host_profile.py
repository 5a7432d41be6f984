"""cProfile of the host side of a small training step (c1 shape: 32 unit map graphs, [64,64,64])."""
import sys, os, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sldm_gnn_b200 as sg
from workloads import unit_map_graphs
dev = torch.device("cuda:0")
ei, _, N = unit_map_graphs(32, seed=0)
ei = ei.to(dev)
blk = sg.SageBlock([64, 64, 64], negative_slope=0.1).to(dev)
x = torch.randn(N, 64, device=dev, requires_grad=True)
w = torch.randn(N, 64, device=dev)
def step():
    blk.clear_cache()
    blk.zero_grad(set_to_none=True)
    x.grad = None
    y = blk(x, ei)
    y.backward(w)
for _ in range(20): step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(200): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host enqueue %.1f us/step, with drain %.1f us/step" % ((t1 - t0) / 200 * 1e6, (t2 - t0) / 200 * 1e6))
pr = cProfile.Profile(); pr.enable()
for _ in range(200): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)

# ---- collate: ours vs a torch.cat restatement on the device (what PyG's collate launches) -------------------------
from workloads import unit_map_graphs as _umg
items = []
for g in range(32):
    eg, _, ng = _umg(1, seed=100 + g)
    items.append(sg.GraphData(x=torch.randn(ng, 16, 6, device=dev), edge_index=eg.to(dev), xsttype=torch.randint(0, 5, (ng,), device=dev),
                              xdims=torch.randn(ng, 2, device=dev), pos_raw=torch.randn(ng, 16, 2, device=dev), y=torch.zeros(1, 4, device=dev)))
def torch_collate():
    nodes = [int(d.x.size(0)) for d in items]
    ptr = torch.zeros(len(items) + 1, dtype=torch.long); ptr[1:] = torch.cumsum(torch.tensor(nodes), 0)
    out = {k: torch.cat([getattr(d, k) for d in items], 0) for k in ("x", "xsttype", "xdims", "pos_raw", "y")}
    out["edge_index"] = torch.cat([d.edge_index + int(ptr[g]) for g, d in enumerate(items)], -1)
    out["batch"] = torch.repeat_interleave(torch.arange(len(items), device=dev), torch.tensor(nodes, device=dev))
    return out
for name, fn in (("sg.collate", lambda: sg.collate(items)), ("torch.cat restatement", torch_collate)):
    for _ in range(10): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(100): fn()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print("collate of 32 graphs, %-22s %.1f us per batch" % (name + ":", (t1 - t0) / 100 * 1e6))

# ---- map attention: fused kernel vs the reference's torch ops on the device ---------------------------------------
import torch.nn.functional as F
B, S, D, K = 200_000, 2048, 32, 5
g = torch.Generator().manual_seed(0)
cent = (torch.rand(S, 2, generator=g) * 2000).to(dev); pos = (torch.rand(B, 2, generator=g) * 2000).to(dev)
emb = torch.randn(S, D, generator=g).to(dev)
att = sg.MapSpatialAttention(cent, K).to(dev)
def ref_ops():
    dists = torch.norm(pos.unsqueeze(1) - cent.unsqueeze(0), dim=2)
    nd, idx = torch.topk(-dists, k=K, dim=1)
    w = F.softmax(att.attn_mlp((-nd).unsqueeze(2)).squeeze(2), dim=1).unsqueeze(2)
    return torch.sum(emb[idx, :] * w, dim=1)
for name, fn in (("fused kernel", lambda: att(pos, emb)), ("reference torch ops on the GPU", ref_ops)):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): fn()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print("map attention forward B=%d S=%d, %-32s %.3f ms" % (B, S, name + ":", (t1 - t0) / 10 * 1e3))

# ---- proximity edges: device build (the comparison with the reference's Python loop lives in tests/test_edges.py) ------
V, T = 300, 16
gx = torch.Generator().manual_seed(0)
xt = torch.zeros(V, T, 6)
xt[:, :, :2] = (torch.rand(V, 1, 2, generator=gx) - 0.5) * 400 + (torch.rand(V, 1, 2, generator=gx) - 0.5) * 4 * torch.arange(T).view(1, T, 1)
xt[:, :, 4] = (torch.rand(V, T, generator=gx) < 0.9).float()
xd = xt.to(dev)
for _ in range(3): sg.build_proximity_edges(xd, 30.0)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(50): ei_g, _ = sg.build_proximity_edges(xd, 30.0)
torch.cuda.synchronize(); t1 = time.perf_counter()
print("proximity edges V=%d T=%d E=%d: device build %.3f ms (incl. the host read of E)" % (V, T, ei_g.size(1), (t1 - t0) / 50 * 1e3))
