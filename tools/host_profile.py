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
pstats.Stats(pr).sort_stats("tottime").print_stats(30)

