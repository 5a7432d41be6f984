"""Small end-to-end run of every kernel family for compute-sanitizer (memcheck / racecheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sldm_gnn_b200 as sg
from workloads import unit_map_graphs, skewed_graph

dev = torch.device("cuda:0")
torch.manual_seed(0)
for hdims, graphs in (([128, 128, 128], 6), ([64, 96, 32], 3), ([16, 32, 32], 2), ([13, 7], 2)):
    ei, batch, N = unit_map_graphs(graphs, seed=1)
    blk = sg.SageBlock(hdims, negative_slope=0.1).to(dev)
    x = torch.randn(N, hdims[0], device=dev, requires_grad=True)
    y = blk(x, ei.to(dev))
    r = sg.global_mean_max_pool(y, batch.to(dev), graphs)
    r.square().sum().backward()
# hub rows + unsorted sources + several radix passes
N = 70000
ei = skewed_graph(N, 200000, seed=2).to(dev)
blk = sg.SageBlock([32, 32], negative_slope=None).to(dev)
x = torch.randn(N, 32, device=dev, requires_grad=True)
blk(x, ei).sum().backward()
torch.cuda.synchronize()
print("SANITIZE_RUN_OK")
