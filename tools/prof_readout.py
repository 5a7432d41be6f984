"""Forward + backward of the 'double' readout at the bench shape (4096 graphs, 0.82 M nodes, 128 features), for ncu."""
import sys, os
sys.path.insert(0, os.getcwd())
import torch, sldm_gnn_b200 as sg
from workloads import unit_map_graphs
dev = torch.device("cuda:0")
ei, bv, N = unit_map_graphs(4096, seed=0)
bv = bv.to(dev)
x = torch.randn(N, 128, device=dev, requires_grad=True)
up = torch.randn(4096, 256, device=dev)
for _ in range(3):
    out = sg.global_mean_max_pool(x, bv, 4096)
    torch.autograd.grad(out, x, up)
torch.cuda.synchronize()
