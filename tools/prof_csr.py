"""CSR build only, a few repetitions (for `ncu --metrics gpu__time_duration.sum`).  usage: prof_csr.py [batch|c4]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sldm_gnn_b200 as sg
from workloads import unit_map_graphs, skewed_graph
kind = sys.argv[1] if len(sys.argv) > 1 else "batch"
dev = torch.device("cuda:0")
if kind == "batch":
    ei, _, N = unit_map_graphs(4096, seed=0)
else:
    N = 1_000_000
    ei = skewed_graph(N, 10_000_000, seed=0)
ei = ei.to(dev)
for _ in range(3):
    csr = sg.build_csr(ei, N)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); csr = sg.build_csr(ei, N); b.record(); b.synchronize()
print(kind, "csr_build %.4f ms" % a.elapsed_time(b), csr.status())
