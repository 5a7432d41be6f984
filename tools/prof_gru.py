"""One training forward + backward of the fused GRU for ncu (python tools/prof_gru.py [N]); default N = 4 waves of 64-sequence tiles."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sldm_gnn_b200.gru import gru_last_hidden

dev = torch.device("cuda:0")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64 * 148 * 4
torch.manual_seed(0)
gru = torch.nn.GRU(6, 96, 1, batch_first=True).to(dev)
x = torch.randn(N, 16, 6, device=dev)
h = gru_last_hidden(gru, x)
h.backward(torch.randn_like(h))
torch.cuda.synchronize()
print("ok", N)
