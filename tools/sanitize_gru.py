"""Small forward + backward of the fused GRU kernels for compute-sanitizer (memcheck / racecheck): the time loops of
k_gru_fwd / k_gru_bwd synchronise with __syncwarp only (a warp owns its rows), which racecheck verifies."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sldm_gnn_b200.gru import gru_last_hidden

dev = torch.device("cuda:0")
torch.manual_seed(0)
for N, T, I, H in ((130, 5, 6, 96), (70, 4, 8, 64), (33, 3, 2, 32)):
    gru = torch.nn.GRU(I, H, 1, batch_first=True).to(dev)
    x = torch.randn(N, T, I, device=dev, requires_grad=True)
    h = gru_last_hidden(gru, x)
    h.square().sum().backward()
torch.cuda.synchronize()
print("SANITIZE_RUN_OK")
