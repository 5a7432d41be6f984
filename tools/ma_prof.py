import sys, os
sys.path.insert(0, os.getcwd())
import torch, sldm_gnn_b200 as sg
dev = torch.device("cuda:0")
B, S, D, K = 200_000, 2048, 32, 5
g = torch.Generator().manual_seed(0)
cent = (torch.rand(S, 2, generator=g) * 2000).to(dev); pos = (torch.rand(B, 2, generator=g) * 2000).to(dev)
emb = torch.randn(S, D, generator=g).to(dev)
att = sg.MapSpatialAttention(cent, K).to(dev)
for _ in range(3): att(pos, emb)
torch.cuda.synchronize()
