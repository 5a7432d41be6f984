"""Forward + backward of MapSpatialAttention at the bench shape, for ncu (tools/ma_prof.py [B])."""
import sys, os
sys.path.insert(0, os.getcwd())
import torch, sldm_gnn_b200 as sg
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 821_772
S, D, K = 2048, 32, 5
g = torch.Generator().manual_seed(0)
cent = (torch.rand(S, 2, generator=g) * 2000).to(dev); pos = (torch.rand(B, 2, generator=g) * 2000).to(dev)
emb = torch.randn(S, D, generator=g).to(dev).requires_grad_(True)
att = sg.MapSpatialAttention(cent, K).to(dev)
up = torch.randn(B, D, device=dev)
for _ in range(3):
    out = att(pos, emb)
    torch.autograd.grad(out, [emb] + list(att.parameters()), up)
torch.cuda.synchronize()
