"""Debug helper: one bf16 layer against the emulation, error pattern by row / column."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sldm_gnn_b200 as sg
from sldm_gnn_b200 import ops
from oracle.sage_oracle import SageBlockBf16Oracle
from workloads import unit_map_graphs

dev = torch.device("cuda:0")
for hd in ([128, 128], [64, 64], [128, 32]):
    torch.manual_seed(0)
    ei, _, N = unit_map_graphs(6, seed=3)
    emu = SageBlockBf16Oracle(hd, negative_slope=0.1)
    ours = sg.SageBlock(hd, negative_slope=0.1); ours.load_state_dict(emu.state_dict()); ours.to(dev)
    x = torch.randn(N, hd[0]).to(torch.bfloat16)
    want = emu(x.float(), ei).detach()
    csr = sg.build_csr(ei.to(dev), N)
    conv, ln = ours.convs[0], ours.posts[0][0]
    _, out, agg, xhat, rstd = ops.layer_forward(x.to(dev), csr, conv.lin_l.weight, conv.lin_l.bias, conv.lin_r.weight, ln.weight, ln.bias, 1e-5, 0.1, True)
    agg_want = emu.convs[0].aggregate(x.float(), ei).to(torch.bfloat16)
    print(hd, "N", N, "agg equal:", float((agg.cpu() == agg_want).float().mean()))
    err = (out.float().cpu() - want).abs()
    bad = err > 0.05
    print("  bad frac", float(bad.float().mean()), "max err", float(err.max()))
    if bad.any():
        rows = bad.any(1).nonzero().flatten()
        cols = bad.any(0).nonzero().flatten()
        print("  bad rows (count, first 20, mod 128):", rows.numel(), rows[:20].tolist(), sorted(set((rows % 128).tolist()))[:40])
        print("  bad cols:", cols.numel(), cols[:40].tolist())
        r = int(rows[0])
        print("  row", r, "got", out[r, :8].float().cpu().tolist(), "want", want[r, :8].tolist())
