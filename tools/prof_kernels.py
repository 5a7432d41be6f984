"""Runs the individual kernel groups a few times on the bench 'batch' shapes (for ncu / quick timing)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sldm_gnn_b200 as sg
from sldm_gnn_b200 import ops
from workloads import unit_map_graphs

which = sys.argv[1] if len(sys.argv) > 1 else "all"
G = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
F = int(sys.argv[3]) if len(sys.argv) > 3 else 128
dev = torch.device("cuda:0")
ei, _, N = unit_map_graphs(G, seed=0)
ei = ei.to(dev)
torch.manual_seed(0)
x = torch.randn(N, F, device=dev)
blk = sg.SageBlock([F, F], negative_slope=0.1).to(dev)
conv, ln = blk.convs[0], blk.posts[0][0]
p = (conv.lin_l.weight, conv.lin_l.bias, conv.lin_r.weight, ln.weight, ln.bias)
csr = sg.build_csr(ei, N)
_, out, agg, xhat, rstd = ops.layer_forward(x, csr, *p, ln.eps, 0.1, True)
dout = torch.randn_like(out)
torch.cuda.synchronize()

def timeit(name, fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort()
    print(f"{name:24s} N={N} F={F}: median {ts[len(ts)//2]:.4f} ms  min {ts[0]:.4f} ms", flush=True)

if which in ("all", "fwd"):
    timeit("project_forward", lambda: ops.project_forward(agg, x, *p, ln.eps, 0.1, True))
if which in ("all", "bwd"):
    timeit("layer_backward", lambda: ops.layer_backward(dout, x, agg, xhat, rstd, csr, p[0], p[2], p[3], p[4], 0.1, True))
if which in ("all", "seg"):
    timeit("segment_mean", lambda: sg.segment_reduce(x, csr))
if which in ("all", "csr"):
    timeit("csr_build", lambda: sg.build_csr(ei, N))
