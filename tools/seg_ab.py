"""A/B harness for the segment gather: prints timing and a hash of the output bytes (variants must agree bit for bit).
usage: python tools/seg_ab.py [batch|c4] [F]     (knobs via env: SLDM_SEG_TILE, SLDM_SEG_TILE_ROWS, SLDM_SEG_TILE_SMEM_KB)"""
import sys, os, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sldm_gnn_b200 as sg
from workloads import unit_map_graphs, skewed_graph

kind = sys.argv[1] if len(sys.argv) > 1 else "batch"
F = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = torch.device("cuda:0")
if kind == "batch":
    ei, _, N = unit_map_graphs(4096, seed=0)
else:
    N = 1_000_000
    ei = skewed_graph(N, 10_000_000, seed=0)
ei = ei.to(dev)
x = torch.randn(N, F, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
y = torch.randn(N, F, device=dev, generator=torch.Generator(device=dev).manual_seed(2))
csr = sg.build_csr(ei, N)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

def h(t):
    return hashlib.sha1(t.cpu().numpy().tobytes()).hexdigest()[:12]

def timeit(fn, reps=7):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]

tag = "lean=%s" % (os.environ.get("SLDM_SEG_LEAN", "-"),)
o1 = sg.segment_reduce(x, csr)
o2 = sg.segment_reduce(y, csr, transpose=True, mean=False, addend=x)
m1 = timeit(lambda: sg.segment_reduce(x, csr))
m2 = timeit(lambda: sg.segment_reduce(y, csr, transpose=True, mean=False, addend=x))
print(f"{kind} F={F} [{tag}] fwd_mean {m1[0]:.4f} ms (min {m1[1]:.4f}) hash {h(o1)} | bwd_sum+addend {m2[0]:.4f} ms (min {m2[1]:.4f}) hash {h(o2)}", flush=True)
