"""A/B timing of the tensor-path kernels (projection forward, dgrad, wgrad, LN backward) on the bench shapes.
usage: [SLDM_LIB_PATH=build/ab/NAME.so] python tools/ab_tc.py [batch|c4] [tag]
Prints one JSON line: median ms per kernel (256 MB L2 flush between repetitions) and a hash of every output, so that
variants can be compared for speed AND bit equality."""
import hashlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sldm_gnn_b200 as sg
from sldm_gnn_b200 import ops, _lib
from workloads import unit_map_graphs, skewed_graph

kind = sys.argv[1] if len(sys.argv) > 1 else "batch"
tag = sys.argv[2] if len(sys.argv) > 2 else os.environ.get("SLDM_LIB_PATH", "default")
F = 128
dev = torch.device("cuda:0")
if kind == "batch":
    ei, _, N = unit_map_graphs(4096, seed=0)
else:
    N = 1_000_000
    ei = skewed_graph(N, 10_000_000, seed=0)
ei = ei.to(dev)
E = ei.size(1)
torch.manual_seed(0)
x = torch.randn(N, F, device=dev)
blk = sg.SageBlock([F, F], negative_slope=0.1).to(dev)
conv, ln = blk.convs[0], blk.posts[0][0]
p = (conv.lin_l.weight, conv.lin_l.bias, conv.lin_r.weight, ln.weight, ln.bias)
csr = sg.build_csr(ei, N)
_, out, agg, xhat, rstd = ops.layer_forward(x, csr, *p, ln.eps, 0.1, True)
dout = torch.randn_like(out)
bb = ops.backward_buffers(N, F, F, E, dev, True)
bargs = (dout, x, agg, xhat, rstd, csr, p[0], p[2], p[3], p[4], 0.1, True)
ops.layer_backward(*bargs, bufs=bb)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()


def h(*ts):
    m = hashlib.sha1()
    for t in ts:
        m.update(t.detach().cpu().numpy().tobytes())
    return m.hexdigest()[:10]


def timeit(fn, reps=9):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort()
    return round(ts[len(ts) // 2], 4)


res = {"tag": tag, "kind": kind, "N": N, "E": E}
res["proj_fwd_train"] = timeit(lambda: ops.project_forward(agg, x, *p, ln.eps, 0.1, True))
res["proj_fwd_infer"] = timeit(lambda: ops.project_forward(agg, x, *p, ln.eps, 0.1, False))
res["ln_bwd"] = timeit(lambda: ops.layer_backward(*bargs, stages=_lib.BWD_STAGE_LN, bufs=bb))
res["dgrad"] = timeit(lambda: ops.layer_backward(*bargs, stages=_lib.BWD_STAGE_DGRAD, bufs=bb))
res["wgrad"] = timeit(lambda: ops.layer_backward(*bargs, stages=_lib.BWD_STAGE_WGRAD, bufs=bb))
res["gather_bwd"] = timeit(lambda: ops.layer_backward(*bargs, stages=_lib.BWD_STAGE_GATHER, bufs=bb))
res["layer_bwd"] = timeit(lambda: ops.layer_backward(*bargs, bufs=bb))
res["seg_fwd"] = timeit(lambda: sg.segment_reduce(x, csr))
res["csr"] = timeit(lambda: sg.build_csr(ei, N))
o2 = ops.project_forward(agg, x, *p, ln.eps, 0.1, True)
res["hash_fwd"] = h(*o2)
res["hash_bwd"] = h(bb["dx"], bb["dW_l"], bb["dW_r"], bb["db_l"], bb["dln_w"], bb["dln_b"], bb["dagg"], bb["dxroot"])
print(json.dumps(res), flush=True)
