"""Which stage of the C2 model breaks a CUDA-graph capture?  usage: dbg_capture_c2.py STAGE [G]   (forward + backward of one stage)"""
import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sldm_gnn_b200 as sg
import bench_c2 as bc
stage, G = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = sg.GruSage(**bc.MODEL_KW, map_tensors=bc.make_map()).to(dev)
d, N, E = bc.make_batch(G, 0)
data = bc.Bag({k: v.to(dev) for k, v in d.items()}, G)
params = [p for p in model.parameters()]

def run():
    if stage == "gru":
        y = model._last_hidden(data.x)
    elif stage == "front":
        y = bc._front(model, data)
    elif stage == "mapenc":
        y = model.map_encoder()
    elif stage == "attn":
        y = model.map_attention(data.pos_raw[:, -1, :], model.map_encoder())
    elif stage == "sage":
        model.sage.clear_cache()
        x = torch.randn(N, 128, device=dev, requires_grad=True)
        y = model.sage(x, data.edge_index)
        model.sage.clear_cache()
    elif stage == "readout":
        x = torch.randn(N, 96, device=dev, requires_grad=True)
        y = model.global_pool(x, data.batch, G)
    else:
        model.sage.clear_cache()
        y = model(data)
        model.sage.clear_cache()
    used = [p for p in params if p.requires_grad]
    g = torch.autograd.grad(y.sum(), used, allow_unused=True)
    return y, g

try:
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            run()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    sg.ops.index_checks.poll(block=True)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = run()
    graph.replay(); torch.cuda.synchronize()
    print(stage, "capture + replay ok")
except Exception as e:
    tb = traceback.format_exc().strip().splitlines()
    print(stage, "FAILED:", repr(e)[:200]); print("\n".join(tb[-14:]))
