"""Generates tests/golden/*.pt.  Run HERE (authoring container), never on the GPU box:

    python tests/golden/make_golden.py

It imports the reference's OWN SageBlock class from /root/reference/src/models/blocks/
sageblock.py.  That file does `from torch_geometric.nn import SAGEConv`, and
torch-geometric 2.7.0 is not installed in this image (no network), so a stub module
`torch_geometric.nn` is registered first whose SAGEConv is the restatement in
oracle/sage_oracle.py.  What the fixtures therefore pin is: the reference's real module
structure (names, LayerNorm/activation/dropout order, state-dict keys) around the
restated conv.  The conv arithmetic itself stays "parity unpinned" (see DESIGN.md).
"""
import importlib.util
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.sage_oracle import SAGEConvOracle  # noqa: E402

REF = "/root/reference/src/models/blocks/sageblock.py"


def load_reference_sageblock():
    tg = types.ModuleType("torch_geometric")
    tgnn = types.ModuleType("torch_geometric.nn")
    tgnn.SAGEConv = SAGEConvOracle
    tg.nn = tgnn
    sys.modules["torch_geometric"] = tg
    sys.modules["torch_geometric.nn"] = tgnn
    spec = importlib.util.spec_from_file_location("ref_sageblock", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.SageBlock


def case(name, hdims, N, E, slope, seed, edge_kind="random"):
    SageBlock = load_reference_sageblock()
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    blk = SageBlock(hdims, dropout=None, negative_slope=slope)
    x = torch.randn(N, hdims[0], generator=g)
    if edge_kind == "random":
        ei = torch.randint(0, N, (2, E), generator=g)
    elif edge_kind == "star":  # every node -> node 0, plus a self loop and a duplicate
        src = torch.arange(1, N)
        ei = torch.stack([torch.cat([src, torch.tensor([0, 1])]), torch.cat([torch.zeros(N - 1, dtype=torch.long), torch.tensor([0, 0])])])
    elif edge_kind == "empty":
        ei = torch.empty((2, 0), dtype=torch.long)
    xr = x.clone().requires_grad_(True)
    y = blk(xr, ei)
    w = torch.randn(y.shape, generator=g)
    (y * w).sum().backward()
    out = {
        "hdims": hdims, "slope": slope, "x": x, "edge_index": ei, "w": w,
        "state_dict": {k: v.detach().clone() for k, v in blk.state_dict().items()},
        "y": y.detach().clone(), "dx": xr.grad.clone(),
        "grads": {k: p.grad.clone() for k, p in blk.named_parameters()},
    }
    torch.save(out, os.path.join(ROOT, "tests", "golden", f"{name}.pt"))
    print(name, "N", N, "E", ei.size(1), "keys", list(out["state_dict"].keys())[:3], "...")


if __name__ == "__main__":
    case("ref_small_leaky", [8, 16, 16], 40, 160, 0.1, 1)
    case("ref_vehicle_dims_relu", [128, 96, 96], 64, 320, None, 2)
    case("ref_map_dims", [16, 32, 32], 50, 120, 0.1, 3)
    case("ref_star_dup_selfloop", [4, 8], 12, 0, 0.1, 4, "star")
    case("ref_empty_edges", [8, 8, 8], 5, 0, 0.1, 5, "empty")
    case("ref_odd_dims", [13, 7, 5], 33, 100, 0.2, 6)
