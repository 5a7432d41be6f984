"""Golden vectors for the model around the block, produced by THE REFERENCE'S OWN CLASSES (GruSage, MapEncoder,
MapZscoreNorm, MapSpatialAttention, SageBlock -- imported unmodified from /root/reference) in the authoring container:

    python tests/golden/make_golden_grusage.py        # needs /root/reference; writes tests/golden/grusage/*.pt

torch_geometric is not installable here, so the three names the reference imports from it (SAGEConv, global_mean_pool,
global_max_pool) are bound to the CPU restatements of oracle/sage_oracle.py; everything else is reference code.
Each fixture: constructor arguments, map tensors, the state dict, one mini-batch, the logits, the BCEWithLogits loss
(pos_weight as in src/utils.py:165) and the gradient of every parameter -- in fp32 and, for adjudication, in fp64."""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.sage_oracle import SAGEConvOracle, global_max_pool_oracle, global_mean_pool_oracle   # noqa: E402

pyg = types.ModuleType("torch_geometric")
pyg.__path__ = []
pyg_nn = types.ModuleType("torch_geometric.nn")
pyg_nn.SAGEConv, pyg_nn.global_mean_pool, pyg_nn.global_max_pool = SAGEConvOracle, global_mean_pool_oracle, global_max_pool_oracle
sys.modules["torch_geometric"], sys.modules["torch_geometric.nn"] = pyg, pyg_nn
sys.path.insert(0, "/root/reference")
from src.models.grusage import GruSage   # noqa: E402  (the reference class, unmodified)

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "grusage")


def make_batch(g, G, T, F, num_st, extent):
    sizes = torch.randint(8, 16, (G,), generator=g).tolist()
    N = sum(sizes)
    ei, batch, off = [], [], 0
    for gi, n in enumerate(sizes):
        e = 4 * n
        src = torch.randint(0, n, (e,), generator=g)
        dst = (src + 1 + torch.randint(0, n - 1, (e,), generator=g)) % n
        ei.append(torch.stack([src, dst]) + off)
        batch += [gi] * n
        off += n
    x = torch.randn(N, T, F, generator=g)
    return dict(x=x, edge_index=torch.cat(ei, 1), edge_attr=torch.randn(sum(4 * n for n in sizes), 4, generator=g),
                xsttype=torch.randint(0, num_st, (N,), generator=g), xdims=torch.randn(N, 2, generator=g),
                pos_raw=torch.rand(N, T, 2, generator=g) * extent, batch=torch.tensor(batch),
                y=(torch.rand(G, 1, generator=g) > 0.5).float(), num_graphs=G)


def make_map(g, S, extent):
    e = 4 * S
    return dict(float_features=torch.randn(S, 5, generator=g) * 3 + 1, bool_features=torch.rand(S, 2, generator=g) > 0.5,
                lane_type_cats=torch.randint(0, 4, (S,), generator=g),
                mgraph_edge_indexes=torch.stack([torch.randint(0, S, (e,), generator=g), torch.randint(0, S, (e,), generator=g)]),
                mseg_centroids=torch.rand(S, 2, generator=g) * extent)


CASES = {
    "map_tensors_leaky_double": dict(kw=dict(dynamic_features_num=6, frames_num=8, gru_hidden_size=16, gru_num_layers=2, fc1dims=[24, 20],
                                             sage_hidden_dims=[32, 32], fc2dims=[16], out_dim=1, num_st_types=8, emb_dim=4,
                                             negative_slope=0.1, global_pooling="double", mapenc_sage_hdims=[8, 8], map_attention_topk=5),
                                     map="tensors", G=6, pos_weight=2.5),
    "map_embeddings_relu_mean": dict(kw=dict(dynamic_features_num=5, frames_num=6, gru_hidden_size=12, gru_num_layers=1, fc1dims=[16],
                                             sage_hidden_dims=[24], fc2dims=[10, 6], out_dim=2, num_st_types=5, emb_dim=3,
                                             negative_slope=None, global_pooling="mean", map_attention_topk=3),
                                     map="embeddings", G=4, pos_weight=1.0),
    "map_tensors_relu_max": dict(kw=dict(dynamic_features_num=6, frames_num=4, gru_hidden_size=8, gru_num_layers=1, fc1dims=[12],
                                         sage_hidden_dims=[16, 16, 16], fc2dims=[8], out_dim=1, num_st_types=6, emb_dim=2,
                                         negative_slope=None, global_pooling="max", mapenc_sage_hdims=[6], mapenc_lane_embdim=3,
                                         map_attention_topk=4),
                                 map="tensors", G=5, pos_weight=0.7),
}


class Bag:
    def __init__(self, d):
        self.__dict__.update(d)


def to64(v):
    return v.double() if isinstance(v, torch.Tensor) and v.is_floating_point() else v


for name, c in CASES.items():
    g = torch.Generator().manual_seed(len(name))
    kw, extent = dict(c["kw"]), 300.0
    mapd = make_map(g, 60, extent)
    extra = {}
    if c["map"] == "tensors":
        extra["map_tensors"] = mapd
    else:
        extra["map_embeddings"], extra["map_centroids"] = torch.randn(60, 7, generator=g), mapd["mseg_centroids"]
    data = make_batch(g, c["G"], kw["frames_num"], kw["dynamic_features_num"], kw["num_st_types"], extent)
    torch.manual_seed(11)
    model = GruSage(**kw, **extra)
    y = data["y"].view(c["G"], kw["out_dim"]) if kw["out_dim"] == 1 else (torch.rand(c["G"], kw["out_dim"], generator=g) > 0.5).float()
    data["y"] = y
    crit = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor(c["pos_weight"]))
    logits = model(Bag(data))
    loss = crit(logits, y)
    loss.backward()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    grads = {k: p.grad.clone() for k, p in model.named_parameters()}
    # fp64 run of the same reference classes (adjudicates fp32 differences between CPU ATen and cuDNN / our kernels)
    extra64 = {k: ({kk: to64(vv) for kk, vv in v.items()} if isinstance(v, dict) else to64(v)) for k, v in extra.items()}
    m64 = GruSage(**kw, **extra64).double()
    m64.load_state_dict({k: to64(v) for k, v in sd.items()})
    l64 = m64(Bag({k: to64(v) for k, v in data.items()}))
    torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor(c["pos_weight"], dtype=torch.float64))(l64, y.double()).backward()
    torch.save({"kwargs": kw, "map_mode": c["map"], "map": mapd, "extra_emb": extra.get("map_embeddings"), "data": data,
                "pos_weight": c["pos_weight"], "state_dict": sd, "logits": logits.detach(), "loss": loss.detach(), "grads": grads,
                "logits64": l64.detach(), "grads64": {k: p.grad.clone() for k, p in m64.named_parameters()}},
               os.path.join(OUT, f"ref_{name}.pt"))
    print(name, tuple(logits.shape), float(loss))
