"""The model around the block (SURVEY 8d configuration C2): GruSage / MapEncoder / MapZscoreNorm as drop-ins.

Golden vectors under tests/golden/grusage come from the reference's own GruSage, MapEncoder, MapSpatialAttention and
SageBlock classes (tests/golden/make_golden_grusage.py; torch_geometric's three names stood in for by the oracle).
CPU: the oracle composition reproduces them bit for bit and our module tree matches the reference's state dict.
GPU: our model, with the reference's weights loaded strictly, reproduces logits, loss and every gradient to
rtol 1e-5 / atol 1e-6 -- adjudicated against the reference's own fp64 run where fp32 itself is no more accurate than
that.  The model is ~12 layers deep and starts with a cuDNN GRU, so the noise floor is measured, not assumed: it is the
larger of (i) the fp32 CPU reference's error against fp64 and (ii) the error of the SAME plain-torch composition
(oracle/grusage_oracle.py) executed on the GPU's fp32 library kernels (cuBLAS, ATen GRU cell / scatter; cuDNN off,
its GRU runs TF32) against fp64.  Our error
against fp64 must stay within the bar or within 8x that floor on the same tensor."""
import glob
import os

import pytest
import torch

from oracle.grusage_oracle import GruSageOracle

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "grusage", "*.pt")))
RTOL, ATOL = 1e-5, 1e-6


class Bag:
    def __init__(self, d):
        self.__dict__.update(d)

    def to(self, dev):
        return Bag({k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in self.__dict__.items()})


def _build(cls, fx, dev="cpu"):
    kw = dict(fx["kwargs"])
    if fx["map_mode"] == "tensors":
        kw["map_tensors"] = {k: v.to(dev) for k, v in fx["map"].items()}
    else:
        kw["map_embeddings"], kw["map_centroids"] = fx["extra_emb"].to(dev), fx["map"]["mseg_centroids"].to(dev)
    m = cls(**kw)
    m.load_state_dict(fx["state_dict"], strict=True)
    return m.to(dev)


def _step(model, fx, dev="cpu"):
    data = Bag(fx["data"]).to(dev)
    logits = model(data)
    loss = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor(fx["pos_weight"], device=dev))(logits, data.y)
    loss.backward()
    return logits.detach(), loss.detach(), {k: p.grad.detach().clone() for k, p in model.named_parameters()}


def test_golden_files_exist():
    assert len(GOLDEN) == 3


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_composition_reproduces_reference_golden(path):
    fx = torch.load(path, weights_only=False)
    logits, loss, grads = _step(_build(GruSageOracle, fx), fx)
    assert torch.equal(logits, fx["logits"]) and torch.equal(loss, fx["loss"])
    assert set(grads) == set(fx["grads"])
    for k, g in grads.items():
        assert torch.equal(g, fx["grads"][k]), k


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_module_tree_matches_reference_state_dict(path):
    import sldm_gnn_b200 as sg
    fx = torch.load(path, weights_only=False)
    m = _build(sg.GruSage, fx)                       # strict load inside
    assert list(m.state_dict().keys()) == list(fx["state_dict"].keys())
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v.shape) for k, v in fx["state_dict"].items()}
    assert all(not k.startswith("map_encoder") for k in m.state_dict_no_mapenc())
    assert set(m.config_dict) >= {"dynamic_features_num", "sage_hidden_dims", "global_pooling", "map_attention_topk"}


def test_zscore_norm_matches_oracle():
    import sldm_gnn_b200 as sg
    from oracle.grusage_oracle import zscore_oracle
    f = torch.randn(50, 4) * 7 + 3
    f[:, 2] = 1.5                                    # constant feature: sigma clamps at 1e-8
    assert torch.equal(sg.MapZscoreNorm.onfly(f), zscore_oracle(f))


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_cuda_model_matches_reference_golden(path):
    import sldm_gnn_b200 as sg
    dev = torch.device("cuda:0")
    fx = torch.load(path, weights_only=False)
    model = _build(sg.GruSage, fx, dev)
    logits, loss, grads = _step(model, fx, dev)
    with torch.backends.cudnn.flags(enabled=False):    # (cuDNN's GRU runs TF32 by default: 2e-4 noise would make the floor meaningless)
        lib_logits, _, lib_grads = _step(_build(GruSageOracle, fx, dev), fx, dev)  # plain torch on the GPU's fp32 library kernels

    def adjudicated(got, ref32, lib32, ref64, what):
        got, ref32, lib32 = got.cpu().double(), ref32.double(), lib32.cpu().double()
        mag = max(1.0, float(ref64.abs().max()))
        e_g = (got - ref64).abs()
        floor = max(float((ref32 - ref64).abs().max()), float((lib32 - ref64).abs().max()))
        fine = (e_g <= ATOL * mag + RTOL * ref64.abs()) | (e_g <= 8.0 * floor)
        assert bool(fine.all()), f"{what}: ours vs fp64 {float(e_g.max()):.3e}, fp32 noise floor {floor:.3e}"

    adjudicated(logits, fx["logits"], lib_logits, fx["logits64"], "logits")
    assert set(grads) == set(fx["grads"])
    for k in grads:
        adjudicated(grads[k], fx["grads"][k], lib_grads[k], fx["grads64"][k], "grad " + k)
    total, per_group = model.grads()
    assert total > 0 and set(per_group) == {"StType Embedding", "GRU Layer", "FC Layers before SAGE", "GraphSAGE Layers",
                                            "FC Layers after SAGE", "Final Output Layer"}
    ipd = model.input_params_dict()
    assert ipd["map_embeddings"].shape[0] == 60 and ipd["map_centroids"].shape == (60, 2)


@pytest.mark.gpu
def test_cuda_model_without_map_and_through_collate():
    """map_included=False (the reference cannot run this case, see oracle/grusage_oracle.py) against the oracle composition,
    with the mini-batch assembled by our collate from per-graph GraphData."""
    import sldm_gnn_b200 as sg
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    kw = dict(dynamic_features_num=6, frames_num=5, gru_hidden_size=10, gru_num_layers=1, fc1dims=[14], sage_hidden_dims=[20, 20],
              fc2dims=[9], out_dim=1, num_st_types=7, emb_dim=3, negative_slope=0.2, global_pooling="double", map_included=False)
    torch.manual_seed(3)
    orc = GruSageOracle(**kw)
    ours = sg.GruSage(**kw)
    ours.load_state_dict(orc.state_dict(), strict=True)
    ours = ours.to(dev)
    graphs = []
    for _ in range(7):
        n = int(torch.randint(5, 12, (1,), generator=g))
        src = torch.randint(0, n, (3 * n,), generator=g)
        graphs.append(dict(x=torch.randn(n, 5, 6, generator=g), edge_index=torch.stack([src, (src + 1) % n]),
                           xsttype=torch.randint(0, 7, (n,), generator=g), xdims=torch.randn(n, 2, generator=g),
                           pos_raw=torch.randn(n, 5, 2, generator=g), y=torch.zeros(1, 1)))
    batch = sg.collate([sg.GraphData(**{k: v.to(dev) for k, v in d.items()}) for d in graphs])
    from oracle.collate_oracle import collate_oracle
    cpu_batch = Bag(collate_oracle([Bag(d) for d in graphs]))
    out_o = orc(cpu_batch)
    out_g = ours(batch)
    assert torch.allclose(out_g.cpu(), out_o, rtol=1e-4, atol=1e-5), float((out_g.cpu() - out_o).abs().max())
    with torch.inference_mode():
        assert torch.equal(ours(batch), out_g)


def _c2_like(dev, dropout=None, seed=0):
    """A small full model (map encoder + attention + GRU head) and two batches of different size."""
    import sldm_gnn_b200 as sg
    from types import SimpleNamespace
    g = torch.Generator().manual_seed(100 + seed)
    S = 48
    mt = dict(float_features=torch.randn(S, 6, generator=g), bool_features=torch.rand(S, 3, generator=g) > 0.5,
              lane_type_cats=torch.randint(0, 4, (S,), generator=g),
              mgraph_edge_indexes=torch.stack([torch.randint(0, S, (4 * S,), generator=g), torch.randint(0, S, (4 * S,), generator=g)]),
              mseg_centroids=torch.rand(S, 2, generator=g) * 100.0)
    kw = dict(dynamic_features_num=6, frames_num=8, gru_hidden_size=32, gru_num_layers=1, fc1dims=[24], sage_hidden_dims=[32, 32],
              fc2dims=[16], out_dim=1, num_st_types=9, emb_dim=4, dropout=dropout, negative_slope=0.1, global_pooling="double",
              mapenc_lane_embdim=4, mapenc_sage_hdims=[16, 16], map_attention_topk=3)
    torch.manual_seed(seed)
    model = sg.GruSage(**kw, map_tensors=mt).to(dev)

    def batch(graphs, s):
        gg = torch.Generator().manual_seed(s)
        ns = [int(torch.randint(4, 11, (1,), generator=gg)) for _ in range(graphs)]
        off, eis, bv = 0, [], []
        for i, n in enumerate(ns):
            src = torch.randint(0, n, (3 * n,), generator=gg)
            eis.append(torch.stack([src, (src + 1 + torch.randint(0, n - 1, (3 * n,), generator=gg)) % n]) + off)
            bv += [i] * n
            off += n
        N = off
        d = dict(x=torch.randn(N, 8, 6, generator=gg), xdims=torch.randn(N, 2, generator=gg),
                 xsttype=torch.randint(0, 9, (N,), generator=gg), pos_raw=torch.rand(N, 8, 2, generator=gg) * 100.0,
                 edge_index=torch.cat(eis, 1), batch=torch.tensor(bv), y=(torch.rand(graphs, 1, generator=gg) > 0.5).float())
        return SimpleNamespace(**{k: v.to(dev) for k, v in d.items()}, num_graphs=graphs)

    return model, batch


@pytest.mark.gpu
def test_graphed_grusage_training_matches_the_uncaptured_model():
    """GraphedGruSage (forward + backward CUDA graphs over padded static buffers) against the same model called
    directly: logits of the real graphs and every parameter gradient, on two batches of different size through one
    bucket, the larger one first (stale rows beyond N must not matter)."""
    from sldm_gnn_b200.grusage import GraphedGruSage
    dev = torch.device("cuda:0")
    model, batch = _c2_like(dev, dropout=None)
    crit = torch.nn.BCEWithLogitsLoss()
    g = GraphedGruSage(model, max_nodes=128, max_edges=512, max_graphs=12, training=True)
    for graphs, seed in ((11, 1), (5, 2), (8, 3)):
        data = batch(graphs, seed)
        model.zero_grad(set_to_none=True)
        ref_logits = model(data)
        crit(ref_logits, data.y).backward()
        ref = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
        model.zero_grad(set_to_none=True)
        out = g(data)
        assert out.shape == ref_logits.shape
        assert torch.allclose(out, ref_logits, rtol=1e-5, atol=1e-6), float((out - ref_logits).abs().max())
        crit(out, data.y).backward()
        got = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
        assert set(got) == set(ref)
        for k in ref:
            scale = max(1.0, float(ref[k].abs().max()))
            assert torch.allclose(got[k], ref[k], rtol=1e-4, atol=2e-6 * scale), (k, float((got[k] - ref[k]).abs().max()))
    with pytest.raises(RuntimeError, match="exceed the bucket"):
        g(batch(12, 4))


@pytest.mark.gpu
def test_graphed_grusage_inference_and_dropout_training():
    from sldm_gnn_b200.grusage import GraphedGruSage
    dev = torch.device("cuda:0")
    model, batch = _c2_like(dev, dropout=0.25, seed=1)
    data = batch(6, 7)
    model.eval()
    with torch.inference_mode():
        want = model(data)
    gi = GraphedGruSage(model, max_nodes=96, max_edges=400, max_graphs=8, training=False)
    got = gi(data)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-6), float((got - want).abs().max())
    # training with dropout: finite, different draws on successive replays, a loss that an optimizer step lowers
    gt = GraphedGruSage(model, max_nodes=96, max_edges=400, max_graphs=8, training=True)
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    crit = torch.nn.BCEWithLogitsLoss()
    a, b = gt(data).detach().clone(), gt(data).detach().clone()
    assert torch.isfinite(a).all() and not torch.equal(a, b), "dropout must draw fresh masks on every replay"
    losses = []
    for _ in range(30):
        opt.zero_grad()
        loss = crit(gt(data), data.y)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert sum(losses[-5:]) < sum(losses[:5]), (losses[:5], losses[-5:])


@pytest.mark.gpu
def test_graphed_train_step_follows_the_eager_loop():
    """GraphedTrainStep (zero_grad + forward + loss + backward + Adam as ONE CUDA graph) against the reference's loop
    run eagerly on a copy of the model: same first-step gradients, the same loss curve over batches of changing size,
    and construction leaves parameters and optimizer state untouched."""
    import copy
    import torch.nn.functional as Fn
    from sldm_gnn_b200.grusage import GraphedTrainStep
    dev = torch.device("cuda:0")
    model, batch = _c2_like(dev, dropout=None, seed=2)
    twin = copy.deepcopy(model)
    pw = torch.tensor(1.5, device=dev)
    crit = torch.nn.BCEWithLogitsLoss(pos_weight=pw)
    loss_fn = lambda logits, y, w: Fn.binary_cross_entropy_with_logits(logits, y, weight=w, pos_weight=pw, reduction="sum")
    opt_e = torch.optim.Adam(twin.parameters(), lr=1e-3, weight_decay=5e-5)
    opt_g = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=5e-5, capturable=True)
    before = {k: v.clone() for k, v in model.state_dict().items()}
    step = GraphedTrainStep(model, opt_g, loss_fn, max_nodes=128, max_edges=512, max_graphs=12)
    for k, v in model.state_dict().items():
        assert torch.equal(v, before[k]), f"construction changed {k}"
    assert all(float(s["step"]) == 0 and not bool(s["exp_avg"].any()) for s in opt_g.state.values())
    losses_e, losses_g = [], []
    for i, (graphs, seed) in enumerate(((11, 1), (5, 2), (8, 3), (11, 4), (3, 5), (9, 6))):
        data = batch(graphs, seed)
        opt_e.zero_grad()
        le = crit(twin(data), data.y)
        le.backward()
        opt_g.zero_grad(set_to_none=True)        # a caller's own zero_grad must not detach the graph's gradient tensors
        lg = step(data, data.y)
        if i == 0:
            ge = dict(twin.named_parameters())
            for k, p in model.named_parameters():
                scale = max(1.0, float(ge[k].grad.abs().max()))
                assert torch.allclose(p.grad, ge[k].grad, rtol=1e-4, atol=2e-6 * scale), (k, float((p.grad - ge[k].grad).abs().max()))
        opt_e.step()
        losses_e.append(float(le)); losses_g.append(float(lg))
    assert all(abs(a - b) <= 2e-3 * max(1.0, abs(a)) for a, b in zip(losses_e, losses_g)), (losses_e, losses_g)
    with pytest.raises(ValueError, match="capturable"):
        GraphedTrainStep(model, torch.optim.Adam(model.parameters()), loss_fn, 128, 512, 12)
