"""CPU tests of the oracle itself: hand-computed known answers, the two restatements
(torch ops / plain C) against each other, the C backward against torch autograd in
fp64, and the committed golden fixtures.  No GPU, no product code."""
import ctypes as C
import glob
import os

import numpy as np
import pytest
import torch

from oracle.sage_oracle import SAGEConvOracle, SageBlockOracle, csr_oracle, scatter_mean, check_edge_index, layer_fwd_bwd_chunked

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.pt")))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def c_forward(lib, x, ei, conv, ln, slope, real="f64"):
    N, Fin = x.shape
    Fout = conv.lin_l.weight.shape[0]
    dt = np.float64 if real == "f64" else np.float32
    xn = np.ascontiguousarray(x.numpy(), dtype=np.float32)
    ein = np.ascontiguousarray(ei.numpy(), dtype=np.int64)
    W_l, b_l, W_r = (np.ascontiguousarray(t.detach().numpy(), dtype=np.float32)
                     for t in (conv.lin_l.weight, conv.lin_l.bias, conv.lin_r.weight))
    g, b = (np.ascontiguousarray(t.detach().numpy(), dtype=np.float32) for t in (ln.weight, ln.bias))
    out = np.zeros((N, Fout), dt); agg = np.zeros((N, Fin), dt)
    xhat = np.zeros((N, Fout), dt); rstd = np.zeros((N,), dt)
    fn = getattr(lib, f"oracle_layer_forward_{real}")
    fn.restype = C.c_int
    rc = fn(_p(xn), C.c_int64(N), C.c_int(Fin), C.c_int(Fout), _p(ein), C.c_int64(ei.shape[1]),
            _p(W_l), _p(b_l), _p(W_r), _p(g), _p(b), C.c_double(ln.eps), C.c_double(slope),
            _p(out), _p(agg), _p(xhat), _p(rstd))
    assert rc == 0
    return out, agg, xhat, rstd


def c_backward(lib, dout, x, agg, xhat, rstd, ei, conv, ln, slope, real="f64"):
    N, Fin = x.shape
    Fout = conv.lin_l.weight.shape[0]
    dt = np.float64 if real == "f64" else np.float32
    xn = np.ascontiguousarray(x.numpy(), dtype=np.float32)
    ein = np.ascontiguousarray(ei.numpy(), dtype=np.int64)
    W_l, W_r = (np.ascontiguousarray(t.detach().numpy(), dtype=np.float32) for t in (conv.lin_l.weight, conv.lin_r.weight))
    g, b = (np.ascontiguousarray(t.detach().numpy(), dtype=np.float32) for t in (ln.weight, ln.bias))
    dx = np.zeros((N, Fin), dt); dW_l = np.zeros((Fout, Fin), dt); dW_r = np.zeros((Fout, Fin), dt)
    db_l = np.zeros(Fout, dt); dg = np.zeros(Fout, dt); db = np.zeros(Fout, dt)
    fn = getattr(lib, f"oracle_layer_backward_{real}")
    fn.restype = C.c_int
    rc = fn(_p(np.ascontiguousarray(dout, dt)), _p(xn), _p(agg), _p(xhat), _p(rstd), C.c_int64(N), C.c_int(Fin), C.c_int(Fout),
            _p(ein), C.c_int64(ei.shape[1]), _p(W_l), _p(W_r), _p(g), _p(b), C.c_double(slope),
            _p(dx), _p(dW_l), _p(db_l), _p(dW_r), _p(dg), _p(db))
    assert rc == 0
    return dx, dW_l, db_l, dW_r, dg, db


# ------------------------------------------------------------------ known answers --
def test_kat_aggregation_path_star_isolated_dup_selfloop():
    # 5 nodes, features = one-hot-ish so sums are readable
    x = torch.tensor([[1., 0.], [0., 2.], [4., 4.], [8., 0.], [0., 16.]])
    #        path 0->1->2, star {0,1,3}->2 (1->2 duplicated), self loop 3->3, node 4 isolated
    ei = torch.tensor([[0, 1, 0, 1, 3, 3], [1, 2, 2, 2, 2, 3]])
    conv = SAGEConvOracle(2, 2)
    agg = conv.aggregate(x, ei)
    want = torch.tensor([
        [0., 0.],                      # node 0: no in-edges -> 0 (count clamped to 1)
        [1., 0.],                      # node 1: from 0
        [(0 + 1 + 0 + 8) / 4, (2 + 0 + 2 + 0) / 4],  # node 2: 1,0,1(dup),3 -> mean counts multiplicity
        [8., 0.],                      # node 3: self loop is an ordinary edge
        [0., 0.],                      # node 4: isolated
    ])
    assert torch.equal(agg, want)
    # isolated node output == b_l + W_r x  (lin_l(0) = bias)
    out = conv(x, ei)
    want4 = conv.lin_l.bias + conv.lin_r.weight @ x[4]
    assert torch.allclose(out[4], want4, rtol=0, atol=1e-6)


def test_kat_empty_edges_and_single_node():
    conv = SAGEConvOracle(3, 4)
    x = torch.randn(1, 3)
    ei = torch.empty((2, 0), dtype=torch.long)
    out = conv(x, ei)
    assert torch.allclose(out[0], conv.lin_l.bias + conv.lin_r.weight @ x[0], atol=1e-6)
    blk = SageBlockOracle([3])          # len(hdims) == 1 -> identity
    assert blk(x, ei) is x


def test_kat_layernorm_activation_order():
    """LayerNorm sits between the conv and the activation."""
    torch.manual_seed(0)
    blk = SageBlockOracle([4, 6], negative_slope=0.1)
    x = torch.randn(7, 4)
    ei = torch.randint(0, 7, (2, 20))
    z = blk.convs[0](x, ei)
    ln = blk.posts[0][0]
    want = torch.nn.functional.leaky_relu(torch.nn.functional.layer_norm(z, (6,), ln.weight, ln.bias, 1e-5), 0.1)
    assert torch.equal(blk(x, ei), want)
    relu = SageBlockOracle([4, 6], negative_slope=None)
    assert isinstance(relu.posts[0][1], torch.nn.ReLU)
    assert isinstance(SageBlockOracle([4, 6], dropout=0.25).posts[0][2], torch.nn.Dropout)
    assert isinstance(SageBlockOracle([4, 6]).posts[0][2], torch.nn.Identity)


def test_edge_index_checks_raise_value_error():
    for bad in (torch.zeros((2, 3), dtype=torch.int32), torch.zeros((3,), dtype=torch.long), torch.zeros((3, 3), dtype=torch.long)):
        with pytest.raises(ValueError):
            check_edge_index(bad)


def test_init_is_uniform_inv_sqrt_fan_in():
    conv = SAGEConvOracle(64, 32)
    bound = 1 / 8
    for t in (conv.lin_l.weight, conv.lin_l.bias, conv.lin_r.weight):
        assert t.abs().max() <= bound and t.abs().max() > 0.8 * bound
    assert conv.lin_r.bias is None


# ------------------------------------------- the two restatements agree (fp64 judge) --
@pytest.mark.parametrize("N,E,Fin,Fout,slope", [(50, 200, 8, 16, 0.1), (33, 100, 13, 7, 0.0), (64, 0, 4, 4, 0.2), (1, 3, 5, 3, 0.1)])
def test_c_oracle_matches_torch_oracle(coracle, N, E, Fin, Fout, slope):
    torch.manual_seed(N + E)
    blk = SageBlockOracle([Fin, Fout], negative_slope=slope if slope else None)
    with torch.no_grad():
        blk.posts[0][0].weight.uniform_(0.5, 1.5)
        blk.posts[0][0].bias.uniform_(-0.5, 0.5)
    x = torch.randn(N, Fin)
    ei = torch.randint(0, N, (2, E))
    xr = x.clone().requires_grad_(True)
    y = blk(xr, ei)
    w = torch.randn_like(y)
    (y * w).sum().backward()
    conv, ln = blk.convs[0], blk.posts[0][0]
    out64, agg64, xhat64, rstd64 = c_forward(coracle, x, ei, conv, ln, slope, "f64")
    assert np.allclose(out64, y.detach().numpy(), rtol=1e-5, atol=1e-6)
    out32, agg32, _, _ = c_forward(coracle, x, ei, conv, ln, slope, "f32")
    # aggregation: same sequential fp32 order as torch's CPU scatter_add_ -> bit equal
    assert np.array_equal(agg32, conv.aggregate(x, ei).numpy())
    assert np.allclose(out32, y.detach().numpy(), rtol=1e-5, atol=1e-6)
    grads = c_backward(coracle, w.numpy().astype(np.float64), x, agg64, xhat64, rstd64, ei, conv, ln, slope, "f64")
    want = (xr.grad, conv.lin_l.weight.grad, conv.lin_l.bias.grad, conv.lin_r.weight.grad, ln.weight.grad, ln.bias.grad)
    for got, ref in zip(grads, want):
        scale = max(1.0, float(ref.abs().max()))
        assert np.allclose(got, ref.numpy(), rtol=1e-5, atol=2e-6 * scale)


def test_c_backward_is_the_gradient_fp64(coracle):
    """Finite-difference check of the restated backward formulas (SURVEY 8c iii)."""
    torch.manual_seed(7)
    N, E, Fin, Fout, slope = 9, 30, 3, 4, 0.1
    blk = SageBlockOracle([Fin, Fout], negative_slope=slope).double()
    x = torch.randn(N, Fin).double()
    ei = torch.randint(0, N, (2, E))
    params = [blk.convs[0].lin_l.weight, blk.convs[0].lin_l.bias, blk.convs[0].lin_r.weight,
              blk.posts[0][0].weight, blk.posts[0][0].bias]
    xr = x.clone().requires_grad_(True)

    def f(xx, *ps):
        c = blk.convs[0]
        agg = scatter_mean(xx.index_select(0, ei[0]), ei[1], N)
        z = torch.nn.functional.linear(agg, ps[0], ps[1]) + torch.nn.functional.linear(xx, ps[2])
        return torch.nn.functional.leaky_relu(torch.nn.functional.layer_norm(z, (Fout,), ps[3], ps[4], 1e-5), slope)

    assert torch.autograd.gradcheck(f, (xr, *params), eps=1e-6, atol=1e-5)


# --------------------------------------------------------------------- index oracle --
@pytest.mark.parametrize("N,E", [(1, 0), (1, 4), (10, 37), (300, 5000), (70000, 1000)])
def test_csr_oracles_agree(coracle, N, E):
    g = torch.Generator().manual_seed(N * 7 + E)
    ei = torch.randint(0, N, (2, E), generator=g)
    want = csr_oracle(ei, N)
    ein = np.ascontiguousarray(ei.numpy())
    rp_d = np.zeros(N + 1, np.int32); rp_s = np.zeros(N + 1, np.int32)
    col_s = np.zeros(max(E, 1), np.int32); col_d = np.zeros(max(E, 1), np.int32)
    coracle.oracle_csr_build.restype = C.c_int
    assert coracle.oracle_csr_build(_p(ein), C.c_int64(E), C.c_int64(N), _p(rp_d), _p(col_s), _p(rp_s), _p(col_d)) == 0
    assert np.array_equal(rp_d, want[0].numpy()) and np.array_equal(col_s[:E], want[1].numpy())
    assert np.array_equal(rp_s, want[2].numpy()) and np.array_equal(col_d[:E], want[3].numpy())
    # sortedness + segment membership property
    dst_sorted = ei[1][torch.sort(ei[1], stable=True).indices]
    assert torch.equal(torch.repeat_interleave(torch.arange(N), (want[0][1:] - want[0][:-1]).long()), dst_sorted)


# ------------------------------------------------------------------ golden fixtures --
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_reproduces_golden(path):
    """Fixtures were produced by the reference's own SageBlock class (tests/golden/make_golden.py)."""
    g = torch.load(path)
    blk = SageBlockOracle(g["hdims"], dropout=None, negative_slope=g["slope"])
    assert list(blk.state_dict().keys()) == list(g["state_dict"].keys())
    blk.load_state_dict(g["state_dict"], strict=True)
    x = g["x"].clone().requires_grad_(True)
    y = blk(x, g["edge_index"])
    (y * g["w"]).sum().backward()
    assert torch.equal(y.detach(), g["y"])
    assert torch.equal(x.grad, g["dx"])
    for k, p in blk.named_parameters():
        assert torch.equal(p.grad, g["grads"][k]), k


def test_against_real_pyg_if_present():
    pyg = pytest.importorskip("torch_geometric", reason="PyG absent -- the restatement is the oracle (parity unpinned)")
    from torch_geometric.nn import SAGEConv
    torch.manual_seed(0)
    real = SAGEConv(8, 16)
    ours = SAGEConvOracle(8, 16)
    ours.load_state_dict(real.state_dict(), strict=True)
    x = torch.randn(40, 8)
    ei = torch.randint(0, 40, (2, 200))
    assert torch.allclose(real(x, ei), ours(x, ei), rtol=1e-6, atol=1e-7)


def test_chunked_big_graph_restatement_equals_the_oracle():
    """layer_fwd_bwd_chunked (the fp64 adjudicator of the 1 M-node test) is the same arithmetic as SageBlockOracle."""
    torch.manual_seed(3)
    N, E, hdims, slope = 700, 9000, [24, 40], 0.1
    ei = torch.randint(0, N, (2, E)); ei[1, : E // 3] = 5          # a hub
    x = torch.randn(N, hdims[0]); w = torch.randn(N, hdims[1])
    ref = SageBlockOracle(hdims, negative_slope=slope).double()
    xd = x.double().requires_grad_(True)
    y = ref(xd, ei); y.backward(w.double())
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    y2, dx2, g2 = layer_fwd_bwd_chunked(x, ei, sd, hdims, slope, w, torch.float64, chunk=1000)
    assert torch.allclose(y2, y.detach(), rtol=1e-12, atol=1e-12) and torch.allclose(dx2, xd.grad, rtol=1e-11, atol=1e-12)
    for k, p in ref.named_parameters():
        assert torch.allclose(g2[k], p.grad, rtol=1e-10, atol=1e-11), k
    # fp32, chunked in edge order == the sequential scatter_add_ of the oracle, bit for bit, for the aggregation part
    y3, _, _ = layer_fwd_bwd_chunked(x, ei, {k: v.float() for k, v in sd.items()}, hdims, slope, w, torch.float32, chunk=1000)
    ref32 = SageBlockOracle(hdims, negative_slope=slope); ref32.load_state_dict({k: v.float() for k, v in sd.items()})
    assert torch.allclose(y3, ref32(x, ei).detach(), rtol=1e-5, atol=1e-6)
