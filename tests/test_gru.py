"""Fused GRU sequence head (csrc/gru.cu, sldm_gnn_b200/gru.py) against the reference's own layer.

The reference computes this step with `torch.nn.GRU` (src/models/grusage.py:55-60, 160-161), so torch's CPU GRU *is*
the reference implementation here: parity is pinned against it directly, in fp32 with the fp64 run as adjudicator
(bar: rtol 1e-5 / atol 1e-6; parameter gradients, sums over all N*T rows, with atol scaled by their magnitude).
"""
import os

import pytest
import torch

import sldm_gnn_b200 as sg
from sldm_gnn_b200 import _lib
from sldm_gnn_b200.gru import fused_gru_eligible, gru_last_hidden

RTOL, ATOL = 1e-5, 1e-6
GOLDEN = sorted(__import__("glob").glob(os.path.join(os.path.dirname(__file__), "golden", "gru", "*.pt")))


# ------------------------------------------------------------------- CPU side --
def test_support_queries_without_gpu():
    lib = _lib.lib
    assert lib.sldm_gru_supported(16, 6, 96) == 1          # the reference's configuration (main.py:42-44)
    assert lib.sldm_gru_supported(16, 8, 64) == 1 and lib.sldm_gru_supported(1, 1, 32) == 1
    assert lib.sldm_gru_supported(16, 6, 128) == 0         # hidden size outside the tile design
    assert lib.sldm_gru_supported(16, 9, 96) == 0          # input too wide
    assert lib.sldm_gru_supported(400, 8, 96) == 0         # x tile does not fit shared memory
    assert lib.sldm_gru_supported(0, 6, 96) == 0
    assert lib.sldm_gru_partial_rows(0) == 0 and lib.sldm_gru_partial_rows(64) == 1 and lib.sldm_gru_partial_rows(65) == 2
    assert lib.sldm_gru_partial_rows(-1) == -1
    assert lib.sldm_gru_partial_width(96) == 28 * 96 and lib.sldm_gru_partial_width(33) == -1


@pytest.mark.parametrize("H,I", [(96, 6), (64, 8), (32, 1)])
def test_partial_layout_decode(H, I):
    """Host half of the backward: the [28*H/32][32] per-tile layout of include/sldm_sage.h, written here element by
    element the way k_gru_bwd's lanes do, decodes to the torch parameter layouts."""
    from sldm_gnn_b200.gru import decode_partials
    U = H // 32
    g = torch.Generator().manual_seed(H + I)
    dW = torch.randn(3 * H, 8, generator=g)
    dW[:, I:] = 0
    dbi, dbn = torch.randn(3 * H, generator=g), torch.randn(H, generator=g)
    P = torch.zeros(int(_lib.lib.sldm_gru_partial_width(H)))
    for u in range(U):
        for lane in range(32):
            for gate in range(3):
                for i in range(8):
                    P[((u * 3 + gate) * 8 + i) * 32 + lane] = dW[gate * H + 32 * u + lane, i]
                P[(24 * U + u * 3 + gate) * 32 + lane] = dbi[gate * H + 32 * u + lane]
            P[(27 * U + u) * 32 + lane] = dbn[32 * u + lane]
    dW_ih, db_ih, db_hh = decode_partials(P, H, I)
    assert dW_ih.shape == (3 * H, I) and torch.equal(dW_ih, dW[:, :I])
    assert torch.equal(db_ih, dbi) and torch.equal(db_hh, torch.cat([dbi[:2 * H], dbn]))


def test_cpu_tensors_are_not_eligible_and_raise():
    gru = torch.nn.GRU(6, 96, 1, batch_first=True)
    x = torch.randn(4, 16, 6)
    assert not fused_gru_eligible(gru, x)
    with pytest.raises(RuntimeError, match="CUDA only"):
        gru_last_hidden(gru, x)


def test_unsupported_shapes_return_eunsupported_without_gpu():
    rc = _lib.lib.sldm_gru_forward(None, 4, 16, 6, 128, None, None, None, None, None, None, None)
    assert rc == _lib.EUNSUPPORTED
    with pytest.raises(NotImplementedError, match="hidden size 128"):
        _lib.check(rc)
    assert _lib.lib.sldm_gru_forward(None, -1, 16, 6, 96, None, None, None, None, None, None, None) == _lib.EINVAL


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-3] for p in GOLDEN])
def test_golden_fixtures_are_what_torch_gru_computes(path):
    """The committed vectors (tests/golden/make_golden_gru.py) are reproduced by this image's torch.nn.GRU on the CPU:
    a drift of the library under the fixtures would show here, not as a false alarm in the GPU parity test."""
    fix = torch.load(path)
    c = fix["cfg"]
    g = torch.nn.GRU(c["I"], c["H"], 1, batch_first=True)
    g.load_state_dict(fix["state_dict"])
    x = fix["x"].clone().requires_grad_(True)
    h = g(x)[1][-1]
    h.backward(fix["up"])
    assert torch.allclose(h, fix["f32"]["h"], rtol=1e-6, atol=1e-7)
    assert torch.allclose(x.grad, fix["f32"]["dx"], rtol=1e-5, atol=1e-6)
    for k, p in g.named_parameters():
        want = fix["f32"]["grads"][k]
        assert torch.allclose(p.grad, want, rtol=1e-5, atol=1e-6 * max(1.0, float(want.abs().max()))), k
    assert len(GOLDEN) == 3


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-3] for p in GOLDEN])
def test_numpy_oracle_reproduces_the_golden_vectors(path):
    """oracle/gru_oracle.py (the recurrence and the backward decomposition csrc/gru.cu implements, as numpy loops) against
    the vectors of the reference's layer, in fp64."""
    from oracle.gru_oracle import gru_last_hidden_oracle, gru_last_hidden_backward_oracle
    fix = torch.load(path)
    sd = {k: v.double().numpy() for k, v in fix["state_dict"].items()}
    x = fix["x"].double().numpy()
    args = (sd["weight_ih_l0"], sd["weight_hh_l0"], sd["bias_ih_l0"], sd["bias_hh_l0"])
    h, tape = gru_last_hidden_oracle(x, *args)
    want = fix["f64"]
    assert abs(h - want["h"].numpy()).max() < 1e-12
    dx, dW_ih, dW_hh, db_ih, db_hh = gru_last_hidden_backward_oracle(fix["up"].double().numpy(), x, args[0], args[1], tape)
    got = {"weight_ih_l0": dW_ih, "weight_hh_l0": dW_hh, "bias_ih_l0": db_ih, "bias_hh_l0": db_hh}
    for k, g in got.items():
        w = want["grads"][k].numpy()
        assert abs(g - w).max() <= 1e-10 * max(1.0, abs(w).max()), k
    assert abs(dx - want["dx"].numpy()).max() < 1e-10


# ------------------------------------------------------------------- GPU side --
def _close(got, want, want64, what, scale_atol=False):
    got, want, w64 = got.detach().cpu(), want.detach(), want64.detach().double()
    atol = ATOL * (max(1.0, float(want.abs().max())) if scale_atol else 1.0)
    err = (got - want).abs()
    bad = err > atol + RTOL * want.abs()
    if not bad.any():
        return
    e_g = (got.double() - w64).abs()[bad]
    e_r = float((want.double() - w64).abs().max())
    ok = (e_g <= atol + RTOL * w64.abs()[bad]) | (e_g <= 1.5 * e_r)
    assert ok.all(), (f"{what}: {int(bad.sum())}/{bad.numel()} outside tolerance, max err {float(err.max()):.3e}; "
                      f"vs fp64: ours {float(e_g.max()):.3e}, torch fp32 {e_r:.3e}")


def _reference(gru, x, up, need_dx):
    """torch's own GRU on the CPU, fp32 and fp64: (h, grads..., dx) each."""
    out = []
    for dt in (torch.float32, torch.float64):
        g = torch.nn.GRU(gru.input_size, gru.hidden_size, 1, batch_first=True).to(dt)
        g.load_state_dict({k: v.detach().cpu().to(dt) for k, v in gru.state_dict().items()})
        xi = x.detach().cpu().to(dt).requires_grad_(need_dx)
        h = g(xi)[1][-1]
        h.backward(up.detach().cpu().to(dt))
        out.append((h.detach(), {k: p.grad for k, p in g.named_parameters()}, xi.grad))
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("N,T,I,H,need_dx", [
    (1, 1, 1, 32, False), (5, 3, 6, 96, True), (64, 16, 6, 96, False), (65, 16, 6, 96, True), (1000, 16, 6, 96, False),
    (300, 7, 8, 64, True), (130, 16, 3, 32, False), (4099, 16, 6, 96, False), (63, 25, 5, 96, True),
])
def test_fused_gru_matches_torch_gru(N, T, I, H, need_dx):
    dev = torch.device("cuda:0")
    torch.manual_seed(N * 131 + T * 7 + I + H)
    gru = torch.nn.GRU(I, H, 1, batch_first=True).to(dev)
    x = torch.randn(N, T, I, device=dev, requires_grad=need_dx)
    up = torch.randn(N, H, device=dev)
    assert fused_gru_eligible(gru, x)
    before = _lib.lib.sldm_launch_count()
    h = gru_last_hidden(gru, x)
    h.backward(up)
    torch.cuda.synchronize()
    assert _lib.lib.sldm_launch_count() - before == 3, "forward, reverse recurrence, dW_hh: one kernel each"
    (h32, g32, dx32), (h64, g64, dx64) = _reference(gru, x, up, need_dx)
    _close(h, h32, h64, "h_last")
    for k, p in gru.named_parameters():
        _close(p.grad, g32[k], g64[k], f"grad {k}", scale_atol=True)
    if need_dx:
        _close(x.grad, dx32, dx64, "dx")
    else:
        assert x.grad is None


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-3] for p in GOLDEN])
def test_fused_gru_matches_golden(path):
    """Against the committed vectors of the reference's layer (no torch GRU call on this side)."""
    dev = torch.device("cuda:0")
    fix = torch.load(path)
    c = fix["cfg"]
    gru = torch.nn.GRU(c["I"], c["H"], 1, batch_first=True)
    gru.load_state_dict(fix["state_dict"])
    gru = gru.to(dev)
    x = fix["x"].to(dev).requires_grad_(True)
    h = gru_last_hidden(gru, x)
    h.backward(fix["up"].to(dev))
    _close(h, fix["f32"]["h"], fix["f64"]["h"], "h_last")
    _close(x.grad, fix["f32"]["dx"], fix["f64"]["dx"], "dx")
    for k, p in gru.named_parameters():
        _close(p.grad, fix["f32"]["grads"][k], fix["f64"]["grads"][k], f"grad {k}", scale_atol=True)


@pytest.mark.gpu
def test_fused_gru_inference_equals_training_forward_and_is_deterministic():
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    gru = torch.nn.GRU(6, 96, 1, batch_first=True).to(dev)
    x = torch.randn(777, 16, 6, device=dev)
    up = torch.randn(777, 96, device=dev)
    with torch.no_grad():
        h0 = gru_last_hidden(gru, x)
    with torch.inference_mode():
        h1 = gru_last_hidden(gru, x)
    grads = []
    for _ in range(2):
        gru.zero_grad()
        h2 = gru_last_hidden(gru, x)
        h2.backward(up)
        grads.append([p.grad.clone() for p in gru.parameters()])
    assert torch.equal(h0, h1) and torch.equal(h0, h2.detach())
    assert all(torch.equal(a, b) for a, b in zip(*grads)), "parameter gradients are bit-identical run to run"


@pytest.mark.gpu
def test_grusage_uses_the_fused_gru_and_matches_the_library_path(monkeypatch):
    """The model's sequence head goes through the kernels (launch counter) and agrees with torch's library GRU."""
    dev = torch.device("cuda:0")
    torch.manual_seed(11)
    m = sg.GruSage(dynamic_features_num=6, frames_num=16, gru_hidden_size=96, gru_num_layers=1, fc1dims=[96],
                   sage_hidden_dims=[96, 96], fc2dims=[32], dropout=None, negative_slope=0.1,
                   map_included=False).to(dev)
    x = torch.randn(500, 16, 6, device=dev)
    before = _lib.lib.sldm_launch_count()
    h = m._last_hidden(x)
    assert _lib.lib.sldm_launch_count() - before == 1
    monkeypatch.setenv("SLDM_DISABLE_FUSED_GRU", "1")
    before = _lib.lib.sldm_launch_count()
    h_lib = m._last_hidden(x)
    assert _lib.lib.sldm_launch_count() == before
    assert torch.allclose(h, h_lib, rtol=1e-5, atol=2e-6), float((h - h_lib).abs().max())


@pytest.mark.gpu
def test_ineligible_modules_stay_on_the_library():
    dev = torch.device("cuda:0")
    x = torch.randn(8, 16, 6, device=dev)
    assert not fused_gru_eligible(torch.nn.GRU(6, 96, 2, batch_first=True).to(dev), x)
    assert not fused_gru_eligible(torch.nn.GRU(6, 96, 1, batch_first=False).to(dev), x)
    assert not fused_gru_eligible(torch.nn.GRU(6, 128, 1, batch_first=True).to(dev), x)
    assert not fused_gru_eligible(torch.nn.GRU(6, 96, 1, batch_first=True, bidirectional=True).to(dev), x)
    assert not fused_gru_eligible(torch.nn.GRU(6, 96, 1, batch_first=True).to(dev), x.double())
    with pytest.raises(NotImplementedError):
        gru_last_hidden(torch.nn.GRU(6, 128, 1, batch_first=True).to(dev), x)
