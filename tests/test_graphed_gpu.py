"""GraphedSageBlock: inference through one captured CUDA graph per size bucket (the reference's online path calls the
model on ONE small graph from a worker thread: rcv.py:77-84, :107).  Rows [0, N) must equal the eager module's bit for
bit for every (N, E) that fits the bucket."""
import threading

import pytest
import torch

import sldm_gnn_b200 as sg
from workloads import unit_map_graphs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _case(n_graphs, seed, F, dev):
    ei, _, N = unit_map_graphs(n_graphs, seed=seed)
    x = torch.randn(N, F, generator=torch.Generator().manual_seed(seed))
    return x.to(dev), ei.to(dev)


@pytest.mark.parametrize("hdims", [[64, 64, 64], [128, 96, 96], [16, 32, 32]])
def test_graphed_equals_eager_bit_for_bit(dev, hdims):
    torch.manual_seed(1)
    blk = sg.SageBlock(hdims, dropout=0.25, negative_slope=0.1).to(dev).eval()
    g = blk.graphed(max_nodes=1024, max_edges=5000)
    for n_graphs, seed in ((1, 0), (3, 1), (4, 2), (1, 3)):          # different N and E through the same captured graph
        x, ei = _case(n_graphs, seed, hdims[0], dev)
        with torch.inference_mode():
            want = blk(x, ei)
        got = g(x, ei)
        assert got.shape == want.shape and torch.equal(got, want), (n_graphs, seed)
    x, ei = _case(1, 5, hdims[0], dev)
    got = g(x, ei[:, :0])                                             # E = 0 is legal (torch.empty((2, 0)) in rcv.py)
    with torch.inference_mode():
        assert torch.equal(got, blk(x, ei[:, :0]))
    with pytest.raises(RuntimeError, match="exceed the bucket"):
        g(*_case(8, 0, hdims[0], dev))


def test_graphed_from_worker_threads(dev):
    torch.manual_seed(2)
    blk = sg.SageBlock([64, 64, 64], negative_slope=0.1).to(dev).eval()
    g = blk.graphed(max_nodes=512, max_edges=2048)
    cases = [_case(1, s, 64, dev) for s in range(4)]
    with torch.inference_mode():
        want = [blk(x, ei) for x, ei in cases]
    got = [None] * len(cases)

    def work(i):
        for _ in range(5):
            got[i] = g(*cases[i])

    ts = [threading.Thread(target=work, args=(i,)) for i in range(len(cases))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for a, b in zip(got, want):
        assert torch.equal(a, b)


def test_graphed_training_matches_eager(dev):
    """Training through the forward / backward graph pair: output and dx equal the eager module's bit for bit (same
    kernels on padded buffers, every kernel row-wise independent); the parameter gradients are split-K sums whose
    slab boundaries depend on the (padded) row count, so they agree to fp32 rounding, not bit for bit."""
    torch.manual_seed(3)
    hdims = [64, 64, 64]
    blk = sg.SageBlock(hdims, negative_slope=0.1).to(dev).train()
    g = blk.graphed(max_nodes=1024, max_edges=5000, training=True)
    for n_graphs, seed in ((2, 0), (4, 1), (1, 2)):
        x, ei = _case(n_graphs, seed, hdims[0], dev)
        w = torch.randn(x.size(0), hdims[-1], device=dev, generator=torch.Generator(device=dev).manual_seed(seed))
        res = []
        for fn in (blk, g):
            blk.zero_grad(set_to_none=True)
            xg = x.clone().requires_grad_(True)
            y = fn(xg, ei)
            (y * w).sum().backward()
            res.append([y.detach().clone(), xg.grad.clone()] + [p.grad.clone() for p in blk.parameters()])
        for a, b in zip(res[0][:2], res[1][:2]):
            assert torch.equal(a, b), (n_graphs, seed)
        for a, b in zip(res[0][2:], res[1][2:]):
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-6 * max(1.0, float(a.abs().max()))), (n_graphs, seed)
    opt = torch.optim.SGD(blk.parameters(), lr=0.1)         # parameters are updated in place: the graphs see the new values
    x, ei = _case(2, 5, hdims[0], dev)
    blk.zero_grad(set_to_none=True)
    g(x, ei).square().mean().backward()
    opt.step()
    with torch.no_grad():
        assert torch.equal(g(x, ei), blk(x, ei))


def test_dropout_matches_torch_functional_dropout(dev):
    """The block's dropout is torch.native_dropout: the kernel nn.Dropout / F.dropout dispatch to on CUDA, so the global
    Philox stream is consumed exactly as by the reference's posts[i][2] (src/models/blocks/sageblock.py:13)."""
    t = torch.randn(1000, 64, device=dev)
    torch.manual_seed(77)
    a = torch.nn.functional.dropout(t, 0.25, True)
    torch.manual_seed(77)
    b, mask = torch.native_dropout(t, 0.25, True)
    assert torch.equal(a, b) and torch.equal(mask, a != 0)
    # and through the module: same mask as running torch's Dropout on the un-dropped activation of the layer
    ei, _, N = unit_map_graphs(2, seed=2)
    blk = sg.SageBlock([16, 32], dropout=0.5, negative_slope=0.1).to(dev)
    x = torch.randn(N, 16, device=dev)
    blk.eval()
    clean = blk(x, ei.to(dev))
    blk.train()
    torch.manual_seed(5)
    got = blk(x, ei.to(dev))
    torch.manual_seed(5)
    want = torch.nn.functional.dropout(clean, 0.5, True)
    assert torch.equal(got, want)
