"""Graph readout (global_mean_pool / global_max_pool / 'double'): oracle known answers on the CPU, CUDA parity on the GPU.

Reference call sites: src/models/grusage.py:113-120 (choice), :185 (x = self.global_pool(x, batch)).
Tolerance: mean within rtol 1e-5 / atol 1e-6 of the fp32 oracle (the oracle adds a graph's rows sequentially, the
kernel as 8 interleaved partial sums combined in fixed order); max is exact; gradients of max are exact given the
same ties, gradients of mean within the same tolerance."""
import pytest
import torch

from oracle.sage_oracle import global_mean_pool_oracle, global_max_pool_oracle, global_double_pool_oracle

RTOL, ATOL = 1e-5, 1e-6


# ------------------------------------------------------------------ CPU: the oracle itself --
def test_kat_mean_max_empty_graph_and_duplicates():
    x = torch.tensor([[1., -2.], [3., 4.], [5., -6.], [7., 8.]])
    batch = torch.tensor([0, 0, 2, 2])          # graph 1 is empty, size 4 adds an empty graph 3
    mean = global_mean_pool_oracle(x, batch, 4)
    mx = global_max_pool_oracle(x, batch, 4)
    assert torch.equal(mean, torch.tensor([[2., 1.], [0., 0.], [6., 1.], [0., 0.]]))
    assert torch.equal(mx, torch.tensor([[3., 4.], [0., 0.], [7., 8.], [0., 0.]]))     # empty -> 0, not -inf
    assert torch.equal(global_double_pool_oracle(x, batch, 4), torch.cat([mean, mx], 1))
    assert global_mean_pool_oracle(x, batch).shape == (3, 2)                           # size = batch.max() + 1


def test_kat_batch_none_is_one_graph():
    x = torch.arange(12.).view(4, 3)
    assert torch.equal(global_mean_pool_oracle(x, None), x.mean(0, keepdim=True))
    assert torch.equal(global_max_pool_oracle(x, None), x.max(0, keepdim=True)[0])


def test_kat_max_backward_shares_gradient_among_ties():
    x = torch.tensor([[1., 5.], [1., 2.], [0., 5.], [9., 9.]], requires_grad=True)
    batch = torch.tensor([0, 0, 0, 1])
    global_max_pool_oracle(x, batch, 2).sum().backward()
    assert torch.equal(x.grad, torch.tensor([[.5, .5], [.5, 0.], [0., .5], [1., 1.]]))


def test_kat_unsorted_batch():
    x = torch.tensor([[1.], [10.], [2.], [20.]])
    batch = torch.tensor([1, 0, 1, 0])
    assert torch.equal(global_mean_pool_oracle(x, batch, 2), torch.tensor([[15.], [1.5]]))
    assert torch.equal(global_max_pool_oracle(x, batch, 2), torch.tensor([[20.], [2.]]))


# ----------------------------------------------------------------------------- GPU parity --
def _case(kind, seed):
    g = torch.Generator().manual_seed(seed)
    if kind == "sorted":
        sizes = torch.randint(1, 40, (50,), generator=g)
        batch = torch.repeat_interleave(torch.arange(50), sizes)
        return batch, 50
    if kind == "with_empty":      # graphs 3 and 7 have no node, two trailing empty graphs
        sizes = torch.randint(1, 30, (10,), generator=g)
        sizes[3] = 0; sizes[7] = 0
        return torch.repeat_interleave(torch.arange(10), sizes), 12
    if kind == "unsorted":
        batch = torch.randint(0, 17, (900,), generator=g)
        return batch, 17
    if kind == "big_graphs":      # more than 256 nodes per graph: split rows in the membership CSR
        sizes = torch.tensor([700, 1, 3000, 257, 256])
        return torch.repeat_interleave(torch.arange(5), sizes), 5
    if kind == "single":
        return torch.zeros(33, dtype=torch.long), 1
    raise ValueError(kind)


def _close(a, b, what):
    a, b = a.detach().cpu(), b.detach().cpu()
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    err = (a - b).abs()
    assert bool((err <= ATOL + RTOL * b.abs()).all()), f"{what}: max err {float(err.max()):.3e}"


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["sorted", "with_empty", "unsorted", "big_graphs", "single"])
@pytest.mark.parametrize("F", [128, 96, 13, 260])
def test_readout_matches_oracle(kind, F):
    import sldm_gnn_b200 as sg
    dev = torch.device("cuda:0")
    batch, G = _case(kind, seed=F)
    N = batch.numel()
    x = torch.randn(N, F, generator=torch.Generator().manual_seed(F + 1))
    x[::7] = x[0]                                  # repeated rows: tied maxima inside a graph
    w = torch.randn(G, 2 * F, generator=torch.Generator().manual_seed(F + 2))
    xr = x.clone().requires_grad_(True)
    ref = global_double_pool_oracle(xr, batch, G)
    (ref * w).sum().backward()
    xg = x.to(dev).requires_grad_(True)
    got = sg.global_mean_max_pool(xg, batch.to(dev), G)
    (got * w.to(dev)).sum().backward()
    _close(got[:, :F], ref[:, :F], "mean")
    assert torch.equal(got[:, F:].cpu(), ref[:, F:].detach()), "max must be exact"
    _close(xg.grad, xr.grad, "dx")
    # the single-statistic entry points and their gradients
    for fn, orc, sl in ((sg.global_mean_pool, global_mean_pool_oracle, slice(0, F)), (sg.global_max_pool, global_max_pool_oracle, slice(F, 2 * F))):
        xr2 = x.clone().requires_grad_(True)
        (orc(xr2, batch, G) * w[:, sl]).sum().backward()
        xg2 = x.to(dev).requires_grad_(True)
        o = fn(xg2, batch.to(dev), G)
        (o * w[:, sl].to(dev)).sum().backward()
        _close(o, orc(x, batch, G), fn.__name__)
        _close(xg2.grad, xr2.grad, fn.__name__ + " dx")


@pytest.mark.gpu
def test_readout_size_none_batch_none_and_modes():
    import sldm_gnn_b200 as sg
    dev = torch.device("cuda:0")
    x = torch.randn(40, 32)
    batch = torch.repeat_interleave(torch.arange(4), 10)
    _close(sg.global_mean_pool(x.to(dev), batch.to(dev)), global_mean_pool_oracle(x, batch), "size=None")
    _close(sg.global_mean_pool(x.to(dev), None), global_mean_pool_oracle(x, None), "batch=None mean")
    assert torch.equal(sg.global_max_pool(x.to(dev), None).cpu(), global_max_pool_oracle(x, None))
    with torch.inference_mode():
        o = sg.global_mean_max_pool(x.to(dev), batch.to(dev), 4)
    assert o.shape == (4, 64)
    with pytest.raises(RuntimeError):
        sg.global_mean_pool(x, batch)                      # CPU tensors: no fallback
    with pytest.raises(ValueError):
        sg.global_mean_pool(x.to(dev), batch.int().to(dev))
    e = sg.global_mean_max_pool(torch.empty(0, 8, device=dev), torch.empty(0, dtype=torch.long, device=dev), 3)
    assert e.shape == (3, 16) and float(e.abs().sum()) == 0.0


@pytest.mark.gpu
def test_readout_full_batch_size_properties():
    """4096 unit map graphs (bench 'batch' shape): linearity of the mean, max >= mean, sum of means weighted by counts."""
    import sldm_gnn_b200 as sg
    from workloads import unit_map_graphs
    dev = torch.device("cuda:0")
    _, batch, N = unit_map_graphs(4096, seed=0)
    batch = batch.to(dev)
    x = torch.randn(N, 128, device=dev, generator=torch.Generator(device=dev).manual_seed(0))
    out = sg.global_mean_max_pool(x, batch, 4096)
    mean, mx = out[:, :128], out[:, 128:]
    assert bool((mx >= mean - 1e-6).all())
    cnt = torch.bincount(batch, minlength=4096).float()
    total = (mean.double() * cnt.double()[:, None]).sum(0)
    assert torch.allclose(total, x.double().sum(0), rtol=1e-6, atol=1e-3)
    out2 = sg.global_mean_max_pool(2.0 * x, batch, 4096)
    assert torch.equal(out2, 2.0 * out)                     # scaling by 2 is exact in fp32
    assert torch.equal(sg.global_mean_max_pool(x, batch, 4096), out)   # deterministic
