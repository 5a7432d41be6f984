"""Parity AT the configurations bench.py quotes (not only on small graphs):

  * `batch`  -- the default bench line: 4096 unit map graphs (~0.82 M nodes, ~4.1 M edges), SageBlock([128,128,128]),
               the exact generator / seeds / upstream gradient of bench.py: output, dx and all 10 parameter gradients
               against the fp32 oracle in full, fp64 oracle as adjudicator; the inference_mode path on the same inputs;
  * `c4`     -- 1 M nodes / 10 M skewed edges, SageBlock([128,128]) (BASELINE configs[3]): output, dx (every row, so
               the hottest source, the 1e5-edge hub and the hub's neighbours are all covered) and every parameter
               gradient in full -- the split-K weight gradient over 148 CTAs x 1 M rows and dx through the hub rows.

Reference lines: src/models/blocks/sageblock.py:16-20 (the loop), src/utils.py:225 (loss.backward()).
The oracle costs ~1 s (batch) / ~4 s (c4) per fwd+bwd on the box's host cores.
"""
import pytest
import torch

import bench
import sldm_gnn_b200 as sg
from oracle.sage_oracle import SageBlockOracle, layer_fwd_bwd_chunked as _big_graph_oracle
from parity_util import assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _blocks(hdims, dev, slope=bench.SLOPE):
    """The bench's own initialisation (torch.manual_seed(0), then the constructor) for the CUDA block; the fp32 and
    fp64 oracles load its state dict."""
    torch.manual_seed(0)
    ours = sg.SageBlock(hdims, dropout=None, negative_slope=slope)
    ref = SageBlockOracle(hdims, dropout=None, negative_slope=slope)
    ref.load_state_dict(ours.state_dict(), strict=True)
    ref64 = SageBlockOracle(hdims, dropout=None, negative_slope=slope).double()
    ref64.load_state_dict({k: v.double() for k, v in ours.state_dict().items()})
    return ours.to(dev), ref, ref64


def _compare_all(ours, ref, ref64, got, want, want64):
    (yg, dxg), (yr, dxr), (yd, dxd) = got, want, want64
    stats = {"output": assert_close(yg, yr, "output", want64=yd),
             "dx": assert_close(dxg, dxr, "dx", scale_atol=True, want64=dxd)}
    rp, dp = dict(ref.named_parameters()), dict(ref64.named_parameters())
    for k, p in ours.named_parameters():
        assert p.grad is not None, k
        stats[k] = assert_close(p.grad, rp[k].grad, f"grad {k}", scale_atol=True, want64=dp[k].grad)
    return stats


def test_batch_workload_fwd_bwd_and_inference(dev):
    wl = bench.WORKLOADS["batch"]
    hdims = wl["hdims"]
    x, ei, N, graphs, _ = bench.make_inputs(wl, 0)
    assert graphs == 4096 and hdims == [128, 128, 128]
    w = torch.randn(N, hdims[-1], generator=torch.Generator().manual_seed(7 + N))   # bench.py's upstream gradient
    ours, ref, ref64 = _blocks(hdims, dev)

    xr = x.clone().requires_grad_(True)
    yr = ref(xr, ei)
    yr.backward(w)
    xd = x.double().requires_grad_(True)
    yd = ref64(xd, ei)
    yd.backward(w.double())

    eid = ei.to(dev)
    xg = x.to(dev).requires_grad_(True)
    yg = ours(xg, eid)
    yg.backward(w.to(dev))
    _compare_all(ours, ref, ref64, (yg, xg.grad), (yr, xr.grad), (yd.detach(), xd.grad))

    # the inference path (test.py:136-139) on the same batch: same numbers as the training forward, nothing saved
    ours.clear_cache()
    with torch.inference_mode():
        yi = ours(x.to(dev), eid)
    assert_close(yi, yr, "inference_mode output", want64=yd.detach())
    assert torch.equal(yi, yg.detach()), "inference and training forward must agree bit for bit"


def test_batch_workload_without_the_activation_kink(dev):
    """The same batch with negative_slope = 1 (LeakyReLU becomes the identity).  At the fp32 noise floor a handful of
    the 1e8 pre-activations change sign between two correct fp32 evaluations; with slope 0.1 each such element moves
    its gradient by a factor 10, which is what the adjudicated counts of the test above are made of (the fp32 oracle
    misses the bar against fp64 just as often).  Without the kink every kernel of the step -- gather, tcgen05
    projection, LayerNorm, dgrad, split-K wgrad over 0.82 M rows, transpose gather -- is compared at face value:
    no tensor may need more than 1e-3 of its elements adjudicated (measured: dx and all six weight gradients need
    none, the output 2.6e-5), except where the fp32 oracle ITSELF misses the bar against fp64 on as many elements:
    torch's fp32 column sums over 0.82 M rows -- the LayerNorm gradients -- are 10-20x further from fp64 than ours."""
    wl = bench.WORKLOADS["batch"]
    hdims = wl["hdims"]
    x, ei, N, _, _ = bench.make_inputs(wl, 0)
    w = torch.randn(N, hdims[-1], generator=torch.Generator().manual_seed(7 + N))
    ours, ref, ref64 = _blocks(hdims, dev, slope=1.0)
    xr = x.clone().requires_grad_(True)
    yr = ref(xr, ei)
    yr.backward(w)
    xd = x.double().requires_grad_(True)
    yd = ref64(xd, ei)
    yd.backward(w.double())
    xg = x.to(dev).requires_grad_(True)
    yg = ours(xg, ei.to(dev))
    yg.backward(w.to(dev))
    stats = _compare_all(ours, ref, ref64, (yg, xg.grad), (yr, xr.grad), (yd.detach(), xd.grad))
    for k, (adj, total, oracle_misses) in stats.items():
        assert adj <= max(2, 1e-3 * total) or adj <= (oracle_misses or 0) + 2, \
            f"{k}: {adj}/{total} elements outside rtol 1e-5 / atol 1e-6 (scaled); the fp32 oracle misses {oracle_misses}"


def test_c4_full_backward_parity(dev):
    import psutil
    if psutil.virtual_memory().available < 40 * 2 ** 30:
        pytest.skip("needs ~25 GB of host memory for the fp32 oracle's [E, F] intermediates")
    wl = bench.WORKLOADS["c4"]
    hdims = wl["hdims"]
    x, ei, N, _, _ = bench.make_inputs(wl, 0)
    assert N == 1_000_000 and ei.size(1) == 10_000_000
    w = torch.randn(N, hdims[-1], generator=torch.Generator().manual_seed(7 + N))
    ours, ref, _ = _blocks(hdims, dev)

    xr = x.clone().requires_grad_(True)
    yr = ref(xr, ei)                      # the oracle proper: index_select + scatter_add_ over all 10 M edges
    yr.backward(w)
    state = {k: v.detach().cpu() for k, v in ours.state_dict().items()}
    yd, dxd, gd = _big_graph_oracle(x, ei, state, hdims, bench.SLOPE, w, torch.float64)
    # the chunked restatement is the same arithmetic: in fp32 it must reproduce the oracle (checks the adjudicator)
    y32, dx32, _ = _big_graph_oracle(x, ei, state, hdims, bench.SLOPE, w, torch.float32)
    assert torch.allclose(y32, yr.detach(), rtol=1e-5, atol=1e-5) and torch.allclose(dx32, xr.grad, rtol=1e-4, atol=1e-4)

    eid = ei.to(dev)
    xg = x.to(dev).requires_grad_(True)
    yg = ours(xg, eid)
    yg.backward(w.to(dev))
    assert_close(yg, yr, "c4 output", want64=yd)
    assert_close(xg.grad, xr.grad, "c4 dx", scale_atol=True, want64=dxd)
    rp = dict(ref.named_parameters())
    for k, p in ours.named_parameters():
        assert_close(p.grad, rp[k].grad, f"c4 grad {k}", scale_atol=True, want64=gd[k])
    # the rows the verdict names, explicitly: hottest destination (the ~1e5-edge hub), its in-neighbours, hottest source
    deg_in, deg_out = torch.bincount(ei[1], minlength=N), torch.bincount(ei[0], minlength=N)
    hub, hot_src = int(deg_in.argmax()), int(deg_out.argmax())
    rows = torch.unique(torch.cat([ei[0][ei[1] == hub][:2000], torch.tensor([hub, hot_src])]))
    assert int(deg_in[hub]) > 50_000
    assert_close(xg.grad.cpu()[rows], xr.grad[rows], "c4 dx (hub, hub neighbours, hottest source)", scale_atol=True,
                 want64=dxd[rows])
    assert_close(yg.detach().cpu()[hub:hub + 1], yr.detach()[hub:hub + 1], "c4 output (hub row)", want64=yd[hub:hub + 1])
