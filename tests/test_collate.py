"""Mini-batch assembly: oracle known answers on the CPU, bit-exact CUDA parity on the GPU.
Reference call sites: main.py:166-167 (DataLoader), src/utils.py:218-223 and src/models/grusage.py:153-173 (fields)."""
import pytest
import torch

from oracle.collate_oracle import collate_oracle


class _D:   # minimal stand-in for a graph item (the oracle only reads attributes)
    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)


def _graphs(sizes, seed, T=4):
    g = torch.Generator().manual_seed(seed)
    out = []
    for i, (n, e) in enumerate(sizes):
        out.append(dict(
            x=torch.randn(n, T, 6, generator=g),
            edge_index=torch.randint(0, max(n, 1), (2, e), generator=g) if n > 0 else torch.empty((2, 0), dtype=torch.long),
            edge_attr=torch.randn(e, 3, generator=g),
            xsttype=torch.randint(0, 5, (n,), generator=g),
            xdims=torch.randn(n, 2, generator=g),
            pos_raw=torch.randn(n, T, 2, generator=g),
            y=torch.randint(0, 2, (1, 4), generator=g).float(),
            flag=torch.rand(n, generator=g) > 0.5,          # 1-byte dtype: unaligned chunk sizes
        ))
    return out


def test_kat_collate_oracle():
    a = _D(x=torch.zeros(2, 1), edge_index=torch.tensor([[0, 1], [1, 0]]), y=torch.tensor([[1.]]))
    b = _D(x=torch.ones(3, 1), edge_index=torch.tensor([[0, 2], [2, 1]]), y=torch.tensor([[0.]]))
    o = collate_oracle([a, b])
    assert torch.equal(o["edge_index"], torch.tensor([[0, 1, 2, 4], [1, 0, 4, 3]]))
    assert torch.equal(o["batch"], torch.tensor([0, 0, 1, 1, 1])) and torch.equal(o["ptr"], torch.tensor([0, 2, 5]))
    assert o["x"].shape == (5, 1) and torch.equal(o["y"], torch.tensor([[1.], [0.]])) and o["num_graphs"] == 2


@pytest.mark.gpu
@pytest.mark.parametrize("sizes", [
    [(5, 12), (7, 0), (1, 3)],                                   # a graph without edges
    [(200, 1000)] * 32,                                          # the reference's batch: 32 graphs
    [(0, 0), (3, 4), (0, 0), (17, 40)],                          # graphs without nodes
    [(33, 129)],                                                 # a single graph
    [(n, 5 * n) for n in range(150, 250, 3)],                    # ragged unit map graphs
])
def test_collate_is_bit_exact(sizes):
    import sldm_gnn_b200 as sg
    dev = torch.device("cuda:0")
    items = _graphs(sizes, seed=len(sizes))
    want = collate_oracle([_D(**d) for d in items])
    got = sg.collate([sg.GraphData(**{k: v.to(dev) for k, v in d.items()}) for d in items])
    torch.cuda.synchronize()
    for k in ("x", "edge_index", "edge_attr", "xsttype", "xdims", "pos_raw", "y", "flag", "batch", "ptr"):
        g, w = getattr(got, k).cpu(), want[k]
        assert g.dtype == w.dtype and g.shape == w.shape, (k, g.dtype, w.dtype, g.shape, w.shape)
        assert torch.equal(g, w), k
    assert got.num_graphs == want["num_graphs"]


@pytest.mark.gpu
def test_collate_feeds_the_block_and_the_readout():
    """collate -> SageBlock -> readout equals running every graph alone (block-diagonal batches do not mix graphs)."""
    import sldm_gnn_b200 as sg
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    blk = sg.SageBlock([16, 32, 32], negative_slope=0.1).to(dev).eval()
    items = [sg.GraphData(x=torch.randn(n, 16, device=dev), edge_index=torch.randint(0, n, (2, 4 * n), device=dev))
             for n in (9, 30, 17)]
    b = sg.collate(items)
    with torch.no_grad():
        pooled = sg.global_mean_max_pool(blk(b.x, b.edge_index), b.batch, b.num_graphs)
        for g, it in enumerate(items):
            alone = sg.global_mean_max_pool(blk(it.x, it.edge_index), None)
            assert torch.allclose(pooled[g], alone[0], rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_collate_errors():
    import sldm_gnn_b200 as sg
    dev = torch.device("cuda:0")
    with pytest.raises(ValueError):
        sg.collate([])
    with pytest.raises(RuntimeError):
        sg.collate([sg.GraphData(x=torch.randn(3, 2), edge_index=torch.zeros((2, 0), dtype=torch.long))])   # CPU tensors
    with pytest.raises(ValueError):
        sg.collate([sg.GraphData(x=torch.randn(3, 2, device=dev), edge_index=torch.zeros((2, 1), dtype=torch.int32, device=dev))])
