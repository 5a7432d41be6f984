"""world_size-2 gloo test (CPU) of the graph-sharded data-parallel wrapper.  The module
being wrapped here is the CPU oracle -- this exercises the host logic only (sharding,
flat bucket, one weighted all-reduce), which is all the N>1 path adds to the kernels."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _make_graph_batch(graph_ids, F, seed):
    """Block-diagonal batch of small graphs; graph g has 5+g%4 nodes, 3x as many edges."""
    xs, eis, sizes, off = [], [], [], 0
    for g in graph_ids:
        gen = torch.Generator().manual_seed(seed * 1000 + g)
        n = 5 + g % 4
        xs.append(torch.randn(n, F, generator=gen))
        eis.append(torch.randint(0, n, (2, 3 * n), generator=gen) + off)
        sizes.append(n); off += n
    return torch.cat(xs), torch.cat(eis, 1), sizes


def _loss(model, x, ei, sizes):
    y = model(x, ei)
    per_graph = torch.stack([c.square().mean() for c in torch.split(y, sizes)])
    return per_graph.mean()          # mean over LOCAL graphs, like BCE 'mean' in src/utils.py:224


def _worker(rank, world, port, num_graphs, out):
    sys.path.insert(0, ROOT)
    from sldm_gnn_b200.parallel import GraphDataParallel, shard_graphs
    from oracle.sage_oracle import SageBlockOracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)            # replicas start DIFFERENT: broadcast must fix it
        model = SageBlockOracle([6, 8, 4], negative_slope=0.1)
        ddp = GraphDataParallel(model)
        assert ddp.num_buckets == 2              # one bucket per SageBlock layer, last layer first (backward order)
        assert [len(b["params"]) for b in ddp._buckets] == [5, 5]
        assert ddp._buckets[0]["params"][0] is model.convs[1].lin_l.weight
        # after the broadcast every rank holds rank 0's parameters
        flat = ddp._flat.clone()
        ref = flat.clone(); dist.broadcast(ref, 0)
        assert torch.equal(flat, ref)
        mine = shard_graphs(num_graphs, rank, world)
        x, ei, sizes = _make_graph_batch(list(mine), 6, seed=7)
        ddp.zero_grad()
        _loss(ddp, x, ei, sizes).backward()
        ddp.sync_gradients(local_weight=len(mine))
        # single-process truth: all graphs in one batch, same parameters
        torch.manual_seed(100)
        full = SageBlockOracle([6, 8, 4], negative_slope=0.1)
        full.load_state_dict(model.state_dict())
        xf, eif, sf = _make_graph_batch(list(range(num_graphs)), 6, seed=7)
        _loss(full, xf, eif, sf).backward()
        for (k, p), (_, q) in zip(model.named_parameters(), full.named_parameters()):
            assert torch.allclose(p.grad, q.grad, rtol=1e-5, atol=1e-7), k
        # gradients still live in the flat bucket and an optimizer step keeps replicas equal
        assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(ddp._params, ddp._views))
        torch.optim.SGD(ddp.parameters(), lr=0.1).step()
        after = ddp._flat.clone(); ref = after.clone(); dist.broadcast(ref, 0)
        assert torch.equal(after, ref)
        out[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_graph_data_parallel_world2_gloo():
    world, num_graphs = 2, 7                      # uneven split: 4 + 3 graphs
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), num_graphs, out), nprocs=world, join=True)
    assert dict(out) == {0: "ok", 1: "ok"}


def test_bucket_keys_follow_backward_order():
    from sldm_gnn_b200.parallel import default_bucket_key as key
    names = ["sage.convs.0.lin_l.weight", "sage.posts.0.0.bias", "sage.convs.1.lin_r.weight", "sage.posts.1.0.weight",
             "fc1.0.weight", "map_encoder.sage.convs.0.lin_l.bias"]
    assert key(names[0]) == key(names[1]) != key(names[2]) == key(names[3])
    assert key(names[4]) == ("", -1) and key(names[5]) == ("map_encoder.sage.", 0)
    order = sorted({key(n) for n in names}, reverse=True)
    assert order.index(key(names[2])) < order.index(key(names[0]))     # layer 1 is exchanged before layer 0


def test_shard_graphs_partition():
    from sldm_gnn_b200.parallel import shard_graphs
    for n in (0, 1, 7, 8, 4096, 65536):
        for w in (1, 2, 4, 8):
            parts = [shard_graphs(n, r, w) for r in range(w)]
            assert sum(len(p) for p in parts) == n
            assert [i for p in parts for i in p] == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_single_process_wrapper_is_transparent():
    from sldm_gnn_b200.parallel import GraphDataParallel
    from oracle.sage_oracle import SageBlockOracle
    torch.manual_seed(0)
    m = SageBlockOracle([4, 4])
    keys = list(m.state_dict().keys())
    w = GraphDataParallel(m)
    assert list(m.state_dict().keys()) == keys
    x, ei, sizes = _make_graph_batch([0, 1], 4, seed=1)
    _loss(w, x, ei, sizes).backward()
    w.sync_gradients()                            # no process group: no-op
    assert float(w.flat_grad.abs().sum()) > 0
    w.zero_grad()
    assert float(w.flat_grad.abs().sum()) == 0


def test_direct_write_claims_only_whole_zeroed_buckets(monkeypatch):
    """Host logic of the direct-write path (parallel.GraphDataParallel.claim): the block's backward may write a
    layer's gradients straight into the bucket views only while the wrapper is armed (zero_grad -> expect_sync), for
    exactly one whole bucket, once; everything else falls back to autograd accumulation."""
    import sldm_gnn_b200.parallel as par
    from oracle.sage_oracle import SageBlockOracle
    torch.manual_seed(0)
    m = SageBlockOracle([4, 6, 3])
    w = par.GraphDataParallel(m)
    w._overlap = True                                   # pretend CUDA + a process group (claim() itself launches nothing)
    monkeypatch.setattr(w, "_ready", lambda: True)
    layer = lambda l: [m.convs[l].lin_l.weight, m.convs[l].lin_l.bias, m.convs[l].lin_r.weight, m.posts[l][0].weight, m.posts[l][0].bias]
    assert w.claim(layer(1)) is None                    # not armed
    w.zero_grad(); w.expect_sync(local_weight=3.0)
    assert par.active_wrapper() is w
    bi, views = w.claim(layer(1))
    assert bi == 0 and all(v.data_ptr() == p.grad.data_ptr() for v, p in zip(views, layer(1)))   # last layer = first bucket
    assert w.claim(layer(0))[0] == 1
    assert w.claim(layer(1)[:4]) is None                # not a whole bucket
    assert w.claim(layer(0)[:3] + layer(1)[3:]) is None # parameters of two buckets
    w._launched.add(0)
    assert w.claim(layer(1)) is None                    # already exchanged
    m.convs[0].lin_l.weight.grad = torch.zeros_like(m.convs[0].lin_l.weight)
    assert w.claim(layer(0)) is None                    # someone replaced .grad: not the bucket view any more
    monkeypatch.setattr(w, "_ready", lambda: False)
    w.sync_gradients(local_weight=3.0)                  # disarms
    assert par.active_wrapper() is None and w.claim(layer(1)) is None
    monkeypatch.setattr(w, "_ready", lambda: True)
    w.expect_sync(local_weight=3.0)                     # gradients were not zeroed since the last backward: stays off
    assert par.active_wrapper() is None and w.claim(layer(1)) is None
