"""Golden vectors for MapSpatialAttention produced by THE REFERENCE ITSELF (plain torch, runs in the authoring container):

    python tests/golden/make_golden_map_attention.py        # needs /root/reference; writes tests/golden/map_attention/*.pt

Each fixture holds the inputs, the reference module's state dict, its output and the gradients of
sum(output * upstream) with respect to the map embeddings and the four MLP tensors."""
import os
import sys

import torch

sys.path.insert(0, "/root/reference")
from src.models.map.mapattention import MapSpatialAttention   # noqa: E402  (the reference class, unmodified)

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "map_attention")
CASES = {"small_k5": (40, 64, 32, 5), "k1": (17, 9, 8, 1), "s_equals_k": (11, 5, 16, 5), "wide": (300, 700, 96, 3)}

for name, (B, S, D, K) in CASES.items():
    g = torch.Generator().manual_seed(len(name) * 1000 + B)
    cent = torch.rand(S, 2, generator=g) * 200.0 - 100.0
    pos = torch.rand(B, 2, generator=g) * 200.0 - 100.0
    emb = torch.randn(S, D, generator=g).requires_grad_(True)
    up = torch.randn(B, D, generator=g)
    torch.manual_seed(7)
    ref = MapSpatialAttention(cent, k_neighbors=K)
    out = ref(pos, emb)
    (out * up).sum().backward()
    torch.save({"centroids": cent, "pos": pos, "emb": emb.detach(), "upstream": up, "k": K,
                "state_dict": {k: v.detach().clone() for k, v in ref.state_dict().items()},
                "out": out.detach(), "demb": emb.grad.clone(),
                "dparams": {k: p.grad.clone() for k, p in ref.named_parameters()}},
               os.path.join(OUT, f"ref_{name}.pt"))
    print(name, tuple(out.shape))
