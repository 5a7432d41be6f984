"""CPU oracle for the proximity-edge construction -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/sage_oracle.py).

PARITY UNPINNED: the reference's builders (src/gbuilder.py) import pandas / torch_geometric and cannot be imported
here; the edge loop itself (gbuilder.py:88-112, identical at :244-268) is numpy on torch slices and is restated below
with the same numpy calls, so the float32 arithmetic (np.linalg.norm, min, max, mean, mean of squares) is numpy's own.
"""
import numpy as np
import torch


def proximity_edges_oracle(x: torch.Tensor, m_radius: float):
    """x: [V, T, F] float32 CPU tensor (0 = X, 1 = Y, 4 = presence).  Returns (edge_index [2,E] int64, edge_attr [E,4] f32)."""
    pairs, attrs = [], []
    V = x.shape[0]
    xy = x[:, :, :2].numpy()
    present = (x[:, :, 4] > 0.5).numpy()
    for i in range(V):                                        # gbuilder.py:91
        for j in range(V):                                    # :94
            if i == j:
                continue
            d = np.linalg.norm(xy[i] - xy[j], axis=1)         # :99  (float32)
            d = d[present[i] & present[j]]                    # :101-102
            if d.size != 0 and d.min() <= m_radius:           # :103
                pairs.append([i, j])                          # :109
                attrs.append([d.min(), d.max(), d.mean(), (d ** 2).mean()])   # :104-111
    ei = torch.tensor(pairs, dtype=torch.long).t().contiguous() if pairs else torch.empty((2, 0), dtype=torch.long)
    ea = torch.tensor(np.array(attrs, dtype=np.float32)) if attrs else torch.empty((0, 4), dtype=torch.float32)
    return ei, ea
