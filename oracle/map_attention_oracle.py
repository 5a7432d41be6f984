"""CPU oracle for MapSpatialAttention -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/sage_oracle.py).

PARITY PINNED: unlike the SageBlock path, this component of the reference is plain torch
(src/models/map/mapattention.py:5-56), imports nothing but torch, and therefore runs in this container.  The
restatement below follows it line by line; tests/test_map_attention.py checks it against the imported reference class
when /root/reference is present, and against golden vectors the reference itself produced
(tests/golden/map_attention/*.pt, made by tests/golden/make_golden_map_attention.py).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class MapSpatialAttentionOracle(nn.Module):
    def __init__(self, map_centroids: torch.Tensor, k_neighbors=5):          # mapattention.py:6-19
        super().__init__()
        self.register_buffer("map_centroids", map_centroids, persistent=False)
        self.k = k_neighbors
        self.attn_mlp = nn.Sequential(nn.Linear(1, 16), nn.ReLU(), nn.Linear(16, 1))

    def forward(self, vehicle_last_positions, map_embeddings):               # mapattention.py:21-56
        diff = vehicle_last_positions.unsqueeze(1) - self.map_centroids.unsqueeze(0)     # :33
        dists = torch.norm(diff, dim=2)                                                   # :34
        neg_dists, indices = torch.topk(-dists, k=self.k, dim=1)                          # :39
        k_dists = -neg_dists                                                              # :40
        batch_map_embeds = map_embeddings[indices, :]                                     # :45
        attn_scores = self.attn_mlp(k_dists.unsqueeze(2)).squeeze(2)                      # :50
        weights = F.softmax(attn_scores, dim=1).unsqueeze(2)                              # :51
        return torch.sum(batch_map_embeds * weights, dim=1)                               # :55

    def neighbours(self, vehicle_last_positions):
        """(distances, indices) of the K nearest segments plus the (K+1)-th distance (tests use it to skip near ties)."""
        d = torch.norm(vehicle_last_positions.unsqueeze(1) - self.map_centroids.unsqueeze(0), dim=2)
        kk = min(self.k + 1, d.size(1))
        nd, idx = torch.topk(-d, k=kk, dim=1)
        return -nd, idx
