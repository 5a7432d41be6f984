"""CPU oracle for the model around the block (GruSage, MapEncoder, MapZscoreNorm) -- TEST INFRASTRUCTURE, NOT PRODUCT
CODE.  Only tests/ and bench.py's cpu_baseline / --impl reference legs may import it.

What is restated: src/models/grusage.py:12-195 (constructor tree and forward), src/models/map/mapencoder.py:6-38,
src/models/map/mapInputNorm.py:12-23, with the graph operators taken from the other oracle files (SageBlockOracle,
MapSpatialAttentionOracle, the pool oracles) and torch's own Embedding / GRU / Linear, which is what the reference uses.

PARITY: the composition is PINNED -- tests/golden/grusage/*.pt were produced by the reference's own GruSage /
MapEncoder / MapSpatialAttention / SageBlock classes run unmodified in the authoring container
(tests/golden/make_golden_grusage.py), and tests/test_grusage.py checks this file against them bit for bit.  The one
thing the reference could not supply is torch_geometric itself (not installable here): its SAGEConv and pools were
stood in for by the restatements of oracle/sage_oracle.py, whose own parity stays UNPINNED as that file says.

The reference cannot run with map_included=False (its forward reads self.map_provided, which is only defined when a
map is given, grusage.py:74-76,164); here that case means "no map context", the evident intent.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .map_attention_oracle import MapSpatialAttentionOracle
from .sage_oracle import SageBlockOracle, global_double_pool_oracle, global_max_pool_oracle, global_mean_pool_oracle


def zscore_oracle(f: torch.Tensor) -> torch.Tensor:                       # mapInputNorm.py:12-23
    mu = torch.sum(f, dim=0, keepdim=True) / f.shape[0]
    sigma = torch.sqrt(torch.sum((f - mu) ** 2, dim=0, keepdim=True) / f.shape[0]).clamp(min=1e-8)
    return (f - mu) / sigma


def _act(slope):
    return nn.LeakyReLU(negative_slope=slope) if slope is not None else nn.ReLU()


def _drop(p):
    return nn.Dropout(p=p) if p is not None else nn.Identity()


class MapEncoderOracle(nn.Module):                                        # mapencoder.py:6-38
    def __init__(self, float_features, bool_features, lane_type_cats, edge_index, lane_embed_dim, sage_hidden_dims, dropout, negative_slope):
        super().__init__()
        self.register_buffer("map_float_features", torch.cat([float_features, bool_features.to(dtype=float_features.dtype)], dim=1), persistent=False)
        self.register_buffer("lane_type_cats", lane_type_cats, persistent=False)
        self.register_buffer("graph_edge_indexes", edge_index, persistent=False)
        self.lane_embedding = nn.Embedding(int(torch.max(lane_type_cats).item()) + 1, lane_embed_dim)
        self.sage = SageBlockOracle([self.map_float_features.shape[1] + lane_embed_dim] + sage_hidden_dims, dropout=dropout, negative_slope=negative_slope)
        self.out_dim = sage_hidden_dims[-1]

    def forward(self):
        x = torch.cat([self.map_float_features, self.lane_embedding(self.lane_type_cats)], dim=1)
        return self.sage(x, self.graph_edge_indexes)


class GruSageOracle(nn.Module):
    def __init__(self, dynamic_features_num, frames_num, gru_hidden_size, gru_num_layers, fc1dims, sage_hidden_dims=[128, 128],
                 fc2dims=[50, 50], out_dim=1, num_st_types=256, emb_dim=12, dropout=None, negative_slope=None,
                 global_pooling="double", map_included=True, *, map_tensors=None, mapenc_sage_hdims=[8, 8],
                 mapenc_lane_embdim=2, map_attention_topk=5, map_embeddings=None, map_centroids=None):
        super().__init__()
        self.st_emb = nn.Embedding(num_st_types, emb_dim)                                     # grusage.py:50
        self.gru = nn.GRU(input_size=dynamic_features_num, hidden_size=gru_hidden_size, num_layers=gru_num_layers, batch_first=True)   # :53-58
        d1 = [gru_hidden_size + 2 + emb_dim] + fc1dims                                        # :61-64
        self.fc1s = nn.ModuleList([nn.Sequential(nn.Linear(d1[i], d1[i + 1]), _act(negative_slope), _drop(dropout)) for i in range(len(d1) - 1)])
        last = d1[-1]
        self.map_provided, self.map_tensors = bool(map_included), map_included and map_tensors is not None
        if map_included:                                                                      # :74-105
            if map_tensors is not None:
                self.map_encoder = MapEncoderOracle(zscore_oracle(map_tensors["float_features"]), map_tensors["bool_features"],
                                                    map_tensors["lane_type_cats"], map_tensors["mgraph_edge_indexes"],
                                                    mapenc_lane_embdim, mapenc_sage_hdims, dropout, negative_slope)
                self.map_attention = MapSpatialAttentionOracle(map_tensors["mseg_centroids"], map_attention_topk)
                last += self.map_encoder.out_dim
            else:
                self.register_buffer("map_embeddings", map_embeddings, persistent=False)
                self.map_attention = MapSpatialAttentionOracle(map_centroids, map_attention_topk)
                last += map_embeddings.shape[1]
        self.sage = SageBlockOracle([last] + sage_hidden_dims, dropout=dropout, negative_slope=negative_slope)   # :108-110
        last = sage_hidden_dims[-1]
        self.global_pool = {"mean": global_mean_pool_oracle, "max": global_max_pool_oracle, "double": global_double_pool_oracle}[global_pooling]
        if global_pooling == "double":
            last *= 2                                                                         # :113-122
        d2 = [last] + fc2dims                                                                 # :126-135
        self.fc2s = nn.ModuleList([nn.Sequential(nn.Linear(d2[i], d2[i + 1]), _act(negative_slope), _drop(dropout)) for i in range(len(d2) - 1)])
        self.linout = nn.Linear(d2[-1], out_dim)                                              # :138

    def forward(self, data):                                                                  # :152-195
        x = self.gru(data.x)[1][-1, :, :]
        x = torch.cat([x, data.xdims, self.st_emb(data.xsttype)], dim=1)
        for fc in self.fc1s:
            x = fc(x)
        if self.map_provided:
            emb = self.map_encoder() if self.map_tensors else self.map_embeddings
            x = torch.cat([x, self.map_attention(data.pos_raw[:, -1, :], emb)], dim=1)
        x = self.sage(x, data.edge_index)
        x = self.global_pool(x, data.batch)
        for fc in self.fc2s:
            x = fc(x)
        return self.linout(x)
