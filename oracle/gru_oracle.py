"""CPU oracle for the sequence head (single-layer GRU, last hidden state) -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
Only tests/ may import it.

What is restated: `gru_out, hlast = self.gru(x); x = hlast[-1,:,:]` (src/models/grusage.py:160-161) for the layer built
at src/models/grusage.py:55-60 (`nn.GRU(input_size, hidden_size, num_layers=1, batch_first=True)`, h0 = 0), as plain
numpy loops over time, forward AND backward, with the formulas of ATen's fused cell (aten/src/ATen/native/cuda/RNN.cu
gru_cell_forward / gru_cell_backward) -- the same decomposition csrc/gru.cu implements:

    r = sigmoid(W_ir x + b_ir + W_hr h + b_hr)        z = sigmoid(W_iz x + b_iz + W_hz h + b_hz)
    hn = W_hn h + b_hn                                n = tanh(W_in x + b_in + r * hn)
    h' = n + z * (h - n)
    backward, given dh' :  dz~ = dh' (h - n)(1 - z) z ;  dn~ = dh' (1 - z)(1 - n^2) ;  dhn = dn~ r ;
                           dr~ = dn~ hn (1 - r) r ;  dh = dh' z + [dr~ dz~ dhn] W_hh
                           dW_ih += [dr~ dz~ dn~]^T x ;  dW_hh += [dr~ dz~ dhn]^T h ;  db_ih += [dr~ dz~ dn~] ;
                           db_hh += [dr~ dz~ dhn] ;  dx = [dr~ dz~ dn~] W_ih

PARITY: PINNED.  The reference's arithmetic for this step is torch.nn.GRU itself, which runs in this image: the vectors
in tests/golden/gru/*.pt come from it (tests/golden/make_golden_gru.py) and tests/test_gru.py checks this file against
them in fp64 (forward 1e-12, every gradient 1e-10 relative).
"""
from __future__ import annotations

import numpy as np


def _sigmoid(v):
    return 1.0 / (1.0 + np.exp(-v))


def gru_last_hidden_oracle(x, W_ih, W_hh, b_ih, b_hh):
    """x [N,T,I]; weights in torch's layout (gates r, z, n stacked on the rows).  Returns (h_last [N,H], tape)."""
    x = np.asarray(x)
    N, T, _ = x.shape
    H = W_hh.shape[1]
    h = np.zeros((N, H), dtype=x.dtype)
    tape = []
    for t in range(T):
        gi = x[:, t, :] @ W_ih.T + b_ih                    # [N,3H]
        gh = h @ W_hh.T + b_hh
        r = _sigmoid(gi[:, :H] + gh[:, :H])
        z = _sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
        hn = gh[:, 2 * H:]
        n = np.tanh(gi[:, 2 * H:] + r * hn)
        tape.append((h, r, z, n, hn))
        h = n + z * (h - n)
    return h, tape


def gru_last_hidden_backward_oracle(dh_last, x, W_ih, W_hh, tape):
    """Returns (dx [N,T,I], dW_ih, dW_hh, db_ih, db_hh) for the upstream gradient dh_last [N,H]."""
    x = np.asarray(x)
    N, T, I = x.shape
    H = W_hh.shape[1]
    dh = np.asarray(dh_last).copy()
    dx = np.zeros_like(x)
    dW_ih, dW_hh = np.zeros_like(W_ih), np.zeros_like(W_hh)
    db_ih, db_hh = np.zeros(3 * H, dtype=x.dtype), np.zeros(3 * H, dtype=x.dtype)
    for t in range(T - 1, -1, -1):
        h_prev, r, z, n, hn = tape[t]
        dz = dh * (h_prev - n) * (1.0 - z) * z
        dn = dh * (1.0 - z) * (1.0 - n * n)
        dhn = dn * r
        dr = dn * hn * (1.0 - r) * r
        dgi = np.concatenate([dr, dz, dn], axis=1)
        dgh = np.concatenate([dr, dz, dhn], axis=1)
        dW_ih += dgi.T @ x[:, t, :]
        dW_hh += dgh.T @ h_prev
        db_ih += dgi.sum(axis=0)
        db_hh += dgh.sum(axis=0)
        dx[:, t, :] = dgi @ W_ih
        dh = dh * z + dgh @ W_hh
    return dx, dW_ih, dW_hh, db_ih, db_hh
