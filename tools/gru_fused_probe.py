"""Fused GRU (csrc/gru.cu) vs torch's library GRU (ATen native path) at the C2 shape: time and agreement.
usage: python tools/gru_fused_probe.py [N]      (default 205587 sequences = 1024 unit graphs; T=16, I=6, H=96)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sldm_gnn_b200.gru import gru_last_hidden
from sldm_gnn_b200 import _lib

dev = torch.device("cuda:0")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 205_587
T, I, H = 16, 6, 96
torch.manual_seed(0)
gru = torch.nn.GRU(I, H, 1, batch_first=True).to(dev)
x = torch.randn(N, T, I, device=dev)
up = torch.randn(N, H, device=dev)


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out


def lib_fwd():
    with torch.backends.cudnn.flags(enabled=False):
        return gru(x)[1][-1]


def step(fwd):
    gru.zero_grad()
    h = fwd()
    h.backward(up)
    return h.detach(), [p.grad.clone() for p in gru.parameters()]


with torch.no_grad():
    ms_f_lib, _ = timed(lib_fwd)
    ms_f_fused, _ = timed(lambda: gru_last_hidden(gru, x))
ms_lib, (h_lib, g_lib) = timed(lambda: step(lib_fwd), 3)
ms_fused, (h_f, g_f) = timed(lambda: step(lambda: gru_last_hidden(gru, x)), 3)
ms_ftrain, _ = timed(lambda: gru_last_hidden(gru, x))   # training forward (saves the gates)
# backward pieces: the kernel alone (direct C call on preallocated buffers) and the dW_hh GEMM alone
from sldm_gnn_b200.ops import _stream
saved = torch.empty(T, N, 5, H, device=dev); h_last = torch.empty(N, H, device=dev)
Wih, Whh, bih, bhh = gru.weight_ih_l0, gru.weight_hh_l0, gru.bias_ih_l0, gru.bias_hh_l0
_lib.check(_lib.lib.sldm_gru_forward(x.data_ptr(), N, T, I, H, Wih.data_ptr(), Whh.data_ptr(), bih.data_ptr(), bhh.data_ptr(),
                                     h_last.data_ptr(), saved.data_ptr(), _stream(dev)))
dgh = torch.empty(T, N, 3 * H, device=dev)
rows, width = _lib.lib.sldm_gru_partial_rows(N), _lib.lib.sldm_gru_partial_width(H)
parts = torch.empty(rows, width, device=dev)
ms_bk, _ = timed(lambda: _lib.check(_lib.lib.sldm_gru_backward(x.data_ptr(), N, T, I, H, Whh.data_ptr(), up.data_ptr(),
                 saved.data_ptr(), dgh.data_ptr(), None, parts.data_ptr(), rows, _stream(dev))))
ms_mm, _ = timed(lambda: dgh.view(T * N, 3 * H).t() @ saved.view(T * N, 5 * H)[:, :H])
ms_ps, _ = timed(lambda: parts.sum(dim=0))
tiles = _lib.lib.sldm_gru_wgrad_tiles(N, T)
wparts = torch.empty(tiles, 3 * H, H, device=dev)
ms_wg, _ = timed(lambda: _lib.check(_lib.lib.sldm_gru_wgrad(dgh.data_ptr(), saved.data_ptr(), N, T, H, wparts.data_ptr(), tiles, _stream(dev))))
ms_ws, dW = timed(lambda: wparts.sum(dim=0))
ref = dgh.view(T * N, 3 * H).t() @ saved.view(T * N, 5 * H)[:, :H]
print(f"backward pieces: k_gru_bwd {ms_bk:.2f} ms, dW_hh: k_gru_wgrad {ms_wg:.2f} ms + tile sum {ms_ws:.3f} ms (cuBLAS GEMM {ms_mm:.2f} ms, "
      f"max rel diff {float((dW - ref).abs().max() / ref.abs().max()):.2e}), partial sum {ms_ps:.3f} ms")
del saved, dgh
flops = 2.0 * N * T * 3 * H * (H + I)
print(f"N={N} T={T} I={I} H={H}")
print(f"forward (no grad): library {ms_f_lib:.2f} ms, fused {ms_f_fused:.2f} ms ({flops / ms_f_fused / 1e9:.1f} TFLOP/s fp32)")
print(f"forward (training, saves 5 x [N,T,H]): fused {ms_ftrain:.2f} ms")
print(f"forward + backward: library {ms_lib:.2f} ms, fused {ms_fused:.2f} ms")
print("max |h - h_lib|", float((h_f - h_lib).abs().max()),
      "max rel grad diff", max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(g_f, g_lib)))
print("launches so far", _lib.lib.sldm_launch_count(), "peak mem GiB", torch.cuda.max_memory_allocated() / 2**30)
