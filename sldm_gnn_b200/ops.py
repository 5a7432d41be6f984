"""Host-side wrappers over the C-ABI: torch owns device memory and streams, the
library does the work.  Every function here launches on torch's current stream of
the tensors' device and never synchronises the host.
"""
from __future__ import annotations

import collections
import os
import threading

import torch

from . import _lib
from ._lib import lib, check


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _ptr(t) -> int | None:
    return None if t is None else t.data_ptr()


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"sldm_gnn_b200: `{name}` is on {t.device}; the SageBlock hot path is CUDA only "
            "(sm_100a kernels, no CPU fallback)")


def check_edge_index(edge_index) -> None:
    """Same checks, same exception type as PyG 2.7.0 MessagePassing._check_input."""
    if not isinstance(edge_index, torch.Tensor):
        raise ValueError("`edge_index` must be a torch.Tensor of dtype torch.long and shape [2, num_edges]")
    if edge_index.dtype != torch.long:
        raise ValueError(f"Expected 'edge_index' to be of integer type (got '{edge_index.dtype}')")
    if edge_index.dim() != 2:
        raise ValueError(f"Expected 'edge_index' to be two-dimensional (got {edge_index.dim()} dimensions)")
    if edge_index.size(0) != 2:
        raise ValueError(f"Expected 'edge_index' to have size '2' in the first dimension (got '{edge_index.size(0)}')")


class _IndexChecks:
    """Out-of-range node / graph ids without a host sync on the hot path.

    The reference stack fails on them (CPU: IndexError from index_select / scatter_add_; CUDA: a device-side assert
    that surfaces at a later synchronisation).  The CSR build clamps such ids so that no kernel leaves its buffers and
    raises meta[2]; sldm_csr_status_async copies that word into a slot of a page-locked ring behind the build (the
    slot is preset to -1 on the host) and the slot is looked at once it has landed -- at the latest on the next call
    into this package -- where it raises IndexError: the error is deferred like a CUDA device assert, never silent.
    Cost on the hot path: one 4-byte cudaMemcpyAsync per new edge_index, no event, no synchronisation.
    SLDM_CHECK_INDICES=1 checks synchronously at the build (one host sync per new edge_index), =0 switches it off.
    """

    SLOTS = 256

    def __init__(self):
        self.lock = threading.Lock()
        self.pending = collections.deque()      # (slot index, description), oldest first
        self.ring = None                        # pinned int32 tensor, its numpy view, its address
        self.next = 0

    @staticmethod
    def mode() -> str:
        return os.environ.get("SLDM_CHECK_INDICES", "deferred")

    def watch(self, csr_buf: torch.Tensor, what: str) -> None:
        mode = self.mode()
        if mode == "0":
            return
        if mode == "1":
            if int(csr_buf[2]) != 0:
                raise IndexError(f"{what}: index out of range")
            return
        if torch.cuda.is_current_stream_capturing():
            return
        with self.lock:
            if self.ring is None:
                t = torch.full((self.SLOTS,), -1, dtype=torch.int32).pin_memory()
                self.ring = (t, t.numpy(), t.data_ptr())
            if len(self.pending) >= self.SLOTS - 1:
                self._drain(block=True)
            i = self.next % self.SLOTS
            self.next += 1
            self.ring[1][i] = -1
            check(lib.sldm_csr_status_async(csr_buf.data_ptr(), self.ring[2] + 4 * i, _stream(csr_buf.device)))
            self.pending.append((i, what))

    def _drain(self, block: bool) -> None:
        bad = None
        arr = self.ring[1]
        while self.pending:
            i, what = self.pending[0]
            v = int(arr[i])
            if v < 0:
                if not block:
                    break
                torch.cuda.synchronize()
                v = int(arr[i])
            self.pending.popleft()
            if v > 0 and bad is None:
                bad = what
        if bad is not None:
            raise IndexError(f"{bad}: index out of range (reported by an earlier device-side CSR build; "
                             "SLDM_CHECK_INDICES=1 raises at the call that passed it)")

    def poll(self, block: bool = False) -> None:
        if not self.pending or torch.cuda.is_current_stream_capturing():   # (no host syncs during a capture)
            return
        with self.lock:
            self._drain(block)


index_checks = _IndexChecks()


class Csr:
    """Device CSR object built from one edge_index (layout: include/sldm_sage.h)."""

    __slots__ = ("buf", "N", "E", "layout", "device")

    def __init__(self, buf, N, E, layout):
        self.buf, self.N, self.E, self.layout, self.device = buf, N, E, layout, buf.device

    def _view(self, name, n):
        o = self.layout[name]
        return self.buf[o:o + n]

    @property
    def meta(self): return self._view("meta", 64)
    @property
    def rowptr_dst(self): return self._view("rowptr_dst", self.N + 1)
    @property
    def col_src(self): return self._view("col_src", self.E)
    @property
    def rowptr_src(self): return self._view("rowptr_src", self.N + 1)
    @property
    def col_dst(self): return self._view("col_dst", self.E)

    def status(self) -> dict:
        """Host-synchronising read of the build flags (debug / tests)."""
        m = self.meta[:5].cpu().tolist()
        return {"hub_chunks_dst": m[0], "hub_chunks_src": m[1], "index_out_of_range": bool(m[2]),
                "src_sorted": not m[3], "dst_sorted": not m[4]}


def build_csr(edge_index: torch.Tensor, num_nodes: int) -> Csr:
    check_edge_index(edge_index)
    _require_cuda(edge_index, "edge_index")
    index_checks.poll()
    ei = edge_index if edge_index.is_contiguous() else edge_index.contiguous()
    E = int(ei.size(1))
    N = int(num_nodes)
    dev = ei.device
    layout = _lib.csr_layout(N, E)
    with torch.cuda.device(dev):
        buf = torch.empty(layout["total"], dtype=torch.int32, device=dev)
        wsb = int(lib.sldm_csr_workspace_bytes(N, E))
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev) if E > 0 else None
        check(lib.sldm_csr_build(_ptr(ei) if E > 0 else None, E, N, buf.data_ptr(), _ptr(ws), wsb if E > 0 else 0,
                                 _stream(dev)))
        csr = Csr(buf, N, E, layout)
        if E > 0:
            index_checks.watch(buf, f"edge_index (num_nodes = {N})")
    return csr


def segment_reduce(src: torch.Tensor, csr: Csr, *, transpose: bool = False, mean: bool = True,
                   addend: torch.Tensor | None = None) -> torch.Tensor:
    _require_cuda(src, "src")
    src = src.contiguous()
    N, F = src.shape
    if N != csr.N:
        raise RuntimeError(f"segment_reduce: src has {N} rows but the CSR was built for {csr.N} nodes")
    dev = src.device
    with torch.cuda.device(dev):
        out = torch.empty_like(src)
        wsb = int(lib.sldm_segment_workspace_bytes(N, csr.E, F))
        ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device=dev)
        check(lib.sldm_segment_reduce(src.data_ptr(), N, F, csr.buf.data_ptr(), csr.E, int(transpose), int(mean),
                                      _ptr(addend), out.data_ptr(), ws.data_ptr(), wsb, _stream(dev)))
    return out


def segment_mean_bf16(src: torch.Tensor, csr: Csr) -> torch.Tensor:
    """Forward segment mean over bfloat16 rows (fp32 accumulation, one rounding): the aggregation of the bf16 feature mode."""
    _require_cuda(src, "src")
    src = src.contiguous()
    N, F = src.shape
    dev = src.device
    with torch.cuda.device(dev):
        out = torch.empty_like(src)
        wsb = int(lib.sldm_segment_workspace_bytes(N, csr.E, F))
        ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device=dev)
        check(lib.sldm_segment_mean_bf16(src.data_ptr(), N, F, csr.buf.data_ptr(), csr.E, out.data_ptr(), ws.data_ptr(),
                                         wsb, _stream(dev)))
    return out


def project_forward(agg, x, W_l, b_l, W_r, ln_w, ln_b, eps, slope, save: bool):
    N, Fin = x.shape
    Fout = W_l.shape[0]
    dev = x.device
    bf16 = x.dtype == torch.bfloat16
    with torch.cuda.device(dev):
        if bf16:
            out = torch.empty((N, Fout), dtype=torch.bfloat16, device=dev)
            xhat = torch.empty((N, Fout), dtype=torch.float32, device=dev) if save else None
            rstd = torch.empty((N,), dtype=torch.float32, device=dev) if save else None
            wsb = int(lib.sldm_sage_project_workspace_bytes(N, Fin, Fout))
            ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device=dev)
            check(lib.sldm_sage_project_forward_bf16(agg.data_ptr(), x.data_ptr(), N, Fin, Fout, W_l.data_ptr(),
                                                     b_l.data_ptr(), W_r.data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(),
                                                     float(eps), float(slope), out.data_ptr(), _ptr(xhat), _ptr(rstd),
                                                     ws.data_ptr(), wsb, _stream(dev)))
            return out, xhat, rstd
        out = torch.empty((N, Fout), dtype=torch.float32, device=dev)
        xhat = torch.empty((N, Fout), dtype=torch.float32, device=dev) if save else None
        rstd = torch.empty((N,), dtype=torch.float32, device=dev) if save else None
        wsb = int(lib.sldm_sage_project_workspace_bytes(N, Fin, Fout))
        ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device=dev)
        check(lib.sldm_sage_project_forward(agg.data_ptr(), x.data_ptr(), N, Fin, Fout, W_l.data_ptr(), b_l.data_ptr(),
                                            W_r.data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(), float(eps), float(slope),
                                            out.data_ptr(), _ptr(xhat), _ptr(rstd), ws.data_ptr(), wsb, _stream(dev)))
    return out, xhat, rstd


def bf16_supported(Fin: int, Fout: int) -> bool:
    """Layer shapes the bf16 feature-storage kernels cover (include/sldm_sage.h)."""
    return bool(lib.sldm_sage_bf16_supported(int(Fin), int(Fout)))


def layer_forward(x, csr: Csr, W_l, b_l, W_r, ln_w, ln_b, eps: float, slope: float, save: bool):
    """One SageBlock layer.  Returns (x_used, out, agg, xhat, rstd); xhat/rstd are None unless `save`.

    bf16 feature storage: a bfloat16 `x` runs the bf16 kernels (out and agg come back as bfloat16, xhat / rstd stay
    float32) when the layer shape is covered; otherwise it is converted to float32 for this layer (x_used is then the
    float32 copy -- the tensor the backward has to be given) and only `out` is rounded to bfloat16."""
    N, Fin = x.shape
    Fout = W_l.shape[0]
    dev = x.device
    if x.dtype == torch.bfloat16:
        if not bf16_supported(Fin, Fout):
            x32, out, agg, xhat, rstd = layer_forward(x.float(), csr, W_l, b_l, W_r, ln_w, ln_b, eps, slope, save)
            return x32, out.to(torch.bfloat16), agg, xhat, rstd
        with torch.cuda.device(dev):
            out = torch.empty((N, Fout), dtype=torch.bfloat16, device=dev)
            agg = torch.empty((N, Fin), dtype=torch.bfloat16, device=dev)
            xhat = torch.empty((N, Fout), dtype=torch.float32, device=dev) if save else None
            rstd = torch.empty((N,), dtype=torch.float32, device=dev) if save else None
            wsb = int(lib.sldm_sage_layer_fwd_workspace_bytes(N, csr.E, Fin, Fout))
            ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device=dev)
            check(lib.sldm_sage_layer_forward_bf16(x.data_ptr(), N, Fin, Fout, csr.buf.data_ptr(), csr.E,
                                                   W_l.data_ptr(), b_l.data_ptr(), W_r.data_ptr(), ln_w.data_ptr(),
                                                   ln_b.data_ptr(), float(eps), float(slope),
                                                   out.data_ptr(), agg.data_ptr(), _ptr(xhat), _ptr(rstd),
                                                   ws.data_ptr(), wsb, _stream(dev)))
        return x, out, agg, xhat, rstd
    with torch.cuda.device(dev):
        out = torch.empty((N, Fout), dtype=torch.float32, device=dev)
        agg = torch.empty((N, Fin), dtype=torch.float32, device=dev)
        xhat = torch.empty((N, Fout), dtype=torch.float32, device=dev) if save else None
        rstd = torch.empty((N,), dtype=torch.float32, device=dev) if save else None
        wsb = int(lib.sldm_sage_layer_fwd_workspace_bytes(N, csr.E, Fin, Fout))
        ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device=dev)
        check(lib.sldm_sage_layer_forward(x.data_ptr(), N, Fin, Fout, csr.buf.data_ptr(), csr.E,
                                          W_l.data_ptr(), b_l.data_ptr(), W_r.data_ptr(), ln_w.data_ptr(),
                                          ln_b.data_ptr(), float(eps), float(slope),
                                          out.data_ptr(), agg.data_ptr(), _ptr(xhat), _ptr(rstd),
                                          ws.data_ptr(), wsb, _stream(dev)))
    return x, out, agg, xhat, rstd


def backward_buffers(N: int, Fin: int, Fout: int, E: int, dev, need_dx: bool, grad_out=None) -> dict:
    """Outputs, scratch and workspace of one layer_backward call.  grad_out: five preallocated fp32 tensors
    (dW_l, db_l, dW_r, dln_w, dln_b) the kernels write the parameter gradients into (a data-parallel bucket)."""
    with torch.cuda.device(dev):
        f32 = dict(dtype=torch.float32, device=dev)
        if grad_out is not None:
            b = dict(dW_l=grad_out[0], db_l=grad_out[1], dW_r=grad_out[2], dln_w=grad_out[3], dln_b=grad_out[4])
        else:
            b = dict(dW_l=torch.empty((Fout, Fin), **f32), dW_r=torch.empty((Fout, Fin), **f32),
                     db_l=torch.empty((Fout,), **f32), dln_w=torch.empty((Fout,), **f32), dln_b=torch.empty((Fout,), **f32))
        b.update(dz=torch.empty((N, Fout), **f32), dx=None, dagg=None, dxroot=None)
        if need_dx:
            b.update(dx=torch.empty((N, Fin), **f32), dagg=torch.empty((N, Fin), **f32),
                     dxroot=torch.empty((N, Fin), **f32))
        wsb = int(lib.sldm_sage_layer_bwd_workspace_bytes(N, E, Fin, Fout))
        b["wsb"] = wsb
        b["ws"] = torch.empty(max(wsb, 1), dtype=torch.uint8, device=dev)
    return b


def layer_backward(dout, x, agg, xhat, rstd, csr: Csr, W_l, W_r, ln_w, ln_b, slope: float, need_dx: bool,
                   stages: int = _lib.BWD_STAGE_ALL, bufs: dict | None = None, on_param_grads=None, grad_out=None):
    """Returns (dx | None, dW_l, db_l, dW_r, dln_w, dln_b).

    `stages` / `bufs` are for profiling (bench.py): launch only the masked kernels on the buffers of an earlier
    full call (include/sldm_sage.h, SLDM_BWD_STAGE_*).  grad_out: write the parameter gradients into these five
    tensors; on_param_grads(event): called with an event recorded when the parameter gradients are final on the
    stream, BEFORE the dx gather is launched (a data-parallel wrapper starts the layer's exchange there)."""
    N, Fin = x.shape
    Fout = W_l.shape[0]
    dev = x.device
    dout = dout.contiguous()
    b = bufs if bufs is not None else backward_buffers(N, Fin, Fout, csr.E, dev, need_dx, grad_out)

    entry = lib.sldm_sage_layer_backward_bf16 if x.dtype == torch.bfloat16 else lib.sldm_sage_layer_backward_stages
    if x.dtype != agg.dtype or dout.dtype != torch.float32:
        raise RuntimeError(f"layer_backward: x is {x.dtype}, agg is {agg.dtype}, dout is {dout.dtype}")

    def launch(mask):
        check(entry(
            dout.data_ptr(), x.data_ptr(), agg.data_ptr(), xhat.data_ptr(), rstd.data_ptr(), N, Fin, Fout,
            csr.buf.data_ptr(), csr.E, W_l.data_ptr(), W_r.data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(), float(slope),
            _ptr(b["dx"]), b["dW_l"].data_ptr(), b["db_l"].data_ptr(), b["dW_r"].data_ptr(), b["dln_w"].data_ptr(),
            b["dln_b"].data_ptr(), b["dz"].data_ptr(), _ptr(b["dagg"]), _ptr(b["dxroot"]), b["ws"].data_ptr(), b["wsb"],
            _stream(dev), int(mask)))

    with torch.cuda.device(dev):
        if on_param_grads is not None and stages == _lib.BWD_STAGE_ALL:
            # the parameter gradients are final before the dx gather starts: mark that point on the stream
            launch(_lib.BWD_STAGE_ALL & ~_lib.BWD_STAGE_GATHER)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            on_param_grads(ev)
            if need_dx and N > 0:
                launch(_lib.BWD_STAGE_GATHER)
        else:
            launch(stages)
    return b["dx"], b["dW_l"], b["db_l"], b["dW_r"], b["dln_w"], b["dln_b"]
