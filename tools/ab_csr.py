"""A/B of CSR-build variants: bit-exactness against torch's stable sort + device time per build.
usage: ab_csr.py [batch|c4|c1|mid] ...   (env: SLDM_LIB_PATH, SLDM_CSR_DIGIT_BITS)   -> one JSON line per workload"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sldm_gnn_b200 as sg
from workloads import unit_map_graphs, skewed_graph

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def make(kind):
    if kind == "batch":
        ei, _, N = unit_map_graphs(4096, seed=0)
    elif kind == "c1":
        ei, _, N = unit_map_graphs(32, seed=0)
    elif kind == "mid":                      # 24-bit keys, both rows unsorted
        N = 3_000_000
        g = torch.Generator().manual_seed(1)
        ei = torch.randint(0, N, (2, 6_000_000), generator=g)
    else:
        N = 1_000_000
        ei = skewed_graph(N, 10_000_000, seed=0)
    return ei.to(dev), N


for kind in (sys.argv[1:] or ["batch", "c4"]):
    ei, N = make(kind)
    csr = sg.build_csr(ei, N)
    od = torch.sort(ei[1], stable=True).indices
    os_ = torch.sort(ei[0], stable=True).indices
    ok = (torch.equal(csr.col_src.long(), ei[0][od]) and torch.equal(csr.col_dst.long(), ei[1][os_])
          and torch.equal(csr.rowptr_dst.long()[1:] - csr.rowptr_dst.long()[:-1], torch.bincount(ei[1], minlength=N))
          and torch.equal(csr.rowptr_src.long()[1:] - csr.rowptr_src.long()[:-1], torch.bincount(ei[0], minlength=N))
          and int(csr.rowptr_dst[0]) == 0 and int(csr.rowptr_src[0]) == 0)
    for _ in range(3):
        sg.build_csr(ei, N)
    ts = []
    for _ in range(15):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); sg.build_csr(ei, N); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    print(json.dumps({"workload": kind, "N": N, "E": int(ei.size(1)), "bit_exact": ok, "ms_median": round(ts[len(ts) // 2], 4),
                      "ms_min": round(ts[0], 4), "lib": os.environ.get("SLDM_LIB_PATH", "in-tree"),
                      "digit_bits": os.environ.get("SLDM_CSR_DIGIT_BITS", "auto"), "status": csr.status()}), flush=True)
