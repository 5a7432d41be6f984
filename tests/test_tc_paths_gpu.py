"""Both arithmetic paths of the projection / backward GEMMs are covered: the tcgen05 3xTF32 kernels
(default when the widths allow) and the FP32 SIMT kernels (SLDM_DISABLE_TC=1, read once per process,
hence the subprocess)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys, torch
sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests')
import sldm_gnn_b200 as sg
from workloads import unit_map_graphs
from test_gpu_parity import run_pair, check_pair
dev = torch.device('cuda:0')
for hdims, slope in (([128, 128, 128], 0.1), ([64, 96, 32], None), ([32, 16, 48], 0.2), ([96, 128], 0.1)):
    ei, _, N = unit_map_graphs(5, seed=hdims[0])
    check_pair(*run_pair(dev, hdims, slope, ei, N))
# tile tails: N not a multiple of 128 / 32, tiny N
for N, E in ((1, 0), (33, 100), (129, 700), (4097, 30000)):
    ei = torch.randint(0, N, (2, E), generator=torch.Generator().manual_seed(N))
    check_pair(*run_pair(dev, [64, 64], 0.1, ei, N))
print('PATH_OK')
""" % (ROOT, ROOT)


@pytest.mark.parametrize("disable_tc", ["0", "1"], ids=["tcgen05", "simt"])
def test_both_gemm_paths_match_oracle(disable_tc):
    env = dict(os.environ, SLDM_DISABLE_TC=disable_tc)
    r = subprocess.run([sys.executable, "-c", SCRIPT], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "PATH_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_tc_kernels_are_in_the_library():
    """SASS evidence: the shipped .so contains tcgen05 MMA / TMEM / TMA instructions."""
    from sldm_gnn_b200 import _lib
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "STTM", "UTMALDG"):      # (the epilogue stores are plain full-row STG since round 2)
        assert mnemonic in sass, mnemonic


GATHER_SCRIPT = r"""
import hashlib, torch
import sldm_gnn_b200 as sg
from workloads import unit_map_graphs, skewed_graph
dev = torch.device("cuda:0")
out = []
for kind, F in (("batch", 128), ("batch", 96), ("batch", 64), ("batch", 32), ("batch", 16), ("skew", 128), ("skew", 40)):
    if kind == "batch":
        ei, _, N = unit_map_graphs(64, seed=3)
    else:
        N = 20000
        ei = skewed_graph(N, 300000, seed=3)
    ei = ei.to(dev)
    g = torch.Generator(device=dev).manual_seed(5)
    x = torch.randn(N, F, device=dev, generator=g)
    y = torch.randn(N, F, device=dev, generator=g)
    csr = sg.build_csr(ei, N)
    a = sg.segment_reduce(x, csr)
    b = sg.segment_reduce(y, csr, transpose=True, mean=False, addend=x)
    out.append(hashlib.sha1(a.cpu().numpy().tobytes() + b.cpu().numpy().tobytes()).hexdigest())
print("HASHES " + " ".join(out))
"""


def test_lean_and_generic_gather_are_bit_identical():
    """The lean gather (per-warp column window, exact-count load pieces, exact power-of-two mean) must produce the
    same bytes as the generic kernel: same values, same order of additions (SLDM_SEG_LEAN=0 selects the generic one)."""
    res = []
    for lean in ("0", "4", "8"):
        env = dict(os.environ, SLDM_SEG_LEAN=lean)
        r = subprocess.run([sys.executable, "-c", GATHER_SCRIPT], env=env, capture_output=True, text=True, timeout=600,
                           cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        assert r.returncode == 0 and "HASHES" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
        res.append(r.stdout.strip().splitlines()[-1])
    assert res[0] == res[1] == res[2], res
