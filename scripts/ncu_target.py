"""Tiny driver for ncu captures: launches each hot kernel a few times (development aid)."""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nums_b200 import cuda_compute as cc
from nums_b200.cuda_system import CudaSystem

which = sys.argv[1] if len(sys.argv) > 1 else "gemm"
system = CudaSystem(); system.init()
dev = torch.device("cuda", 0)
if which == "gemm":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    A = torch.randn((n, n), dtype=torch.float64, device=dev)
    B = torch.randn((n, n), dtype=torch.float64, device=dev)
    for _ in range(4):
        C = system.bop("tensordot", A, B, (n, n), (n, n), False, False, axes=1, syskwargs={})
elif which == "grouped":
    from nums_b200 import blocks
    n, bs = 8192, 2048
    app = blocks.ArrayApp(system)
    A = blocks.BlockArray(blocks.ArrayGrid((n, n), (bs, bs), "float64"), system)
    B = blocks.BlockArray(blocks.ArrayGrid((n, n), (bs, bs), "float64"), system)
    for e in A.grid.get_entry_iterator():
        A.blocks[e].oid = torch.randn((bs, bs), dtype=torch.float64, device=dev)
        B.blocks[e].oid = torch.randn((bs, bs), dtype=torch.float64, device=dev)
    for _ in range(3):
        (A @ B).touch()
elif which == "bop":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000_000
    u = torch.rand(n, dtype=torch.float64, device=dev); v = torch.rand(n, dtype=torch.float64, device=dev)
    for _ in range(4):
        w = system.bop("add", u, v, (n,), (n,), False, False, axes=None, syskwargs={})
elif which == "lr":
    n, d = 11_000_000, 28
    X = torch.randn((n, d), dtype=torch.float64, device=dev)
    y = (torch.rand(n, device=dev) < 0.5).to(torch.float64)
    beta = torch.randn(d, dtype=torch.float64, device=dev) / 5
    for _ in range(4):
        out = cc.lr_grad_hess(X, y, beta)
elif which == "qrh":          # the Householder leaf (blocked WY / DMMA kernel) on one config-3 block
    X = torch.randn((2_097_152, 128), dtype=torch.float64, device=dev)
    for _ in range(2):
        r = cc._householder_r(X)
elif which == "rowcol":       # (n, 1) * (n, 28): the s * X of the LR Hessian (glms.py:236) on one config-4 block
    n, d = 1_375_000, 28
    X = torch.randn((n, d), dtype=torch.float64, device=dev)
    col = torch.randn((n, 1), dtype=torch.float64, device=dev)
    for _ in range(4):
        w = system.bop("mul", col, X, (n, 1), (n, d), False, False, axes=None, syskwargs={})
elif which == "skinny":       # X^T (s X) of the LR Hessian (glms.py:232-238) on one config-4 block: 28 x n_b by n_b x 28
    n, d = 1_375_000, 28
    X = torch.randn((n, d), dtype=torch.float64, device=dev)
    S = torch.randn((n, d), dtype=torch.float64, device=dev)
    for _ in range(4):
        h = system.bop("tensordot", X, S, (d, n), (n, d), True, False, axes=1, syskwargs={})
elif which == "gemvn":        # X beta (glms.py:140-143) on one config-4 block
    n, d = 1_375_000, 28
    X = torch.randn((n, d), dtype=torch.float64, device=dev)
    b = torch.randn((d,), dtype=torch.float64, device=dev)
    for _ in range(4):
        z = system.bop("tensordot", X, b, (n, d), (d,), False, False, axes=1, syskwargs={})
elif which == "syrk":         # Gram matrix of one config-3 block (the TSQR leaf on the Gram path)
    X = torch.randn((2_097_152, 128), dtype=torch.float64, device=dev)
    for _ in range(4):
        gmat = cc._gram_of(X)
elif which == "tall":         # Q = X R^-1 of one config-3 block
    X = torch.randn((2_097_152, 128), dtype=torch.float64, device=dev)
    Rinv = torch.randn((128, 128), dtype=torch.float64, device=dev)
    for _ in range(4):
        qmat = cc.tensordot(X, Rinv, 1)
elif which == "qr":
    X = torch.randn((262144, 128), dtype=torch.float64, device=dev)
    for _ in range(2):
        r = cc.qr_r(X)
torch.cuda.synchronize()
print("done", which)
