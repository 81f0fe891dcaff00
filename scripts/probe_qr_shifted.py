"""Times qr(mode="r") of one ill-conditioned 2 097 152 x 128 block: iterated shifted Cholesky passes (default) against the
Householder kernel (NUMS_QR_SHIFTED=0).  Development aid, run under gpurun."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nums_b200 import cuda_compute as cc  # noqa: E402
from nums_b200.cuda_system import CudaSystem  # noqa: E402

system = CudaSystem()
system.init()
m, n = 2_097_152, 128
for kappa in (1e3, 1e9, 1e13):
    # graded spectrum built on the device: orthonormalise a Gaussian block with our own QR, scale, rotate
    g = torch.randn((m, n), dtype=torch.float64, device="cuda")
    r = cc.qr_r(g)
    q = torch.empty_like(g)
    cc.gemm_into(q, g, False, n, cc._inv_nocheck(r), False, n, m, n, n)
    sv = torch.logspace(0, -float(np.log10(kappa)), n, dtype=torch.float64, device="cuda")
    v, _ = torch.linalg.qr(torch.randn((n, n), dtype=torch.float64, device="cuda"))
    x = (q * sv) @ v.T
    del g, q
    before = dict(cc.QR_STATS)
    times = []
    for it in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rr = cc.qr_r(x)
        b.record()
        b.synchronize()
        if it:
            times.append(a.elapsed_time(b))
    gram = cc._gram_of(x)
    rtr = torch.empty((n, n), dtype=torch.float64, device="cuda")
    cc.gemm_into(rtr, rr, True, n, rr, False, n, n, n, n)
    err = float((torch.linalg.norm(rtr - gram) / torch.linalg.norm(gram)).cpu())
    ran = {k: cc.QR_STATS[k] - before[k] for k in before if cc.QR_STATS[k] != before[k]}
    print(json.dumps({"kappa": kappa, "ms": float(np.median(times)), "branches": ran, "gram_rel_err": err}))
    del x
