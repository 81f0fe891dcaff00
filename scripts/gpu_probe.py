"""Micro-benchmarks of the individual kernels on one B200 (development aid, not bench.py).

Writes gpurun_out/probe.json.  Every figure is CUDA-event timed on the launching stream after
warm-up; inputs are larger than L2 or L2 is flushed between iterations.
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nums_b200 import cuda_compute as cc  # noqa: E402
from nums_b200.cuda_system import CudaSystem  # noqa: E402

OUT = {}
dev = torch.device("cuda", 0)
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def flush_l2():
    flush_buf.zero_()


def timeit(fn, iters=10, warmup=3, flush=True):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(iters):
        if flush:
            flush_l2()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        times.append(s.elapsed_time(e) * 1e-3)
    times.sort()
    return times[len(times) // 2], times[0]


def record(name, **kw):
    OUT[name] = kw
    print(name, json.dumps(kw), flush=True)


def main():
    system = CudaSystem()
    system.init()
    system.contractions.enabled = False   # time the individual kernels, not the deferred chains
    which = set(sys.argv[1:]) or {"copy", "bop", "gemm", "reduce", "lr", "qr", "gemv"}
    props = torch.cuda.get_device_properties(0)
    record("device", gpu=props.name, sms=props.multi_processor_count, mem_gb=props.total_memory / 2 ** 30)

    if "copy" in which:
        a = torch.empty(1 << 28, dtype=torch.float64, device=dev)  # 2 GiB
        b = torch.empty_like(a)
        med, best = timeit(lambda: b.copy_(a), flush=False)
        record("torch_copy_2GiB", gbs_med=2 * a.numel() * 8 / med / 1e9, gbs_best=2 * a.numel() * 8 / best / 1e9)
        del a, b

    if "bop" in which:
        for n in (12_500_000, 100_000_000):
            u = torch.rand(n, dtype=torch.float64, device=dev)
            v = torch.rand(n, dtype=torch.float64, device=dev)
            for op in ("add", "mul"):
                fn = lambda: system.bop(op, u, v, (n,), (n,), False, False, axes=None, syskwargs={})
                med, best = timeit(fn)
                record("bop_%s_%d" % (op, n), ms=med * 1e3, gbs_med=24 * n / med / 1e9, gbs_best=24 * n / best / 1e9)
            med, best = timeit(lambda: torch.add(u, v))
            record("torch_add_%d" % n, ms=med * 1e3, gbs_med=24 * n / med / 1e9, gbs_best=24 * n / best / 1e9)
            one = system.put(np.array(1.0, dtype=np.float32))
            med, best = timeit(lambda: system.bop("add", u, one, (n,), (), False, False, axes=None, syskwargs={}))
            record("bop_add_scalar_%d" % n, ms=med * 1e3, gbs_med=16 * n / med / 1e9)
            med, best = timeit(lambda: system.map_uop("exp", u, (), {}, syskwargs={}))
            record("uop_exp_%d" % n, ms=med * 1e3, gbs_med=16 * n / med / 1e9)
            del u, v
        n, d = 1_375_000, 28
        X = torch.rand((n, d), dtype=torch.float64, device=dev)
        s = torch.rand((n, 1), dtype=torch.float64, device=dev)
        med, best = timeit(lambda: system.bop("mul", s, X, (n, 1), (n, d), False, False, axes=None, syskwargs={}))
        record("bop_colbcast_%dx%d" % (n, d), ms=med * 1e3, gbs_med=(16 * n * d + 8 * n) / med / 1e9)
        del X, s

    if "reduce" in which:
        n = 100_000_000
        u = torch.rand(n, dtype=torch.float64, device=dev)
        med, best = timeit(lambda: system.reduce_axis("sum", u, None, False, False, syskwargs={}))
        record("reduce_sum_all_%d" % n, ms=med * 1e3, gbs_med=8 * n / med / 1e9)
        med, best = timeit(lambda: system.reduce_axis("max", u, None, False, False, syskwargs={}))
        record("reduce_max_all_%d" % n, ms=med * 1e3, gbs_med=8 * n / med / 1e9)
        X = u[:1_375_000 * 28].view(1_375_000, 28)
        for axis in (0, 1):
            med, best = timeit(lambda: system.reduce_axis("sum", X, axis, False, False, syskwargs={}))
            record("reduce_sum_axis%d_1375000x28" % axis, ms=med * 1e3, gbs_med=8 * X.numel() / med / 1e9)
        del u, X

    if "gemm" in which:
        for n in (2048, 4096, 8192):
            A = torch.randn((n, n), dtype=torch.float64, device=dev)
            B = torch.randn((n, n), dtype=torch.float64, device=dev)
            iters = 10 if n <= 4096 else 4
            med, best = timeit(lambda: torch.matmul(A, B), iters=iters, warmup=2)
            record("cublas_dgemm_%d" % n, ms=med * 1e3, tflops_med=2 * n ** 3 / med / 1e12, tflops_best=2 * n ** 3 / best / 1e12)
            for ta, tb in ((False, False), (True, False), (False, True), (True, True)):
                fn = lambda: system.bop("tensordot", A, B, (n, n), (n, n), ta, tb, axes=1, syskwargs={})
                med, best = timeit(fn, iters=iters, warmup=2)
                record("nums_dgemm_%d_%s%s" % (n, "T" if ta else "N", "T" if tb else "N"), ms=med * 1e3,
                       tflops_med=2 * n ** 3 / med / 1e12, tflops_best=2 * n ** 3 / best / 1e12)
            C = system.bop("tensordot", A, B, (n, n), (n, n), False, False, axes=1, syskwargs={})
            ref = torch.matmul(A, B)
            record("nums_dgemm_%d_err" % n, rel_fro=float((C - ref).norm() / ref.norm()))
            del A, B, C, ref
        # tall-skinny shapes of TSQR's Q = X R^-1 and the Gram matrix
        m, n = 2_097_152, 128
        X = torch.randn((m, n), dtype=torch.float64, device=dev)
        Rm = torch.randn((n, n), dtype=torch.float64, device=dev)
        med, best = timeit(lambda: system.bop("tensordot", X, Rm, (m, n), (n, n), False, False, axes=1, syskwargs={}), iters=5)
        record("nums_dgemm_tallskinny_XR", ms=med * 1e3, tflops_med=2 * m * n * n / med / 1e12, gbs=16 * m * n / med / 1e9)
        med, best = timeit(lambda: torch.matmul(X, Rm), iters=5)
        record("cublas_dgemm_tallskinny_XR", ms=med * 1e3, tflops_med=2 * m * n * n / med / 1e12)
        med, best = timeit(lambda: system.bop("tensordot", X, X, (n, m), (m, n), True, False, axes=1, syskwargs={}), iters=5)
        record("nums_dgemm_gram_XtX", ms=med * 1e3, tflops_med=2 * m * n * n / med / 1e12, gbs=8 * m * n / med / 1e9)
        med, best = timeit(lambda: torch.matmul(X.T, X), iters=5)
        record("cublas_dgemm_gram_XtX", ms=med * 1e3, tflops_med=2 * m * n * n / med / 1e12)
        del X, Rm

    if "gemv" in which:
        n, d = 11_000_000, 28
        X = torch.randn((n, d), dtype=torch.float64, device=dev)
        beta = torch.randn(d, dtype=torch.float64, device=dev)
        w = torch.randn(n, dtype=torch.float64, device=dev)
        med, best = timeit(lambda: system.bop("tensordot", X, beta, (n, d), (d,), False, False, axes=1, syskwargs={}))
        record("gemv_Xb_11Mx28", ms=med * 1e3, gbs_med=8 * n * (d + 1) / med / 1e9)
        med, best = timeit(lambda: system.bop("tensordot", X, w, (d, n), (n,), True, False, axes=1, syskwargs={}))
        record("gemv_Xtw_11Mx28", ms=med * 1e3, gbs_med=8 * n * (d + 1) / med / 1e9)
        med, best = timeit(lambda: system.bop("tensordot", X, X, (d, n), (n, d), True, False, axes=1, syskwargs={}), iters=5)
        record("gemm_XtX_11Mx28", ms=med * 1e3, gbs_med=8 * n * d / med / 1e9)
        del X, beta, w

    if "lr" in which:
        for n in (1_375_000, 11_000_000):
            d = 28
            X = torch.randn((n, d), dtype=torch.float64, device=dev)
            y = (torch.rand(n, device=dev) < 0.5).to(torch.float64)
            beta = torch.randn(d, dtype=torch.float64, device=dev) / 5
            med, best = timeit(lambda: cc.lr_grad_hess(X, y, beta))
            record("lr_fused_%dx%d" % (n, d), ms=med * 1e3, gbs_med=8 * n * (d + 1) / med / 1e9,
                   gbs_best=8 * n * (d + 1) / best / 1e9)
            del X, y

    if "qr" in which:
        for (m, n) in ((2_097_152, 128), (2_097_152, 32), (262_144, 128)):
            X = torch.randn((m, n), dtype=torch.float64, device=dev)
            t0 = time.time()
            med, best = timeit(lambda: cc.qr_r(X), iters=3, warmup=1)
            flops = 2 * m * n * n - 2 * n ** 3 / 3
            record("qr_r_%dx%d" % (m, n), ms=med * 1e3, tflops_med=flops / med / 1e12, wall_s=time.time() - t0)
            del X

    if "qrparts" in which:       # where one Gram-path qr_r of a 2M x 128 block spends its time
        from nums_b200 import _lib
        LIB = _lib.LIB
        m, n = 2_097_152, 128
        X = torch.randn((m, n), dtype=torch.float64, device=dev)
        gram = torch.empty((n, n), dtype=torch.float64, device=dev)
        med, _ = timeit(lambda: cc.gemm_into(gram, X, True, n, X, False, n, n, n, m), iters=5, warmup=2)
        record("qrparts_gram_gemm", ms=med * 1e3, tflops=2.0 * m * n * n / med / 1e12)
        low = torch.empty((n, n), dtype=torch.float64, device=dev)
        info = torch.empty((), dtype=torch.int32, device=dev)
        med, _ = timeit(lambda: LIB.call_ws(LIB.dll.nums_cholesky, X.device, ((_lib.F64, n, gram.data_ptr(), n, low.data_ptr(), n,
                                                                                 info.data_ptr()), (cc._stream(),))), flush=False)
        record("qrparts_cholesky_128", us=med * 1e6)
        med, _ = timeit(lambda: cc._inv_nocheck(low), flush=False)
        record("qrparts_inv_128", us=med * 1e6)
        med, _ = timeit(lambda: cc._norm_1_inf(low), flush=False)
        record("qrparts_norms", us=med * 1e6)
        med, _ = timeit(lambda: cc._gram_factor(X), iters=5, warmup=2)
        record("qrparts_gram_factor_total", ms=med * 1e3)
        med, _ = timeit(lambda: cc.qr_r(X), iters=5, warmup=2)
        record("qrparts_qr_r_total", ms=med * 1e3)
        del X

    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/probe.json", "w") as f:
        json.dump(OUT, f, indent=1)


if __name__ == "__main__":
    main()
