"""Times A^T A of one config-3 block (2 097 152 x 128 float64): the streaming SYRK kernel (default) or, with
NUMS_SYRK_STREAM=0 in the environment, the general DMMA GEMM (development aid, run under gpurun)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nums_b200 import cuda_compute as cc  # noqa: E402
from nums_b200.cuda_system import CudaSystem  # noqa: E402

system = CudaSystem()
system.init()
m, n = 2_097_152, 128
x = torch.randn((m, n), dtype=torch.float64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
times = []
for it in range(8):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    g = cc._gram_of(x)
    b.record()
    b.synchronize()
    if it >= 3:
        times.append(a.elapsed_time(b) * 1e-3)
t = sorted(times)[len(times) // 2]
print(json.dumps({"kernel": "syrk_stream" if os.environ.get("NUMS_SYRK_STREAM", "1") != "0" else "gemm",
                  "m": m, "n": n, "ms": t * 1e3, "tflops_2mn2": 2.0 * m * n * n / t / 1e12,
                  "GBps": 8.0 * m * n / t / 1e9}))

from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    g = cc._gram_of(x)
    torch.cuda.synchronize()
for evt in prof.key_averages():
    total = getattr(evt, "device_time_total", None) or getattr(evt, "cuda_time_total", 0.0)
    if total:
        sys.stdout.write("%-60s %4d %10.1f us\n" % (evt.key[:60], evt.count, float(total)))
