"""Where does an iteration of the reference's glms.newton go when it runs over CudaSystem?  (development aid)

Config 4 shape (11M x 28 float64, 8 row blocks).  Prints, per Newton iteration: host enqueue time (no sync),
device time, the kernel-call count, a cProfile of the host side (top functions by own time) and the CUPTI kernel
table of one iteration (name, launches, total / mean microseconds).  Writes gpurun_out/probe_newton_path.json.
"""
import cProfile
import json
import os
import pstats
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nums_b200.host import HostLayers  # noqa: E402
from nums_b200._lib import LIB  # noqa: E402


def main():
    quick = "--quick" in sys.argv
    host = HostLayers()
    dev = torch.device("cuda", 0)
    N, d, G = (1_100_000 if quick else 11_000_000), 28, 8
    X = host.from_blocks((N, d), (N // G, d), lambda _e, s: torch.randn(s, dtype=torch.float64, device=dev))
    y = host.from_blocks((N,), (N // G,), lambda _e, s: (torch.rand(s, dtype=torch.float64, device=dev) < 0.5).to(torch.float64))
    model = host.logistic_model()
    out = {"host_layers": host.kind, "rows": N, "cols": d, "blocks": G}

    def run(iters):
        host.launch(host.newton(model, X, y, 1e-300, iters))

    run(2)
    torch.cuda.synchronize()
    iters = 4
    calls0 = LIB.dll.nums_launch_count()
    t0 = time.perf_counter()
    run(iters)
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    out["host_enqueue_ms_per_iter"] = t_host / iters * 1e3
    out["wall_ms_per_iter"] = t_all / iters * 1e3
    out["kernel_launches_per_iter"] = (LIB.dll.nums_launch_count() - calls0) / iters
    print(json.dumps(out), flush=True)

    pr = cProfile.Profile()
    pr.enable()
    run(iters)
    pr.disable()
    torch.cuda.synchronize()
    stats = pstats.Stats(pr, stream=sys.stdout)
    stats.sort_stats("tottime").print_stats(35)

    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        run(1)
        torch.cuda.synchronize()
    rows = []
    for evt in prof.key_averages():
        total = getattr(evt, "device_time_total", None)
        if total is None:
            total = getattr(evt, "cuda_time_total", 0.0)
        if total:
            rows.append((evt.key, evt.count, float(total), float(total) / max(evt.count, 1)))
    rows.sort(key=lambda r: -r[2])
    print("%-70s %6s %12s %10s" % ("kernel", "count", "total us", "mean us"))
    for name, count, total, mean in rows[:30]:
        print("%-70s %6d %12.1f %10.1f" % (name[:70], count, total, mean))
    out["kernels_one_iteration"] = [{"name": n, "count": c, "total_us": t, "mean_us": m} for n, c, t, m in rows[:30]]
    out["device_us_one_iteration"] = sum(r[2] for r in rows)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "probe_newton_path.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
