#!/usr/bin/env python
"""bench.py -- the reference's headline workloads on B200 through ``cuda_compute``.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Headline line (one JSON object on stdout, printed by rank 0): BASELINE.json configs[1] --
blocked float64 matmul 16384 x 16384 @ 16384 x 16384 on an 8 x 8 grid of 2048 x 2048 blocks,
``metric`` = FP64 TFLOP/s of the whole job.  A "step" is one full ``C = A @ B`` over the block
grid through the per-block kernel interface (512 ``bop('tensordot')`` + 448 ``bop('add')``
calls at N = 1; SUMMA at N > 1).  The other parts of BASELINE.json's metric (bop
GB/s, TSQR TFLOP/s, Newton-LR s/iter) are measured in the same run and reported under
``"workloads"``.

``--impl reference`` times the reference's CPU implementation of the same path (the NumPy /
OpenBLAS kernels it calls, driven by the same block-level call sequence) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_MATMUL = 16384
BLOCK = 2048
GRID = N_MATMUL // BLOCK
FLOPS_PER_STEP = 2.0 * N_MATMUL ** 3
FLOPS_PER_BLOCK_GEMM = 2.0 * BLOCK ** 3
NOMINAL_FP64_TFLOPS = 40.0     # B200 datasheet FP64 / FP64-tensor figure (not measured)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler(object):
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.thread, self.index = [], None, None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark(self):
        return time.time()

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:  # noqa: BLE001
                self.proc.kill()

    def summary(self, t0, t1):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if ts < t0 or ts > t1 + 0.2:
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# data
# ------------------------------------------------------------------------------------------------
def matmul_blocks_host(seed_a=3, seed_b=4, pinned=True):
    """The 8 x 8 grids of 2048 x 2048 float64 blocks of A and B (standard normal, seeded per block
    so that no 2 GiB host array has to be generated serially)."""
    import torch
    out = {}
    for name, seed in (("A", seed_a), ("B", seed_b)):
        blocks = {}
        for i in range(GRID):
            for j in range(GRID):
                rng = np.random.default_rng([seed, i, j])
                arr = rng.standard_normal((BLOCK, BLOCK))
                if pinned:
                    t = torch.empty((BLOCK, BLOCK), dtype=torch.float64, pin_memory=True)
                    t.numpy()[...] = arr
                    arr = t.numpy()
                blocks[(i, j)] = arr
        out[name] = blocks
    return out["A"], out["B"]


_HOST_BLOCKS = {}


def host_block(name, i, j, pinned=True):
    """One 2048 x 2048 block of A or B on the host (same values as matmul_blocks_host), generated on demand
    and cached -- at N > 1 a rank only ever touches the blocks it owns."""
    import torch
    key = (name, i, j)
    if key not in _HOST_BLOCKS:
        rng = np.random.default_rng([3 if name == "A" else 4, i, j])
        arr = rng.standard_normal((BLOCK, BLOCK))
        if pinned:
            t = torch.empty((BLOCK, BLOCK), dtype=torch.float64, pin_memory=True)
            t.numpy()[...] = arr
            arr = t.numpy()
        _HOST_BLOCKS[key] = arr
    return _HOST_BLOCKS[key]


def blockarray_from_blocks(host, host_blocks, entries=None, into=None):
    """BlockArray (of the host layers `host`) whose blocks are system.put() from the host dict, `entries`
    (default: all) in order."""
    ba = into if into is not None else host.blockarray((N_MATMUL, N_MATMUL), (BLOCK, BLOCK), "float64")
    for entry in (entries if entries is not None else host_blocks):
        ba.blocks[entry].oid = host.system.put(host_blocks[entry])
    return ba


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's NumPy kernels driven by the same block-level call sequence
# ------------------------------------------------------------------------------------------------
def reference_cpu_app():
    """The reference's own ArrayApplication over SerialSystem + numpy_compute (kind "reference"), imported from
    the installed copy (baseline/_ref, /root/reference); None when no NumS installation is present."""
    from nums_b200 import reference_compat
    if not reference_compat.available():
        return None
    try:
        from oracle import ref_loader
        return ref_loader.serial_app()
    except Exception as exc:  # noqa: BLE001
        sys.stderr.write("[bench] reference not importable (%s); CPU arm falls back to the oracle port\n" % exc)
        return None


def cpu_matmul_sample(c_blocks, a_host, b_host, app=None):
    """Computes `c_blocks` C blocks of the 8 x 8 blocked matmul on the host CPU.  With `app` (the reference's
    ArrayApplication over SerialSystem + numpy_compute) the blocks are whole block rows of C computed by the
    reference's own BlockArray.__matmul__ on the leading block rows of A; otherwise the oracle port is driven with
    the same per-block call sequence (8 tensordot + 7 add kernel calls per C block, blockarray.py:460-472).
    Returns (seconds, C blocks computed)."""
    if app is not None:
        from nums.core.array.blockarray import BlockArray
        from nums.core.storage.storage import ArrayGrid
        rows = max(1, min(GRID, (c_blocks + GRID - 1) // GRID))
        A = BlockArray(ArrayGrid((rows * BLOCK, N_MATMUL), (BLOCK, BLOCK), "float64"), app.system)
        B = BlockArray(ArrayGrid((N_MATMUL, N_MATMUL), (BLOCK, BLOCK), "float64"), app.system)
        for (i, k) in A.grid.get_entry_iterator():
            A.blocks[i, k].oid = app.system.put(a_host[(i, k)])
        for (k, j) in B.grid.get_entry_iterator():
            B.blocks[k, j].oid = app.system.put(b_host[(k, j)])
        t0 = time.perf_counter()
        C = A @ B
        C.touch()
        return time.perf_counter() - t0, rows * GRID
    from oracle.cpu_system import OracleSystem
    system = OracleSystem()
    shape = (BLOCK, BLOCK)
    t0 = time.perf_counter()
    done = 0
    for i in range(GRID):
        for j in range(GRID):
            if done >= c_blocks:
                break
            acc = None
            for k in range(GRID):
                sk = {"grid_entry": (i, j), "grid_shape": (GRID, GRID)}
                dot = system.bop("tensordot", a_host[(i, k)], b_host[(k, j)], shape, shape, False, False, axes=1, syskwargs=sk)
                acc = dot if acc is None else system.bop("add", acc, dot, shape, shape, False, False, axes=None, syskwargs=sk)
            done += 1
    return time.perf_counter() - t0, done


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=len(os.sched_getaffinity(0)), user_api="blas")
    except Exception:  # noqa: BLE001
        pass


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        blas = [p for p in threadpool_info() if p.get("user_api") == "blas"]
        if blas:
            return int(blas[0]["num_threads"]), blas[0].get("internal_api", "blas")
    except Exception:  # noqa: BLE001
        pass
    return len(os.sched_getaffinity(0)), "unknown"


def _cpu_arm_description(app, blocks_per_step, api, threads):
    if app is not None:
        return ("%d of 8 block rows of C (%d of 64 C blocks; 8 tensordot + 7 add kernel calls each) of the 16384^2 blocked "
                "matmul, computed by the UNMODIFIED reference: BlockArray.__matmul__ over SerialSystem + numpy_compute "
                "(NumPy %s / %s, %d threads)" % (blocks_per_step // GRID, blocks_per_step, np.__version__, api, threads))
    return ("%d of 64 C blocks (8 tensordot + 7 add kernel calls each) of the 16384^2 blocked matmul, oracle port of "
            "numpy_compute driven by the BlockArray._tensordot call sequence (NumPy %s / %s, %d threads)"
            % (blocks_per_step, np.__version__, api, threads))


def measure_cpu(a_host, b_host, target_seconds=12.0):
    """Bounded sample of the CPU path: as many C blocks as fit in about `target_seconds`."""
    use_all_host_threads()
    app = reference_cpu_app()
    probe, probe_blocks = cpu_matmul_sample(1, a_host, b_host, app)
    per_block = probe / probe_blocks
    blocks = int(max(1, min(GRID * GRID, target_seconds // max(per_block, 1e-3))))
    seconds, done = cpu_matmul_sample(blocks, a_host, b_host, app) if blocks > probe_blocks else (probe, probe_blocks)
    flops = done * GRID * FLOPS_PER_BLOCK_GEMM
    threads, api = cpu_threads()
    return {"value": flops / seconds / 1e12, "unit": "TFLOP/s", "cores": threads,
            "kind": "reference" if app is not None else "port",
            "sample": _cpu_arm_description(app, done, api, threads) + ", %.1f s" % seconds}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    use_all_host_threads()
    app = reference_cpu_app()
    a_host, b_host = matmul_blocks_host(pinned=False)
    probe, probe_blocks = cpu_matmul_sample(1, a_host, b_host, app)
    per_block = probe / probe_blocks
    per_step_blocks = int(max(1, min(GRID * GRID, 40.0 // max(per_block, 1e-3) // max(args.steps + args.warmup, 1))))
    done_per_step = 0
    for _ in range(args.warmup):
        _, done_per_step = cpu_matmul_sample(per_step_blocks, a_host, b_host, app)
    t0 = time.perf_counter()
    done = 0
    for _ in range(args.steps):
        _, done_per_step = cpu_matmul_sample(per_step_blocks, a_host, b_host, app)
        done += done_per_step
    seconds = time.perf_counter() - t0
    threads, api = cpu_threads()
    value = done * GRID * FLOPS_PER_BLOCK_GEMM / seconds / 1e12
    sample = "each step = " + _cpu_arm_description(app, done_per_step, api, threads)
    line = {
        "impl": "reference", "metric": "blocked_matmul_fp64_tflops", "value": value, "unit": "TFLOP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": seconds / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "blocked matmul float64 16384x16384 @ 16384x16384, 8x8 grid of 2048x2048 blocks "
                               "(BASELINE.json configs[1]); CPU arm: " + sample + "; the rate is per C block, so the "
                               "sampled rate equals the whole-product rate",
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": "TFLOP/s", "cores": threads,
                         "kind": "reference" if app is not None else "port", "sample": sample},
        "e2e": {"value": value, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def cuda_time(fn, sync):
    import torch
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    start.record()
    fn()
    end.record()
    end.synchronize()
    sync()
    return start.elapsed_time(end) * 1e-3


def _guarded(out, name, fn):
    """Run one secondary workload; a failure is recorded under its name instead of costing the whole line."""
    import gc
    import traceback
    import torch
    try:
        fn()
    except Exception as exc:  # noqa: BLE001 -- reported in the JSON line, the other workloads still run
        out[name] = {"error": "%s: %s" % (type(exc).__name__, exc), "traceback": traceback.format_exc()[-1500:]}
        sys.stderr.write("[bench] workload %s failed: %s\n" % (name, exc))
    gc.collect()
    torch.cuda.empty_cache()


def other_workloads(host, quick, out=None):
    """bop (cfg1), TSQR (cfg3) and Newton LR (cfg4) on one GPU through the host layers `host` (the reference's
    BlockArray / ArrayApplication / glms when installed); device-timed, synthetic data."""
    import torch
    from nums_b200 import cuda_compute as cc
    out = {} if out is None else out
    dev = torch.device("cuda", torch.cuda.current_device())
    system, app = host.system, host.app

    def timed(fn, iters):
        fn()
        torch.cuda.synchronize()
        times = []
        for _ in range(iters):
            times.append(cuda_time(fn, torch.cuda.synchronize))
        return float(np.median(times))

    def device_blockarray(shape, block_shape, fill):
        return host.from_blocks(shape, block_shape, lambda _entry, block_shape_: fill(block_shape_))

    def cfg1():
        # u + v, u * v on two 1e8-element vectors in 8 blocks (24 B / element)
        n = 100_000_000
        U = device_blockarray((n,), (n // 8,), lambda s: torch.rand(s, dtype=torch.float64, device=dev))
        V = device_blockarray((n,), (n // 8,), lambda s: torch.rand(s, dtype=torch.float64, device=dev))
        for name, fn in (("add", lambda: host.launch(U + V)), ("mul", lambda: host.launch(U * V))):
            t = timed(fn, 5 if quick else 20)
            out["bop_" + name] = {"value": 24.0 * n / t / 1e9, "unit": "GB/s", "ms": t * 1e3,
                                  "workload": "float64 %s of two 1e8-element BlockArrays, 8 blocks (inputs 1.6 GB > L2)" % name}

    def cfg4():
        # Newton LR 11M x 28, 8 row blocks
        N, d, G = 11_000_000, 28, 8
        nb_rows = N // G
        X = device_blockarray((N, d), (nb_rows, d), lambda s: torch.randn(s, dtype=torch.float64, device=dev))
        theta = torch.randn(d, dtype=torch.float64, device=dev) / np.sqrt(d)
        y = host.blockarray((N,), (nb_rows,), "float64")
        for (i,) in y.grid.get_entry_iterator():
            xb = X.blocks[i, 0].oid
            p = torch.sigmoid(xb @ theta)
            y.blocks[i].oid = (torch.rand(xb.shape[0], dtype=torch.float64, device=dev) < p).to(torch.float64)
        model = host.logistic_model()
        iters_unfused = 2 if quick else 6

        def unfused():
            host.launch(host.newton(model, X, y, 1e-300, iters_unfused))
        t = timed(unfused, 1 if quick else 3)
        out["newton_lr_interface_path"] = {
            "value": t / iters_unfused, "unit": "s/iter",
            "workload": "glms.newton (~15 kernel calls per block per iteration) on 11M x 28 float64, 8 row blocks, host layers: "
                        + host.kind,
            "algorithmic_GBps": 8.0 * N * (d + 1) / (t / iters_unfused) / 1e9,
            "unfused_traffic_GBps": 14.8e9 / (t / iters_unfused) / 1e9,
            "note": "the unfused call sequence streams X about six times per iteration (14.8 GB, SURVEY.md 8d); "
                    "unfused_traffic_GBps divides that figure by the time"}
        from nums_b200 import multi_gpu
        comm = multi_gpu.Comm()
        xs = [X.blocks[i, 0].oid for i in range(G)]
        ys = [y.blocks[i].oid for i in range(G)]
        iters_fused = 10

        def fused():
            multi_gpu.newton_lr(system, comm, xs, ys, d, 0.0, iters_fused, cc.lr_grad_hess_blocks, step=cc.newton_step)
        t = timed(fused, 2 if quick else 3)
        out["newton_lr_fused"] = {
            "value": t / iters_fused, "unit": "s/iter",
            "workload": "Newton iteration = fused gradient+Hessian kernel + fused update kernel, 11M x 28 float64, 8 row blocks, "
                        "10 iterations, one 16-byte read-back each",
            "algorithmic_GBps": 8.0 * N * (d + 1) / (t / iters_fused) / 1e9}
        if host.kind == "reference":
            from nums_b200 import glms_fused

            def fused_glms():
                beta0 = app.zeros((d,), (d,), np.float64)
                host.launch(glms_fused.newton(app, model, beta0, X, y, app.scalar(0.0), iters_fused))
            t = timed(fused_glms, 2 if quick else 3)
            out["newton_lr_fused_glms"] = {
                "value": t / iters_fused, "unit": "s/iter",
                "workload": "nums_b200.glms_fused.newton (the drop-in for glms.newton: lr_grad_hess per block + sum_reduce + "
                            "newton_step through the kernel interface, status read one iteration late), 11M x 28 float64, "
                            "8 row blocks, 10 iterations",
                "algorithmic_GBps": 8.0 * N * (d + 1) / (t / iters_fused) / 1e9}

    def cfg3():
        # TSQR 16M x 128 (17.2 GB), 8 row blocks: R only, and (Q, R)
        m, ncol, G = 16_777_216, 128, 8
        X = device_blockarray((m, ncol), (m // G, ncol), lambda s: torch.randn(s, dtype=torch.float64, device=dev))
        flops_r = 2.0 * m * ncol ** 2 - 2.0 * ncol ** 3 / 3.0
        t = timed(lambda: host.launch(app.indirect_tsr(X)), 1 if quick else 3)
        out["tsqr_r"] = {"value": flops_r / t / 1e12, "unit": "TFLOP/s", "ms": t * 1e3,
                         "workload": "indirect_tsr on 16777216 x 128 float64, 8 row blocks (2mn^2 - 2n^3/3 flop)"}
        t = timed(lambda: host.launch(app.indirect_tsqr(X)[0]), 1 if quick else 3)
        out["tsqr_qr"] = {"value": (flops_r + 2.0 * m * ncol ** 2) / t / 1e12, "unit": "TFLOP/s", "ms": t * 1e3,
                          "workload": "indirect_tsqr (Q = X R^-1, R) on 16777216 x 128 float64, 8 row blocks"}

    def csv():
        out["csv_ingest"] = csv_workload(system, quick)

    _guarded(out, "bop", cfg1)
    _guarded(out, "newton_lr", cfg4)
    _guarded(out, "tsqr", cfg3)
    _guarded(out, "csv_ingest", csv)
    return out


def csv_workload(system, quick):
    """read_csv_block on a HIGGS-shaped text file (29 columns, np.savetxt's %.18e).  Reports
    GB/s of text for the two kernels alone (text resident in HBM), end to end from the file (page cache ->
    page-locked buffer -> H2D -> kernels -> shape read-back), and the reference's Python parser on a bounded
    sample of the same file (1 core)."""
    import tempfile
    import torch
    from nums_b200 import cuda_compute as cc
    from oracle import csv_oracle
    rng = np.random.default_rng(9)
    block_rows, cols = 4000, 29
    x = rng.standard_normal((block_rows, cols))
    x[:, 0] = rng.integers(0, 2, block_rows)
    chunk = "".join(",".join("%.18e" % v for v in row) + "\n" for row in x).encode()
    repeat = 16 if quick else 64                      # 46 MB / 185 MB of text
    path = os.path.join(tempfile.mkdtemp(), "higgs_like.csv")
    with open(path, "wb") as f:
        for _ in range(repeat):
            f.write(chunk)
    size = os.path.getsize(path)
    rows = block_rows * repeat
    # end to end, through the public entry point
    cc.read_csv_block(path, 0, size, np.float64, ",", False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    block, shape = cc.read_csv_block(path, 0, size, np.float64, ",", False)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    assert tuple(shape) == (rows, cols)
    got = system.get(block)
    assert np.array_equal(got[:block_rows], x) and np.array_equal(got[-block_rows:], x)   # %.18e round-trips exactly
    del block, got
    # kernels alone
    host = torch.empty(((size + 63) // 32 * 32,), dtype=torch.uint8, pin_memory=True)
    with open(path, "rb") as f:
        f.readinto(memoryview(host.numpy())[:size])
    text = host.cuda()
    t_kernel = min(cuda_time(lambda: cc.parse_csv_text(text, 0, size, ord(","), np.float64, cols), torch.cuda.synchronize)
                   for _ in range(3))
    del text, host
    # reference parser (oracle port of read_csv_block) on the first ~3 MB
    sample_end = min(size, 3 << 20)
    t0 = time.perf_counter()
    ref, _shape = csv_oracle.read_csv_block(path, 0, sample_end, np.float64, ",", False)
    t_cpu = time.perf_counter() - t0
    os.remove(path)
    return {"value": size / t_kernel / 1e9, "unit": "GB/s of text", "ms": t_kernel * 1e3,
            "workload": "read_csv_block on %d rows x %d columns of %%.18e text (%.0f MB), float64" % (rows, cols, size / 1e6),
            "kernels": "nums_csv_index + nums_csv_parse incl. the two 32-byte status read-backs",
            "e2e_GBps": size / t_e2e / 1e9, "e2e_ms": t_e2e * 1e3,
            "cpu_baseline": {"value": sample_end / t_cpu / 1e9, "unit": "GB/s of text", "cores": 1, "kind": "port",
                             "sample": "first %.1f MB (%d rows) through the oracle port of read_csv_block" % (sample_end / 1e6,
                                                                                                          ref.shape[0])},
            "parity": "first and last 4000 rows bit-identical to the generated doubles"}


def sharded_workloads(system, comm, quick, out=None):
    """bop / TSQR / Newton LR with the block grid dealt over the ranks: elementwise work is
    shard-local, TSQR reduces R over a send/recv tree, Newton LR all-reduces g | H once per iteration.
    Every figure is the whole job (all ranks), timed on the device, max over ranks."""
    import torch
    import torch.distributed as dist
    from nums_b200 import cuda_compute as cc
    from nums_b200 import multi_gpu
    dev = torch.device("cuda", torch.cuda.current_device())
    world, rank = comm.world, comm.rank
    out = {} if out is None else out
    G = 8
    mine = [i for i in range(G) if i % world == rank]

    def timed(fn, iters):
        fn()
        times = []
        for _ in range(iters):
            torch.cuda.synchronize()
            comm.barrier()
            t = torch.tensor([cuda_time(fn, torch.cuda.synchronize)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            times.append(float(t.item()))
        return float(np.median(times))

    def cfg1():
        # 8 blocks of 12.5M, round robin
        n = 100_000_000
        us = [torch.rand(n // G, dtype=torch.float64, device=dev) for _ in mine]
        vs = [torch.rand(n // G, dtype=torch.float64, device=dev) for _ in mine]
        shape = (n // G,)

        def bop_add():      # (cuda_time brackets the launches with events and waits for the end event itself)
            for u, v, i in zip(us, vs, mine):
                system.bop("add", u, v, shape, shape, False, False, axes=None, syskwargs={"grid_entry": (i,), "grid_shape": (G,)})
        t = timed(bop_add, 5 if quick else 20)
        out["bop_add"] = {"value": 24.0 * n / t / 1e9, "unit": "GB/s", "ms": t * 1e3,
                          "workload": "float64 add of two 1e8-element arrays, 8 blocks dealt over %d GPU(s), no exchange" % world}

    def cfg3():
        # TSQR 16M x 128, 8 row blocks
        m, ncol = 16_777_216, 128
        xs = [torch.randn((m // G, ncol), dtype=torch.float64, device=dev) for _ in mine]
        flops_r = 2.0 * m * ncol ** 2 - 2.0 * ncol ** 3 / 3.0

        def tsqr_r():
            multi_gpu.tsqr_r_tree(system, comm, xs, ncol)
            torch.cuda.synchronize()
        t = timed(tsqr_r, 2 if quick else 3)
        out["tsqr_r"] = {"value": flops_r / t / 1e12, "unit": "TFLOP/s", "ms": t * 1e3,
                         "workload": "R of 16777216 x 128 float64, 8 row blocks over %d GPU(s), binary tree over the per-rank R" % world}

        def tsqr_qr():
            r = multi_gpu.tsqr_r_tree(system, comm, xs, ncol)
            r_inv = system.inv(r, syskwargs={"grid_entry": (0, 0), "grid_shape": (1, 1)})
            qs = [multi_gpu.tsqr_q(system, x, r_inv, (i, 0), (G, 1)) for x, i in zip(xs, mine)]
            system.flush()
            torch.cuda.synchronize()
            return qs
        t = timed(tsqr_qr, 2 if quick else 3)
        out["tsqr_qr"] = {"value": (flops_r + 2.0 * m * ncol ** 2) / t / 1e12, "unit": "TFLOP/s", "ms": t * 1e3,
                          "workload": "(Q = X R^-1, R) of 16777216 x 128 float64 over %d GPU(s)" % world}

    def cfg4():
        # Newton LR 11M x 28, 8 row blocks
        N, d = 11_000_000, 28
        xs = [torch.randn((N // G, d), dtype=torch.float64, device=dev) for _ in mine]
        theta = torch.ones(d, dtype=torch.float64, device=dev) / np.sqrt(d)
        ys = [(torch.rand(x.shape[0], dtype=torch.float64, device=dev) < torch.sigmoid(x @ theta)).to(torch.float64) for x in xs]
        iters = 10

        def newton():
            multi_gpu.newton_lr(system, comm, xs, ys, d, 0.0, iters, cc.lr_grad_hess_blocks, step=cc.newton_step)
        t = timed(newton, 2 if quick else 3)
        out["newton_lr_fused"] = {"value": t / iters, "unit": "s/iter",
                                  "workload": "Newton LR 11M x 28 float64, 8 row blocks over %d GPU(s), fused g|H kernel + one "
                                              "all-reduce of 812 doubles per iteration" % world,
                                  "algorithmic_GBps": 8.0 * N * (d + 1) / (t / iters) / 1e9}

    _guarded(out, "bop_add", cfg1)
    _guarded(out, "tsqr_r", cfg3)
    _guarded(out, "newton_lr_fused", cfg4)
    return out


def api_workloads(host, quick, out=None):
    """cfg1 / cfg3 / cfg4 at N > 1 THROUGH THE PLUGIN API: the reference's BlockArray operators,
    ArrayApplication.indirect_tsr and glms.newton over SpmdSystem, blocks living on their owners.  Whole-job figures,
    device-timed, max over ranks (same conventions as sharded_workloads, which calls the drivers directly)."""
    import torch
    import torch.distributed as dist
    system, app = host.system, host.app
    world, rank = system.world_size, system.rank
    dev = torch.device("cuda", torch.cuda.current_device())
    out = {} if out is None else out
    G = 8

    def timed(fn, iters):
        fn()
        times = []
        for _ in range(iters):
            system.synchronize()
            dist.barrier()
            t = torch.tensor([cuda_time(fn, torch.cuda.synchronize)], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            times.append(float(t.item()))
        return float(np.median(times))

    def distributed(shape, block_shape, fill):
        ba = host.blockarray(shape, block_shape, "float64")
        gshape = ba.grid.grid_shape
        for entry in ba.grid.get_entry_iterator():
            bshape = ba.grid.get_block_shape(entry)
            mine = system.owner(entry, gshape) == rank
            ba.blocks[entry].oid = system.put_at(fill(bshape) if mine else None, entry, gshape, shape=bshape, dtype=np.float64)
        return ba

    def cfg1():
        n = 100_000_000
        U = distributed((n,), (n // G,), lambda s: torch.rand(s, dtype=torch.float64, device=dev))
        V = distributed((n,), (n // G,), lambda s: torch.rand(s, dtype=torch.float64, device=dev))
        t = timed(lambda: host.launch(U + V), 5 if quick else 20)
        out["bop_add_api"] = {"value": 24.0 * n / t / 1e9, "unit": "GB/s", "ms": t * 1e3,
                              "workload": "BlockArray.__add__ of two 1e8-element float64 arrays, 8 blocks over %d GPU(s) through "
                                          "SpmdSystem (shard-local, no exchange)" % world}

    def cfg3():
        m, ncol = 16_777_216, 128
        X = distributed((m, ncol), (m // G, ncol), lambda s: torch.randn(s, dtype=torch.float64, device=dev))
        flops_r = 2.0 * m * ncol ** 2 - 2.0 * ncol ** 3 / 3.0
        t = timed(lambda: (host.launch(app.indirect_tsr(X)), torch.cuda.synchronize()), 2 if quick else 3)
        out["tsqr_r_api"] = {"value": flops_r / t / 1e12, "unit": "TFLOP/s", "ms": t * 1e3,
                             "workload": "ArrayApplication.indirect_tsr on 16777216 x 128 float64, 8 row blocks over %d GPU(s): "
                                         "per-block Gram matrices summed on their owners, one all-reduce, one factorization "
                                         "(Householder tree over the ranks when the condition bound refuses), R replicated" % world}
        t = timed(lambda: (host.launch(app.indirect_tsqr(X)[0]), torch.cuda.synchronize()), 2 if quick else 3)
        out["tsqr_qr_api"] = {"value": (flops_r + 2.0 * m * ncol ** 2) / t / 1e12, "unit": "TFLOP/s", "ms": t * 1e3,
                              "workload": "ArrayApplication.indirect_tsqr (Q = X R^-1 on every block's owner, R replicated) on "
                                          "16777216 x 128 float64, 8 row blocks over %d GPU(s)" % world}

    def cfg4():
        N, d = 11_000_000, 28
        X = distributed((N, d), (N // G, d), lambda s: torch.randn(s, dtype=torch.float64, device=dev))
        y = distributed((N,), (N // G,), lambda s: (torch.rand(s, dtype=torch.float64, device=dev) < 0.5).to(torch.float64))
        model = host.logistic_model()
        iters = 2 if quick else 4
        t = timed(lambda: (host.launch(host.newton(model, X, y, 1e-300, iters)), torch.cuda.synchronize()), 1 if quick else 2)
        out["newton_lr_interface_path"] = {
            "value": t / iters, "unit": "s/iter",
            "workload": "glms.newton on 11M x 28 float64, 8 row blocks over %d GPU(s) through SpmdSystem: ~15 kernel calls per block "
                        "per iteration on the block's owner, g and H all-reduced, beta replicated" % world,
            "algorithmic_GBps": 8.0 * N * (d + 1) / (t / iters) / 1e9}
        from nums_b200 import glms_fused
        iters_fused = 10

        def fused_glms():
            beta0 = app.zeros((d,), (d,), np.float64)
            host.launch(glms_fused.newton(app, model, beta0, X, y, app.scalar(0.0), iters_fused))
            torch.cuda.synchronize()
        t = timed(fused_glms, 2 if quick else 3)
        out["newton_lr_fused_glms"] = {
            "value": t / iters_fused, "unit": "s/iter",
            "workload": "nums_b200.glms_fused.newton through SpmdSystem on %d GPU(s): lr_grad_hess per block on its owner, g | H "
                        "summed locally and all-reduced (one 812-double NCCL all-reduce per iteration), newton_step replicated, "
                        "status read one iteration late; 11M x 28 float64, 8 row blocks, 10 iterations" % world,
            "algorithmic_GBps": 8.0 * N * (d + 1) / (t / iters_fused) / 1e9}

    def cfg5():
        out["large_matmul_65536"] = large_matmul_workload(host, steps=2)

    _guarded(out, "bop_add_api", cfg1)
    _guarded(out, "tsqr_r_api", cfg3)
    _guarded(out, "newton_lr_api", cfg4)
    if world >= 8 and not quick:
        _guarded(out, "large_matmul_65536", cfg5)
    return out


def large_matmul_workload(host, steps=2, n=65536, bs=8192, samples=8):
    """BASELINE.json configs[4] through the plugin API: float64 65536 x 65536 @ 65536 x 65536 on an 8 x 8 grid of
    8192 x 8192 blocks (34.4 GB per operand), BlockArray.__matmul__ over SpmdSystem (or CudaSystem at N = 1).
    Blocks are generated on their owners' devices (seeded per block).  Parity: a 128 x 128 corner of `samples`
    different C blocks against NumPy on the host, from the matching slices of the operand blocks."""
    import torch
    import torch.distributed as dist
    system = host.system
    world, rank = getattr(system, "world_size", 1), getattr(system, "rank", 0)
    spmd = world > 1
    dev = torch.device("cuda", torch.cuda.current_device())
    g = n // bs

    def block(seed, i, j):
        gen = torch.Generator(device=dev)
        gen.manual_seed(seed * 1000 + i * g + j)
        return torch.randn((bs, bs), dtype=torch.float64, device=dev, generator=gen)

    def operand(seed):
        ba = host.blockarray((n, n), (bs, bs), "float64")
        for (i, j) in ba.grid.get_entry_iterator():
            if spmd:
                mine = system.owner((i, j), (g, g)) == rank
                ba.blocks[i, j].oid = system.put_at(block(seed, i, j) if mine else None, (i, j), (g, g),
                                                    shape=(bs, bs), dtype=np.float64)
            else:
                ba.blocks[i, j].oid = block(seed, i, j)
        return ba

    def sync():
        system.synchronize()
        if spmd:
            dist.barrier()
            torch.cuda.synchronize()
    A, B = operand(3), operand(4)

    def product():
        if spmd:
            system.evict_copies()
        c = A @ B
        system.flush()
        return c
    c = product()                       # warm-up (also stages the operands into the peer arenas at N > 1)
    sync()
    c = None
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(steps):
        c = None
        c = product()
    end.record()
    end.synchronize()
    sync()
    t = torch.tensor([start.elapsed_time(end) * 1e-3], dtype=torch.float64, device=dev)
    if spmd:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    seconds = float(t.item()) / steps
    # parity on sampled corners: the owner of each block cuts the slice, rank 0 receives it
    def corner(ba, i, j, rows, cols):
        oid = ba.blocks[i, j].oid
        if not spmd:
            return system.get(system.contractions.resolve(oid)[rows, cols])
        home = oid.home
        piece = torch.empty((128, bs) if cols == slice(None) else ((bs, 128) if rows == slice(None) else (128, 128)),
                            dtype=torch.float64, device=dev)
        if rank == home:
            piece.copy_(system.backend.settle(oid.value)[rows, cols])
        dist.broadcast(piece, src=home)
        return piece.cpu().numpy()
    errs = []
    top = slice(0, 128)
    for s_ in range(samples):
        i, j = (3 * s_ + 1) % g, (5 * s_ + 2) % g
        got = corner(c, i, j, top, top)
        ref = np.zeros((128, 128))
        for k in range(g):
            ref += corner(A, i, k, top, slice(None)) @ corner(B, k, j, slice(None), top)
        errs.append(float(np.linalg.norm(got - ref) / np.linalg.norm(ref)))
    resident = torch.cuda.max_memory_allocated() / 1e9
    del A, B, c
    torch.cuda.empty_cache()
    flops = 2.0 * n ** 3
    return {"value": flops / seconds / 1e12, "unit": "TFLOP/s", "ms": seconds * 1e3, "steps": steps,
            "workload": "blocked matmul float64 65536x65536 @ 65536x65536, 8x8 grid of 8192x8192 blocks (BASELINE.json "
                        "configs[4]) through BlockArray.__matmul__ on %d GPU(s)" % world,
            "parity_rel_err_max": max(errs), "parity_samples": samples, "parity_bar": 1e-10,
            "parity_what": "128x128 corner of %d different C blocks vs NumPy on the host" % samples,
            "peak_allocated_gb_per_gpu": resident}


def run_large(args):
    """BASELINE.json configs[4]: float64 65536 x 65536 @ 65536 x 65536, 8 x 8 grid of 8192 x 8192 blocks
    (34.4 GB per operand, 103 GB resident).  Blocks are generated on the device, block by block (seeded);
    parity is checked on a sampled 512 x 512 corner of one C block against NumPy on the host."""
    import torch
    import torch.distributed as dist
    from nums_b200 import _lib, multi_gpu, reference_compat
    from nums_b200.cuda_system import CudaSystem
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), pg_options=multi_gpu.nccl_options())
    if reference_compat.available() and not args.summa_driver:
        # through the plugin API: the reference's BlockArray.__matmul__ over CudaSystem (N = 1) / SpmdSystem (N > 1)
        from nums_b200.host import HostLayers
        host = HostLayers()
        w = large_matmul_workload(host, steps=args.steps)
        if rank == 0:
            emit({"metric": "blocked_matmul_fp64_tflops", "value": w["value"], "unit": "TFLOP/s", "n_gpus": world,
                  "steps": args.steps, "warmup": 1, "ms_per_step": w["ms"], "higher_is_better": True, "scaling": "strong",
                  "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                  "config": {"workload": w["workload"], "host_layers": host.description,
                             "parity_rel_err": w["parity_rel_err_max"]},
                  "parity_check": {"what": w["parity_what"], "rel_err": w["parity_rel_err_max"], "bar": 1e-10},
                  "resident_gb_per_gpu": w["peak_allocated_gb_per_gpu"]})
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    comm = multi_gpu.Comm()
    system = CudaSystem(rank=rank, world_size=world)
    system.init()
    n, bs, g = 65536, 8192, 8
    dev = torch.device("cuda", local_rank)

    def block(seed, i, j):
        gen = torch.Generator(device=dev)
        gen.manual_seed(seed * 1000 + i * 8 + j)
        return torch.randn((bs, bs), dtype=torch.float64, device=dev, generator=gen)
    like = torch.empty((1,), dtype=torch.float64, device=dev)
    summa = multi_gpu.SummaMatmul(system, comm, g, bs, like)
    # fill the packed per-k panels directly (no second copy of the operands: 103 GB must fit one GPU)
    pa, pb = {}, {}
    for k in range(g):
        if k % summa.pc == summa.c:
            pa[k] = torch.empty((len(summa.my_i), bs, bs), dtype=torch.float64, device=dev)
            for idx, i in enumerate(summa.my_i):
                pa[k][idx].copy_(block(3, i, k))
        if k % summa.pr == summa.r:
            pb[k] = torch.empty((len(summa.my_j), bs, bs), dtype=torch.float64, device=dev)
            for idx, j in enumerate(summa.my_j):
                pb[k][idx].copy_(block(4, k, j))
    packed = (pa, pb)
    torch.cuda.empty_cache()

    def sync_all():
        torch.cuda.synchronize()
        comm.barrier()
        torch.cuda.synchronize()
    c = None
    for _ in range(max(1, args.warmup)):
        c = None
        c = summa.run(packed)
    sync_all()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(args.steps):
        c = None                   # release the previous result before the next one is allocated
        c = summa.run(packed)
    end.record()
    end.synchronize()
    sync_all()
    elapsed = torch.tensor([start.elapsed_time(end) * 1e-3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
    seconds = float(elapsed.item())
    flops = 2.0 * n ** 3
    rel = None
    if rank == 0:
        (i, j) = sorted(c)[0]
        got = system.get(system.contractions.resolve(c[(i, j)])[:512, :512])
        ref = np.zeros((512, 512))
        for k in range(g):
            a = block(3, i, k)[:512].cpu().numpy()
            bcol = block(4, k, j)[:, :512].contiguous().cpu().numpy()
            ref += a @ bcol
        rel = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
        line = {"metric": "blocked_matmul_fp64_tflops", "value": flops * args.steps / seconds / 1e12, "unit": "TFLOP/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(1, args.warmup),
                "ms_per_step": seconds / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "blocked matmul float64 65536x65536 @ 65536x65536, 8x8 grid of 8192x8192 blocks "
                                       "(BASELINE.json configs[4]), SUMMA on a %dx%d device grid" % multi_gpu.device_grid(world)},
                "parity_check": {"what": "512x512 corner of one C block vs NumPy on the host", "rel_err": rel, "bar": 1e-10},
                "resident_gb_per_gpu": torch.cuda.max_memory_allocated() / 1e9}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from nums_b200 import _lib, multi_gpu
    from nums_b200.cuda_system import CudaSystem

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), pg_options=multi_gpu.nccl_options())
    comm = multi_gpu.Comm()
    system = None
    LIB = _lib.LIB

    def sync_all():
        torch.cuda.synchronize()
        comm.barrier()
        torch.cuda.synchronize()

    # ---- inputs (host, pinned) -----------------------------------------------------------------------
    from nums_b200 import reference_compat
    api_path = world == 1 or (reference_compat.available() and not args.mirror and not args.summa_driver)
    summa = packed = None
    if world == 1:
        a_host, b_host = matmul_blocks_host(pinned=True)
        from nums_b200.host import HostLayers
        host = HostLayers(prefer_reference=not args.mirror)
        system = host.system            # (a ReferenceCudaSystem when the reference's host layers are driving)
        A = blockarray_from_blocks(host, a_host)
        B = blockarray_from_blocks(host, b_host)

        def step_resident():
            # BlockArray.__matmul__ of the host layers enqueues the whole step (512 tensordot + 448 add kernel
            # calls -> one grouped launch); the timed region is bracketed by a device synchronisation on both
            # sides, so the host may run ahead of the GPU between steps
            c = A @ B
            system.flush()
            return c

        def step_e2e(reference_get=False):
            # puts are asynchronous (upload stream) and get() drains finished block rows while later ones
            # compute.  Streaming order of the 128 puts: row 0 of A, then B column by column, then the other
            # rows of A -- C(0, j) can start as soon as column j of B has landed, C(i, :) as soon as row i of A.
            # (Measured alternative: row i of A / column i of B alternately, which feeds the GEMM in L-shaped fronts
            # and never starves it -- but then no block ROW of C is complete before the very end, the drain of
            # get_assembled cannot overlap the compute any more, and the step takes 306 ms instead of 272.)
            a = blockarray_from_blocks(host, a_host, [(0, k) for k in range(GRID)])
            b = blockarray_from_blocks(host, b_host, [(k, j) for j in range(GRID) for k in range(GRID)])
            blockarray_from_blocks(host, a_host, [(i, k) for i in range(1, GRID) for k in range(GRID)], into=a)
            c = a @ b
            return c.get() if reference_get else host.get(c)
        parallelism = ("1 GPU; host layers = %s; BlockArray.__matmul__ -> _tensordot issues 512 tensordot + 448 add kernel "
                       "calls, deferred by CudaSystem into one grouped DMMA launch per step" % host.description)
    elif api_path:
        # N > 1 through the plugin API: every rank runs the reference's BlockArray.__matmul__ over SpmdSystem
        # (nums_b200/spmd.py); blocks enter on their owners (put_at), operand copies are evicted before every step
        # so that each product pays for its exchange
        from nums_b200.host import HostLayers
        host = HostLayers()
        system = host.system
        grid_shape = (GRID, GRID)

        def distributed_operand(name):
            ba = host.blockarray((N_MATMUL, N_MATMUL), (BLOCK, BLOCK), "float64")
            for entry in ba.grid.get_entry_iterator():
                mine = system.owner(entry, grid_shape) == rank
                ba.blocks[entry].oid = system.put_at(host_block(name, *entry) if mine else None, entry, grid_shape,
                                                     shape=(BLOCK, BLOCK), dtype=np.float64)
            return ba
        A, B = distributed_operand("A"), distributed_operand("B")

        def step_resident():
            system.evict_copies()
            c = A @ B
            system.flush()
            return c

        def step_e2e(reference_get=False):
            a, b = distributed_operand("A"), distributed_operand("B")
            c = a @ b
            return system.get_owned([c.blocks[e].oid for e in c.grid.get_entry_iterator()])
        pr, pc = system.device_grid
        parallelism = ("%d GPUs, one process each; host layers = %s over SpmdSystem: every rank runs BlockArray.__matmul__ "
                       "(512 tensordot + 448 add kernel calls), C(i,j) is computed on rank (i mod %d)*%d + (j mod %d), remote "
                       "A(i,k) / B(k,j) blocks are exchanged in one batched NCCL point-to-point transfer per product (cached "
                       "copies evicted before every step) and each rank contracts its blocks in one grouped DMMA launch"
                       % (world, host.description, pr, pc, pc))
    else:
        a_host, b_host = matmul_blocks_host(pinned=True)
        system = CudaSystem(rank=rank, world_size=world)
        system.init()
        pr, pc = multi_gpu.device_grid(world)
        like = torch.empty((1,), dtype=torch.float64, device="cuda")
        summa = multi_gpu.SummaMatmul(system, comm, GRID, BLOCK, like)
        mine_a = {e: system.put(v) for e, v in a_host.items() if summa.owner_a(*e) == rank}
        mine_b = {e: system.put(v) for e, v in b_host.items() if summa.owner_b(*e) == rank}
        packed = summa.pack(mine_a, mine_b)      # resident layout: one contiguous panel per k
        del mine_a, mine_b

        flush_every = int(os.environ.get("NUMS_SUMMA_FLUSH", "0"))   # 0: SummaMatmul.run's own schedule

        def step_resident():
            return summa.run(packed, flush_every=flush_every)

        def step_e2e():
            la = {e: system.put(v) for e, v in a_host.items() if summa.owner_a(*e) == rank}
            lb = {e: system.put(v) for e, v in b_host.items() if summa.owner_b(*e) == rank}
            c = summa.run(la, lb)
            return {e: system.get(v) for e, v in c.items()}
        exchange = ("each rank pulls the A(:,k) panel of its device row and the B(k,:) panel of its device column out of "
                    "the owners' symmetric (peer-mapped) memory with copy-engine transfers over NVLink, all queued up front; "
                    "step 0 and steps 1..7 are two grouped DMMA launches"
                    if isinstance(packed, multi_gpu._PublishedPanels) else
                    "one NCCL broadcast of the A(:,k) panel per device row and of the B(k,:) panel per device column per "
                    "step, prefetched behind the grouped local GEMM")
        parallelism = "SUMMA on a %dx%d device grid: %s" % (pr, pc, exchange)

    # ---- timed: HBM-resident ---------------------------------------------------------------------------
    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    sync_all()
    launches0 = LIB.dll.nums_launch_count()
    t_mark0 = sampler.mark()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    host_step_ms = []
    for _ in range(args.steps):
        t_host = time.perf_counter()
        step_resident()
        host_step_ms.append((time.perf_counter() - t_host) * 1e3)
    end.record()
    end.synchronize()
    sync_all()
    t_mark1 = sampler.mark()
    if os.environ.get("NUMS_TRACE") == "1":
        sys.stderr.write("TRACE[rank %d] timed region: device %.1f ms, host enqueue per step: %s\n"
                         % (rank, start.elapsed_time(end), " ".join("%.1f" % m for m in host_step_ms)))
    launches = LIB.dll.nums_launch_count() - launches0
    elapsed = torch.tensor([start.elapsed_time(end) * 1e-3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
    seconds = float(elapsed.item())
    value = FLOPS_PER_STEP * args.steps / seconds / 1e12

    if os.environ.get("NUMS_TRACE") == "1":
        from nums_b200 import trace
        trace.report()
        host_ms = []
        for _ in range(3):
            t0 = time.perf_counter()
            step_resident()
            host_ms.append((time.perf_counter() - t0) * 1e3)
        rows = trace.report()
        if rank == 0:
            sys.stderr.write("TRACE host ms per step_resident call: %s\n" % ", ".join("%.1f" % m for m in host_ms))
            for label, t, dt, host_t in rows:
                sys.stderr.write("TRACE %9.3f ms (+%8.3f) host %8.1f  %s\n" % (t, dt, host_t, label))
    if summa is not None and os.environ.get("NUMS_SUMMA_TRACE"):
        summa.trace = []
        t_start = torch.cuda.Event(enable_timing=True)
        t_start.record()
        step_resident()
        step_resident()
        torch.cuda.synchronize()
        if rank == 0:
            prev = t_start
            for label, k, ev in summa.trace:
                sys.stderr.write("TRACE %-13s k=%d  +%.3f ms  (t=%.3f)\n" % (label, k, prev.elapsed_time(ev), t_start.elapsed_time(ev)))
                prev = ev
        summa.trace = None

    # ---- timed: end to end (pinned host -> HBM -> kernels -> host) ------------------------------------------
    e2e_steps = max(1, min(args.steps, 3))
    step_e2e()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    sync_all()
    e2e_seconds = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_seconds, op=dist.ReduceOp.MAX)
    e2e_value = FLOPS_PER_STEP * e2e_steps / float(e2e_seconds.item()) / 1e12
    bytes_matrix = 8 * N_MATMUL * N_MATMUL
    e2e_blockarray_get = None
    if world == 1 and host.kind == "reference":
        # the same step with the result fetched by the reference's own BlockArray.get() (base.py:348-360): one
        # system.get of all blocks, then a single-threaded host copy of every block into a fresh np.zeros array
        step_e2e(reference_get=True)
        t0 = time.perf_counter()
        step_e2e(reference_get=True)
        e2e_blockarray_get = FLOPS_PER_STEP / (time.perf_counter() - t0) / 1e12

    # ---- parity spot check of this run's output (rank 0): one C block against NumPy on the host ------
    verified = None
    if api_path:
        c_dev = system.get((A @ B).blocks[0, 0].oid)        # at N > 1 a broadcast from the block's home: all ranks call it
        if rank == 0:
            ref = np.zeros((BLOCK, BLOCK))
            for k in range(GRID):
                ref += host_block("A", 0, k, pinned=False) @ host_block("B", k, 0, pinned=False)
            verified = float(np.linalg.norm(c_dev - ref) / np.linalg.norm(ref))
    elif rank == 0:
        c_all = summa.run(packed)
        i, j = key = sorted(c_all)[0]
        c_dev = system.get(c_all[key])
        ref = np.zeros((BLOCK, BLOCK))
        for k in range(GRID):
            ref += a_host[(i, k)] @ b_host[(k, j)]
        verified = float(np.linalg.norm(c_dev - ref) / np.linalg.norm(ref))
    else:
        summa.run(packed)              # collective: every rank takes part in the verification pass
    sync_all()

    # ---- everything the headline needs is measured: from here on a watchdog guarantees the JSON line ----------
    # The secondary workloads are collective at N > 1; should one of them fail or stall on some rank, rank 0
    # still prints the headline (with the failure recorded under "workloads") and every rank exits.
    bytes_matrix = 8 * N_MATMUL * N_MATMUL
    clocks = None
    if rank == 0:
        sampler.stop()
        clocks = sampler.summary(t_mark0, t_mark1)
    state = {"workloads": None, "roofline": None, "cpu_baseline": None, "done": False}

    def compose():
        return {
            "metric": "blocked_matmul_fp64_tflops", "value": value, "unit": "TFLOP/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": seconds / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "blocked matmul float64 16384x16384 @ 16384x16384, 8x8 grid of 2048x2048 blocks "
                                   "(BASELINE.json configs[1])",
                       "parallelism": parallelism,
                       "parity_rel_err": verified,
                       "l2_policy": "inputs (2 x 2.1 GB) and output (2.1 GB) exceed the 126 MB L2; no flush needed"},
            "e2e": {"value": e2e_value, "unit": "TFLOP/s",
                    "h2d_bytes_per_step": 2 * bytes_matrix, "d2h_bytes_per_step": bytes_matrix, "steps": e2e_steps,
                    "what": ("pinned host blocks -> system.put (async, upload stream) -> A @ B through the block kernel interface "
                             "(deferred, launched in groups as operands land) -> CudaSystem.get_assembled on the host (block rows "
                             "drained on a download stream); wall clock, host<->device copies inside") if world == 1 else
                            ("every rank: pinned host blocks it owns -> put_at (async upload) -> BlockArray.__matmul__ over SpmdSystem "
                             "(exchange + grouped launch) -> get_owned: the C blocks it owns back to its host; wall clock, "
                             "host<->device copies inside, max over ranks") if api_path else
                            "every rank: put of its blocks -> SummaMatmul driver -> get of its C blocks; wall clock, max over ranks",
                    "through_reference_BlockArray_get": e2e_blockarray_get},
            "gpu_launches": int(launches),
            "parity_check": {"what": "one 2048x2048 block of C vs NumPy on the host (relative Frobenius error)",
                             "rel_err": verified, "bar": 1e-10},
            "clocks": clocks,
            "roofline": state["roofline"],
            "cpu_baseline": state["cpu_baseline"],
            "workloads": state["workloads"],
        }

    def watchdog_fired():
        if state["done"]:
            return
        state["done"] = True
        if rank == 0:
            note = {"error": "secondary workloads did not finish within %d s; headline figures above are complete" % args.watchdog}
            if isinstance(state["workloads"], dict):
                state["workloads"]["_watchdog"] = note
            else:
                state["workloads"] = note
            emit(compose())
        else:
            time.sleep(2.0)
        os._exit(0)
    watchdog = threading.Timer(float(args.watchdog), watchdog_fired)
    watchdog.daemon = True
    watchdog.start()

    sharded = None
    if world > 1 and not args.skip_workloads:
        packed = None
        if api_path:
            del A, B
        torch.cuda.empty_cache()
        raw_system = system.local if api_path else system
        state["workloads"] = sharded = {}
        sharded.update(sharded_workloads(raw_system, comm, quick=args.quick))   # collective: all ranks
        if api_path:
            api_workloads(host, quick=args.quick, out=sharded)

    if rank != 0:
        state["done"] = True
        watchdog.cancel()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- rank 0, N == 1 extras: roofline of the dominant kernel, CPU baseline, other workloads ----------------
    roofline = None
    if world == 1:
        # Dominant kernel: the grouped DMMA launch that a step's dot/add chain collapses into
        # (one launch per step, see nums_b200/deferred.py).  Event-timed here, in isolation.
        # The host enqueue of the 960 deferred calls (~5 ms) happens BEFORE the start event: the events bracket
        # the flush only, i.e. the grouped GEMM launch plus the 6 us kernel that copies its descriptor tables.
        def one_launch():
            c = A @ B
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            system.flush()
            ev1.record()
            ev1.synchronize()
            del c
            return ev0.elapsed_time(ev1) * 1e-3
        one_launch()
        launches_before = LIB.dll.nums_launch_count()
        t_kernel = min(one_launch() for _ in range(3))
        launches_per_step = (LIB.dll.nums_launch_count() - launches_before) / 3.0
        big = torch.randn((8192, 8192), dtype=torch.float64, device="cuda")
        torch.matmul(big, big)
        t_cublas = min(cuda_time(lambda: torch.matmul(big, big), torch.cuda.synchronize) for _ in range(3))
        del big
        cublas = 2.0 * 8192 ** 3 / t_cublas / 1e12
        # FP64 tensor pipe: every sub-partition issues one DMMA.8x8x4 (256 FMA) per 16 cycles (DESIGN.md section 4;
        # ncu shows the dmma sub-pipe 98.8 % busy at 97.6 % of this figure)
        sm_mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0
        peak = torch.cuda.get_device_properties(0).multi_processor_count * 4 * (256 / 16.0) * 2 * sm_mhz * 1e6 / 1e12
        achieved = FLOPS_PER_STEP / t_kernel / 1e12
        roofline = {"bound": "tensor",
                    "kernel": "dgemm_dmma_tma_kernel (FP64 DMMA m8n8k4, 128x128x32 tiles fed by 2-D tensor-map TMA from a "
                              "producer warp, grouped over the 64 result blocks x 8 k-terms of one step)",
                    "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "peak_source": "FP64 tensor-pipe issue limit = SMs x 4 sub-partitions x 256 FMA per 16 cycles x the median SM "
                                   "clock sampled under load (%.0f MHz); MEASURED_PEAKS.json and the profiling guide hold no FP64 "
                                   "figure (only bf16 and HBM). cuBLAS DGEMM measured live in this run is reported next to it"
                                   % sm_mhz,
                    "cublas_dgemm_tflops": cublas, "fp64_dgemm_tflops_measured": cublas, "frac_of_cublas": achieved / cublas,
                    "frac_of_nominal_40TF": achieved / NOMINAL_FP64_TFLOPS,
                    "algorithmic_flops_per_launch": FLOPS_PER_STEP, "gemm_launches_per_step": 1,
                    "all_launches_per_step": launches_per_step,
                    "launch_note": "one dgemm_dmma_tma_kernel launch per step; the other launch is table_copy_kernel (6 us), "
                                   "which moves the grouped launch's descriptor tables",
                    "avg_launch_ms": t_kernel * 1e3, "traffic": None}
        try:
            with open(os.path.join(ROOT, "profiles", "dgemm_traffic.json")) as f:
                roofline["traffic"] = json.load(f).get("dram_bytes_per_launch")
        except Exception:  # noqa: BLE001
            pass
        state["roofline"] = roofline
        del A, B
        torch.cuda.empty_cache()
        if not args.skip_cpu:
            try:
                state["cpu_baseline"] = measure_cpu(a_host, b_host)
            except Exception as exc:  # noqa: BLE001 -- reported, the headline stands
                state["cpu_baseline"] = {"error": "%s: %s" % (type(exc).__name__, exc)}
        if not args.skip_workloads:
            state["workloads"] = {}
            other_workloads(host, quick=args.quick, out=state["workloads"])
    else:
        try:    # per-GPU rate of the whole product (both grouped launches + whatever transfer time is exposed)
            sm_mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0
            peak = torch.cuda.get_device_properties(local_rank).multi_processor_count * 4 * (256 / 16.0) * 2 * sm_mhz * 1e6 / 1e12
            roofline = {"bound": "tensor",
                        "kernel": "dgemm_dmma_tma_kernel, the grouped launches of one SUMMA product on each rank",
                        "achieved": value / world, "peak": peak, "unit": "TFLOP/s", "frac": value / world / peak,
                        "scope": "per GPU over the whole timed step (max over ranks), not a kernel in isolation",
                        "peak_source": "FP64 tensor-pipe issue limit = SMs x 4 sub-partitions x 256 FMA per 16 cycles x the SM clock "
                                       "sampled on rank 0 (%.0f MHz)" % sm_mhz,
                        "algorithmic_flops_per_launch": FLOPS_PER_STEP / world, "traffic": None}
        except Exception:  # noqa: BLE001 -- reporting only
            roofline = None
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:  # noqa: BLE001
        pass
    workloads = state["workloads"]
    if isinstance(workloads, dict) and peaks.get("hbm_gbs"):
        for key in ("bop_add", "bop_mul", "bop_add_api"):
            if key in workloads and "value" in workloads[key]:
                workloads[key]["frac_of_measured_hbm"] = workloads[key]["value"] / (peaks["hbm_gbs"] * world)

    state["roofline"] = roofline
    state["done"] = True
    watchdog.cancel()
    emit(compose())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def emit(line):
    """The single JSON line goes to the real stdout; everything else was redirected to stderr."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = os.dup(1)


def main():
    os.dup2(2, 1)      # NCCL prints its version banner to stdout: keep fd 1 for the JSON line only
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="nums_b200", choices=["nums_b200", "reference"])
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--skip-workloads", action="store_true", help="only the headline matmul")
    ap.add_argument("--quick", action="store_true", help="fewer repetitions of the secondary workloads")
    ap.add_argument("--watchdog", type=int, default=420,
                    help="seconds the secondary workloads may take before the headline line is printed without them")
    ap.add_argument("--large", action="store_true", help="config 5 instead: 65536^2 matmul (one-off record run)")
    ap.add_argument("--summa-driver", action="store_true",
                    help="N > 1: the hand-called SummaMatmul driver (round 1) instead of BlockArray.__matmul__ over SpmdSystem")
    ap.add_argument("--mirror", action="store_true", help="drive nums_b200.blocks even when the reference's host layers are installed")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.large:
        run_large(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
