"""Timeline of one end-to-end blocked matmul step (development aid): when each upload lands, when each
launch group of the deferred flush finishes, and where the host spends its time."""
import functools
import os
import sys
import time

import torch

_RealEvent = torch.cuda.Event
torch.cuda.Event = functools.partial(_RealEvent, enable_timing=True)   # every internal event gets a timestamp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from nums_b200.cuda_system import CudaSystem  # noqa: E402

order = sys.argv[1] if len(sys.argv) > 1 else "b_then_a"
system = CudaSystem()
system.init()
a_host, b_host = bench.matmul_blocks_host(pinned=True)
G = bench.GRID


def put_order():
    if order == "b_then_a":
        return [("B", e) for e in sorted(b_host)] + [("A", e) for e in sorted(a_host)]
    if order == "a_then_b":
        return [("A", e) for e in sorted(a_host)] + [("B", e) for e in sorted(b_host)]
    # "wavefront": A row 0, then the columns of B, then the remaining rows of A
    seq = [("A", (0, k)) for k in range(G)]
    for j in range(G):
        seq += [("B", (k, j)) for k in range(G)]
    for i in range(1, G):
        seq += [("A", (i, k)) for k in range(G)]
    return seq


def step(trace):
    from nums_b200.blocks import BlockArray
    from nums_b200.grid import ArrayGrid
    t0 = time.perf_counter()
    start = torch.cuda.Event()
    start.record()
    A = BlockArray(ArrayGrid((bench.N_MATMUL,) * 2, (bench.BLOCK,) * 2, "float64"), system)
    B = BlockArray(ArrayGrid((bench.N_MATMUL,) * 2, (bench.BLOCK,) * 2, "float64"), system)
    for name, e in put_order():
        (A if name == "A" else B).blocks[e].oid = system.put((a_host if name == "A" else b_host)[e])
    t1 = time.perf_counter()
    C = A @ B
    t2 = time.perf_counter()
    ups = [(n, e, (A if n == "A" else B).blocks[e].oid._nums_ready[1]) for n, e in put_order()]
    out = C.get()
    t3 = time.perf_counter()
    if trace:
        print("host: puts %.1f ms, A@B call sequence %.1f ms, get %.1f ms, total %.1f ms"
              % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t3 - t0) * 1e3))
        marks = [ups[i] for i in (0, 7, 15, 63, 64, 71, 127)]
        for n, e, ev in marks:
            print("  upload %s%s landed at %.1f ms" % (n, e, start.elapsed_time(ev)))
        seen = {}
        for e in sorted(C.grid.get_entry_iterator()):
            ev = getattr(system.contractions.resolve(C.blocks[e].oid), "_nums_done", None)
            if ev is not None and id(ev) not in seen:
                seen[id(ev)] = True
                print("  launch group ending with C%s done at %.1f ms" % (e, start.elapsed_time(ev)))
    return out, t3 - t0


step(False)
step(False)
_, dt = step(True)
flops = 2.0 * bench.N_MATMUL ** 3
print("order=%s e2e %.1f ms = %.2f TFLOP/s" % (order, dt * 1e3, flops / dt / 1e12))
