"""Times grouped DMMA launches shaped like one SUMMA step (development aid)."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nums_b200.cuda_system import CudaSystem

system = CudaSystem(); system.init()
dev = torch.device("cuda", 0)
bs = 2048
shape = (bs, bs)

def run(nprob, nterm, with_cin, distinct_b):
    A = [torch.randn(shape, dtype=torch.float64, device=dev) for _ in range(nprob * nterm if not distinct_b else (nprob // 4 + 1) * nterm)]
    B = [torch.randn(shape, dtype=torch.float64, device=dev) for _ in range(4 * nterm)]
    prev = [torch.randn(shape, dtype=torch.float64, device=dev) for _ in range(nprob)] if with_cin else None
    def build():
        outs = []
        for p in range(nprob):
            acc = prev[p] if with_cin else None
            for t in range(nterm):
                a = A[(p // 4) * nterm + t] if distinct_b else A[p * nterm + t]
                dot = system.bop("tensordot", a, B[(p % 4) * nterm + t], shape, shape, False, False, axes=1, syskwargs={})
                acc = dot if acc is None else system.bop("add", acc, dot, shape, shape, False, False, axes=None, syskwargs={})
            outs.append(acc)
        return outs
    for _ in range(2):
        keep = build(); system.flush()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        keep = build()                      # host work first: only the launch is timed
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); system.flush(); e.record(); e.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    flops = nprob * nterm * 2.0 * bs ** 3
    print("nprob=%2d nterm=%d cin=%d sharedAB=%d: %.3f ms  %.2f TFLOP/s" % (nprob, nterm, with_cin, distinct_b, ts[2], flops / ts[2] / 1e9), flush=True)

cases = [(16, 1, 0, 1), (16, 1, 1, 1), (16, 8, 0, 1), (64, 8, 0, 0)]
if len(sys.argv) > 1 and sys.argv[1] == "summa":
    # the two launches of one SUMMA product per rank at N = 8 (8 problems), N = 4 (16) and N = 2 (32):
    # (problems, terms, addend)   expected pro rata from the 64 x 8 launch
    cases = [(32, 1, 0, 1), (32, 7, 1, 1), (8, 7, 1, 1)]
for args in cases:
    run(*args)
