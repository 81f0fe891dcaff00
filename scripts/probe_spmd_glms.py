"""Per-iteration cost of nums_b200.glms_fused.newton over SpmdSystem (torchrun, development aid): multi-block
launches on / off, host enqueue vs wall time, cProfile of rank 0."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from nums_b200.host import HostLayers
    from nums_b200 import glms_fused
    host = HostLayers()
    system, app = host.system, host.app
    rank = system.rank
    dev = torch.device("cuda", torch.cuda.current_device())
    N, d, G = 11_000_000, 28, 8
    def distributed(shape, block_shape, fill):
        ba = host.blockarray(shape, block_shape, "float64")
        gshape = ba.grid.grid_shape
        for entry in ba.grid.get_entry_iterator():
            bshape = ba.grid.get_block_shape(entry)
            mine = system.owner(entry, gshape) == rank
            ba.blocks[entry].oid = system.put_at(fill(bshape) if mine else None, entry, gshape, shape=bshape, dtype=np.float64)
        return ba
    X = distributed((N, d), (N // G, d), lambda s: torch.randn(s, dtype=torch.float64, device=dev))
    y = distributed((N,), (N // G,), lambda s: (torch.rand(s, dtype=torch.float64, device=dev) < 0.5).to(torch.float64))
    model = host.logistic_model()
    iters = 20

    def run():
        beta0 = app.zeros((d,), (d,), np.float64)
        host.launch(glms_fused.newton(app, model, beta0, X, y, app.scalar(0.0), iters))

    for label, drop in (("multi", False), ("per-block", True)):
        saved = None
        if drop:
            saved = system.methods.pop("lr_grad_hess_multi", None)
        run()
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        run()
        t_host = time.perf_counter() - t0
        torch.cuda.synchronize()
        t_all = time.perf_counter() - t0
        if rank == 0:
            print("%s: host enqueue %.1f us/iter, wall %.1f us/iter" % (label, t_host / iters * 1e6, t_all / iters * 1e6), flush=True)
        if rank == 0 and not drop:
            pr = cProfile.Profile(); pr.enable(); run(); pr.disable()
            torch.cuda.synchronize()
            pstats.Stats(pr, stream=sys.stdout).sort_stats("tottime").print_stats(14)
        else:
            run(); torch.cuda.synchronize()
        dist.barrier()
        if saved is not None:
            system.methods["lr_grad_hess_multi"] = saved
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
