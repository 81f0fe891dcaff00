"""Multi-GPU parity check of the plugin path, run under torchrun (one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
        scripts/spmd_gpu_check.py [--out gpurun_out/spmd_check_nN.json]

Every rank runs the reference's unmodified host layers (BlockArray / ArrayApplication / glms.newton) over
``SpmdSystem(local = CudaSystem)``; the results -- C = A @ B, X^T X, elementwise, reductions, R and Q of TSQR,
beta of Newton LR, argmax / where -- are fetched on every rank and compared with NumPy (the reference's
numpy_compute arithmetic) at the BASELINE.json tolerances.  Rank 0 prints one JSON object; exit code 1 on a
parity failure on any rank.  Used by tests/test_gpu_spmd.py when the box has >= 2 GPUs.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def canon_r(R):
    s = np.sign(np.diag(R)).copy()
    s[s == 0] = 1
    return R * s[:, None]


def rel(got, want):
    return float(np.linalg.norm(np.asarray(got, dtype=np.float64) - want) / max(np.linalg.norm(want), 1e-300))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    os.environ.setdefault("NUMS_SPMD_CHECK", "1")
    import torch
    import torch.distributed as dist
    from nums_b200 import reference_compat
    from nums_b200._lib import LIB
    app = reference_compat.cuda_app()
    system = app.system
    rank, world = system.rank, system.world_size
    assert world > 1 and dist.get_backend() == "nccl"
    from nums_b200.spmd import REPLICATED
    launches0 = LIB.dll.nums_launch_count()
    rng = np.random.default_rng(11)
    checks = {}

    def distributed(arr, block_shape):
        """BlockArray whose blocks live on their owners (put replicates; adding 0 runs on the owner)."""
        return app.array(arr, block_shape) + app.zero

    # elementwise (config 1 shape class): shard-local
    u, v = rng.random(1 << 20), rng.random(1 << 20)
    U, V = distributed(u, (1 << 17,)), distributed(v, (1 << 17,))
    moved = system.stats["moved_bytes"]
    checks["add_exact"] = bool(np.array_equal((U + V).get(), u + v))
    checks["mul_exact"] = bool(np.array_equal((U * V).get(), u * v))
    checks["elementwise_moved_bytes"] = system.stats["moved_bytes"] - moved

    # blocked matmul (config 2 shape class): 2-D grid, operands exchanged, grouped DMMA launch per rank
    n, bs = 2048, 256
    A, B = rng.standard_normal((n, n)), rng.standard_normal((n, n))
    Ab, Bb = distributed(A, (bs, bs)), distributed(B, (bs, bs))
    C = Ab @ Bb
    pr, pc = system.device_grid
    checks["matmul_homes_ok"] = all(C.blocks[i, j].oid.home == (i % pr) * pc + (j % pc)
                                    for (i, j) in C.grid.get_entry_iterator())
    checks["matmul_rel"] = rel(C.get(), A @ B)
    system.evict_copies()
    checks["matmul_again_rel"] = rel((Ab @ Bb).get(), A @ B)
    checks["gram_rel"] = rel((Ab.T @ Ab).get(), A.T @ A)

    # reductions: local partials + all-reduce
    X = rng.standard_normal((8000, 28))
    Xb = distributed(X, (1000, 28))
    got = app.sum(Xb, axis=0).get()
    checks["sum_axis0_rel"] = float(np.max(np.abs(got - X.sum(axis=0)) / np.abs(X.sum(axis=0))))
    checks["max_exact"] = bool(np.array_equal(app.max(Xb, axis=0).get(), X.max(axis=0)))
    checks["sum_all_rel"] = float(abs(app.sum(Xb).get() - X.sum()) / abs(X.sum()))

    # TSQR (config 3 shape class)
    T = rng.standard_normal((1 << 16, 64))
    Tb = distributed(T, (1 << 13, 64))
    R = app.indirect_tsr(Tb)
    checks["R_replicated"] = R.blocks[0, 0].oid.home == REPLICATED
    checks["R_rel"] = rel(canon_r(R.get()), canon_r(np.linalg.qr(T, mode="r")))
    Q, R2 = app.indirect_tsqr(Tb)
    Qh, R2h = Q.get(), R2.get()
    checks["QR_rel"] = rel(Qh @ R2h, T)
    checks["Q_orth"] = float(np.linalg.norm(Qh.T @ Qh - np.eye(64)))

    # Newton LR (config 4 shape class) through glms.newton
    from nums.core import application_manager
    from nums.models.glms import LogisticRegression, newton
    application_manager.set_instance(app)
    N, d = 1 << 17, 28
    Xl = rng.standard_normal((N, d))
    theta = rng.standard_normal(d) / np.sqrt(d)
    yl = (rng.random(N) < 1.0 / (1.0 + np.exp(-Xl @ theta))).astype(np.float64)
    Xn, yn = distributed(Xl, (N // 8, d)), distributed(yl, (N // 8,))
    model = LogisticRegression(solver="newton", penalty="none")
    model._app = app
    moved = system.stats["moved_bytes"]
    beta = newton(app, model, app.zeros((d,), (d,), dtype=np.float64), Xn, yn, app.scalar(1e-8), 8).get()
    checks["lr_moved_bytes"] = system.stats["moved_bytes"] - moved
    ref = np.zeros(d)
    for _ in range(8):
        mu = 1.0 / (1.0 + np.exp(-(Xl @ ref)))
        g = Xl.T @ (mu - yl)
        H = Xl.T @ ((mu * (1 - mu))[:, None] * Xl)
        ref = ref - np.linalg.inv(H) @ g
        if np.max(np.abs(g)) <= 1e-8:
            break
    checks["beta_rel"] = rel(beta, ref)
    # the same fit through the fused kernels behind glms (one lr_grad_hess per block on its owner, g | H all-reduced)
    from nums_b200 import glms_fused
    moved = system.stats["moved_bytes"]
    beta_f = glms_fused.newton(app, model, app.zeros((d,), (d,), dtype=np.float64), Xn, yn, app.scalar(1e-8), 8).get()
    checks["fused_lr_moved_bytes"] = system.stats["moved_bytes"] - moved
    checks["fused_beta_rel"] = rel(beta_f, ref)

    # carried state / dynamic sizes
    vals = rng.standard_normal(100_000)
    vb = distributed(vals, (12_500,))
    checks["argmax_exact"] = int(app.argop("argmax", vb, axis=0).get()) == int(np.argmax(vals))
    w = app.where(vb > app.scalar(1.5))
    checks["where_exact"] = bool(np.array_equal(w[0].get(), np.where(vals > 1.5)[0]))

    # shuffles (test_np_random.py:46-106): chains of update_block_along_axis -> packed rows, one all-to-all
    M = rng.standard_normal((4096, 96))
    Mb = distributed(M, (512, 32))
    perm0 = np.random.default_rng(5).permutation(4096)
    perm1 = np.random.default_rng(6).permutation(96)[:50]
    moved, rows = system.stats["moved_bytes"], system.stats["scatter_moved_bytes"]
    S0 = Mb._advanced_single_array_subscript((perm0,), axis=0)
    S1 = Mb._advanced_single_array_subscript((perm1,), axis=1)
    checks["shuffle_rows_exact"] = bool(np.array_equal(S0.get(), M[perm0]))
    checks["shuffle_cols_exact"] = bool(np.array_equal(S1.get(), M[:, perm1]))
    checks["shuffle_whole_block_bytes"] = system.stats["moved_bytes"] - moved
    checks["shuffle_row_bytes"] = system.stats["scatter_moved_bytes"] - rows
    checks["shuffle_bytes_bound"] = (4096 * 96 + 4096 * 50) * 8
    ivals = np.arange(20_000, dtype=np.int64)            # (the reference groups index pairs in Python loops: keep it small)
    ib = app.array(ivals, (2_500,)) + app.scalar(0)       # stays int64, blocks on their owners
    permv = np.random.default_rng(9).permutation(20_000)
    checks["shuffle_int_vector_exact"] = bool(np.array_equal(ib[permv].get(), ivals[permv]))

    ok = (checks["add_exact"] and checks["mul_exact"] and checks["elementwise_moved_bytes"] == 0
          and checks["shuffle_rows_exact"] and checks["shuffle_cols_exact"] and checks["shuffle_int_vector_exact"]
          and checks["shuffle_whole_block_bytes"] == 0 and 0 < checks["shuffle_row_bytes"] <= checks["shuffle_bytes_bound"]
          and checks["matmul_homes_ok"] and checks["matmul_rel"] <= 1e-10 and checks["matmul_again_rel"] <= 1e-10
          and checks["gram_rel"] <= 1e-10 and checks["sum_axis0_rel"] <= 1e-12 and checks["max_exact"]
          and checks["sum_all_rel"] <= 1e-12 and checks["R_replicated"] and checks["R_rel"] <= 1e-10
          and checks["QR_rel"] <= 1e-12 and checks["Q_orth"] <= 1e-10 and checks["lr_moved_bytes"] == 0
          and checks["beta_rel"] <= 1e-10 and checks["fused_beta_rel"] <= 1e-10 and checks["fused_lr_moved_bytes"] == 0
          and checks["argmax_exact"] and checks["where_exact"])
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    all_ok = bool(int(flag.item()))
    record = {"world": world, "ok": all_ok, "rank0_checks": checks, "stats": dict(system.stats),
              "kernels_launched_rank0": int(LIB.dll.nums_launch_count() - launches0)}
    if not ok:
        sys.stderr.write("[rank %d] parity failure: %s\n" % (rank, json.dumps(checks)))
    if rank == 0:
        print(json.dumps(record))
        if args.out:
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            with open(args.out, "w") as fh:
                json.dump(record, fh, indent=1)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if all_ok else 1)


if __name__ == "__main__":
    main()
