"""Host logic of the multi-GPU drivers on CPU: world_size = 2 and 4 over gloo.

The drivers in nums_b200/multi_gpu.py only talk to a `system` (kernel interface) and to
torch.distributed, so here they run over the oracle-backed CPU system with NumPy blocks: SUMMA
ownership / broadcast schedule, the TSQR tree and the Newton all-reduce are checked against a
single-process NumPy computation.
"""
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import canon_r, rel_fro


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _np_grad_hess_blocks(xs, ys, beta):
    return sum(_np_grad_hess(x, y, beta) for x, y in zip(xs, ys))


def _np_grad_hess(X, y, beta):
    mu = 1.0 / (1.0 + np.exp(-(X @ beta)))
    g = X.T @ (mu - y)
    H = X.T @ ((mu * (1.0 - mu))[:, None] * X)
    return np.concatenate([g, H.reshape(-1)])


def _np_newton_step(gh, beta):
    d = beta.shape[0]
    g, H = gh[:d], gh[d:].reshape(d, d)
    return beta - np.linalg.solve(H, g), np.array([np.max(np.abs(g)), 0.0])


def _worker(rank, world, port, results):
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        from nums_b200 import multi_gpu
        from oracle.cpu_system import OracleSystem
        system = OracleSystem()
        comm = multi_gpu.Comm()
        assert comm.rank == rank and comm.world == world
        out = {}

        # ---- SUMMA: 4 x 4 grid of 16 x 16 blocks ------------------------------------------------
        g, bs = 4, 16
        rng = np.random.default_rng(100)
        A = rng.standard_normal((g * bs, g * bs))
        B = rng.standard_normal((g * bs, g * bs))
        summa = multi_gpu.SummaMatmul(system, comm, g, bs, np.empty((1,), dtype=np.float64))
        blk = lambda M, i, j: np.ascontiguousarray(M[i * bs:(i + 1) * bs, j * bs:(j + 1) * bs])  # noqa: E731
        mine_a = {(i, k): blk(A, i, k) for i in range(g) for k in range(g) if summa.owner_a(i, k) == rank}
        mine_b = {(k, j): blk(B, k, j) for k in range(g) for j in range(g) if summa.owner_b(k, j) == rank}
        c = summa.run(mine_a, mine_b)
        assert sorted(c) == sorted((i, j) for i in range(g) for j in range(g) if summa.owner_c(i, j) == rank)
        out["summa"] = {e: v for e, v in c.items()}
        out["summa_counts"] = (len(mine_a), len(mine_b), len(c))

        # ---- TSQR tree: 8 row blocks of 40 x 6, dealt round-robin ------------------------------------
        X = np.random.default_rng(101).standard_normal((8 * 40, 6))
        mine = [X[i * 40:(i + 1) * 40] for i in range(8) if i % world == rank]
        out["tsqr_r"] = multi_gpu.tsqr_r_tree(system, comm, mine, 6)

        # ---- Newton LR: 8 row blocks of 50 x 5 ----------------------------------------------------------
        rng = np.random.default_rng(102)
        Xl = rng.standard_normal((400, 5))
        yl = (rng.random(400) < 0.5).astype(np.float64)
        xs = [Xl[i * 50:(i + 1) * 50] for i in range(8) if i % world == rank]
        ys = [yl[i * 50:(i + 1) * 50] for i in range(8) if i % world == rank]
        beta, iters = multi_gpu.newton_lr(system, comm, xs, ys, 5, 1e-10, 6, _np_grad_hess_blocks)
        out["beta"], out["iters"] = np.asarray(beta), iters
        # the same iteration with the update fused into one `step` call (NumPy statement of newton_step)
        beta_f, iters_f = multi_gpu.newton_lr(system, comm, xs, ys, 5, 1e-10, 6, _np_grad_hess_blocks, step=_np_newton_step)
        assert iters_f == iters and np.allclose(np.asarray(beta_f), np.asarray(beta), rtol=1e-12, atol=1e-14)
        results[rank] = out
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_drivers_over_gloo(world):
    manager = mp.Manager()
    results = manager.dict()
    mp.spawn(_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    assert sorted(results.keys()) == list(range(world))
    g, bs = 4, 16
    rng = np.random.default_rng(100)
    A = rng.standard_normal((g * bs, g * bs))
    B = rng.standard_normal((g * bs, g * bs))
    C = A @ B
    seen = set()
    for rank in range(world):
        for (i, j), v in results[rank]["summa"].items():
            assert (i, j) not in seen
            seen.add((i, j))
            assert rel_fro(v, C[i * bs:(i + 1) * bs, j * bs:(j + 1) * bs]) < 1e-12
    assert len(seen) == g * g
    # every rank holds 1/world of each operand and of the result
    for rank in range(world):
        assert results[rank]["summa_counts"] == (g * g // world,) * 3
    X = np.random.default_rng(101).standard_normal((8 * 40, 6))
    R = np.linalg.qr(X, mode="r")
    for rank in range(world):
        assert rel_fro(canon_r(results[rank]["tsqr_r"]), canon_r(R)) < 1e-12
    rng = np.random.default_rng(102)
    Xl = rng.standard_normal((400, 5))
    yl = (rng.random(400) < 0.5).astype(np.float64)
    beta = np.zeros(5)
    for _ in range(results[0]["iters"]):
        gh = _np_grad_hess(Xl, yl, beta)
        beta = beta - np.linalg.inv(gh[5:].reshape(5, 5)) @ gh[:5]
    for rank in range(world):
        assert rel_fro(results[rank]["beta"], beta) < 1e-10
        assert results[rank]["iters"] == results[0]["iters"]


def test_device_grid_and_owner_rules():
    from nums_b200 import multi_gpu
    from nums_b200.cuda_system import CudaSystem
    assert [multi_gpu.device_grid(n) for n in (1, 2, 4, 8)] == [(1, 1), (1, 2), (2, 2), (2, 4)]
    flat = CudaSystem(world_size=8, placement="flat")
    assert [flat.owner((i, 0), (16, 1)) for i in range(10)] == [0, 1, 2, 3, 4, 5, 6, 7, 0, 1]
    cyc = CudaSystem(world_size=8, device_grid=(2, 4), placement="cyclic")
    assert cyc.owner((3, 6), (8, 8)) == (3 % 2) * 4 + (6 % 4)      # schedulers.py:170-191
    assert multi_gpu.local_entries((8,), 1, 4) == [(1,), (5,)]
