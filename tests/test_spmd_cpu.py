"""``SpmdSystem`` on CPU: world_size 2 and 4 over gloo, the REFERENCE's unmodified host layers on top.

Every rank runs the same program -- ``BlockArray`` operators, ``ArrayApplication.indirect_tsqr``,
``glms.newton`` from the reference -- over ``SpmdSystem(local = oracle system)`` (NumPy blocks), exactly as
the GPU ranks do over ``SpmdSystem(local = CudaSystem)``.  Checked on every rank against single-process
NumPy: results, which rank ran what (placement), and that cross-rank sums went through all-reduces rather
than block shipping.  NUMS_SPMD_CHECK=1 asserts that the inferred result shapes / dtypes equal the actual ones.
"""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from nums_b200 import reference_compat
from tests.helpers import canon_r, rel_fro

pytestmark = pytest.mark.skipif(not reference_compat.available(), reason="reference not present")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _app():
    """The reference's ArrayApplication over SpmdSystem(oracle system)."""
    reference_compat.load_reference()
    from nums.core.array.application import ArrayApplication
    from nums.core.systems.filesystem import FileSystem
    from nums.core.systems.systems import SerialSystem
    from nums_b200.spmd import SpmdSystem
    from oracle import np_oracle
    from oracle.cpu_system import OracleSystem

    class OracleSpmdSystem(SpmdSystem, SerialSystem):
        pass

    local = OracleSystem()
    # NumPy statements of the optional fused-LR kernels (cuda_compute.EXTRA_KERNELS), so that nums_b200.glms_fused
    # and its grouping of row blocks by owner run over gloo as they do over NCCL
    def lr_grad_hess(X, y, beta):
        mu = 1.0 / (1.0 + np.exp(-(X @ beta)))
        return np.concatenate([X.T @ (mu - y), (X.T @ ((mu * (1.0 - mu))[:, None] * X)).ravel()])

    def lr_grad_hess_multi(*blocks):
        beta = blocks[-1]
        return sum(lr_grad_hess(blocks[i], blocks[i + 1], beta) for i in range(0, len(blocks) - 1, 2))

    def newton_step(gh, beta):
        d = beta.shape[0]
        g, H = gh[:d], gh[d:].reshape(d, d)
        return beta - np.linalg.inv(H) @ g, np.array([np.max(np.abs(g)), 0.0])
    for fn in (lr_grad_hess, lr_grad_hess_multi, newton_step):
        setattr(local.imp, fn.__name__, fn)
    system = OracleSpmdSystem(local, check=True)
    system.rng_cls = np_oracle.RNG
    system.init()
    return ArrayApplication(system=system, filesystem=FileSystem(system))


def _worker(rank, world, port, results):
    os.environ["NUMS_SPMD_CHECK"] = "1"
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        from nums_b200.spmd import REPLICATED
        app = _app()
        system = app.system
        out = {}
        rng = np.random.default_rng(7)

        # ---- elementwise on row-blocked vectors: shard-local, no traffic ----------------------------------
        u, v = rng.random(4000), rng.random(4000)
        U, V = app.array(u, (500,)), app.array(v, (500,))           # put -> replicated inputs
        before = dict(system.stats)
        W = U + V
        homes = [W.blocks[i].oid.home for i in range(8)]
        assert homes == [i % world for i in range(8)], homes          # flattened grid entry mod world
        Z = W * U
        assert [Z.blocks[i].oid.home for i in range(8)] == homes      # compute where the operands live
        assert system.stats["moves"] == before["moves"]
        out["add"] = W.get()
        out["mul"] = Z.get()
        executed = system.stats["executed"] - before["executed"]
        assert executed == 2 * len([i for i in range(8) if i % world == rank])

        # ---- reductions: per-block partials where the blocks are, cross-block stage = all-reduce ------------
        X = rng.standard_normal((800, 12))
        Xb = app.array(X, (100, 12)) + app.zero                      # distribute (results are owned, not replicated)
        before = dict(system.stats)
        out["sum0"] = app.sum(Xb, axis=0).get()
        out["sum_all"] = app.sum(Xb).get()
        out["max0"] = app.max(Xb, axis=0).get()
        out["mean1"] = app.mean(Xb, axis=1).get()
        assert system.stats["all_reduces"] > before["all_reduces"]

        # ---- blocked matmul, 2-D grid: owner-computes on the device grid, operands exchanged in batches -----
        A, B = rng.standard_normal((512, 384)), rng.standard_normal((384, 448))
        Ab, Bb = app.array(A, (128, 128)) + app.zero, app.array(B, (128, 128)) + app.zero
        C = Ab @ Bb
        pr, pc = system.device_grid
        for (i, j) in C.grid.get_entry_iterator():
            assert C.blocks[i, j].oid.home == (i % pr) * pc + (j % pc)
        out["matmul"] = C.get()
        out["gram"] = (Ab.T @ Ab).get()

        # ---- TSQR: local R per block, stacked-R qr as a tree over the ranks, replicated R ---------------------
        T = rng.standard_normal((1600, 16))
        Tb = app.array(T, (200, 16)) + app.zero
        R = app.indirect_tsr(Tb)
        assert R.blocks[0, 0].oid.home == REPLICATED
        out["R"] = R.get()
        Q, R2 = app.indirect_tsqr(Tb)
        assert [Q.blocks[i, 0].oid.home for i in range(8)] == [i % world for i in range(8)]
        out["Q"], out["R2"] = Q.get(), R2.get()
        Qd, Rd = app.direct_tsqr(Tb)
        out["Qd"], out["Rd"] = Qd.get(), Rd.get()

        # ---- Newton logistic regression through glms: g and H all-reduced, beta replicated --------------------
        from nums.core import application_manager
        from nums.models.glms import LogisticRegression, newton
        application_manager.set_instance(app)
        n, d = 1600, 8
        Xl = rng.standard_normal((n, d))
        theta = rng.standard_normal(d) / np.sqrt(d)
        yl = (rng.random(n) < 1.0 / (1.0 + np.exp(-Xl @ theta))).astype(np.float64)
        Xn, yn = app.array(Xl, (200, d)) + app.zero, app.array(yl, (200,)) + app.zero
        model = LogisticRegression(solver="newton", penalty="none")
        model._app = app
        before = dict(system.stats)
        beta = newton(app, model, app.zeros((d,), (d,), dtype=np.float64), Xn, yn, app.scalar(1e-10), 6)
        assert beta.blocks[0].oid.home == REPLICATED
        out["beta"] = beta.get()
        out["lr_moved_bytes"] = system.stats["moved_bytes"] - before["moved_bytes"]
        out["lr_all_reduces"] = system.stats["all_reduces"] - before["all_reduces"]
        # the same fit through the drop-in on the fused kernels: one multi-block call per owner, one all-reduce / iteration
        from nums_b200 import glms_fused
        before = dict(system.stats)
        executed = system.stats["executed"]
        beta_f = glms_fused.newton(app, model, app.zeros((d,), (d,), dtype=np.float64), Xn, yn, app.scalar(1e-10), 6)
        out["beta_fused"] = beta_f.get()
        out["fused_moved_bytes"] = system.stats["moved_bytes"] - before["moved_bytes"]
        out["fused_all_reduces"] = system.stats["all_reduces"] - before["all_reduces"]

        # ---- dynamic-size / carried-state kernels (description broadcast from the executing rank) --------------
        vals = rng.standard_normal(1000)
        vb = app.array(vals, (125,)) + app.zero
        out["argmax"] = int(app.argop("argmax", vb, axis=0).get())
        out["where"] = [w.get() for w in app.where(vb > app.scalar(0.5))]
        rs = app.random_state(1337)
        out["random"] = rs.random((64, 4), (16, 4)).get()
        # ---- shuffles (test_np_random.py:46-106): update_block_along_axis chains -> packed rows, one all-to-all ----
        M = rng.standard_normal((96, 40))
        Mb = app.array(M, (12, 10)) + app.zero                       # 8 x 4 blocks living on their owners
        perm0 = np.random.default_rng(5).permutation(96)
        perm1 = np.random.default_rng(6).permutation(40)[:25]
        before = dict(system.stats)
        S0 = Mb._advanced_single_array_subscript((perm0,), axis=0)
        S1 = Mb._advanced_single_array_subscript((perm1,), axis=1)
        system.flush()
        out["shuffle_homes"] = [S0.blocks[e].oid.home for e in S0.grid.get_entry_iterator()]   # (get() replicates small blocks)
        out["shuffle0"], out["shuffle1"] = S0.get(), S1.get()
        out["_M"] = M
        out["shuffle_whole_block_moves"] = system.stats["moved_bytes"] - before["moved_bytes"]
        out["shuffle_row_bytes"] = system.stats["scatter_moved_bytes"] - before["scatter_moved_bytes"]
        out["shuffle_exchanges"] = system.stats["scatter_exchanges"] - before["scatter_exchanges"]
        out["shuffle_owners"] = [system.owner(e, S0.grid.grid_shape) for e in S0.grid.get_entry_iterator()]
        Vb = app.array(vals, (125,)) + app.zero
        out["shuffle_vec"] = Vb[np.random.default_rng(8).permutation(1000)].get()
        out["shuffle_then_sum"] = app.sum(S0 + S0, axis=0).get()      # a pending chain consumed by another kernel
        # a shuffle of a shuffle with nothing materialised in between, other block sizes, and a 3-D array along every axis
        perm2 = np.random.default_rng(9).permutation(96)[:50]
        out["shuffle_twice"] = Mb._advanced_single_array_subscript((perm0,), axis=0) \
            ._advanced_single_array_subscript((perm2,), block_size=7, axis=0).get()
        T3 = rng.standard_normal((12, 10, 9))
        T3b = app.array(T3, (4, 5, 3)) + app.zero
        out["_T3"] = T3
        out["shuffle_3d"] = [T3b._advanced_single_array_subscript((np.random.default_rng(20 + ax).permutation(T3.shape[ax]),),
                                                                  axis=ax).get() for ax in range(3)]
        out["stats"] = dict(system.stats)
        results[rank] = out
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_reference_host_layers_over_spmd(world):
    manager = mp.Manager()
    results = manager.dict()
    mp.spawn(_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    assert sorted(results.keys()) == list(range(world))
    rng = np.random.default_rng(7)
    u, v = rng.random(4000), rng.random(4000)
    X = rng.standard_normal((800, 12))
    A, B = rng.standard_normal((512, 384)), rng.standard_normal((384, 448))
    T = rng.standard_normal((1600, 16))
    n, d = 1600, 8
    Xl = rng.standard_normal((n, d))
    theta = rng.standard_normal(d) / np.sqrt(d)
    yl = (rng.random(n) < 1.0 / (1.0 + np.exp(-Xl @ theta))).astype(np.float64)
    vals = rng.standard_normal(1000)
    beta = np.zeros(d)
    for _ in range(6):
        mu = 1.0 / (1.0 + np.exp(-(Xl @ beta)))
        g = Xl.T @ (mu - yl)
        H = Xl.T @ ((mu * (1 - mu))[:, None] * Xl)
        beta = beta - np.linalg.inv(H) @ g
        if np.max(np.abs(g)) <= 1e-10:
            break
    Rref = np.linalg.qr(T, mode="r")
    for rank in range(world):
        r = results[rank]
        assert np.array_equal(r["add"], u + v) and np.array_equal(r["mul"], (u + v) * u)
        assert np.allclose(r["sum0"], X.sum(axis=0), rtol=1e-12, atol=1e-12)
        assert np.allclose(r["sum_all"], X.sum(), rtol=1e-12, atol=1e-12)
        assert np.array_equal(r["max0"], X.max(axis=0))
        assert np.allclose(r["mean1"], X.mean(axis=1), rtol=1e-12, atol=1e-13)
        assert rel_fro(r["matmul"], A @ B) < 1e-12 and rel_fro(r["gram"], A.T @ A) < 1e-12
        assert rel_fro(canon_r(r["R"]), canon_r(Rref)) < 1e-11
        assert rel_fro(r["Q"] @ r["R2"], T) < 1e-12 and rel_fro(r["Qd"] @ r["Rd"], T) < 1e-12
        assert np.linalg.norm(r["Q"].T @ r["Q"] - np.eye(16)) < 1e-10
        assert rel_fro(r["beta"], beta) < 1e-9
        assert rel_fro(r["beta_fused"], beta) < 1e-9
        assert r["fused_moved_bytes"] == 0 and 1 <= r["fused_all_reduces"] <= 7
        assert r["argmax"] == int(np.argmax(vals))
        assert len(r["where"]) == 1 and np.array_equal(r["where"][0], np.where(vals > 0.5)[0])
        assert np.array_equal(r["random"], results[0]["random"])          # same stream on every rank
        assert np.array_equal(r["shuffle0"], r["_M"][np.random.default_rng(5).permutation(96)])
        assert np.array_equal(r["shuffle1"], r["_M"][:, np.random.default_rng(6).permutation(40)[:25]])
        assert np.array_equal(r["shuffle_vec"], vals[np.random.default_rng(8).permutation(1000)])
        assert np.allclose(r["shuffle_then_sum"], 2 * r["_M"].sum(axis=0), rtol=1e-12, atol=1e-12)
        assert np.array_equal(r["shuffle_twice"], r["_M"][np.random.default_rng(5).permutation(96)][
            np.random.default_rng(9).permutation(96)[:50]])
        for ax in range(3):
            idx = [slice(None)] * 3
            idx[ax] = np.random.default_rng(20 + ax).permutation(r["_T3"].shape[ax])
            assert np.array_equal(r["shuffle_3d"][ax], r["_T3"][tuple(idx)])
        assert r["shuffle_homes"] == r["shuffle_owners"]                  # destination blocks live on their owners
        assert r["shuffle_whole_block_moves"] == 0                        # no source block travelled whole
        assert 0 < r["shuffle_row_bytes"] <= (96 * 40 + 96 * 25) * 8      # every row crossed the links at most once
        assert r["shuffle_exchanges"] <= 3
        # Newton LR: nothing but all-reduces crosses ranks (X and y never move)
        assert r["lr_moved_bytes"] == 0, r["lr_moved_bytes"]
        assert r["lr_all_reduces"] >= 6
