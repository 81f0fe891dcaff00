"""TEST INFRASTRUCTURE ONLY -- small hot-path scenarios runnable on either host layer.

Each scenario takes an ``api`` adapter (``RefApi`` = the real reference ``ArrayApplication``,
``MirrorApi`` = ``nums_b200.blocks``) and returns a dict of final results as NumPy arrays.  They
are the workloads of BASELINE.json at toy sizes, shaped like the reference's own tests
(tests/core/array/test_bop.py, test_linalg.py, tests/models/test_glms.py), including ragged last
blocks.  ``oracle/make_golden.py`` records the kernel calls the reference issues for them.
"""
import numpy as np


def data(seed, *shape):
    return np.random.default_rng(seed).standard_normal(shape)


def s_elementwise(api):
    u, v = data(1, 1000), data(2, 1000) + 3.0
    U, V = api.array(u, (128,)), api.array(v, (128,))
    X = api.array(data(3, 50, 12), (16, 5))
    col = api.array(data(4, 50, 1), (16, 1))
    row = api.array(data(5, 12), (5,))
    out = {"add": U + V, "mul": U * V, "sub": U - V, "div": U / V, "pow": V ** U,
           "scalar_mul": U * 2.5, "scalar_radd": 1.0 + U, "scalar_rsub": 1.0 - U, "scalar_rdiv": 1.0 / V,
           "neg": -U, "col_bcast": col * X, "row_bcast": X - row, "lt": U < V, "ge": U >= 0.25}
    return {k: api.get(v) for k, v in out.items()}


def s_matmul(api):
    A, B = data(11, 70, 45), data(12, 45, 52)
    a, b = api.array(A, (32, 16)), api.array(B, (16, 20))
    at = api.array(np.ascontiguousarray(A.T), (16, 32))
    bt = api.array(np.ascontiguousarray(B.T), (20, 16))
    x = api.array(data(13, 45), (16,))
    w = api.array(data(14, 70), (32,))
    w2 = api.array(data(15, 70), (32,))
    out = {"ab": a @ b, "atb": at.T @ b, "abt": a @ bt.T, "atbt": at.T @ bt.T, "ax": a @ x, "atw": a.T @ w,
           "ww": w @ w2, "gram": a.T @ a}
    return {k: api.get(v) for k, v in out.items()}


def s_reduce(api):
    X = api.array(data(21, 37, 22), (10, 8))
    v = api.array(data(22, 1000), (128,))
    out = {}
    for op in ("sum", "min", "max"):
        fn = getattr(api, op)
        out[op + "_all"] = fn(X)
        out[op + "_0"] = fn(X, axis=0)
        out[op + "_1"] = fn(X, axis=1)
        out[op + "_1k"] = fn(X, axis=1, keepdims=True)
        out[op + "_vec"] = fn(v)
    out["sum_T0"] = api.sum(X.T, axis=0)
    out["max_abs"] = api.max(api.abs(v))
    return {k: api.get(v) for k, v in out.items()}


def s_tsqr(api):
    X = api.array(data(31, 2345, 9), (123, 9))          # test_linalg.py:120-127 (last block 8 x 9)
    R = api.indirect_tsr(X)
    Q2, R2 = api.indirect_tsqr(X)
    return {"R": api.get(R), "Q_ind": api.get(Q2), "R_ind": api.get(R2)}


def s_tsqr_direct(api):
    """Results only: the reference splits Q2 with its general block-reshape machinery
    (application.py:896-917), the mirror with one create_block per row block -- same data, same
    arithmetic kernels (qr, qr, tensordot), different copy kernels."""
    X = api.array(data(31, 2345, 9), (123, 9))
    Q3, R3 = api.direct_tsqr(X)
    return {"Q_dir": api.get(Q3), "R_dir": api.get(R3)}


def lr_problem(n=1000, d=8, seed=41):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, d))
    theta = rng.standard_normal(d) / np.sqrt(d)
    y = (rng.random(n) < 1.0 / (1.0 + np.exp(-X @ theta))).astype(np.float64)
    return X, y


def s_newton(api):
    Xn, yn = lr_problem()
    X, y = api.array(Xn, (256, 8)), api.array(yn, (256,))
    beta = api.newton(X, y, tol=1e-8, max_iter=4)
    return {"beta": api.get(beta)}


SCENARIOS = {"elementwise": s_elementwise, "matmul": s_matmul, "reduce": s_reduce, "tsqr": s_tsqr,
             "tsqr_direct": s_tsqr_direct, "newton": s_newton}
RESULTS_ONLY = {"tsqr_direct"}


class MirrorApi(object):
    """``nums_b200.blocks`` behind the scenario interface."""

    def __init__(self, system):
        from nums_b200 import blocks
        self.blocks = blocks
        self.app = blocks.ArrayApp(system)

    def array(self, arr, block_shape): return self.app.array(arr, block_shape)
    def get(self, x): return x.get()
    def sum(self, x, axis=None, keepdims=False): return self.app.sum(x, axis, keepdims)
    def min(self, x, axis=None, keepdims=False): return self.app.min(x, axis, keepdims)
    def max(self, x, axis=None, keepdims=False): return self.app.max(x, axis, keepdims)
    def abs(self, x): return self.app.abs(x)
    def indirect_tsr(self, X): return self.app.indirect_tsr(X)
    def indirect_tsqr(self, X): return self.app.indirect_tsqr(X)
    def direct_tsqr(self, X): return self.app.direct_tsqr(X)

    def newton(self, X, y, tol, max_iter):
        model = self.blocks.LogisticRegression(self.app)
        beta = self.app.zeros((X.shape[1],), (X.block_shape[1],), X.dtype)
        beta, _ = self.blocks.newton(self.app, model, beta, X, y, self.app.scalar(tol), max_iter)
        return beta


class RefApi(object):
    """The real reference ``ArrayApplication`` (build container only)."""

    def __init__(self, app):
        self.app = app

    def array(self, arr, block_shape): return self.app.array(arr, block_shape=block_shape)
    def get(self, x): return x.get()
    def sum(self, x, axis=None, keepdims=False): return self.app.sum(x, axis=axis, keepdims=keepdims)
    def min(self, x, axis=None, keepdims=False): return self.app.min(x, axis=axis, keepdims=keepdims)
    def max(self, x, axis=None, keepdims=False): return self.app.max(x, axis=axis, keepdims=keepdims)
    def abs(self, x): return self.app.abs(x)
    def indirect_tsr(self, X): return self.app.indirect_tsr(X)
    def indirect_tsqr(self, X): return self.app.indirect_tsqr(X)
    def direct_tsqr(self, X): return self.app.direct_tsqr(X)

    def newton(self, X, y, tol, max_iter):
        from nums.core import application_manager
        from nums.models import glms
        application_manager.set_instance(self.app)
        model = glms.LogisticRegression(solver="newton")
        beta = self.app.zeros((X.shape[1],), (X.block_shape[1],), dtype=X.dtype)
        return glms.newton(self.app, model, beta, X, y, self.app.scalar(tol), max_iter)
