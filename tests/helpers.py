"""Shared test helpers: golden loaders and an oracle-backed system (CPU checker only)."""
import gzip
import os
import pickle

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with gzip.open(os.path.join(GOLDEN_DIR, name), "rb") as f:
        return pickle.load(f)


class OracleSystem(object):
    """SerialSystem-shaped object over ``oracle.np_oracle.OracleCompute`` that records every call.

    Test infrastructure: lets the host-side drivers (nums_b200.blocks) run on CPU so that their
    kernel-call sequences can be compared with the reference's recorded ones."""

    def __init__(self):
        from oracle.np_oracle import OracleCompute
        self.imp = OracleCompute()
        self.trace = []

    def put(self, value):
        return np.asarray(value)

    def get(self, oids):
        if isinstance(oids, list):
            return [self.get(o) for o in oids]
        return oids

    def call(self, name, *args, **kwargs):
        from oracle.make_golden import call_signature, freeze
        self.trace.append(call_signature(name, freeze(args), freeze(kwargs)))
        kwargs = {k: v for k, v in kwargs.items() if k != "syskwargs"}
        return getattr(self.imp, name)(*args, **kwargs)

    def __getattr__(self, name):
        if name.startswith("_") or not hasattr(type(self.__dict__.get("imp")), name):
            raise AttributeError(name)
        return lambda *a, **k: self.call(name, *a, **k)


def canon_r(R):
    """Row-sign canonical form of an upper-trapezoidal R (QR is unique up to row signs)."""
    R = np.asarray(R)
    k = min(R.shape)
    s = np.sign(np.diag(R[:k, :k]).copy())
    s[s == 0] = 1
    out = R.copy()
    out[:k] = R[:k] * s[:, None]
    return out


def rel_fro(got, want):
    want = np.asarray(want, dtype=np.float64)
    return np.linalg.norm(np.asarray(got, dtype=np.float64) - want) / max(np.linalg.norm(want), 1e-300)
