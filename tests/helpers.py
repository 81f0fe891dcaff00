"""Shared test helpers: golden loaders and an oracle-backed system (CPU checker only)."""
import gzip
import os
import pickle

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with gzip.open(os.path.join(GOLDEN_DIR, name), "rb") as f:
        return pickle.load(f)


from oracle.cpu_system import OracleSystem as _OracleSystem


def OracleSystem():
    """Recording oracle-backed system (see oracle/cpu_system.py)."""
    return _OracleSystem(record=True)


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
