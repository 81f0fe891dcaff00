"""GPU replay of the reference's own kernel calls (tests/golden/, recorded by oracle/make_golden.py).

* every de-duplicated ``ComputeCls`` call the real reference made while running the hot-path
  scenarios and a sweep of its host APIs is replayed through ``CudaSystem`` -> ``cuda_compute`` ->
  the C ABI and compared with the recorded NumPy result;
* the scenarios themselves are run end to end through ``nums_b200.blocks`` on the GPU and compared
  with the reference's final arrays.

Bars: bit-exact for copies / indices / comparisons / integer work / IEEE add, sub, mul, div, sqrt;
<= 1e-12 relative for FP64 transcendental elementwise and sums; <= 1e-10 relative Frobenius for
tensordot, qr (R compared after row-sign canonicalisation) and inv.
"""
import numpy as np
import pytest

from oracle import scenarios
from tests.helpers import canon_r, load_golden, rel_fro

pytestmark = pytest.mark.gpu

EXACT_BOPS = {"add", "sub", "mul", "truediv", "lt", "le", "gt", "ge", "eq", "ne", "fmin", "fmax", "maximum",
              "minimum", "subtract", "multiply", "true_divide"}
EXACT_UOPS = {"abs", "absolute", "sqrt", "negative", "square", "sign", "floor", "ceil", "isnan"}


def to_device(system, x):
    if isinstance(x, np.ndarray):
        return system.put(x)
    if isinstance(x, np.generic):
        return system.put(np.asarray(x))
    return x


def from_device(system, x):
    import torch
    if isinstance(x, torch.Tensor):
        return system.get(x)
    if isinstance(x, tuple):   # also namedtuples (np.linalg.qr's QRResult under the fake system)
        return tuple(from_device(system, v) for v in x)
    if isinstance(x, list):
        return [from_device(system, v) for v in x]
    return x


def exact(got, want, ctx):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, (ctx, got.shape, want.shape)
    assert got.dtype == want.dtype, (ctx, got.dtype, want.dtype)
    assert np.array_equal(got, want, equal_nan=got.dtype.kind == "f"), ctx


def close(got, want, ctx, tol=1e-12, scale=None):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape and got.dtype == want.dtype, (ctx, got.shape, want.shape, got.dtype, want.dtype)
    if got.dtype.kind != "f":
        assert np.array_equal(got, want), ctx
        return
    if got.dtype == np.float32:
        tol = max(tol, 1e-5)
    ref = np.abs(want) if scale is None else scale
    assert np.array_equal(np.isnan(got), np.isnan(want)), ctx
    ok = np.isfinite(want)
    assert (np.abs(got[ok] - want[ok]) <= tol * np.maximum(np.broadcast_to(ref, want.shape)[ok], 1e-300)).all(), \
        (ctx, np.abs(got[ok] - want[ok]).max())


def bound_arguments(name, args, kwargs):
    """Parameter name -> value, however the reference passed it (positionally or by keyword)."""
    import inspect
    from oracle.np_oracle import OracleCompute
    sig = inspect.signature(getattr(OracleCompute(), name))
    try:
        return dict(sig.bind(*args, **kwargs).arguments)
    except TypeError:
        return {}


def check_call(system, c):
    name, args, kwargs, want = c["name"], c["args"], c["kwargs"], c["result"]
    named = bound_arguments(name, args, kwargs)
    dargs = [to_device(system, a) for a in args]
    dkw = {k: to_device(system, v) for k, v in kwargs.items()}
    got = from_device(system, getattr(system, name)(*dargs, syskwargs={}, **dkw))
    ctx = (name, str(args[0])[:40] if args else "")
    if name in ("empty",) or (name == "new_block" and named["op_name"] == "empty"):
        assert np.asarray(got).shape == want.shape and np.asarray(got).dtype == want.dtype
    elif name == "bop":
        op = named["op"]
        if op == "tensordot":
            got, want = np.asarray(got), np.asarray(want)
            assert got.shape == want.shape and got.dtype == want.dtype, ctx
            if want.dtype.kind == "f":
                a, b = np.asarray(named["a1"], dtype=np.float64), np.asarray(named["a2"], dtype=np.float64)
                bound = 1e-10 * max(np.linalg.norm(a) * np.linalg.norm(b), 1e-300)
                assert np.linalg.norm(got - want) <= bound, (ctx, np.linalg.norm(got - want), bound)
            else:
                assert np.array_equal(got, want), ctx
        elif op in EXACT_BOPS:
            exact(got, want, ctx)
        else:
            close(got, want, ctx)
    elif name == "map_uop":
        (exact if named["op_name"] in EXACT_UOPS else close)(got, want, ctx)
    elif name == "reduce_axis":
        op, arr, axis, keepdims, transposed = (named[k] for k in ("op_name", "arr", "axis", "keepdims", "transposed"))
        if op == "sum" and np.asarray(want).dtype.kind == "f":
            src = arr.T if transposed else arr
            close(got, want, ctx, scale=np.abs(src).sum(axis=axis, keepdims=keepdims))
        else:
            exact(got, want, ctx)
    elif name == "qr":
        mode = kwargs.get("mode", "reduced")
        if mode == "r":
            assert rel_fro(canon_r(got), canon_r(want)) <= 1e-10, ctx
        else:
            q, r = got
            stacked = np.concatenate(named["arrays"], axis=kwargs["axis"]) if len(named["arrays"]) > 1 else named["arrays"][0]
            assert q.shape == want[0].shape and r.shape == want[1].shape
            assert rel_fro(q @ r, stacked) <= 1e-10, ctx
            assert np.linalg.norm(q.T @ q - np.eye(q.shape[1])) <= 1e-10, ctx
            assert rel_fro(canon_r(r), canon_r(want[1])) <= 1e-10, ctx
    elif name in ("inv", "cholesky"):
        assert rel_fro(got, want) <= 1e-10 * max(1.0, np.linalg.cond(named["arr"])), ctx
    elif name == "where":
        assert tuple(got[-1]) == tuple(want[-1]), ctx
        for g, w in zip(got[:-1], want[:-1]):
            exact(g, w, ctx)
    elif name == "arg_op":
        assert int(got[0]) == int(want[0]) and got[1] == want[1], ctx
    elif name in ("allclose", "logical_and"):
        assert bool(got) == bool(want), ctx
    else:  # copies, creation, astype, sum_reduce, reshape, diag, arange, random_block, update_block*
        exact(got, want, ctx)


def test_replay_reference_kernel_calls(cuda_system):
    calls = load_golden("ref_calls.pkl.gz")
    failures = []
    counts = {}
    for i, c in enumerate(calls):
        try:
            check_call(cuda_system, c)
            counts[c["name"]] = counts.get(c["name"], 0) + 1
        except Exception as exc:  # noqa: BLE001  -- collect everything, report once
            failures.append((i, c["name"], str(c["args"][0])[:30] if c["args"] else "", type(exc).__name__, str(exc)[:200]))
    assert not failures, "%d of %d replayed calls differ; first: %r" % (len(failures), len(calls), failures[:8])
    assert sum(counts.values()) == len(calls)


@pytest.mark.parametrize("name", sorted(scenarios.SCENARIOS))
def test_scenarios_end_to_end(cuda_system, name):
    want = load_golden("ref_scenarios.pkl.gz")[name]["results"]
    api = scenarios.MirrorApi(cuda_system)
    got = scenarios.SCENARIOS[name](api)
    for key, value in want.items():
        g = np.asarray(got[key])
        assert g.shape == value.shape and g.dtype == value.dtype, key
        if name in ("elementwise",) and key != "pow":
            assert np.array_equal(g, value), key
        elif name == "reduce" and not key.startswith("sum"):
            assert np.array_equal(g, value), key
        elif key.startswith("R"):
            assert rel_fro(canon_r(g), canon_r(value)) <= 1e-10, key
        elif key.startswith("Q"):
            # Q is unique up to column signs once R is canonical
            sg = np.sign(np.sum(g * value, axis=0))
            assert rel_fro(g * sg, value) <= 1e-9, key
        else:
            assert rel_fro(g, value) <= 1e-10, (key, rel_fro(g, value))
