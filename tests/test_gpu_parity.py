"""Parity of ``cuda_compute.ComputeCls`` (through the C ABI) with the CPU oracle, method by method.

Bars (BASELINE.json north_star): bit-exact for integer / bool / index work and for IEEE
add/sub/mul/div/sqrt; <= 1e-12 relative for FP64 elementwise and reductions; <= 1e-10 relative
Frobenius for FP64 contractions and factorizations.  The inputs mirror the shapes the reference's
own tests use (tests/core/array/test_bop.py, test_basic_ops.py, test_ho_ops.py,
tests/numpy/test_arithmetic.py, test_np_reduction.py, test_linalg.py).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

F64_TOL = 1e-12
GEMM_TOL = 1e-10


def _get(system, x):
    return system.get(x)


def assert_exact(got, want):
    got, want = np.asarray(got), np.asarray(want)
    assert got.dtype == want.dtype, (got.dtype, want.dtype)
    assert got.shape == want.shape, (got.shape, want.shape)
    assert np.array_equal(got, want, equal_nan=got.dtype.kind == "f"), np.abs(got.astype(float) - want.astype(float)).max()


def assert_close(got, want, tol=F64_TOL, abs_scale=0.0):
    """Relative bar `tol`, plus 4 ulp of the operand magnitude `abs_scale` for results that are a
    cancelling sum of O(abs_scale) terms (logaddexp near zero etc.)."""
    got, want = np.asarray(got), np.asarray(want)
    assert got.dtype == want.dtype, (got.dtype, want.dtype)
    assert got.shape == want.shape, (got.shape, want.shape)
    if got.dtype.kind != "f":
        assert np.array_equal(got, want)
        return
    if got.dtype == np.float32:
        tol = max(tol, 2e-6)
    finite = np.isfinite(want)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.array_equal(got[~finite & ~np.isnan(want)], want[~finite & ~np.isnan(want)])
    denom = np.maximum(np.abs(want[finite]), 1e-300)
    err = np.abs(got[finite] - want[finite]) / denom
    small = np.abs(want[finite]) < 1e-290
    slack = 4 * np.finfo(got.dtype).eps * abs_scale / denom
    assert (err[~small] <= tol + slack[~small]).all(), err[~small].max()


def rel_fro(got, want):
    return np.linalg.norm(np.asarray(got, dtype=np.float64) - want) / max(np.linalg.norm(want), 1e-300)


# ----------------------------------------------------------------------------------------------------
# bop: elementwise
# ----------------------------------------------------------------------------------------------------
EXACT_BOPS = ["add", "sub", "mul", "truediv", "lt", "le", "gt", "ge", "eq", "ne", "maximum", "minimum",
              "fmax", "fmin", "copysign", "logical_and", "logical_or", "logical_xor", "heaviside",
              "floor_divide", "remainder", "fmod", "nextafter"]
CLOSE_BOPS = ["pow", "float_power", "arctan2", "hypot", "logaddexp", "logaddexp2", "xlogy"]


def _bop(system, oracle, op, a, b, a_T=False, b_T=False, exact=True):
    a_shape = a.T.shape if a_T else a.shape
    b_shape = b.T.shape if b_T else b.shape
    want = oracle.bop(op, a, b, a_shape, b_shape, a_T, b_T, None)
    got = _get(system, system.bop(op, system.put(a), system.put(b), a_shape, b_shape, a_T, b_T, axes=None,
                                  syskwargs={"grid_entry": (0,), "grid_shape": (1,)}))
    if exact:
        assert_exact(got, want)
    else:
        mags = [np.abs(v[np.isfinite(v)]).max() for v in (np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64))
                if np.isfinite(v).any()]
        assert_close(got, want, abs_scale=max(mags + [0.0]))


@pytest.mark.parametrize("op", EXACT_BOPS + CLOSE_BOPS)
def test_bop_f64_same_shape(cuda_system, oracle, op):
    rng = np.random.default_rng(11)
    for n in (1, 7, 1000, 4096 + 3, 100003):
        a = rng.standard_normal(n)
        b = rng.standard_normal(n)
        if op in ("pow", "float_power", "xlogy"):
            a, b = np.abs(a) + 0.1, b
            if op == "xlogy":
                b = np.abs(b) + 0.1
        _bop(cuda_system, oracle, op, a, b, exact=op in EXACT_BOPS)


def test_bop_special_values(cuda_system, oracle):
    vals = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 1e-320, -1e-320, 1e308, 2.5, -2.5, 3.0])
    a, b = np.meshgrid(vals, vals)
    a, b = np.ascontiguousarray(a.ravel()), np.ascontiguousarray(b.ravel())
    with np.errstate(all="ignore"):
        for op in ["add", "sub", "mul", "truediv", "maximum", "minimum", "fmax", "fmin", "lt", "le", "eq", "ne",
                   "copysign", "heaviside", "floor_divide", "remainder", "fmod", "logical_and", "logical_xor"]:
            _bop(cuda_system, oracle, op, a, b, exact=True)
        for op in ["hypot", "arctan2", "logaddexp", "logaddexp2"]:
            _bop(cuda_system, oracle, op, a, b, exact=False)


@pytest.mark.parametrize("dtypes", [(np.float64, np.int64), (np.int64, np.int64), (np.float64, np.float32),
                                     (np.float32, np.float32), (np.int32, np.int64), (np.int32, np.int32),
                                     (np.bool_, np.bool_), (np.bool_, np.int64), (np.int64, np.float64)])
def test_bop_dtype_matrix(cuda_system, oracle, dtypes):
    rng = np.random.default_rng(5)
    n = 5001
    def make(dt):
        if dt == np.bool_:
            return rng.random(n) < 0.5
        if np.dtype(dt).kind == "i":
            return rng.integers(-50, 50, n).astype(dt)
        return (rng.standard_normal(n) * 10).astype(dt)
    a, b = make(dtypes[0]), make(dtypes[1])
    ops = ["add", "mul", "lt", "ge", "eq", "ne", "maximum", "minimum", "fmax", "fmin", "logical_and", "logical_or"]
    if dtypes != (np.bool_, np.bool_):
        ops += ["sub", "truediv", "floor_divide", "remainder", "fmod"]
    with np.errstate(all="ignore"):
        for op in ops:
            exact = not (np.float32 in dtypes and op in ("truediv",)) or True
            _bop(cuda_system, oracle, op, a, b, exact=exact)
    if all(np.dtype(d).kind in "ib" for d in dtypes):
        for op in ["bitwise_and", "bitwise_or", "bitwise_xor"]:
            _bop(cuda_system, oracle, op, a, b)
    if all(np.dtype(d).kind == "i" for d in dtypes):
        a2, b2 = np.abs(a) + 1, np.abs(b) % 5
        for op in ["gcd", "lcm", "left_shift", "right_shift", "pow"]:
            _bop(cuda_system, oracle, op, a2, b2)


def test_bop_test_arithmetic_vectors(cuda_system, oracle):
    """The operand sets of the reference's tests/numpy/test_arithmetic.py:55-78."""
    pairs = [(np.array([.1, 5.0, .3]), np.array([.2, 6.0, .3])),
             (np.array([.1, 5.0, .3]), np.array([4, 2, 6], dtype=np.int64)),
             (np.array([3, 7, 3], dtype=np.int64), np.array([4, 2, 6], dtype=np.int64))]
    names = ["add", "arctan2", "copysign", "divide", "equal", "float_power", "floor_divide", "fmax", "fmin", "fmod",
             "greater", "greater_equal", "heaviside", "hypot", "less", "less_equal", "logaddexp", "logaddexp2",
             "logical_and", "logical_or", "logical_xor", "maximum", "minimum", "mod", "multiply", "nextafter",
             "not_equal", "power", "remainder", "subtract", "true_divide"]
    for name in names:
        for a, b in pairs:
            _bop(cuda_system, oracle, name, a, b, exact=False)
    a, b = pairs[1]
    _bop(cuda_system, oracle, "ldexp", a, b, exact=True)
    ba = np.array([True, False, True, False])
    bb = np.array([True, True, False, False])
    for name in ["bitwise_and", "bitwise_or", "bitwise_xor", "logical_and", "logical_or", "logical_xor"]:
        _bop(cuda_system, oracle, name, ba, bb)
    ia, ib = np.array([8, 3, 7], dtype=np.int64), np.array([4, 12, 13], dtype=np.int64)
    for name in ["gcd", "lcm"]:
        _bop(cuda_system, oracle, name, ia, ib)
    sa, sb = np.array([7000, 8000, 9000], dtype=np.int64), np.array([1, 2, 3], dtype=np.int64)
    for name in ["left_shift", "right_shift"]:
        _bop(cuda_system, oracle, name, sa, sb)


def test_bop_broadcasting(cuda_system, oracle):
    rng = np.random.default_rng(3)
    X = rng.standard_normal((1000, 28))
    col = rng.standard_normal((1000, 1))
    row = rng.standard_normal((1, 28))
    vec = rng.standard_normal(28)
    s32 = np.array(-1.0, dtype=np.float32)   # BlockArray.from_scalar makes f32 0-d blocks (blockarray.py:47-58)
    one = np.array(1.0)
    for op in ("add", "sub", "mul", "truediv"):
        _bop(cuda_system, oracle, op, col, X)
        _bop(cuda_system, oracle, op, X, col)
        _bop(cuda_system, oracle, op, X, row)
        _bop(cuda_system, oracle, op, X, vec)
        _bop(cuda_system, oracle, op, X, s32)
        _bop(cuda_system, oracle, op, one, X)
        _bop(cuda_system, oracle, op, rng.standard_normal((6, 1)), rng.standard_normal(8))  # test_bop.py:140-147
    # 3-D and 4-D
    A = rng.standard_normal((5, 1, 7))
    B = rng.standard_normal((4, 7))
    _bop(cuda_system, oracle, "add", A, B)
    A4 = rng.standard_normal((3, 4, 5, 6))
    _bop(cuda_system, oracle, "mul", A4, rng.standard_normal((5, 1)))
    # scalar stored as (1,) but used as () and vice versa (numpy_compute.py:226-229)
    a = np.array([2.5])
    want = oracle.bop("mul", a, X, (), X.shape, False, False, None)
    got = cuda_system.get(cuda_system.bop("mul", cuda_system.put(a), cuda_system.put(X), (), X.shape, False, False,
                                          axes=None, syskwargs={}))
    assert_exact(got, want)


def test_bop_transposed_operands(cuda_system, oracle):
    rng = np.random.default_rng(4)
    A = rng.standard_normal((37, 53))
    B = rng.standard_normal((53, 37))
    _bop(cuda_system, oracle, "add", A, B, a_T=False, b_T=True)
    _bop(cuda_system, oracle, "sub", A, A, a_T=True, b_T=True)
    v = rng.standard_normal(37)
    _bop(cuda_system, oracle, "mul", A, v, a_T=True)     # (d, n)^T-view (+) (n,)   glms.py:277
    T3 = rng.standard_normal((3, 4, 5))
    _bop(cuda_system, oracle, "add", T3, rng.standard_normal((5, 4, 3)), a_T=True)
    big = rng.standard_normal((300, 200))
    _bop(cuda_system, oracle, "mul", big, rng.standard_normal((200, 300)), a_T=True)


# ----------------------------------------------------------------------------------------------------
# map_uop / astype
# ----------------------------------------------------------------------------------------------------
EXACT_UOPS = ["abs", "absolute", "fabs", "negative", "positive", "sign", "sqrt", "square", "reciprocal", "floor",
              "ceil", "trunc", "rint", "isnan", "isinf", "isfinite", "signbit", "logical_not", "conjugate",
              "spacing", "deg2rad", "rad2deg", "radians", "degrees"]
CLOSE_UOPS = ["cbrt", "exp", "exp2", "expm1", "log", "log2", "log10", "log1p", "sin", "cos", "tan", "arcsin",
              "arccos", "arctan", "sinh", "cosh", "tanh", "arcsinh", "arccosh", "arctanh"]


@pytest.mark.parametrize("op", EXACT_UOPS + CLOSE_UOPS)
def test_map_uop_f64(cuda_system, oracle, op):
    rng = np.random.default_rng(21)
    x = rng.standard_normal(20011) * 3
    if op in ("log", "log2", "log10", "sqrt"):
        x = np.abs(x) + 1e-3
    if op in ("arcsin", "arccos", "arctanh"):
        x = np.clip(x / 4, -0.99, 0.99)
    if op == "arccosh":
        x = np.abs(x) + 1.0
    if op == "log1p":
        x = np.abs(x)
    want = oracle.map_uop(op, x, (), {})
    got = cuda_system.get(cuda_system.map_uop(op, cuda_system.put(x), (), {}, syskwargs={}))
    (assert_exact if op in EXACT_UOPS else assert_close)(got, want)


def test_map_uop_other_dtypes(cuda_system, oracle):
    rng = np.random.default_rng(22)
    xi = rng.integers(-100, 100, 999)
    xb = rng.random(999) < 0.5
    xf = (rng.standard_normal(999) * 2).astype(np.float32)
    for op in ["abs", "negative", "sign", "square", "sqrt", "isnan", "logical_not", "invert", "floor", "exp"]:
        want = oracle.map_uop(op, xi, (), {})
        got = cuda_system.get(cuda_system.map_uop(op, cuda_system.put(xi), (), {}, syskwargs={}))
        assert_close(got, want)
    for op in ["invert", "bitwise_not", "logical_not", "abs", "isfinite"]:
        want = oracle.map_uop(op, xb, (), {})
        got = cuda_system.get(cuda_system.map_uop(op, cuda_system.put(xb), (), {}, syskwargs={}))
        assert_exact(got, want)
    for op in ["exp", "sqrt", "abs", "negative", "tanh", "isnan"]:
        with np.errstate(all="ignore"):
            want = oracle.map_uop(op, xf, (), {})
        got = cuda_system.get(cuda_system.map_uop(op, cuda_system.put(xf), (), {}, syskwargs={}))
        assert_close(got, want)


def test_astype_and_copies(cuda_system, oracle):
    rng = np.random.default_rng(23)
    x = rng.standard_normal((33, 17)) * 100
    for dt in ["int64", "int32", "float32", "bool", "float64", "int", "float"]:
        want = oracle.astype(x, dt)
        got = cuda_system.get(cuda_system.astype(cuda_system.put(x), dt, syskwargs={}))
        assert_exact(got, want)
    xi = rng.integers(-5, 5, (20, 3))
    for dt in ["float64", "bool", "int32"]:
        assert_exact(cuda_system.get(cuda_system.astype(cuda_system.put(xi), dt, syskwargs={})), oracle.astype(xi, dt))
    # transpose / reshape are views that must download correctly
    t = cuda_system.transpose(cuda_system.put(x), syskwargs={})
    assert_exact(cuda_system.get(t), x.T)
    r = cuda_system.reshape(cuda_system.put(x), (17, 33), syskwargs={})
    assert_exact(cuda_system.get(r), x.reshape(17, 33))
    r2 = cuda_system.reshape(t, (33 * 17,), syskwargs={})
    assert_exact(cuda_system.get(r2), x.T.reshape(-1))
    big = rng.standard_normal((257, 131))
    assert_exact(cuda_system.get(cuda_system.reshape(cuda_system.transpose(cuda_system.put(big), syskwargs={}),
                                                     (131 * 257,), syskwargs={})), big.T.reshape(-1))


# ----------------------------------------------------------------------------------------------------
# reductions
# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(1,), (37,), (100003,), (5, 7), (1000, 28), (28, 1000), (3, 50, 7), (2, 3, 4, 5),
                                   (70000, 3), (3, 70000), (300, 300)])
def test_reduce_axis_f64(cuda_system, oracle, shape):
    rng = np.random.default_rng(31)
    x = rng.standard_normal(shape)
    for op in ("sum", "min", "max", "amin", "amax"):
        for axis in [None] + list(range(len(shape))):
            for keepdims in (False, True):
                for transposed in (False, True):
                    want = oracle.reduce_axis(op, x, axis, keepdims, transposed)
                    got = cuda_system.get(cuda_system.reduce_axis(op, cuda_system.put(x), axis, keepdims, transposed,
                                                                  syskwargs={}))
                    want = np.asarray(want)
                    if op == "sum":
                        assert got.shape == want.shape and got.dtype == want.dtype
                        scale = np.abs(x).sum() if axis is None else np.abs(x.T if transposed else x).sum(axis=axis, keepdims=keepdims)
                        assert (np.abs(got - want) <= F64_TOL * np.maximum(scale, 1e-300)).all()
                    else:
                        assert_exact(got, want)


def test_reduce_axis_other_dtypes(cuda_system, oracle):
    rng = np.random.default_rng(32)
    xb = rng.random((301, 17)) < 0.3
    xi = rng.integers(-1000, 1000, (301, 17))
    xi32 = xi.astype(np.int32)
    xf = rng.standard_normal((301, 17)).astype(np.float32)
    for x in (xb, xi, xi32):
        for op in ("sum", "min", "max"):
            for axis in (None, 0, 1):
                want = oracle.reduce_axis(op, x, axis, False, False)
                got = cuda_system.get(cuda_system.reduce_axis(op, cuda_system.put(x), axis, False, False, syskwargs={}))
                assert_exact(got, np.asarray(want))
    for op in ("sum", "min", "max"):
        for axis in (None, 0, 1):
            want = np.asarray(oracle.reduce_axis(op, xf, axis, False, False))
            got = cuda_system.get(cuda_system.reduce_axis(op, cuda_system.put(xf), axis, False, False, syskwargs={}))
            assert got.dtype == want.dtype
            assert np.allclose(got, want, rtol=1e-5, atol=1e-4)
    # NaN propagation of min/max (np.min) -- numpy_compute.py:177-181
    xn = rng.standard_normal((50, 6))
    xn[7, 2] = np.nan
    for op in ("min", "max", "sum"):
        for axis in (None, 0, 1):
            want = np.asarray(oracle.reduce_axis(op, xn, axis, False, False))
            got = cuda_system.get(cuda_system.reduce_axis(op, cuda_system.put(xn), axis, False, False, syskwargs={}))
            assert np.array_equal(np.isnan(got), np.isnan(want))


def test_sum_reduce(cuda_system, oracle):
    rng = np.random.default_rng(33)
    for count in (1, 2, 8, 40):
        arrs = [rng.standard_normal((28,)) for _ in range(count)]
        want = oracle.sum_reduce(*arrs)
        got = cuda_system.get(cuda_system.sum_reduce(*[cuda_system.put(a) for a in arrs], syskwargs={}))
        assert_exact(got, want)   # same left-to-right order => bit-identical
    arrs = [rng.integers(0, 9, (5, 6)) for _ in range(3)]
    assert_exact(cuda_system.get(cuda_system.sum_reduce(*[cuda_system.put(a) for a in arrs], syskwargs={})),
                 oracle.sum_reduce(*arrs))


def test_arg_op(cuda_system, oracle):
    rng = np.random.default_rng(34)
    for n in (1, 5, 1000, 300007):
        for dt in (np.float64, np.int64, np.float32):
            x = (rng.integers(0, 50, n)).astype(dt)   # many ties: first occurrence must win
            for op in ("argmin", "argmax"):
                sl = slice(1000, 1000 + n)
                wi, wv = oracle.arg_op(op, x, sl)
                gi, gv = cuda_system.arg_op(op, cuda_system.put(x), sl, None, None, syskwargs={})
                assert int(cuda_system.get(gi)) == int(wi)
                assert cuda_system.get(gv) == wv
                # carried optimum: ties keep the carried (earlier) one only if strictly better
                for carried in (wv, wv - 1, wv + 1):
                    wi2, wv2 = oracle.arg_op(op, x, sl, 7, dt(carried))
                    gi2, gv2 = cuda_system.arg_op(op, cuda_system.put(x), sl, 7, dt(carried), syskwargs={})
                    assert int(cuda_system.get(gi2)) == int(wi2)
                    assert cuda_system.get(gv2) == wv2


def test_where_allclose_logical_and(cuda_system, oracle):
    rng = np.random.default_rng(35)
    for shape in [(10,), (1000,), (37, 21), (5, 6, 7), (100003,)]:
        x = rng.random(shape) < 0.3
        slices = [(10 * (i + 1), 10 * (i + 1) + s) for i, s in enumerate(shape)]
        want = oracle.where(x.copy(), None, None, slices)
        got = cuda_system.where(cuda_system.put(x), None, None, slices, syskwargs={})
        assert tuple(got[-1]) == tuple(want[-1])
        for g, w in zip(got[:-1], want[:-1]):
            assert_exact(cuda_system.get(g), w)
    a = rng.standard_normal((100, 30))
    b = a + 1e-9
    for (p, q, rtol, atol) in [(a, a, 1e-5, 1e-8), (a, b, 1e-5, 1e-8), (a, a + 1.0, 1e-5, 1e-8), (a, b, 0.0, 1e-12)]:
        want = oracle.allclose(p, q, rtol, atol)
        got = cuda_system.get(cuda_system.allclose(cuda_system.put(p), cuda_system.put(q), rtol, atol, syskwargs={}))
        assert bool(got) == bool(want)
    t = cuda_system.allclose(cuda_system.put(a), cuda_system.put(a), 1e-5, 1e-8, syskwargs={})
    f = cuda_system.allclose(cuda_system.put(a), cuda_system.put(a + 1), 1e-5, 1e-8, syskwargs={})
    assert bool(cuda_system.get(cuda_system.logical_and(t, t, t, syskwargs={}))) is True
    assert bool(cuda_system.get(cuda_system.logical_and(t, f, t, syskwargs={}))) is False


# ----------------------------------------------------------------------------------------------------
# creation / data movement
# ----------------------------------------------------------------------------------------------------
def test_creation_kernels(cuda_system, oracle):
    meta = {"shape": (10, 7), "block_shape": (4, 3), "dtype": "float64"}
    for entry in [(0, 0), (2, 2), (1, 2)]:
        for op in ("zeros", "ones"):
            assert_exact(cuda_system.get(cuda_system.new_block(op, entry, meta, syskwargs={})),
                         oracle.new_block(op, entry, meta))
        assert cuda_system.get(cuda_system.empty(entry, meta, syskwargs={})).shape == oracle.empty(entry, meta).shape
    meta_eye = {"shape": (10, 10), "block_shape": (4, 4), "dtype": "int64"}
    for entry in [(0, 0), (2, 2)]:
        assert_exact(cuda_system.get(cuda_system.new_block("eye", entry, meta_eye, syskwargs={})),
                     oracle.new_block("eye", entry, meta_eye))
    for args in [(0, 10, 1, np.int64), (3, 40, 4, np.int64), (0.0, 1.0, 0.125, np.float64), (5, 5, 1, np.int64)]:
        assert_exact(cuda_system.get(cuda_system.arange(*args, syskwargs={})), oracle.arange(*args))
    v = np.arange(5.0)
    assert_exact(cuda_system.get(cuda_system.diag(cuda_system.put(v), syskwargs={})), oracle.diag(v))
    M = np.arange(12.0).reshape(3, 4)
    assert_exact(cuda_system.get(cuda_system.diag(cuda_system.put(M), syskwargs={})), oracle.diag(M))
    params = (1234, 3)
    for name, args, shape, dt in [("random", ((4, 5),), (4, 5), np.float64), ("normal", (0.0, 1.0, (3, 2)), (3, 2), np.float32),
                                  ("integers", (0, 10, (6,)), (6,), np.int64)]:
        assert_exact(cuda_system.get(cuda_system.random_block(params, name, args, shape, dt, syskwargs={})),
                     oracle.random_block(params, name, args, shape, dt))
    assert_exact(cuda_system.get(cuda_system.permutation(params, 17, syskwargs={})), oracle.permutation(params, 17))


def test_block_copies(cuda_system, oracle):
    rng = np.random.default_rng(41)
    a = rng.standard_normal((6, 8))
    b = rng.standard_normal((8, 6))
    src_params = [((slice(1, 4), slice(0, 5)), False), ((slice(0, 3), slice(1, 6)), True)]
    dst_params = [((slice(0, 3), slice(0, 5)), False), ((slice(3, 6), slice(0, 5)), False)]
    want = oracle.create_block(a, b, src_params=src_params, dst_params=dst_params, dst_shape=(6, 5), dst_shape_bc=None)
    got = cuda_system.get(cuda_system.create_block(cuda_system.put(a), cuda_system.put(b), src_params=src_params,
                                                   dst_params=dst_params, dst_shape=(6, 5), dst_shape_bc=None,
                                                   syskwargs={}))
    assert_exact(got, want)
    dst = rng.standard_normal((6, 8))
    up_src = [((slice(0, 2), slice(0, 3)), None, False)]
    up_dst = [((slice(4, 6), slice(5, 8)), False)]
    want = oracle.update_block(dst, a, src_params=up_src, dst_params=up_dst)
    dput = cuda_system.put(dst)
    got = cuda_system.get(cuda_system.update_block(dput, cuda_system.put(a), src_params=up_src, dst_params=up_dst,
                                                   syskwargs={}))
    assert_exact(got, want)
    assert_exact(cuda_system.get(dput), dst)   # inputs are immutable
    pairs = [((0, 1), (2, 3)), ((5, 7), (0, 0))]
    assert_exact(cuda_system.get(cuda_system.update_block_by_index(cuda_system.put(dst), cuda_system.put(a), pairs,
                                                                   syskwargs={})),
                 oracle.update_block_by_index(dst, a, pairs))
    axis_pairs = [(0, 3), (4, 1)]
    for axis in (0, 1):
        src = a if axis == 0 else rng.standard_normal((6, 8))
        assert_exact(cuda_system.get(cuda_system.update_block_along_axis(cuda_system.put(dst), cuda_system.put(src),
                                                                         axis_pairs, axis, syskwargs={})),
                     oracle.update_block_along_axis(dst, src, axis_pairs, axis))
    parts = cuda_system.split(cuda_system.put(a), 2, 1, False, syskwargs={})
    for g, w in zip(parts, oracle.split(a, 2, 1, False)):
        assert_exact(cuda_system.get(g), w)
    parts = cuda_system.split(cuda_system.put(a), [1, 4], 0, True, syskwargs={})
    for g, w in zip(parts, oracle.split(a, [1, 4], 0, True)):
        assert_exact(cuda_system.get(g), w)


# ----------------------------------------------------------------------------------------------------
# tensordot
# ----------------------------------------------------------------------------------------------------
def _tdot(system, oracle, a, b, axes=1, a_T=False, b_T=False):
    a_shape = a.T.shape if a_T else a.shape
    b_shape = b.T.shape if b_T else b.shape
    want = oracle.bop("tensordot", a, b, a_shape, b_shape, a_T, b_T, axes)
    got = system.get(system.bop("tensordot", system.put(a), system.put(b), a_shape, b_shape, a_T, b_T, axes=axes,
                                syskwargs={}))
    want = np.asarray(want)
    assert got.shape == want.shape and got.dtype == want.dtype
    if got.dtype.kind in "iu":
        assert np.array_equal(got, want)
    else:
        tol = GEMM_TOL if got.dtype == np.float64 else 1e-5
        assert rel_fro(got, want.astype(np.float64)) <= tol, rel_fro(got, want.astype(np.float64))


@pytest.mark.parametrize("mnk", [(128, 128, 16), (128, 128, 64), (256, 384, 128), (100, 130, 70), (129, 127, 33),
                                 (32, 32, 4096), (28, 28, 20000), (64, 2, 10), (2, 64, 10), (512, 512, 512),
                                 (1000, 28, 28), (2048, 128, 128), (32, 28, 50000), (8, 30, 10000), (2, 2, 100000),
                                 (28, 28, 8191), (28, 28, 8193)])
def test_tensordot_f64_matrix(cuda_system, oracle, mnk):
    m, n, k = mnk
    rng = np.random.default_rng(51)
    A = rng.standard_normal((m, k))
    B = rng.standard_normal((k, n))
    _tdot(cuda_system, oracle, A, B)
    _tdot(cuda_system, oracle, np.ascontiguousarray(A.T), B, a_T=True)
    _tdot(cuda_system, oracle, A, np.ascontiguousarray(B.T), b_T=True)
    _tdot(cuda_system, oracle, np.ascontiguousarray(A.T), np.ascontiguousarray(B.T), a_T=True, b_T=True)


def test_tensordot_vector_forms(cuda_system, oracle):
    rng = np.random.default_rng(52)
    for (m, k) in [(1000, 28), (28, 1000), (100003, 28), (333, 70), (64, 64), (5, 300)]:
        A = rng.standard_normal((m, k))
        x = rng.standard_normal(k)
        yv = rng.standard_normal(m)
        _tdot(cuda_system, oracle, A, x)                                   # (m,k).(k,)
        _tdot(cuda_system, oracle, A, yv, a_T=True)                       # (m,k)^T.(m,)  LR gradient
        _tdot(cuda_system, oracle, yv, A)                                  # (m,).(m,k)
        _tdot(cuda_system, oracle, x, A, b_T=True)                         # (k,).(m,k)^T
        _tdot(cuda_system, oracle, A, x.reshape(k, 1))                     # (m,k).(k,1)
    for n in (1, 10, 100003):
        u, v = rng.standard_normal(n), rng.standard_normal(n)
        _tdot(cuda_system, oracle, u, v)


@pytest.mark.parametrize("k", [8192, 8192 + 127, 100001, 1 << 20])
def test_tensordot_skinny_dense_and_narrow_gemv(cuda_system, oracle, k):
    """The streaming kernels behind the unfused LR step (glms.py:222-238): A^T B with M, N <= 32 over a long
    contraction (bulk-copy ring, dgemm_tn_skinny_dense_kernel), A x for a narrow dense A (gemv_rows_f64_narrow_kernel)
    and A^T w (gemv_t_narrow_kernel + the warp-parallel fold); every even width, ragged last chunks, A^T A."""
    rng = np.random.default_rng(k)
    widths = [(28, 28), (30, 2), (2, 32), (32, 32), (16, 24), (12, 20), (4, 6), (26, 18)] if k <= 100001 else [(28, 28), (6, 32)]
    for m, n in widths:
        A = rng.standard_normal((k, m))
        B = rng.standard_normal((k, n))
        _tdot(cuda_system, oracle, A, B, a_T=True)
        t = cuda_system.put(A)                                              # X^T X on ONE block: the operand is loaded once
        gram = cuda_system.get(cuda_system.bop("tensordot", t, t, (m, k), (k, m), True, False, axes=1, syskwargs={}))
        assert rel_fro(gram, A.T @ A) <= GEMM_TOL
    for cols in ([2, 4, 6, 10, 14, 16, 22, 26, 28, 30, 32] if k <= 100001 else [28]):
        A = rng.standard_normal((k, cols))
        x = rng.standard_normal(cols)
        w = rng.standard_normal(k)
        _tdot(cuda_system, oracle, A, x)
        _tdot(cuda_system, oracle, A, w, a_T=True)
    u, v = rng.standard_normal(k), rng.standard_normal(k)
    _tdot(cuda_system, oracle, u, v)


@pytest.mark.parametrize("k", [16384, 16384 + 37, 100_000, 1 << 20])
def test_gram_128_streaming_syrk(cuda_system, k):
    """A^T A of a tall dense 128-column block (the TSQR leaf on the Gram path, cuda_compute._gram_of): the streaming
    kernel computes the 136 blocks on or above the block diagonal and mirrors the rest."""
    from nums_b200 import cuda_compute as cc
    rng = np.random.default_rng(k)
    A = rng.standard_normal((k, 128))
    t = cuda_system.put(A)
    gram = cuda_system.get(cc._gram_of(t))
    want = A.T @ A
    assert rel_fro(gram, want) <= GEMM_TOL
    assert np.array_equal(gram, gram.T)
    # R through the public entry: qr(mode="r") on the Gram path; its inverse comes with the factorization
    Rdev = cuda_system.qr(t, mode="r", syskwargs={})
    Rinv = cuda_system.get(cuda_system.inv(Rdev, syskwargs={}))
    Rhost = cuda_system.get(Rdev)
    assert np.linalg.norm(Rinv @ Rhost - np.eye(128)) <= 1e-10 and np.allclose(np.tril(Rinv, -1), 0)
    R = cuda_system.get(cuda_system.qr(t, mode="r", syskwargs={}))
    Rref = np.linalg.qr(A, mode="r")
    sg = np.sign(np.diag(R)) * np.sign(np.diag(Rref))
    assert rel_fro(R * sg[:, None], Rref) <= 1e-10


@pytest.mark.parametrize("m", [16384, 16384 + 41, 200_003])
def test_tall_times_square_128_stream(cuda_system, oracle, m):
    """(m x 128) . (128 x 128), the explicit Q of TSQR (application.py:833-845): streaming kernel with B resident in
    registers; ragged last chunk; also through indirect_tsqr-like use with a transposed / non-dense operand (fallback)."""
    rng = np.random.default_rng(m)
    X = rng.standard_normal((m, 128))
    Rinv = rng.standard_normal((128, 128))
    _tdot(cuda_system, oracle, X, Rinv)
    _tdot(cuda_system, oracle, X, np.ascontiguousarray(Rinv.T), b_T=True)        # not the streaming shape: tiled GEMM


def test_tensordot_int_f32_nd(cuda_system, oracle):
    A = np.arange(6 * 7, dtype=np.int64).reshape(6, 7)       # test_bop.py:38-42 uses arange matrices
    B = np.arange(7 * 5, dtype=np.int64).reshape(7, 5)
    _tdot(cuda_system, oracle, A, B)
    _tdot(cuda_system, oracle, A, np.arange(7, dtype=np.int64))
    _tdot(cuda_system, oracle, np.arange(10), np.arange(10))
    rng = np.random.default_rng(53)
    Af = rng.standard_normal((123, 45)).astype(np.float32)
    Bf = rng.standard_normal((45, 67)).astype(np.float32)
    _tdot(cuda_system, oracle, Af, Bf)
    T1 = rng.standard_normal((2, 3, 4, 5))
    T2 = rng.standard_normal((4, 5, 6))
    _tdot(cuda_system, oracle, T1, T2, axes=2)                 # test_bop.py:70-81
    _tdot(cuda_system, oracle, T1, rng.standard_normal((5, 2)), axes=1)
    _tdot(cuda_system, oracle, rng.standard_normal((5, 4, 3, 2)), T2, axes=2, a_T=True)
    _tdot(cuda_system, oracle, rng.standard_normal(3), rng.standard_normal(4), axes=0)


# ----------------------------------------------------------------------------------------------------
# factorizations
# ----------------------------------------------------------------------------------------------------
def _canon(R):
    s = np.sign(np.diag(R[:, :R.shape[0]]))
    s[s == 0] = 1
    return R * s[:, None]


@pytest.mark.parametrize("shape", [(9, 9), (123, 9), (8, 9), (1000, 28), (4096, 128), (20000, 64), (300, 100),
                                   (70000, 16), (5, 3), (1, 1)])
def test_qr(cuda_system, oracle, shape):
    rng = np.random.default_rng(61)
    X = rng.standard_normal(shape)
    R_want = oracle.qr(X, mode="r")
    R_got = cuda_system.get(cuda_system.qr(cuda_system.put(X), mode="r", axis=None, syskwargs={}))
    assert R_got.shape == R_want.shape
    assert np.allclose(np.tril(R_got, -1), 0)
    assert rel_fro(_canon(R_got), _canon(R_want)) <= GEMM_TOL, rel_fro(_canon(R_got), _canon(R_want))
    Q, R = cuda_system.qr(cuda_system.put(X), mode="reduced", axis=None, syskwargs={})
    Q, R = cuda_system.get(Q), cuda_system.get(R)
    Qw, Rw = oracle.qr(X, mode="reduced")
    assert Q.shape == Qw.shape and R.shape == Rw.shape
    assert rel_fro(Q @ R, X) <= GEMM_TOL
    assert np.linalg.norm(Q.T @ Q - np.eye(Q.shape[1])) <= 1e-10
    assert rel_fro(_canon(R), _canon(Rw)) <= GEMM_TOL


def test_qr_concatenated_inputs(cuda_system, oracle):
    rng = np.random.default_rng(62)
    cols = [rng.standard_normal((123, 4)), rng.standard_normal((123, 4)), rng.standard_normal((123, 1))]
    want = oracle.qr(*cols, mode="r", axis=1)                 # application.py:784-797
    got = cuda_system.get(cuda_system.qr(*[cuda_system.put(c) for c in cols], mode="r", axis=1, syskwargs={}))
    assert rel_fro(_canon(got), _canon(want)) <= GEMM_TOL
    Rs = [np.triu(rng.standard_normal((9, 9))) for _ in range(19)] + [rng.standard_normal((8, 9))]
    want = oracle.qr(*Rs, mode="r", axis=0)                   # application.py:807-814
    got = cuda_system.get(cuda_system.qr(*[cuda_system.put(r) for r in Rs], mode="r", axis=0, syskwargs={}))
    assert rel_fro(_canon(got), _canon(want)) <= GEMM_TOL


@pytest.mark.parametrize("n", [1, 2, 7, 28, 64, 127, 128])
def test_gram_factor_entry(cuda_system, n):
    """nums_gram_factor (the small-matrix stage of the Gram QR path): L, R = L^T, L^-1 and the four
    norms against NumPy; a non-positive-definite input reports the failing step."""
    import torch
    from nums_b200 import _lib
    LIB = _lib.LIB
    rng = np.random.default_rng(64)
    A = rng.standard_normal((4 * n + 8, n))
    G = A.T @ A
    dev = cuda_system.put(G)
    outs = [torch.empty((n, n), dtype=torch.float64, device=dev.device) for _ in range(3)]
    stats = torch.empty((5,), dtype=torch.float64, device=dev.device)
    stream = torch.cuda.current_stream().cuda_stream
    LIB.check(LIB.dll.nums_gram_factor(n, dev.data_ptr(), n, outs[0].data_ptr(), n, outs[1].data_ptr(), n,
                                       outs[2].data_ptr(), n, stats.data_ptr(), stream))
    L, R, Linv = (t.cpu().numpy() for t in outs)
    info, l1, linf, i1, iinf = stats.cpu().numpy()
    want = np.linalg.cholesky(G)
    assert info == 0
    assert rel_fro(L, want) <= 1e-12
    assert np.array_equal(R, L.T)
    assert rel_fro(Linv, np.linalg.inv(want)) <= 1e-11
    assert np.array_equal(np.triu(L, 1), np.zeros((n, n))) and np.array_equal(np.triu(Linv, 1), np.zeros((n, n)))
    for got, ref in ((l1, np.linalg.norm(want, 1)), (linf, np.linalg.norm(want, np.inf)),
                     (i1, np.linalg.norm(np.linalg.inv(want), 1)), (iinf, np.linalg.norm(np.linalg.inv(want), np.inf))):
        assert abs(got - ref) <= 1e-11 * ref
    assert np.sqrt(l1 * linf * i1 * iinf) >= np.linalg.cond(A) * (1 - 1e-10)     # the bound really bounds
    if n >= 2:
        bad = G.copy()
        bad[n - 1, n - 1] = -1.0
        LIB.check(LIB.dll.nums_gram_factor(n, cuda_system.put(bad).data_ptr(), n, outs[0].data_ptr(), n, outs[1].data_ptr(),
                                           n, outs[2].data_ptr(), n, stats.data_ptr(), stream))
        assert stats.cpu().numpy()[0] == n


@pytest.mark.parametrize("n", [1, 2, 9, 28, 64, 128, 200])
def test_inv_cholesky(cuda_system, oracle, n):
    rng = np.random.default_rng(63)
    for dt, tol in ((np.float64, 1e-10), (np.float32, 2e-3)):
        A = rng.standard_normal((n, n)).astype(dt) + n * np.eye(n, dtype=dt)
        got = cuda_system.get(cuda_system.inv(cuda_system.put(A), syskwargs={}))
        want = oracle.inv(A)
        assert got.dtype == want.dtype
        assert rel_fro(got, want.astype(np.float64)) <= tol
        S = (A @ A.T).astype(dt)
        L = cuda_system.get(cuda_system.cholesky(cuda_system.put(S), syskwargs={}))
        assert rel_fro(L, oracle.cholesky(S).astype(np.float64)) <= tol
    G = rng.standard_normal((n, n))           # general, needs pivoting
    if n > 1:
        G[0, 0] = 0.0
    got = cuda_system.get(cuda_system.inv(cuda_system.put(G), syskwargs={}))
    assert rel_fro(got @ G, np.eye(n)) <= 1e-8 * max(1.0, np.linalg.cond(G))


def test_inv_singular_raises(cuda_system, monkeypatch):
    """np.linalg.LinAlgError like the reference (numpy_compute.py:248-257).  Default: the status word is examined at
    the next host synchronisation point (get / touch / synchronize); "eager": from the call itself."""
    from nums_b200 import cuda_compute as cc
    assert cc.CHECK_FACTORIZATION_STATUS == "deferred"
    bad = cuda_system.inv(cuda_system.put(np.zeros((4, 4))), syskwargs={})
    with pytest.raises(np.linalg.LinAlgError):
        cuda_system.get(bad)
    cuda_system.cholesky(cuda_system.put(-np.eye(3)), syskwargs={})
    with pytest.raises(np.linalg.LinAlgError):
        cuda_system.synchronize()
    good = cuda_system.inv(cuda_system.put(np.eye(4) * 2.0), syskwargs={})
    assert np.array_equal(cuda_system.get(good), np.eye(4) * 0.5)       # an examined failure does not linger
    monkeypatch.setattr(cc, "CHECK_FACTORIZATION_STATUS", "eager")
    with pytest.raises(np.linalg.LinAlgError):
        cuda_system.inv(cuda_system.put(np.zeros((4, 4))), syskwargs={})
    with pytest.raises(np.linalg.LinAlgError):
        cuda_system.cholesky(cuda_system.put(-np.eye(3)), syskwargs={})


def test_integer_power_negative_exponent_raises(cuda_system, oracle):
    """np.power(int, negative int) raises ValueError in the reference (numpy_compute.py:233-238 -> NumPy); here the
    check is a device-side minimum examined at the next host synchronisation."""
    base = np.arange(1, 9, dtype=np.int64)
    good = np.array([0, 1, 2, 3, 0, 1, 2, 3], dtype=np.int64)
    got = cuda_system.get(cuda_system.bop("pow", cuda_system.put(base), cuda_system.put(good), (8,), (8,), False, False,
                                          axes=None, syskwargs={}))
    assert_exact(got, oracle.bop("pow", base, good, (8,), (8,), False, False, None))
    bad = good.copy()
    bad[5] = -1
    with pytest.raises(ValueError):
        oracle.bop("pow", base, bad, (8,), (8,), False, False, None)
    out = cuda_system.bop("pow", cuda_system.put(base), cuda_system.put(bad), (8,), (8,), False, False, axes=None,
                          syskwargs={})
    with pytest.raises(ValueError):
        cuda_system.get(out)
    # float bases take negative exponents like NumPy
    fb = base.astype(np.float64)
    got = cuda_system.get(cuda_system.bop("pow", cuda_system.put(fb), cuda_system.put(bad), (8,), (8,), False, False,
                                          axes=None, syskwargs={}))
    assert np.allclose(got, fb ** bad, rtol=1e-15)


# ----------------------------------------------------------------------------------------------------
# fused LR kernel vs the reference composition (glms.py:213-240)
# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nd", [(1000, 28), (100003, 28), (257, 8), (5000, 30), (4096, 48), (999, 2), (3000, 12),
                                (70001, 20), (1000, 4), (256, 28), (255, 28), (513, 36)])
def test_lr_grad_hess(cuda_system, nd):
    from nums_b200 import cuda_compute
    n, d = nd
    rng = np.random.default_rng(71)
    X = rng.standard_normal((n, d))
    y = (rng.random(n) < 0.5).astype(np.float64)
    beta = rng.standard_normal(d) / np.sqrt(d)
    mu = 1.0 / (1.0 + np.exp(-(X @ beta)))
    g = X.T @ (mu - y)
    H = X.T @ ((mu * (1 - mu))[:, None] * X)
    out = cuda_system.get(cuda_compute.lr_grad_hess(cuda_system.put(X), cuda_system.put(y), cuda_system.put(beta)))
    assert rel_fro(out[:d], g) <= GEMM_TOL
    assert rel_fro(out[d:].reshape(d, d), H) <= GEMM_TOL
    assert np.array_equal(out[d:].reshape(d, d), out[d:].reshape(d, d).T)


# ----------------------------------------------------------------------------------------------------
# deferred dot/add chains (nums_b200/deferred.py): same results as the eager kernels and the oracle
# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shapes", [((512, 384), (384, 640), (128, 128), (128, 128)),
                                     ((300, 260), (260, 280), (128, 128), (128, 128)),
                                     ((256, 1024), (1024, 192), (128, 256), (256, 64)),
                                     ((130, 70), (70, 200), (66, 16), (16, 100))])
def test_deferred_blocked_matmul(cuda_system, shapes):
    from nums_b200 import blocks
    from nums_b200._lib import LIB
    (sa, sb, ba, bb) = shapes
    rng = np.random.default_rng(81)
    A, B = rng.standard_normal(sa), rng.standard_normal(sb)
    app = blocks.ArrayApp(cuda_system)
    a, b = app.array(A, ba), app.array(B, bb)
    want = A @ B
    for label, (x, y, ref) in {"nn": (a, b, want), "tn": (app.array(np.ascontiguousarray(A.T), ba[::-1]).T, b, want),
                               "nt": (a, app.array(np.ascontiguousarray(B.T), bb[::-1]).T, want)}.items():
        cuda_system.contractions.enabled = True
        before = LIB.dll.nums_launch_count()
        lazy = (x @ y).get()
        lazy_launches = LIB.dll.nums_launch_count() - before
        cuda_system.contractions.enabled = False
        try:
            before = LIB.dll.nums_launch_count()
            eager = (x @ y).get()
            eager_launches = LIB.dll.nums_launch_count() - before
        finally:
            cuda_system.contractions.enabled = True
        assert rel_fro(lazy, ref) <= GEMM_TOL, label
        assert rel_fro(eager, ref) <= GEMM_TOL, label
        assert lazy_launches <= eager_launches, (label, lazy_launches, eager_launches)
    # a chain that mixes a concrete addend, deferred terms and a non-contraction consumer
    c = (a @ b) + (a @ b)
    d = c * 2.0
    assert rel_fro(d.get(), 4.0 * want) <= GEMM_TOL


@pytest.mark.parametrize("n", [1, 2, 9, 28, 64, 128, 129])
def test_svd(cuda_system, oracle, n):
    rng = np.random.default_rng(64)
    for make in (lambda: rng.standard_normal((n, n)), lambda: np.triu(rng.standard_normal((n, n)))):
        A = make()
        u, s, vt = (cuda_system.get(x) for x in cuda_system.svd(cuda_system.put(A), syskwargs={}))
        uw, sw, vtw = oracle.svd(A)
        assert u.shape == uw.shape and s.shape == sw.shape and vt.shape == vtw.shape
        assert np.all(np.diff(s) <= 0)
        assert np.abs(s - sw).max() <= 1e-10 * sw[0]
        assert rel_fro((u * s) @ vt, A) <= 1e-10
        assert np.linalg.norm(u.T @ u - np.eye(n)) <= 1e-10 * n
        assert np.linalg.norm(vt @ vt.T - np.eye(n)) <= 1e-10 * n


def test_lr_grad_hess_blocks(cuda_system):
    from nums_b200 import cuda_compute
    rng = np.random.default_rng(72)
    for d, rows in ((28, [1000, 257, 4096, 31]), (12, [300] * 20), (8, [500, 600])):
        xs = [rng.standard_normal((r, d)) for r in rows]
        ys = [(rng.random(r) < 0.5).astype(np.float64) for r in rows]
        beta = rng.standard_normal(d) / np.sqrt(d)
        X, y = np.concatenate(xs), np.concatenate(ys)
        mu = 1.0 / (1.0 + np.exp(-(X @ beta)))
        g = X.T @ (mu - y)
        H = X.T @ ((mu * (1 - mu))[:, None] * X)
        out = cuda_system.get(cuda_compute.lr_grad_hess_blocks([cuda_system.put(x) for x in xs],
                                                               [cuda_system.put(v) for v in ys], cuda_system.put(beta)))
        assert rel_fro(out[:d], g) <= GEMM_TOL
        assert rel_fro(out[d:].reshape(d, d), H) <= GEMM_TOL


@pytest.mark.parametrize("d", [1, 5, 28, 64, 128])
def test_newton_step(cuda_system, d):
    """nums_newton_step against the reference's update (glms.py:362-372): beta - inv(H) @ g, max |g|."""
    from nums_b200 import cuda_compute as cc
    rng = np.random.default_rng(70 + d)
    A = rng.standard_normal((3 * d + 2, d))
    H = A.T @ A
    if d > 1:
        H[[0, d - 1]] = H[[d - 1, 0]]            # not symmetric any more: forces row exchanges
    g = rng.standard_normal(d)
    beta = rng.standard_normal(d)
    out, status = cc.newton_step(cuda_system.put(np.concatenate([g, H.ravel()])), cuda_system.put(beta))
    want = beta - np.linalg.inv(H) @ g
    assert rel_fro(cuda_system.get(out), want) <= 1e-10
    gmax, info = cuda_system.get(status)
    assert gmax == np.max(np.abs(g)) and info == 0
    singular = H.copy()
    singular[:, 0] = 0.0
    _out, status = cc.newton_step(cuda_system.put(np.concatenate([g, singular.ravel()])), cuda_system.put(beta))
    assert cuda_system.get(status)[1] == 1
    g[d // 2] = np.nan
    _out, status = cc.newton_step(cuda_system.put(np.concatenate([g, H.ravel()])), cuda_system.put(beta))
    assert np.isnan(cuda_system.get(status)[0])


def test_batched_scatter_kernels(cuda_system, oracle):
    """update_block_by_index / update_block_along_axis as one launch (nums_scatter_axis): many pairs,
    duplicate destinations (the last pair wins), negative indices, 1-D to 3-D, every dtype, dtype-casting
    assignment, and NumPy's IndexError."""
    rng = np.random.default_rng(71)
    for dt in (np.float64, np.float32, np.int64, np.int32, np.bool_):
        dst = (rng.standard_normal((37, 23)) * 100).astype(dt)
        src = (rng.standard_normal((41, 19)) * 100).astype(dt)
        pairs = [((int(rng.integers(-37, 37)), int(rng.integers(-23, 23))), (int(rng.integers(-41, 41)), int(rng.integers(-19, 19))))
                 for _ in range(900)]                     # 851 cells -> plenty of duplicate destinations
        got = cuda_system.get(cuda_system.update_block_by_index(cuda_system.put(dst), cuda_system.put(src), pairs, syskwargs={}))
        assert_exact(got, oracle.update_block_by_index(dst, src, pairs))
    # permutation of a vector (nums.numpy.random.permutation / shuffle, test_np_random.py:46-106)
    v = rng.standard_normal(5000)
    perm = rng.permutation(5000)
    pairs = [((int(i),), (int(j),)) for i, j in enumerate(perm)]
    got = cuda_system.get(cuda_system.update_block_by_index(cuda_system.put(np.zeros(5000)), cuda_system.put(v), pairs, syskwargs={}))
    assert_exact(got, v[perm])
    # slices along each axis of a 3-D block, duplicates and negative indices included
    dst3 = rng.standard_normal((7, 9, 5))
    for axis, n_dst, n_src in ((0, 7, 11), (1, 9, 4), (2, 5, 6)):
        shape = list(dst3.shape)
        shape[axis] = n_src
        src3 = rng.standard_normal(shape)
        pairs = [(int(rng.integers(-n_dst, n_dst)), int(rng.integers(-n_src, n_src))) for _ in range(20)]
        got = cuda_system.get(cuda_system.update_block_along_axis(cuda_system.put(dst3), cuda_system.put(src3), pairs, axis,
                                                                  syskwargs={}))
        assert_exact(got, oracle.update_block_along_axis(dst3, src3, pairs, axis))
    # assignment casts to the destination dtype
    di, sf = np.arange(12, dtype=np.int64).reshape(3, 4), rng.standard_normal((3, 4)) * 10
    pairs = [((0, 0), (2, 3)), ((2, 1), (0, 0))]
    got = cuda_system.get(cuda_system.update_block_by_index(cuda_system.put(di), cuda_system.put(sf), pairs, syskwargs={}))
    assert_exact(got, oracle.update_block_by_index(di, sf, pairs))
    with pytest.raises(IndexError):
        cuda_system.update_block_by_index(cuda_system.put(di), cuda_system.put(sf), [((3, 0), (0, 0))], syskwargs={})


# ----------------------------------------------------------------------------------------------------
# qr: graded spectra -- every branch of cuda_compute.qr_r_ex (one Cholesky pass, CholeskyQR2, Householder)
# ----------------------------------------------------------------------------------------------------
def _graded(m, n, kappa, seed, dtype=np.float64):
    """m x n matrix with singular values log-spaced from 1 down to 1/kappa."""
    rng = np.random.default_rng(seed)
    U, _ = np.linalg.qr(rng.standard_normal((m, n)))
    V, _ = np.linalg.qr(rng.standard_normal((n, n)))
    return ((U * np.logspace(0, -np.log10(kappa), n)) @ V.T).astype(dtype)


def _qr_r_checks(cuda_system, X, r_tol, gram_tol):
    from nums_b200 import cuda_compute as cc
    before = dict(cc.QR_STATS)
    R = cuda_system.get(cuda_system.qr(cuda_system.put(X), mode="r", axis=None, syskwargs={}))
    ran = {k: cc.QR_STATS[k] - before[k] for k in before}
    X64 = X.astype(np.float64)
    assert R.shape == (X.shape[1], X.shape[1]) and np.allclose(np.tril(R, -1), 0)
    # backward stability, independent of row signs: R^T R = X^T X
    G = X64.T @ X64
    assert np.linalg.norm(R.astype(np.float64).T @ R.astype(np.float64) - G) / np.linalg.norm(G) <= gram_tol
    if r_tol is not None:
        Rw = np.linalg.qr(X64, mode="r")
        assert rel_fro(_canon(R), _canon(Rw)) <= r_tol, rel_fro(_canon(R), _canon(Rw))
    return ran


@pytest.mark.parametrize("kappa,branch,r_tol", [(1e0, "gram", 1e-10), (2e1, "gram2", 1e-10), (1e2, "gram2", 1e-10),
                                               (1e5, "gram2", 1e-10), (1e9, "gram3", 1e-10),
                                               (1e13, "gram3", 1e-10)])
def test_qr_graded_spectrum_branches(cuda_system, kappa, branch, r_tol):
    """Tall float64 blocks of growing condition number: the 1e-10 sign-canonical bar on R against LAPACK's
    Householder R, backward stability (R^T R = X^T X), and the branch the condition bound selects: one Cholesky
    pass, CholeskyQR2, iterated shifted Cholesky passes ("gram3", oracle/shifted_cholqr_study.py)."""
    X = _graded(40_000, 64, kappa, seed=int(np.log10(kappa)) + 7)
    ran = _qr_r_checks(cuda_system, X, r_tol, gram_tol=1e-13)
    assert ran[branch] == 1 and sum(ran.values()) == 1, ran
    Q, R = cuda_system.qr(cuda_system.put(X), mode="reduced", axis=None, syskwargs={})
    Q, R = cuda_system.get(Q), cuda_system.get(R)
    assert rel_fro(Q @ R, X) <= 1e-12
    if kappa <= 1e5:
        assert np.linalg.norm(Q.T @ Q - np.eye(64)) <= 1e-10


def test_qr_rank_deficient_and_float32(cuda_system):
    rng = np.random.default_rng(77)
    X = rng.standard_normal((30_000, 32))
    X[:, 7] = X[:, 3]                        # exactly rank deficient: the Gram matrix is not positive definite
    X[:, 20] = 0.5 * X[:, 1] - 2.0 * X[:, 2]
    ran = _qr_r_checks(cuda_system, X, None, gram_tol=1e-13)
    assert ran["gram3"] + ran["householder"] == 1 and ran["gram"] == 0 and ran["gram2"] == 0, ran
    from nums_b200 import cuda_compute as cc
    saved = cc.QR_SHIFTED_ENABLED
    cc.QR_SHIFTED_ENABLED = False                  # the Householder kernel on the same block
    try:
        ran = _qr_r_checks(cuda_system, X, None, gram_tol=1e-13)
        assert ran["householder"] == 1 and ran["gram3"] == 0, ran
        Xi = _graded(40_000, 64, 1e9, seed=5)
        ran = _qr_r_checks(cuda_system, Xi, None, gram_tol=1e-13)
        assert ran["householder"] == 1, ran
    finally:
        cc.QR_SHIFTED_ENABLED = saved
    X32 = rng.standard_normal((50_000, 48)).astype(np.float32)     # float32 tall block: Gram path on a float64 copy
    ran = _qr_r_checks(cuda_system, X32, 2e-5, gram_tol=2e-6)
    assert ran["gram"] == 1 and ran["householder"] == 0, ran
    R32 = cuda_system.get(cuda_system.qr(cuda_system.put(X32), mode="r", axis=None, syskwargs={}))
    assert R32.dtype == np.float32
    X32w = rng.standard_normal((300, 48)).astype(np.float32)         # not tall enough: Householder kernel, float32 in and out
    ran = _qr_r_checks(cuda_system, X32w, 2e-5, gram_tol=2e-6)
    assert ran["householder"] == 1, ran
    Xz = np.zeros((4096, 16))
    R = cuda_system.get(cuda_system.qr(cuda_system.put(Xz), mode="r", axis=None, syskwargs={}))
    assert np.array_equal(R, np.zeros((16, 16)))


def test_bop_row_and_column_broadcast_large(cuda_system, oracle):
    """(n, d) (+) (n, 1) / (1, d) / (d,) at sizes that use the 128-bit row/column broadcast kernel, both operand
    orders, float64 and float32, comparison results, plus shapes that must fall back to the strided kernel."""
    rng = np.random.default_rng(9)
    for n, d in ((60_000, 28), (4_099, 6), (1_000, 7)):          # d = 7: odd width -> generic strided kernel
        X = rng.standard_normal((n, d))
        col, row, vec = rng.standard_normal((n, 1)), rng.standard_normal((1, d)), rng.standard_normal(d)
        for op in ("add", "sub", "mul", "truediv", "lt", "maximum"):
            for a, b in ((col, X), (X, col), (X, row), (row, X), (X, vec), (vec, X)):
                _bop(cuda_system, oracle, op, a, b)
        _bop(cuda_system, oracle, "mul", col.astype(np.float32), X)          # f32 column, f64 loop
        X32 = X.astype(np.float32)
        _bop(cuda_system, oracle, "mul", col.astype(np.float32), X32)
        _bop(cuda_system, oracle, "sub", X32, row.astype(np.float32))


def test_flush_tile_accounting_and_dangling_chain(cuda_system, monkeypatch):
    """One grouped launch per flush with exactly the expected tile count; a partial k-chain that is still
    referenced at flush time is detected (deferred.py: ContractionQueue._audit)."""
    from nums_b200 import blocks
    rng = np.random.default_rng(5)
    A, B = rng.standard_normal((512, 512)), rng.standard_normal((512, 512))
    app = blocks.ArrayApp(cuda_system)
    a, b = app.array(A, (128, 128)), app.array(B, (128, 128))
    q = cuda_system.contractions
    c = a @ b
    cuda_system.flush()
    assert q.last_flush == {"contractions": 16, "tile_terms": 16 * 4, "redundant_tile_terms": 0}
    assert rel_fro(c.get(), A @ B) <= GEMM_TOL
    # the same chain built by hand, keeping the partial sums alive
    shape, sk = (128, 128), {"grid_entry": (0, 0), "grid_shape": (4, 4)}
    keep, acc = [], None
    for k in range(4):
        dot = cuda_system.bop("tensordot", a.blocks[0, k].oid, b.blocks[k, 0].oid, shape, shape, False, False, axes=1, syskwargs=sk)
        acc = dot if acc is None else cuda_system.bop("add", acc, dot, shape, shape, False, False, axes=None, syskwargs=sk)
        keep.append(acc)
    with pytest.raises(AssertionError, match="partial dot/add chain"):
        cuda_system.flush()
    keep = acc = dot = None
    monkeypatch.setenv("NUMS_DEFERRED_STRICT", "0")
    q._warned = False
    keep, acc = [], None
    for k in range(4):
        dot = cuda_system.bop("tensordot", a.blocks[0, k].oid, b.blocks[k, 0].oid, shape, shape, False, False, axes=1, syskwargs=sk)
        acc = dot if acc is None else cuda_system.bop("add", acc, dot, shape, shape, False, False, axes=None, syskwargs=sk)
        keep.append(acc)
    with pytest.warns(RuntimeWarning, match="partial dot/add chain"):
        cuda_system.flush()
    assert q.last_flush["redundant_tile_terms"] == 1 + 2 + 3
    assert rel_fro(cuda_system.get(acc), (A @ B)[:128, :128]) <= GEMM_TOL
