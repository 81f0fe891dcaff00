"""The reference's own host layers (BlockArray / ArrayApplication / nums.numpy / nums.models.glms, imported
unmodified from baseline/_ref or /root/reference) driven over ``CudaSystem`` + ``cuda_compute`` and compared,
in the same process, with the same layers over the reference's ``SerialSystem`` + ``numpy_compute``.

Complements tests/test_gpu_reference_suite.py (which runs the reference's test files as they are): the tests
below cover what those files cannot on this NumPy -- ``test_ufunc`` (tests/numpy/test_arithmetic.py:20-79 fails
on NumPy 2 because of ufunc aliases the reference API lacks; here the same loop runs over the names it has) --
and the hot-path workloads at the tolerances BASELINE.json states.
"""
import numpy as np
import pytest

from nums_b200 import reference_compat
from tests.helpers import canon_r, rel_fro

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not reference_compat.available(), reason="reference not installed (scripts/install_reference.sh)")]


@pytest.fixture(scope="module")
def apps():
    reference_compat.load_reference()
    from nums.core.array.application import ArrayApplication
    from nums.core.systems import numpy_compute
    from nums.core.systems.filesystem import FileSystem
    from nums.core.systems.systems import SerialSystem
    serial = SerialSystem(compute_module=numpy_compute)
    serial.init()
    return reference_compat.cuda_app(), ArrayApplication(system=serial, filesystem=FileSystem(serial))


@pytest.fixture()
def nps_cuda():
    app = reference_compat.use_cuda()
    import nums.numpy as nps
    yield nps, app
    from nums.core import application_manager, settings
    application_manager.destroy()
    settings.system_name = "serial"


def test_system_is_registered(nps_cuda):
    """NUMS_SYSTEM=cuda route: application_manager.create builds the GPU application."""
    import torch
    from nums.core import application_manager, settings
    from nums.core.systems.systems import SerialSystem
    from nums_b200.cuda_system import CudaSystem
    application_manager.destroy()
    settings.system_name = "cuda"
    app = application_manager.instance()
    assert isinstance(app.system, CudaSystem) and isinstance(app.system, SerialSystem)
    x = app.array(np.arange(6.0).reshape(2, 3), (1, 3))
    assert isinstance(x.blocks[0, 0].oid, torch.Tensor) and x.blocks[0, 0].oid.is_cuda


def test_ufunc_through_nums_numpy(nps_cuda):
    """tests/numpy/test_arithmetic.py:20-79 over the ufuncs nums.numpy implements."""
    nps, _app = nps_cuda
    from nums.numpy import numpy_utils
    uops, bops = numpy_utils.ufunc_op_signatures()
    ran = 0
    for name, _ in sorted(uops):
        if not hasattr(nps, name):
            continue        # NumPy-2 aliases (acos, asin, bitwise_invert ...) the reference API predates
        if name in ("arccosh", "arcsinh"):
            np_val = np.array([np.e])
        elif name == "invert" or name.startswith("bitwise") or name.startswith("logical"):
            np_val = np.array([True, False], dtype=np.bool_)
        else:
            np_val = np.array([.1, .2, .3])
        try:
            ns_ufunc = getattr(nps, name)
            ns_result = ns_ufunc(nps.array(np_val))
        except NotImplementedError as exc:
            if "not yet implemented" in str(exc):
                continue
            raise
        want = getattr(np, name)(np_val)
        got = ns_result.get()
        assert got.dtype == want.dtype, name
        assert np.allclose(want, got, rtol=1e-12, atol=0), name
        ran += 1
    assert ran >= 40

    def check_bop(name, a, b):
        if name == "ldexp" and b.dtype.kind != "i":
            return 0
        want = getattr(np, name)(a, b)
        try:
            got = getattr(nps, name)(nps.array(a), nps.array(b)).get()
        except NotImplementedError as exc:
            if "not yet implemented" in str(exc):      # declared but unimplemented in the reference API (e.g. left_shift)
                return 0
            raise
        assert got.dtype == want.dtype, (name, a.dtype, b.dtype)
        assert np.allclose(want, got, rtol=1e-12, atol=0), (name, a.dtype, b.dtype)
        return 1

    ran = 0
    for name, _ in bops:
        if not hasattr(nps, name):
            continue
        try:
            getattr(nps, name)
        except NotImplementedError:
            continue
        if name.startswith("bitwise") or name.startswith("logical"):
            ran += check_bop(name, np.array([True, False, True, False]), np.array([True, True, False, False]))
        elif name in ("gcd", "lcm"):
            ran += check_bop(name, np.array([8, 3, 7]), np.array([4, 12, 13]))
        elif name.endswith("shift"):
            ran += check_bop(name, np.array([7000, 8000, 9000]), np.array([1, 2, 3]))
        else:
            for a, b in ((np.array([.1, 5.0, .3]), np.array([.2, 6.0, .3])),
                         (np.array([.1, 5.0, .3]), np.array([4, 2, 6])),
                         (np.array([3, 7, 3]), np.array([4, 2, 6]))):
                ran += check_bop(name, a, b)
    assert ran >= 60


def test_blocked_matmul_reference_blockarray(apps):
    cuda_app, serial_app = apps
    rng = np.random.default_rng(3)
    A, B = rng.standard_normal((1024, 768)), rng.standard_normal((768, 640))
    outs = []
    for app in apps:
        a, b = app.array(A, (256, 256)), app.array(B, (256, 256))
        outs.append(((a @ b).get(), (a.T @ a).get(), (a @ a.T).get()))
    for got, want in zip(*outs):
        assert rel_fro(got, want) <= 1e-10
    assert rel_fro(outs[0][0], A @ B) <= 1e-10
    assert cuda_app.system.contractions.flushes > 0      # went through the grouped DMMA launch


def test_elementwise_bit_exact(apps):
    rng = np.random.default_rng(1)
    u, v = rng.random(100_000), rng.random(100_000)
    outs = []
    for app in apps:
        U, V = app.array(u, (12_500,)), app.array(v, (12_500,))
        outs.append(((U + V).get(), (U * V).get(), (U / V).get(), (U - V).get(), (U < V).get()))
    for got, want in zip(*outs):
        assert got.dtype == want.dtype and np.array_equal(got, want)


def test_tsqr_reference_application(apps):
    rng = np.random.default_rng(5)
    X = rng.standard_normal((40_000, 32))
    res = []
    for app in apps:
        x = app.array(X, (5_000, 32))
        R = app.indirect_tsr(x).get()
        Q, R2 = app.indirect_tsqr(x)
        Qd, Rd = app.direct_tsqr(x)
        res.append((R, Q.get(), R2.get(), Qd.get(), Rd.get()))
    (R, Q, R2, Qd, Rd), (Rr, _Qr, _R2r, _Qdr, Rdr) = res
    assert rel_fro(canon_r(R), canon_r(Rr)) <= 1e-10
    assert rel_fro(canon_r(Rd), canon_r(Rdr)) <= 1e-10
    assert rel_fro(Q @ R2, X) <= 1e-12 and rel_fro(Qd @ Rd, X) <= 1e-12
    assert np.linalg.norm(Q.T @ Q - np.eye(32)) <= 1e-10 and np.linalg.norm(Qd.T @ Qd - np.eye(32)) <= 1e-10


def test_newton_logistic_regression_glms(apps):
    """glms.newton (glms.py:362-372) with the reference's LogisticRegression on both back ends."""
    from nums.models.glms import LogisticRegression, newton
    rng = np.random.default_rng(6)
    n, d = 40_000, 28
    X = rng.standard_normal((n, d))
    theta = rng.standard_normal(d) / np.sqrt(d)
    y = (rng.random(n) < 1.0 / (1.0 + np.exp(-X @ theta))).astype(np.float64)
    betas = []
    for app in apps:
        model = LogisticRegression(solver="newton", penalty="none", tol=1e-8, max_iter=10)
        model._app = app
        xb, yb = app.array(X, (5_000, d)), app.array(y, (5_000,))
        beta = newton(app, model, app.zeros((d,), (d,), dtype=np.float64), xb, yb, app.scalar(1e-8), 10)
        betas.append(beta.get())
    assert rel_fro(betas[0], betas[1]) <= 1e-10
    assert np.linalg.norm(betas[0] - theta) < 0.2


def test_reductions_and_argops(nps_cuda):
    nps, _app = nps_cuda
    rng = np.random.default_rng(7)
    x = rng.standard_normal((300, 40))
    ba = nps.array(x).reshape(block_shape=(64, 16))
    for axis in (None, 0, 1):
        for name in ("sum", "min", "max", "mean", "std", "var"):
            got = getattr(nps, name)(ba, axis=axis).get()
            want = getattr(np, name)(x, axis=axis)
            assert np.allclose(got, want, rtol=1e-12, atol=1e-13), (name, axis)
    v = rng.standard_normal(1000)
    bv = nps.array(v).reshape(block_shape=(128,))
    assert int(nps.argmax(bv).get()) == int(np.argmax(v)) and int(nps.argmin(bv).get()) == int(np.argmin(v))


def test_fused_newton_through_glms(apps):
    """nums_b200.glms_fused: the optional lr_grad_hess / newton_step kernels behind glms.newton and
    LogisticRegression.fit, against the reference's own newton on numpy_compute."""
    from nums.models import glms
    from nums.models.glms import LogisticRegression
    from nums_b200 import glms_fused
    from nums_b200._lib import LIB
    cuda_app, serial_app = apps
    rng = np.random.default_rng(16)
    n, d = 64_000, 28
    X = rng.standard_normal((n, d))
    theta = rng.standard_normal(d) / np.sqrt(d)
    y = (rng.random(n) < 1.0 / (1.0 + np.exp(-X @ theta))).astype(np.float64)
    model = LogisticRegression(solver="newton", penalty="none")
    model._app = serial_app
    want = glms.newton(serial_app, model, serial_app.zeros((d,), (d,), dtype=np.float64), serial_app.array(X, (8_000, d)),
                       serial_app.array(y, (8_000,)), serial_app.scalar(1e-8), 10).get()
    model._app = cuda_app
    xb, yb = cuda_app.array(X, (8_000, d)), cuda_app.array(y, (8_000,))
    launches = LIB.dll.nums_launch_count()
    got = glms_fused.newton(cuda_app, model, cuda_app.zeros((d,), (d,), dtype=np.float64), xb, yb,
                            cuda_app.scalar(1e-8), 10).get()
    fused_launches = LIB.dll.nums_launch_count() - launches
    assert rel_fro(got, want) <= 1e-10
    launches = LIB.dll.nums_launch_count()
    unfused = glms.newton(cuda_app, model, cuda_app.zeros((d,), (d,), dtype=np.float64), xb, yb,
                          cuda_app.scalar(1e-8), 10).get()
    assert rel_fro(unfused, want) <= 1e-10
    assert fused_launches * 5 < LIB.dll.nums_launch_count() - launches      # ~10 launches / iteration instead of ~130
    # d = 7 is not served by the fused kernels: the same entry point falls back to the reference's newton
    X7 = X[:, :7].copy()
    b7 = glms_fused.newton(cuda_app, model, cuda_app.zeros((7,), (7,), dtype=np.float64), cuda_app.array(X7, (8_000, 7)),
                           yb, cuda_app.scalar(1e-8), 5).get()
    model._app = serial_app
    w7 = glms.newton(serial_app, model, serial_app.zeros((7,), (7,), dtype=np.float64), serial_app.array(X7, (8_000, 7)),
                     serial_app.array(y, (8_000,)), serial_app.scalar(1e-8), 5).get()
    assert rel_fro(b7, w7) <= 1e-10


def test_fit_and_predict_with_repaired_intercept(nps_cuda):
    """LogisticRegression.fit through the installed fused newton, and the intercept repair of INTEGRATION.md
    section 3: after ``fit_with_intercept`` the model's ``predict`` works (it raises on the unmodified fork,
    SURVEY.md section 0.3)."""
    nps, app = nps_cuda
    from nums.models import glms
    from nums_b200 import glms_fused
    rng = np.random.default_rng(17)
    n, d = 20_000, 12
    X = rng.standard_normal((n, d))
    theta, bias = rng.standard_normal(d), 0.7
    y = (rng.random(n) < 1.0 / (1.0 + np.exp(-(X @ theta + bias)))).astype(np.float64)
    glms_fused.install()
    try:
        assert glms.newton is glms_fused.newton
        xb, yb = app.array(X, (5_000, d)), app.array(y, (5_000,))
        model = glms.LogisticRegression(solver="newton", penalty="none", tol=1e-8, max_iter=10)
        model.fit(xb, yb)                                 # the fork's fit: no intercept column, beta split anyway
        model = glms.LogisticRegression(solver="newton", penalty="none", tol=1e-8, max_iter=10)
        glms_fused.fit_with_intercept(model, xb, yb)
        beta, beta0 = model._beta.get(), float(model._beta0.get())
        assert beta.shape == (d,) and abs(beta0 - bias) < 0.1 and np.linalg.norm(beta - theta) < 0.2 * np.linalg.norm(theta)
        pred = model.predict(xb).get()
        assert pred.shape == (n,) and np.mean(pred == y) > 0.8
    finally:
        glms_fused.uninstall()
