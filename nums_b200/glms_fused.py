"""Fused Newton step for ``nums.models.glms``.

``glms.newton`` (/root/reference/nums/models/glms.py:362-372) spends one iteration in ~15 kernel calls per
row block -- ``forward`` (:140-143), ``gradient`` (:222-227), ``hessian`` (:232-238) -- streaming X about six
times and materialising an X-sized temporary.  When the system offers the optional kernels
``lr_grad_hess`` / ``newton_step`` (cuda_compute.EXTRA_KERNELS), ``newton`` below does the same iteration as

    one ``lr_grad_hess(X_i, y_i, beta)`` per row block (one pass over X_i, on the block's GPU),
    one ``sum_reduce`` of the (d + d*d,) partials (an all-reduce when the blocks live on several GPUs),
    one ``newton_step`` (solve, subtract, max |g|, singularity flag) + one 16-byte read-back,

through the same ``system.<kernel>(..., syskwargs=...)`` seam as every other call, so it runs unchanged on one
GPU (CudaSystem) and on N (SpmdSystem).  Anything the fused kernels do not cover (a penalty term, column-blocked
X, unsupported d, a model that is not logistic regression) falls back to the reference's own ``newton``.

``install()`` rebinds ``nums.models.glms.newton`` so that ``LogisticRegression(solver="newton").fit(X, y)``
takes this path -- the run-time form of the diff in INTEGRATION.md section 3.  ``fit_with_intercept`` is the
repaired ``GLM.fit`` of that diff: the fork commented the intercept column out (glms.py:108-112) but still
splits ``beta[-1]`` off as the intercept (:137-138), which breaks ``predict``.
"""
import numpy as np


def _fusable(app, model, beta, X, y):
    system = app.system
    methods = getattr(system, "methods", {})
    if "lr_grad_hess" not in methods or "newton_step" not in methods:
        return False
    if type(model).__name__ != "LogisticRegression" or getattr(model, "_penalty", None) is not None:
        return False
    if len(X.shape) != 2 or len(y.shape) != 1 or X.grid.grid_shape[1] != 1 or beta.grid.grid_shape != (1,):
        return False
    if X.grid.grid_shape[0] != y.grid.grid_shape[0] or X.block_shape[0] != y.block_shape[0]:
        return False
    d = X.shape[1]
    if d > 48 or d % 2 or np.dtype(X.dtype) != np.float64 or np.dtype(y.dtype) != np.float64:
        return False
    return not any(X.blocks[i, 0].transposed for i in range(X.grid.grid_shape[0]))


def newton(app, model, beta, X, y, tol, max_iter):
    """Drop-in for ``nums.models.glms.newton`` (same arguments, same result)."""
    from nums.core.array.blockarray import BlockArray
    if not _fusable(app, model, beta, X, y):
        return _reference_newton()(app, model, beta, X, y, tol, max_iter)
    system = app.system
    methods = getattr(system, "methods", {})
    G, d = X.grid.grid_shape[0], X.shape[1]
    tol_value = float(np.asarray(tol.get() if hasattr(tol, "get") else tol))
    one = {"grid_entry": (0,), "grid_shape": (1,)}
    beta_oid = beta.blocks[0].oid
    get_async = getattr(system, "get_async", None)
    pending = None            # (beta after that iteration, waiter for its {max |g|, info})

    def converged(entry):
        gmax, info = (float(v) for v in np.asarray(entry[1]()))
        if info != 0:
            raise np.linalg.LinAlgError("Singular matrix")
        return gmax <= tol_value

    # Row blocks that live on the same device go into ONE multi-block launch (up to 16 blocks each): group them by
    # their home (SpmdSystem handles carry it; on a single GPU everything is one group).
    groups = {}
    for i in range(G):
        groups.setdefault(getattr(X.blocks[i, 0].oid, "home", 0), []).append(i)
    multi = "lr_grad_hess_multi" in methods and any(len(g) > 1 for g in groups.values())
    for _ in range(max_iter):
        if multi:
            parts = []
            for members in groups.values():
                for lo in range(0, len(members), 16):
                    chunk = members[lo:lo + 16]
                    flat = []
                    for i in chunk:
                        flat.extend((X.blocks[i, 0].oid, y.blocks[i].oid))
                    parts.append(system.lr_grad_hess_multi(*flat, beta_oid,
                                                           syskwargs={"grid_entry": (chunk[0], 0), "grid_shape": (G, 1)}))
        else:
            parts = [system.lr_grad_hess(X.blocks[i, 0].oid, y.blocks[i].oid, beta_oid,
                                         syskwargs={"grid_entry": (i, 0), "grid_shape": (G, 1)}) for i in range(G)]
        gh = parts[0] if len(parts) == 1 else system.sum_reduce(*parts, syskwargs=one)
        beta_oid, status = system.newton_step(gh, beta_oid, syskwargs=one)
        # The convergence test of iteration i (glms.py:370) is read one iteration late: iteration i + 1 is already
        # enqueued when the host looks at the 16 status bytes of iteration i, so the device never idles on the
        # read-back.  If iteration i had converged, its beta is returned and the speculative step is dropped --
        # the same value the reference's loop returns.
        if get_async is None:
            waiter = (lambda v: (lambda: v))(system.get(status))
        else:
            waiter = get_async(status)
        if pending is not None and converged(pending):
            beta_oid = pending[0]
            pending = None
            break
        pending = (beta_oid, waiter)
    if pending is not None:
        converged(pending)       # surfaces a singular Hessian of the last iteration
    return BlockArray.from_oid(beta_oid, (d,), np.float64, system)


_original = None


def _reference_newton():
    from nums.models import glms
    return _original if _original is not None else glms.newton


def install():
    """``nums.models.glms.newton`` -> the fused version (idempotent)."""
    global _original
    from nums.models import glms
    if getattr(glms.newton, "_nums_b200", False):
        return
    _original = glms.newton
    newton._nums_b200 = True
    glms.newton = newton


def uninstall():
    global _original
    from nums.models import glms
    if _original is not None:
        glms.newton = _original
        _original = None


def fit_with_intercept(model, X, y):
    """``GLM.fit`` (glms.py:103-138) with the intercept column restored, i.e. the commented-out concatenation
    (:108-112) put back, so that ``beta[-1]`` really is the intercept that ``predict`` / ``forward`` add."""
    app = model._app
    ones = app.ones(shape=(X.shape[0], 1), block_shape=(X.block_shape[0], 1), dtype=X.dtype)
    X1 = app.concatenate([X, ones], axis=1, axis_block_size=X.block_shape[1])
    type(model).fit(model, X1, y)
    return model
