"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's per-block kernels.

``oracle/`` is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this file.  ``nums_b200/`` never does.

What is restated: ``nums.core.systems.numpy_compute.ComputeCls``
(``/root/reference/nums/core/systems/numpy_compute.py:84-286``), the 28-method
``ComputeInterface`` (``interfaces.py:73-167``) the new ``cuda_compute`` is a
drop-in for.  The arithmetic itself lives in third-party code that is not part
of the reference tree: NumPy (``setup.py:21`` pins ``numpy>1.18.0,<=1.20.0``;
this image has 2.3.5 on OpenBLAS) and SciPy (unpinned, ``setup.py:23``).  Each
function below names the reference line it follows and then calls the same
NumPy/SciPy routine the reference calls there.

Pinned by: ``tests/test_oracle.py`` replays the golden kernel-call traces that
``oracle/make_golden.py`` recorded from the *real* reference (imported through
``oracle/ref_loader.py`` in the build container) and, when ``/root/reference``
is present, compares every method with the reference's own ``ComputeCls`` on
seeded inputs.  The reference's test-suite holds no golden vectors for this
path, so those recorded traces are the known-answer set.
"""
from functools import reduce as _fold

import numpy as np
import scipy.special

# nums/core/settings.py:48-61 -- short operator names -> NumPy ufunc names.
UFUNC_ALIASES = dict(truediv="true_divide", sub="subtract", pow="power", mult="multiply",
                     mul="multiply", tensordot="multiply", lt="less", le="less_equal",
                     gt="greater", ge="greater_equal", eq="equal", ne="not_equal")


def _axis_batches(dim, step):
    """storage/utils.py:45-62 (Batch.get_batches): [start, stop) pairs along one axis."""
    if dim < step:
        return [(0, dim)]
    edges = list(range(0, dim, step)) + [dim]
    return [(edges[i], edges[i + 1]) for i in range(len(edges) - 1) if edges[i] < edges[i + 1]]


def grid_block_shape(grid_meta, grid_entry):
    """storage.py:29-86 (ArrayGrid.get_block_shape) from a ``to_meta()`` dict."""
    out = []
    for dim, step, idx in zip(grid_meta["shape"], grid_meta["block_shape"], grid_entry):
        lo, hi = _axis_batches(dim, step)[idx]
        out.append(hi - lo)
    return tuple(out)


def _np_dtype(name):
    if isinstance(name, str):
        return {"int": np.int64, "float": np.float64, "bool": np.bool_}.get(name) or getattr(np, name)
    return name


def _rng(seed, jump):
    # numpy_compute.py:29-30
    return np.random.Generator(np.random.PCG64(seed).jumped(jump))


class OracleRNG(object):
    """numpy_compute.py:70-81: (seed, jump_index) hand-out, one jump per sampled block."""

    def __init__(self, seed=None, jump_index=0):
        import random
        self.seed = random.getrandbits(128) if seed is None else seed
        self.rng = np.random.PCG64(self.seed)
        self.jump_index = jump_index

    def new_block_rng_params(self):
        self.jump_index += 1
        return self.seed, self.jump_index - 1


class OracleCompute(object):
    """Same method names / parameters as the reference ``ComputeCls`` (minus ``syskwargs``)."""

    # -- creation ----------------------------------------------------------------
    def touch(self, arr):                                   # :88-89
        return isinstance(arr, np.ndarray)

    def empty(self, grid_entry, grid_meta):                 # :91-94
        return np.empty(grid_block_shape(grid_meta, grid_entry), dtype=_np_dtype(grid_meta["dtype"]))

    def new_block(self, op_name, grid_entry, grid_meta):    # :96-104
        shape = grid_block_shape(grid_meta, grid_entry)
        dt = _np_dtype(grid_meta["dtype"])
        if op_name == "eye":
            if len(set(grid_entry)) > 1:
                raise AssertionError("eye blocks exist on the grid diagonal only")
            return np.eye(*shape, dtype=dt)
        return getattr(np, op_name)(shape, dtype=dt)

    def random_block(self, rng_params, rfunc_name, rfunc_args, shape, dtype):   # :106-113
        sample = getattr(_rng(*rng_params), rfunc_name)(*rfunc_args).reshape(shape)
        return sample if rfunc_name in ("random", "integers") else sample.astype(dtype)

    def permutation(self, rng_params, size):                # :115-117
        return _rng(*rng_params).permutation(size)

    def diag(self, arr):                                    # :171-172
        return np.diag(arr)

    def arange(self, start, stop, step, dtype):             # :174-175
        return np.arange(start, stop, step, dtype)

    # -- data movement -------------------------------------------------------------
    def create_block(self, *src_arrs, src_params, dst_params, dst_shape, dst_shape_bc):   # :119-132
        assert len(src_params) == len(dst_params)
        out = np.empty(dst_shape, dtype=src_arrs[0].dtype)
        target = out if dst_shape_bc is None else out.reshape(dst_shape_bc)
        for src, (src_sel, src_t), (dst_sel, _dst_t) in zip(src_arrs, src_params, dst_params):
            target[dst_sel] = (src.T if src_t else src)[src_sel]
        return out

    def update_block(self, dst_arr, *src_arrs, src_params, dst_params):   # :134-152
        assert len(src_params) == len(dst_params)
        out = np.array(dst_arr, copy=True)
        if dst_params[0][1]:
            out = out.T
        for src, (src_sel, src_bc, src_t), (dst_sel, _t) in zip(src_arrs, src_params, dst_params):
            view = src.T if src_t else src
            if src_bc is not None:
                view = view.reshape(src_bc)
            out[dst_sel] = view[src_sel]
        return out

    def update_block_by_index(self, dst_arr, src_arr, index_pairs):       # :154-158
        out = np.array(dst_arr, copy=True)
        for dst_index, src_index in index_pairs:
            out[tuple(dst_index)] = src_arr[tuple(src_index)]
        return out

    def update_block_along_axis(self, dst_arr, src_arr, index_pairs, axis):   # :160-169
        out = np.array(dst_arr, copy=True)
        for dst_index, src_index in index_pairs:
            d = [slice(None)] * dst_arr.ndim
            s = [slice(None)] * src_arr.ndim
            d[axis], s[axis] = dst_index, src_index
            out[tuple(d)] = src_arr[tuple(s)]
        return out

    def transpose(self, arr):                               # :213-214
        return arr.T

    def reshape(self, arr, shape):                          # :285-286
        return arr.reshape(shape)

    def split(self, arr, indices_or_sections, axis, transposed):   # :216-219
        return np.split(arr.T if transposed else arr, indices_or_sections, axis)

    def astype(self, arr, dtype_str):                       # :206-208
        return arr.astype(getattr(np, dtype_str) if hasattr(np, dtype_str) else _np_dtype(dtype_str))

    # -- elementwise -----------------------------------------------------------------
    def map_uop(self, op_name, arr, args, kwargs):          # :184-186
        return getattr(np, op_name)(arr, *args, **kwargs)

    def xlogy(self, arr_x, arr_y):                          # :203-204
        return scipy.special.xlogy(arr_x, arr_y)

    def bop(self, op, a1, a2, a1_shape, a2_shape, a1_T, a2_T, axes):   # :221-238
        lhs = a1.T if a1_T else a1
        rhs = a2.T if a2_T else a2
        if tuple(lhs.shape) != tuple(a1_shape):
            lhs = lhs.reshape(a1_shape)
        if tuple(rhs.shape) != tuple(a2_shape):
            rhs = rhs.reshape(a2_shape)
        if op == "tensordot":
            return np.tensordot(lhs, rhs, axes=axes)
        name = UFUNC_ALIASES.get(op, op)
        fn = getattr(np, name, None)
        if fn is None:
            fn = getattr(scipy.special, name)
        return fn(lhs, rhs)

    # -- reductions ------------------------------------------------------------------
    def reduce_axis(self, op_name, arr, axis, keepdims, transposed):   # :177-181
        return getattr(np, op_name)(arr.T if transposed else arr, axis=axis, keepdims=keepdims)

    def sum_reduce(self, *arrs):                            # :210-211
        return np.add.reduce(arrs)

    def arg_op(self, op_name, arr, block_slice, other_argoptima=None, other_optima=None):   # :269-283
        if op_name not in ("argmin", "argmax"):
            raise Exception("Unsupported arg op.")
        local = getattr(np, op_name)(arr)          # np.intp, as in the reference: the result is a NumPy scalar
        best = arr[local]
        if other_optima is not None:
            carried_wins = other_optima < best if op_name == "argmin" else other_optima > best
            if carried_wins:
                return other_argoptima, other_optima
        return block_slice.start + local, best

    def where(self, arr, x, y, block_slice_tuples):         # :188-201
        if x is None:
            assert y is None
            res = np.where(arr)
            for axis, (start, _stop) in enumerate(block_slice_tuples):
                res[axis][...] += start
        else:
            assert isinstance(x, np.ndarray) and isinstance(y, np.ndarray)
            res = np.where(arr, x, y)
        return tuple(list(res) + [res[0].shape])

    def allclose(self, a, b, rtol, atol):                   # :261-262
        return np.allclose(a, b, rtol, atol)

    def logical_and(self, *bool_list):                      # :266-267
        return np.all(bool_list)

    # -- dense linear algebra -----------------------------------------------------------
    def qr(self, *arrays, mode="reduced", axis=None):       # :240-246
        if len(arrays) > 1:
            assert axis is not None
            mat = np.concatenate(arrays, axis=axis)
        else:
            mat = arrays[0]
        return np.linalg.qr(mat, mode=mode)

    def cholesky(self, arr):                                # :248-249
        return np.linalg.cholesky(arr)

    def svd(self, arr):                                     # :251-254
        u, s, vt = np.linalg.svd(arr)
        return u[:s.shape[0]], s, vt

    def inv(self, arr):                                     # :256-257
        return np.linalg.inv(arr)


# What the reference System expects of a compute module (systems.py:39-45).
ComputeCls = OracleCompute
RNG = OracleRNG


def sum_list(arrs):
    """Left fold used by callers that emulate BlockArray's '+' chains (blockarray.py:402-407)."""
    return _fold(np.add, arrs)
