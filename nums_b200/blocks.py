"""Minimal block-array host layer for the hot path.

The reference's ``BlockArray`` / ``ArrayApplication`` / ``glms.newton`` (pure Python, ~3 kLoC)
cannot travel to the GPU box, so this module restates -- for the operations on the hot path
only -- *which per-block kernel calls they issue and in which order*:

* elementwise operators, one ``system.bop`` per block pair (blockarray.py:582-610 via
  base.py:167-246), scalars entering as float32 0-d blocks (blockarray.py:47-58);
* ``tensordot`` dispatch to ``_vecdot`` / ``_matvec`` / ``_tensordot`` (blockarray.py:416-580):
  matvec = one ``bop('tensordot')`` per block + one ``sum_reduce`` per block row, general case =
  ``dot`` then ``add`` accumulating left-to-right in k;
* ``reduce_axis`` two-stage (blockarray.py:343-408): per-block ``reduce_axis`` then an ``add``
  chain (sum) or ``fmin``/``fmax`` chain (min/max);
* lazy transpose (base.py:72-85), ``(n,) -> (n,1)`` reshape (blockarray.py:702-854);
* TSQR ``indirect_tsr`` / ``indirect_tsqr`` / ``direct_tsqr`` (application.py:772-933), ``inv``
  (application.py:956-977);
* Newton's method for logistic regression (glms.py:213-240, :362-372).

``tests/test_call_trace.py`` checks that the call sequences produced here are identical to the
ones the real reference produces (recorded in tests/golden/).  The same code drives either
``CudaSystem`` (product) or, in tests and the CPU baseline, an oracle-backed system.
"""
import functools
import itertools

import numpy as np

from nums_b200.grid import ArrayGrid

_PAIRWISE = {"min": "fmin", "amin": "fmin", "max": "fmax", "amax": "fmax"}   # settings.py:63-68
_UFUNC_ALIASES = {"truediv": "true_divide", "sub": "subtract", "pow": "power", "mul": "multiply",
                  "tensordot": "multiply", "lt": "less", "le": "less_equal", "gt": "greater",
                  "ge": "greater_equal", "eq": "equal", "ne": "not_equal"}


@functools.lru_cache(maxsize=None)
def _bop_dtype(op, dt_a, dt_b):
    """array/utils.py:33-42: run the ufunc on two 0-d arrays to learn the output dtype."""
    import scipy.special
    name = _UFUNC_ALIASES.get(op, op)
    fn = getattr(np, name, None) or getattr(scipy.special, name)
    return fn(np.array(1, dtype=dt_a), np.array(2, dtype=dt_b)).dtype.type


@functools.lru_cache(maxsize=None)
def _uop_dtype(op, dt):
    return getattr(np, op)(np.array(1, dtype=dt)).dtype.type


@functools.lru_cache(maxsize=None)
def _reduce_dtype(op, dt):
    return getattr(np, op)(np.array([0, 1], dtype=dt)).dtype.type


class Block(object):
    """A handle on one block: geometry + the opaque object id the system returned."""
    __slots__ = ("grid_entry", "grid_shape", "shape", "dtype", "transposed", "oid", "system")

    def __init__(self, system, grid_entry, grid_shape, shape, dtype, transposed=False, oid=None):
        self.system = system
        self.grid_entry = tuple(grid_entry)
        self.grid_shape = tuple(grid_shape)
        self.shape = tuple(shape)
        self.dtype = dtype
        self.transposed = transposed
        self.oid = oid

    def _sys(self):
        return {"grid_entry": self.grid_entry, "grid_shape": self.grid_shape}

    def transpose(self):
        return Block(self.system, reversed(self.grid_entry), reversed(self.grid_shape), reversed(self.shape),
                     self.dtype, not self.transposed, self.oid)

    def uop(self, op_name):
        out = Block(self.system, self.grid_entry, self.grid_shape, self.shape, _uop_dtype(op_name, self.dtype),
                    self.transposed)
        out.oid = self.system.map_uop(op_name, self.oid, (), {}, syskwargs=out._sys())
        return out

    def astype(self, dtype):
        out = Block(self.system, self.grid_entry, self.grid_shape, self.shape, dtype, self.transposed)
        out.oid = self.system.astype(self.oid, np.dtype(dtype).type.__name__, syskwargs=out._sys())
        return out

    def reduce_axis(self, op_name, axis, keepdims):
        entry, gshape, shape = [], [], []
        for ax in range(len(self.shape)):
            if axis is None or ax == axis:
                if keepdims:
                    entry.append(0); gshape.append(1); shape.append(1)
                continue
            entry.append(self.grid_entry[ax]); gshape.append(self.grid_shape[ax]); shape.append(self.shape[ax])
        out = Block(self.system, entry, gshape, shape, _reduce_dtype(op_name, self.dtype))
        out.oid = self.system.reduce_axis(op_name=op_name, arr=self.oid, axis=axis, keepdims=keepdims,
                                          transposed=self.transposed, syskwargs=out._sys())
        return out

    def bop(self, op, other, axes=None):
        if op == "tensordot":
            entry = self.grid_entry[:len(self.grid_entry) - axes] + other.grid_entry[axes:]
            gshape = self.grid_shape[:len(self.grid_shape) - axes] + other.grid_shape[axes:]
            shape = self.shape[:len(self.shape) - axes] + other.shape[axes:]
        else:
            entry, gshape, shape = [], [], []
            na, nb = len(self.shape), len(other.shape)
            for back in range(1, max(na, nb) + 1):   # broadcasting from the trailing axis
                ia, ib = na - back, nb - back
                mine = ib < 0 or (ia >= 0 and other.shape[ib] < self.shape[ia])
                src, i = (self, ia) if mine else (other, ib)
                entry.append(src.grid_entry[i]); gshape.append(src.grid_shape[i]); shape.append(src.shape[i])
            entry.reverse(); gshape.reverse(); shape.reverse()
        out = Block(self.system, entry, gshape, shape, _bop_dtype(op, self.dtype, other.dtype))
        out.oid = self.system.bop(op, self.oid, other.oid, self.shape, other.shape, self.transposed,
                                  other.transposed, axes=axes, syskwargs=out._sys())
        return out


    # NumPy reduces object arrays of Blocks by calling these (blockarray.py:402-407)
    def __add__(self, other): return self.bop("add", other)
    def __mul__(self, other): return self.bop("mul", other)


class BlockArray(object):

    def __init__(self, grid, system, blocks=None):
        self.grid = grid
        self.system = system
        self.shape = grid.shape
        self.block_shape = grid.block_shape
        self.dtype = grid.dtype
        if blocks is None:
            blocks = np.empty(grid.grid_shape, dtype=object)
            for entry in grid.get_entry_iterator():
                blocks[entry] = Block(system, entry, grid.grid_shape, grid.get_block_shape(entry), grid.dtype)
        self.blocks = blocks

    # -- construction / materialisation ---------------------------------------------------------
    @classmethod
    def from_np(cls, arr, block_shape, system):
        arr = np.asarray(arr)
        grid = ArrayGrid(arr.shape, block_shape, arr.dtype.type.__name__)
        out = cls(grid, system)
        for entry in grid.get_entry_iterator():
            out.blocks[entry].oid = system.put(arr[grid.get_slice(entry)])
        return out

    @classmethod
    def from_scalar(cls, val, system):
        if isinstance(val, (int, np.integer)):
            arr = np.array(val, dtype=np.int64)
        elif isinstance(val, (float, np.floating)):
            arr = np.array(val, dtype=np.float32)   # blockarray.py:50-51
        else:
            arr = np.asarray(val)
        return cls.from_np(arr, (), system)

    @classmethod
    def from_blocks(cls, blocks, system):
        """Rebuild an array from a grid of result blocks (blockarray.py:85-108)."""
        first = blocks.flat[0] if blocks.shape != () else blocks[()]
        if blocks.shape == ():
            grid = ArrayGrid(first.shape, first.shape, np.dtype(first.dtype).type.__name__)
            out = cls(grid, system)
            out.blocks[()] = first
            return out
        shape = []
        for ax in range(blocks.ndim):
            idx = [0] * blocks.ndim
            total = 0
            for i in range(blocks.shape[ax]):
                idx[ax] = i
                total += blocks[tuple(idx)].shape[ax]
            shape.append(total)
        grid = ArrayGrid(tuple(shape), first.shape, np.dtype(first.dtype).type.__name__)
        return cls(grid, system, blocks)

    def get(self):
        entries = list(self.grid.get_entry_iterator())
        if len(entries) > 1 and hasattr(self.system, "get_assembled"):
            # device-side assembly + one transfer (same result as the block-by-block path below)
            return self.system.get_assembled(self.grid, [(e, self.blocks[e].oid) for e in entries])
        out = np.zeros(self.shape, dtype=self.dtype)
        values = self.system.get([self.blocks[e].oid for e in entries])
        for entry, value in zip(entries, values):
            block = self.blocks[entry]
            value = np.asarray(value)
            out[self.grid.get_slice(entry)] = value.reshape(block.shape)
        return out

    def touch(self):
        oids = [self.system.touch(self.blocks[e].oid, syskwargs=self.blocks[e]._sys())
                for e in self.grid.get_entry_iterator()]
        self.system.get(oids)
        return self

    @property
    def T(self):
        grid = ArrayGrid(tuple(reversed(self.shape)), tuple(reversed(self.block_shape)), self.dtype.__name__)
        out = BlockArray(grid, self.system)
        out.blocks = np.copy(self.blocks.T)
        for entry in grid.get_entry_iterator():
            out.blocks[entry] = out.blocks[entry].transpose()
        return out

    # -- elementwise --------------------------------------------------------------------------------
    def _coerce(self, other):
        if isinstance(other, BlockArray):
            return other
        return BlockArray.from_scalar(other, self.system)

    def _elementwise(self, op, other):
        other = self._coerce(other)
        a, b = np.broadcast_arrays(self.blocks, other.blocks) if self.blocks.shape != other.blocks.shape \
            else (self.blocks, other.blocks)
        result = np.empty(a.shape, dtype=object)
        for idx in np.ndindex(*a.shape):
            result[idx] = a[idx].bop(op, b[idx])
        return BlockArray.from_blocks(result, self.system)

    def __add__(self, other): return self._elementwise("add", other)
    def __sub__(self, other): return self._elementwise("sub", other)
    def __mul__(self, other): return self._elementwise("mul", other)
    def __truediv__(self, other): return self._elementwise("truediv", other)
    def __pow__(self, other): return self._elementwise("pow", other)
    __radd__ = __add__      # blockarray.py:668: the array stays the left operand
    def __rsub__(self, other): return self._coerce(other)._elementwise("sub", self)
    __rmul__ = __mul__      # blockarray.py:674
    def __rtruediv__(self, other): return self._coerce(other)._elementwise("truediv", self)
    def __neg__(self): return self._elementwise("mul", -1.0)      # blockarray.py:688-689: -1 * x
    def __le__(self, other): return self._elementwise("le", other)
    def __lt__(self, other): return self._elementwise("lt", other)
    def __ge__(self, other): return self._elementwise("ge", other)
    def __gt__(self, other): return self._elementwise("gt", other)

    def __bool__(self):
        """The one host synchronisation of a Newton iteration (blockarray.py:620-628)."""
        if np.dtype(self.dtype) == np.bool_ and int(np.sum(self.shape)) == len(self.shape):
            return bool(self.get())
        return True

    def ufunc(self, op_name):
        out = BlockArray(ArrayGrid(self.shape, self.block_shape, _uop_dtype(op_name, self.dtype).__name__),
                         self.system)
        for entry in self.grid.get_entry_iterator():
            out.blocks[entry] = self.blocks[entry].uop(op_name)
        return out

    def astype(self, dtype):
        out = BlockArray(ArrayGrid(self.shape, self.block_shape, np.dtype(dtype).type.__name__), self.system)
        for entry in self.grid.get_entry_iterator():
            out.blocks[entry] = self.blocks[entry].astype(dtype)
        return out

    def column_view(self):
        """(n,) -> (n, 1) keeping the blocking (what ``reshape`` does in glms.py:235)."""
        assert len(self.shape) == 1
        grid = ArrayGrid((self.shape[0], 1), (self.block_shape[0], 1), self.dtype.__name__)
        out = BlockArray(grid, self.system)
        for (i,) in self.grid.get_entry_iterator():
            src = self.blocks[i]
            dst = out.blocks[i, 0]
            dst.oid = self.system.reshape(src.oid, dst.shape, syskwargs=dst._sys())
        return out

    # -- reductions ----------------------------------------------------------------------------------
    def reduce_axis(self, op_name, axis=None, keepdims=False):
        partial = np.empty_like(self.blocks, dtype=object)   # keeps the (possibly F-ordered) layout of .T grids
        for entry in self.grid.get_entry_iterator():
            partial[entry] = self.blocks[entry].reduce_axis(op_name, axis, keepdims)
        shape, bshape = [], []
        for ax in range(len(self.shape)):
            if axis is None or ax == axis:
                if keepdims:
                    shape.append(1); bshape.append(1)
                continue
            shape.append(self.shape[ax]); bshape.append(self.block_shape[ax])
        rdtype = _reduce_dtype(op_name, self.dtype)
        result = BlockArray(ArrayGrid(tuple(shape), tuple(bshape), rdtype.__name__), self.system)
        if op_name not in _PAIRWISE:
            # sum: NumPy's own reduction over the object array of Blocks decides the order of the
            # `add` chain (blockarray.py:402-407), so do literally the same thing
            folded = getattr(np, op_name)(partial, axis=axis, keepdims=keepdims)
            if result.shape == ():
                result.blocks[()] = folded
            else:
                result.blocks = folded
            return result
        combine = _PAIRWISE[op_name]   # min/max: explicit fmin/fmax chain (blockarray.py:370-401)
        for rentry in result.grid.get_entry_iterator() if result.shape != () else [()]:
            acc = None
            if axis is None:
                sources = list(self.grid.get_entry_iterator())
            else:
                sources = []
                for i in range(self.grid.grid_shape[axis]):
                    e = list(rentry)
                    if keepdims:
                        e[axis] = i
                    else:
                        e = e[:axis] + [i] + e[axis:]
                    sources.append(tuple(e))
            for e in sources:
                acc = partial[e] if acc is None else acc.bop(combine, partial[e])
            if axis is None and result.shape != ():
                result.blocks[:] = acc
            else:
                result.blocks[rentry] = acc
        return result

    # -- contractions -----------------------------------------------------------------------------------
    def __matmul__(self, other):
        return self.tensordot(other, 2 if len(self.shape) > 2 else 1)

    def tensordot(self, other, axes=1):
        def is_vector(ba, axis):
            if len(ba.shape) == 0:
                return False
            if len(ba.shape) == 1:
                return True
            rest = list(ba.shape[:axis]) + list(ba.shape[axis + 1:])
            return sum(rest) == len(rest) <= 1 < ba.shape[axis]
        if is_vector(self, len(self.shape) - 1) and is_vector(other, 0):
            return self._vecdot(other)
        if len(self.shape) == 2 and (len(other.shape) == 1 or (len(other.shape) == 2 and other.shape[1] == 1)):
            return self._matvec(other)
        return self._tensordot(other, axes)

    def _result(self, shape, block_shape, dtype):
        return BlockArray(ArrayGrid(tuple(shape), tuple(block_shape), np.dtype(dtype).type.__name__), self.system)

    def _tensordot(self, other, axes):
        mine, mine_sum = self.grid.grid_shape[:-axes], self.grid.grid_shape[-axes:]
        theirs, theirs_sum = other.grid.grid_shape[axes:], other.grid.grid_shape[:axes]
        assert mine_sum == theirs_sum
        result = self._result(self.shape[:-axes] + other.shape[axes:], self.block_shape[:-axes] + other.block_shape[axes:],
                              _bop_dtype("tensordot", self.dtype, other.dtype))
        for i in itertools.product(*map(range, mine)):
            for j in itertools.product(*map(range, theirs)):
                acc = None
                for k in itertools.product(*map(range, mine_sum)):
                    dot = self.blocks[i + k].bop("tensordot", other.blocks[k + j], axes=axes)
                    acc = dot if acc is None else acc.bop("add", dot)
                result.blocks[i + j] = acc
        return result

    def _dot_call(self, a, b, sch_entry, sch_shape):
        return self.system.bop("tensordot", a1=a.oid, a2=b.oid, a1_shape=a.shape, a2_shape=b.shape,
                               a1_T=a.transposed, a2_T=b.transposed, axes=1,
                               syskwargs={"grid_entry": sch_entry, "grid_shape": sch_shape})

    def _vecdot(self, other):
        result = self._result(self.shape[:-1] + other.shape[1:], self.block_shape[:-1] + other.block_shape[1:], self.dtype)
        na, nb = len(self.grid.grid_shape), len(other.grid.grid_shape)
        oids = []
        for i in range(self.grid.grid_shape[-1]):
            ea = tuple(i if ax == na - 1 else 0 for ax in range(na))
            eb = tuple(i if ax == 0 else 0 for ax in range(nb))
            a, b = self.blocks[ea], other.blocks[eb]
            if a.transposed != b.transposed and a.transposed:
                sch = (eb, other.grid.grid_shape)
            else:
                sch = (ea, self.grid.grid_shape)
            oids.append(self._dot_call(a, b, *sch))
        rentry = tuple(0 for _ in result.grid.grid_shape)
        result.blocks[rentry].oid = self.system.sum_reduce(
            *oids, syskwargs={"grid_entry": rentry, "grid_shape": result.grid.grid_shape})
        return result

    def _matvec(self, other):
        result = self._result(self.shape[:1] + other.shape[1:], self.block_shape[:1] + other.block_shape[1:], self.dtype)
        for i in range(self.grid.grid_shape[0]):
            row = []
            for j in range(self.grid.grid_shape[1]):
                a = self.blocks[i, j]
                if len(other.shape) == 2:
                    b, rentry = other.blocks[j, 0], (i, 0)
                else:
                    b, rentry = other.blocks[j], (i,)
                if a.transposed:
                    sch = ((j, i), tuple(reversed(self.grid.grid_shape)))
                else:
                    sch = ((i, j), self.grid.grid_shape)
                row.append(self._dot_call(a, b, *sch))
            result.blocks[rentry].oid = self.system.sum_reduce(
                *row, syskwargs={"grid_entry": rentry, "grid_shape": result.grid.grid_shape})
        return result


class ArrayApp(object):
    """The slice of ``ArrayApplication`` the hot path uses (application.py:35-1062)."""

    def __init__(self, system):
        self.system = system
        self.one = self.scalar(1.0)
        self.two = self.scalar(2.0)

    def scalar(self, value):
        return BlockArray.from_scalar(value, self.system)

    def array(self, arr, block_shape):
        return BlockArray.from_np(arr, block_shape, self.system)

    def _new(self, op, shape, block_shape, dtype):
        grid = ArrayGrid(shape, block_shape, np.dtype(dtype).type.__name__)
        out = BlockArray(grid, self.system)
        meta = grid.to_meta()
        for entry in grid.get_entry_iterator():
            out.blocks[entry].oid = self.system.new_block(op, entry, meta,
                                                          syskwargs={"grid_entry": entry, "grid_shape": grid.grid_shape})
        return out

    def zeros(self, shape, block_shape, dtype=np.float64):
        return self._new("zeros", shape, block_shape, dtype)

    def ones(self, shape, block_shape, dtype=np.float64):
        return self._new("ones", shape, block_shape, dtype)

    def exp(self, x): return x.ufunc("exp")
    def log(self, x): return x.ufunc("log")
    def abs(self, x): return x.ufunc("abs")
    def sum(self, x, axis=None, keepdims=False): return x.reduce_axis("sum", axis, keepdims)
    def max(self, x, axis=None, keepdims=False): return x.reduce_axis("max", axis, keepdims)
    def min(self, x, axis=None, keepdims=False): return x.reduce_axis("min", axis, keepdims)

    # -- TSQR (application.py:772-933) --------------------------------------------------------------------
    def indirect_tsr(self, X):
        assert len(X.shape) == 2 and X.block_shape[0] >= X.shape[1]
        g0, g1 = X.grid.grid_shape
        r_oids = []
        for i in range(g0):
            row = [X.blocks[i, j].oid for j in range(g1)]
            r_oids.append(self.system.qr(*row, mode="r", axis=1,
                                         syskwargs={"grid_entry": (i, 0), "grid_shape": (g0, 1),
                                                    "options": {"num_returns": 1}}))
        n = X.shape[1]
        R = BlockArray(ArrayGrid((n, n), (n, n), X.dtype.__name__), self.system)
        R.blocks[0, 0].oid = self.system.qr(*r_oids, mode="r", axis=0,
                                            syskwargs={"grid_entry": (0, 0), "grid_shape": (1, 1),
                                                       "options": {"num_returns": 1}})
        return R

    def inv(self, X):
        assert len(X.shape) == 2 and X.shape[0] == X.shape[1] and X.shape == X.block_shape, \
            "only single-block inverses are on the hot path"
        out = BlockArray(X.grid.copy(), self.system)
        out.blocks[0, 0].oid = self.system.inv(X.blocks[0, 0].oid, syskwargs={"grid_entry": (0, 0), "grid_shape": (1, 1)})
        return out

    def indirect_tsqr(self, X):
        assert X.grid.grid_shape[1] == 1, "column-blocked X needs the reshape path, not on the hot path"
        R = self.indirect_tsr(X)
        Q = X @ self.inv(R)
        return Q, R

    def qr(self, X):
        return self.indirect_tsqr(X)

    def direct_tsqr(self, X):
        g0, g1 = X.grid.grid_shape
        assert g1 == 1
        n = X.shape[1]
        q_oids, r_oids, dims = [], [], []
        for i in range(g0):
            m_i = X.grid.get_block_shape((i, 0))[0]
            k_i = min(m_i, n)
            dims.append((m_i, k_i))
            q, r = self.system.qr(X.blocks[i, 0].oid, mode="reduced", axis=1,
                                  syskwargs={"grid_entry": (i, 0), "grid_shape": (g0, 1), "options": {"num_returns": 2}})
            q_oids.append(q); r_oids.append(r)
        q2, r2 = self.system.qr(*r_oids, mode="reduced", axis=0,
                                syskwargs={"grid_entry": (0, 0), "grid_shape": (1, 1), "options": {"num_returns": 2}})
        Q = BlockArray(ArrayGrid(X.shape, (X.block_shape[0], n), X.dtype.__name__), self.system)
        pos = 0
        for i in range(g0):
            m_i, k_i = dims[i]
            sel = (slice(pos, pos + k_i), slice(0, n))
            q2_i = self.system.create_block(q2, src_params=[(sel, False)],
                                            dst_params=[((slice(0, k_i), slice(0, n)), False)],
                                            dst_shape=(k_i, n), dst_shape_bc=None,
                                            syskwargs={"grid_entry": (i, 0), "grid_shape": (g0, 1)})
            pos += k_i
            Q.blocks[i, 0].oid = self.system.bop("tensordot", q_oids[i], q2_i, a1_shape=(m_i, k_i), a2_shape=(k_i, n),
                                                 a1_T=False, a2_T=False, axes=1,
                                                 syskwargs={"grid_entry": (i, 0), "grid_shape": (g0, 1)})
        R = BlockArray(ArrayGrid((n, n), (n, n), X.dtype.__name__), self.system)
        R.blocks[0, 0].oid = r2
        return Q, R


class LogisticRegression(object):
    """link_inv / gradient / hessian of nums/models/glms.py:213-240 (no penalty)."""

    def __init__(self, app):
        self.app = app

    def forward(self, X, beta):
        return self.link_inv(X @ beta)

    def link_inv(self, eta):
        one = self.app.one
        return one / (one + self.app.exp(-eta))

    def gradient(self, X, y, mu):
        return X.T @ (mu - y)

    def hessian(self, X, y, mu):
        s = (mu * (self.app.one - mu)).column_view()
        return X.T @ (s * X)


def newton(app, model, beta, X, y, tol, max_iter):
    """glms.py:362-372."""
    iters = 0
    for _ in range(max_iter):
        iters += 1
        mu = model.forward(X, beta)
        g = model.gradient(X, y, mu)
        beta = beta + (-(app.inv(model.hessian(X, y, mu))) @ g)
        if app.max(app.abs(g)) <= tol:
            break
    return beta, iters


class FileSystem(object):
    """``FileSystem.read_csv`` of the reference (filesystem.py:402-439): the file is cut into
    ``num_workers`` byte ranges (``Batch.from_num_batches``, storage/utils.py:30-62), every range goes
    to the registered ``read_csv_block`` kernel (one call per range, ``num_returns = 2``), and each
    non-empty result becomes a single-block ``BlockArray``.  ``read_csv_block`` defaults to the device
    parser (``cuda_compute.read_csv_block``); the CPU tests pass the oracle's.  Also mirrors the array
    persistence calls (``write_fs`` / ``read_fs`` / ``delete_fs``) in the reference's on-disk format."""

    def __init__(self, system, read_csv_block=None, block_io=None):
        self.system = system
        if read_csv_block is None or block_io is None:
            from nums_b200 import cuda_compute
            read_csv_block = read_csv_block or cuda_compute.read_csv_block
            block_io = block_io or (cuda_compute.write_block_fs, cuda_compute.read_block_fs, cuda_compute.delete_block_fs)
        self.system.register("read_csv_block", read_csv_block, {})
        for name, func in zip(("write_block_fs", "read_block_fs", "delete_block_fs"), block_io):
            self.system.register(name, func, {})

    # -- array persistence: ArrayApplication.write_fs / read_fs / delete_fs (application.py:154-191,204-219) over
    #    the block functions above and the meta file of filesystem.py:66-93,308-343 --------------------------
    @staticmethod
    def _meta_path(filename):
        import os
        os.makedirs(filename, exist_ok=True)
        return os.path.join(filename, "meta.pkl")

    def write_fs(self, ba, filename):
        import pickle
        for entry in ba.grid.get_entry_iterator():
            self.system.call("write_block_fs", ba.blocks[entry].oid, filename, entry,
                             syskwargs={"grid_entry": entry, "grid_shape": ba.grid.grid_shape})
        meta = {"filename": filename, "grid_meta": ba.grid.to_meta(),
                "addresses": self.system.get_block_addresses(ba.grid)}
        with open(self._meta_path(filename), "wb") as fh:
            pickle.dump(meta, fh)
        return meta

    def read_meta_fs(self, filename):
        import pickle
        with open(self._meta_path(filename), "rb") as fh:
            return pickle.load(fh)

    def read_fs(self, filename):
        meta = self.read_meta_fs(filename)
        grid = ArrayGrid.from_meta(meta["grid_meta"])
        ba = BlockArray(grid, self.system)
        for entry in meta["addresses"]:
            ba.blocks[entry].oid = self.system.call("read_block_fs", filename, entry,
                                                    syskwargs={"grid_entry": entry, "grid_shape": grid.grid_shape})
        return ba

    def delete_fs(self, filename):
        import os
        meta = self.read_meta_fs(filename)
        grid = ArrayGrid.from_meta(meta["grid_meta"])
        for entry in meta["addresses"]:
            self.system.call("delete_block_fs", filename, entry,
                             syskwargs={"grid_entry": entry, "grid_shape": grid.grid_shape})
        os.remove(self._meta_path(filename))

    @staticmethod
    def byte_ranges(total_size, num_batches):
        batch_size = (total_size + num_batches - 1) // num_batches
        if total_size < batch_size:
            return [[0, total_size]]
        starts = list(range(0, total_size, batch_size))
        out = [starts[i:i + 2] for i in range(int(total_size / batch_size))]
        if len(out[-1]) == 1:
            out[-1].append(total_size)
        if out[-1][1] != total_size:
            out.append([out[-1][1], total_size])
        return out

    def read_csv(self, filename, dtype=np.float64, delimiter=",", has_header=False, num_workers=4):
        import os
        ranges = self.byte_ranges(os.path.getsize(filename), num_workers)
        parts = []
        for i, (file_start, file_end) in enumerate(ranges):
            parts.append(self.system.call("read_csv_block", filename, file_start, file_end, dtype, delimiter, has_header,
                                          syskwargs={"grid_entry": (i,), "grid_shape": (num_workers,),
                                                     "options": {"num_returns": 2}}))
        arrays = []
        for block_oid, shape in parts:
            shape = tuple(int(v) for v in self.system.get(shape)) if not isinstance(shape, tuple) else shape
            if shape[0] == 0:
                continue
            grid = ArrayGrid(shape, shape, np.dtype(dtype).type.__name__)
            arr = BlockArray(grid, self.system)
            arr.blocks[next(iter(grid.get_entry_iterator()))].oid = block_oid
            arrays.append(arr)
        return arrays
