"""``cuda_compute`` -- the B200 implementation of NumS' per-block compute interface.

Drop-in beside ``nums.core.systems.numpy_compute``: this module exports ``ComputeCls`` (the 28
methods of ``ComputeInterface``, /root/reference/nums/core/systems/interfaces.py:73-167, with
the parameter names ``check_implementation`` compares, systems/utils.py:59-72) and ``RNG``
(numpy_compute.py:33-81).  Blocks ("oids") are ``torch.Tensor`` objects living in HBM; every
method enqueues hand-written sm_100a kernels from ``libnumscuda.so`` (C ABI in
``include/nums_cuda.h``) on the current CUDA stream and returns immediately -- the only
synchronisation points are ``touch`` and the host reads documented below.  PyTorch is used for
device memory, views (metadata) and streams only.

There is no CPU fallback: anything the kernel library does not implement raises
``NotImplementedError``; a missing ``libnumscuda.so`` raises on first use.
"""
import random

import numpy as np
import scipy.special
import torch

from nums_b200 import _lib
from nums_b200._lib import LIB, describe
from nums_b200.grid import ArrayGrid

try:  # when the reference package is importable, be a real ComputeImp / RNGInterface subclass
    from nums.core.systems.interfaces import ComputeImp as _ComputeImp, RNGInterface as _RNGInterface
except Exception:  # the GPU box has no reference tree
    class _ComputeImp(object):
        pass

    class _RNGInterface(object):
        pass

# nums/core/settings.py:48-61 -- BlockArray operator names -> NumPy ufunc names
_SHORT_OP_NAMES = {
    "truediv": "true_divide", "sub": "subtract", "pow": "power", "mult": "multiply",
    "mul": "multiply", "lt": "less", "le": "less_equal", "gt": "greater", "ge": "greater_equal",
    "eq": "equal", "ne": "not_equal", "divide": "true_divide", "mod": "remainder",
}

# How inv / cholesky report a singular / non-positive-definite block (np.linalg.LinAlgError in the reference):
#   "deferred" (default) -- the kernel's status word stays on the device and is examined at the next point where
#       the host waits for the device anyway (get / touch / synchronize): the error surfaces there, the way a
#       failed remote task of the reference's Ray systems surfaces at ray.get, and a loop such as glms.newton
#       (one inv per iteration, glms.py:368) keeps the host running ahead of the GPU;
#   "eager" -- a 4-byte read-back (one stream synchronisation) inside the call, raising from the call itself
#       exactly where NumPy would (NUMS_FACTORIZATION_STATUS=eager);
#   "off"   -- never checked.
CHECK_FACTORIZATION_STATUS = __import__("os").environ.get("NUMS_FACTORIZATION_STATUS", "deferred")
_PENDING_STATUS = []          # [(device scalar, exception to raise if it is "bad")] of kernels not yet examined
_PENDING_STATUS_MAX = 64


def check_pending_status():
    """Examine the status words of every data-dependent error check launched since the last look (the caller has
    synchronised, or is about to: each read is a few bytes).  Raises the first failure: LinAlgError for a singular /
    non-positive-definite factorization (status word != 0), ValueError for an integer power with a negative
    exponent (minimum of the exponent block < 0)."""
    if not _PENDING_STATUS:
        return
    pending = list(_PENDING_STATUS)
    del _PENDING_STATUS[:]
    for word, bad, exc in pending:
        if bad(word.cpu().item()):
            raise exc


def _nonzero(v):
    return int(v) != 0


def _negative(v):
    return v < 0


class RNG(_RNGInterface):
    """Same hand-out of (seed, jump_index) pairs as the reference (numpy_compute.py:70-81).

    Sampling itself stays on the host with NumPy's PCG64 so that seeded streams are bit-identical
    to the reference (tests/core/array/test_random.py:164-172); blocks are then uploaded.
    """

    def __init__(self, seed=None, jump_index=0):
        if seed is None:
            seed = random.getrandbits(128)
        self.seed = seed
        self.rng = np.random.PCG64(seed)
        self.jump_index = jump_index

    def new_block_rng_params(self):
        params = self.seed, self.jump_index
        self.jump_index += 1
        return params


def block_rng(seed, jump_index):
    return np.random.Generator(np.random.PCG64(seed).jumped(jump_index))


# ---------------------------------------------------------------------------------------------
# plumbing helpers (metadata + allocation only)
# ---------------------------------------------------------------------------------------------
_DEVICES = {}


def _device():
    """torch.device of this process' GPU (one process drives one GPU; cached per current device index --
    this sits on the per-block dispatch path)."""
    try:
        index = _raw_current_device() if (_DEVICES and _raw_current_device is not None) else torch.cuda.current_device()    # (the first call initialises CUDA)
    except Exception:  # noqa: BLE001 -- no driver / no device
        index = None
    dev = _DEVICES.get(index)
    if dev is None:
        if index is None or not torch.cuda.is_available():
            raise _lib.NumsCudaError("cuda_compute needs a CUDA device; there is no CPU fallback")
        dev = _DEVICES[index] = torch.device("cuda", index)
    return dev


class _Transfers(object):
    """Copy streams of the host<->HBM pipeline.

    ``upload`` of a large page-locked array runs on ``up`` and returns at once; the tensor carries
    ``_nums_ready = (sequence number, event)``.  Every kernel launch fetches its stream through
    ``_stream()``, which first makes that stream wait for *all* uploads issued so far -- except the
    deferred-contraction flush (deferred.py), which waits per launch group only for the operands the
    group reads, so the grouped GEMM starts while later blocks are still crossing PCIe.  ``down`` is
    used by ``CudaSystem.get_assembled`` to drain finished block rows while later rows compute.
    """
    up = None
    down = None
    seq = 0            # sequence number of the latest upload
    last_event = None
    awaited = {}       # stream handle -> highest sequence number that stream has waited for


def _upload_stream():
    if _Transfers.up is None:
        _Transfers.up = torch.cuda.Stream()
    return _Transfers.up


def _download_stream():
    if _Transfers.down is None:
        _Transfers.down = torch.cuda.Stream()
    return _Transfers.down


def await_uploads(stream=None, upto=None):
    """Order ``stream`` (default: current) after the uploads up to ``upto`` = (sequence number, event)
    (default: every upload issued so far)."""
    if _Transfers.last_event is None:
        return
    cur = torch.cuda.current_stream() if stream is None else stream
    key = cur.cuda_stream       # torch streams come from a pool and are never destroyed: handles are not recycled
    seq, event = (_Transfers.seq, _Transfers.last_event) if upto is None else upto
    if _Transfers.awaited.get(key, 0) < seq:
        cur.wait_event(event)
        _Transfers.awaited[key] = seq


# Raw handle of the current stream without building a torch.cuda.Stream object (that constructor was the
# largest single item of the per-kernel dispatch cost: ~8 us of ~45); the private accessor is the one
# torch.cuda.current_stream itself is built on, with the public API as the fallback.
try:
    _raw_current_stream = torch._C._cuda_getCurrentRawStream
    _raw_current_device = torch._C._cuda_getDevice
except AttributeError:  # pragma: no cover
    _raw_current_stream = _raw_current_device = None


def _stream_unordered():
    if _raw_current_stream is not None:
        return _raw_current_stream(_raw_current_device())
    return torch.cuda.current_stream().cuda_stream


def _stream():
    """Handle of the current stream, ordered after every upload issued so far (see _Transfers)."""
    if _raw_current_stream is not None:
        key = _raw_current_stream(_raw_current_device())
        if _Transfers.seq and _Transfers.awaited.get(key, 0) < _Transfers.seq:
            torch.cuda.current_stream().wait_event(_Transfers.last_event)
            _Transfers.awaited[key] = _Transfers.seq
        return key
    cur = torch.cuda.current_stream()
    key = cur.cuda_stream
    if _Transfers.seq and _Transfers.awaited.get(key, 0) < _Transfers.seq:
        cur.wait_event(_Transfers.last_event)
        _Transfers.awaited[key] = _Transfers.seq
    return key


def _empty(shape, dtype):
    return torch.empty(tuple(int(s) for s in shape), dtype=_lib.torch_dtype(dtype), device=_device())


def upload(value, dtype=None):
    """Host value -> device tensor on the current stream.

    Large page-locked sources (e.g. arrays backed by ``torch.empty(pin_memory=True)``) are copied
    asynchronously on the upload stream straight from where they are (see ``_Transfers``); pageable
    ones go through the driver's own staging on the current stream.

    Ownership: for the asynchronous (page-locked, >= 1 MiB) path the copy reads the caller's buffer AFTER
    this call returns -- the caller must not modify or free it until the block has been consumed (any
    ``get`` / ``synchronize``), unlike the reference's ``put`` and the pageable path here, which snapshot
    the array at the time of the call.
    """
    if isinstance(value, torch.Tensor):
        return value
    arr = np.asarray(value) if dtype is None else np.asarray(value, dtype=dtype)
    _lib.dtype_code(arr.dtype)  # raises for unsupported dtypes
    if not arr.flags.c_contiguous:
        arr = np.ascontiguousarray(arr)
    if arr.ndim == 0:
        host = torch.from_numpy(arr.reshape(1).copy())
        return host.to(_device()).view(())
    if not arr.flags.writeable:
        arr = arr.copy()
    return _upload_tensor(torch.from_numpy(arr))


def _upload_tensor(host):
    """Host tensor (any torch dtype, e.g. raw uint8 file bytes) -> device tensor."""
    if host.numel() * host.element_size() >= (1 << 20) and host.is_pinned():
        # asynchronous, on the upload stream; consumers order themselves through _stream()
        device = _device()
        home = torch.cuda.current_stream()
        with torch.cuda.stream(_upload_stream()):
            dev = torch.empty(host.shape, dtype=host.dtype, device=device)
            dev.copy_(host, non_blocking=True)
        dev.record_stream(home)
        event = torch.cuda.Event()
        event.record(_Transfers.up)
        _Transfers.seq += 1
        _Transfers.last_event = event
        dev._nums_ready = (_Transfers.seq, event)
        return dev
    return host.to(_device())


def upload_tag(t):
    """``(sequence number, event)`` of the asynchronous upload that fills tensor ``t``, or None.

    The tag is a Python attribute of the tensor object ``_upload_tensor`` returned, so views of it
    (``transpose`` / ``reshape`` / ``split`` results, user slices) do not carry it; for those the base
    tensor is consulted, and a view whose base cannot be identified is conservatively treated as
    depending on the latest upload issued."""
    if not isinstance(t, torch.Tensor):
        return None
    tag = getattr(t, "_nums_ready", None)
    if tag is not None:
        return tag
    base = t._base
    if base is not None:
        tag = getattr(base, "_nums_ready", None)
        if tag is not None:
            return tag
        if _Transfers.last_event is not None:
            # a view of a view-less base without a tag was either computed on the device (no upload to wait
            # for -- but then the kernel that produced it already ordered the stream) or its tagged Python
            # object is gone: wait for everything uploaded so far
            return (_Transfers.seq, _Transfers.last_event)
    return None


def _to_pinned(t):
    """Start an asynchronous D2H copy into page-locked memory (torch's caching host allocator
    recycles the buffers, so steady-state loops do not pay cudaHostAlloc)."""
    if not t.is_contiguous():
        t = _materialize(t)
    await_uploads()
    host = torch.empty(tuple(t.shape), dtype=t.dtype, pin_memory=True)
    host.copy_(t, non_blocking=True)
    return host


def download(t):
    """Device tensor -> numpy array (synchronises the current stream)."""
    if isinstance(t, Touched):
        torch.cuda.current_stream().synchronize()
        check_pending_status()
        return t.ok
    if not isinstance(t, torch.Tensor):
        return t
    if t.numel() * t.element_size() < (1 << 16):
        if not t.is_contiguous():
            t = _materialize(t)
        await_uploads()
        out = t.cpu().numpy()
        check_pending_status()
        return out
    host = _to_pinned(t)
    torch.cuda.current_stream().synchronize()
    check_pending_status()
    return host.numpy()


def download_many(tensors):
    """Several device tensors -> numpy arrays with one synchronisation (BlockArray.get)."""
    staged = [(_to_pinned(t) if isinstance(t, torch.Tensor) and t.numel() * t.element_size() >= (1 << 16) else None)
              for t in tensors]
    torch.cuda.current_stream().synchronize()
    check_pending_status()
    return [h.numpy() if h is not None else download(t) for h, t in zip(staged, tensors)]


def _copy_into(dst, src):
    """dst[...] = src with NumPy broadcasting and dtype conversion (nums_uop COPY)."""
    if dst.numel() == 0:
        return
    LIB.check(LIB.dll.nums_uop(_lib.UOP_CODE["copy"], _lib.dtype_code(dst.dtype), describe(src),
                               describe(dst), _stream()))


def _materialize(view, dtype=None):
    out = torch.empty(tuple(view.shape), dtype=view.dtype if dtype is None else dtype, device=view.device)
    _copy_into(out, view)
    return out


def _ravel_indices(indices, shape):
    """Multi-indices (NumPy semantics: negative values wrap, out of range raises IndexError) -> int64
    offsets into a C-ordered array of `shape`."""
    idx = np.asarray(indices, dtype=np.int64).reshape(len(indices), len(shape))
    dims = np.asarray(shape, dtype=np.int64)
    if np.any(idx < -dims) or np.any(idx >= dims):
        bad = idx[np.any((idx < -dims) | (idx >= dims), axis=1)][0]
        raise IndexError("index %s is out of bounds for shape %s" % (tuple(int(v) for v in bad), tuple(shape)))
    idx = np.where(idx < 0, idx + dims, idx)
    strides = np.ones(len(shape), dtype=np.int64)
    for k in range(len(shape) - 2, -1, -1):
        strides[k] = strides[k + 1] * dims[k + 1]
    return idx @ strides


def _scatter(dst, src, dst_index, src_index, outer, dst_len, src_len, inner):
    """dst[o, dst_index[p], i] = src[o, src_index[p], i] in one launch (nums_scatter_axis).  `dst` is a
    dense array this call may write; duplicate destinations keep their LAST pair, like the sequential
    loops of the reference (numpy_compute.py:154-169)."""
    if src.dtype != dst.dtype or not src.is_contiguous():
        src = _materialize(src, dst.dtype)            # ndarray assignment casts to the destination dtype
    dst_index = np.asarray(dst_index, dtype=np.int64)
    src_index = np.asarray(src_index, dtype=np.int64)
    if np.unique(dst_index).size != dst_index.size:
        last = {}
        for pos, d in enumerate(dst_index.tolist()):
            last[d] = pos
        keep = np.fromiter(sorted(last.values()), dtype=np.int64)
        dst_index, src_index = dst_index[keep], src_index[keep]
    both = upload(np.concatenate([dst_index, src_index]))
    n = int(dst_index.size)
    LIB.check(LIB.dll.nums_scatter_axis(dst.element_size(), int(outer), int(dst_len), int(src_len), int(inner), n,
                                        both.data_ptr(), both.data_ptr() + 8 * n, dst.data_ptr(), src.data_ptr(),
                                        _stream()))


def _carry_tag(view, src):
    """Views are new Python objects: keep the upload tag of the tensor they alias."""
    if view is not src:
        tag = getattr(src, "_nums_ready", None)
        if tag is not None:
            view._nums_ready = tag
    return view


def _transpose_view(t):
    return _carry_tag(t.permute(*reversed(range(t.dim()))), t) if t.dim() > 1 else t


def _reshape(t, shape):
    shape = tuple(int(s) for s in shape)
    if tuple(t.shape) == shape:
        return t
    try:
        return _carry_tag(t.view(shape), t)
    except RuntimeError:
        return _materialize(t).view(shape)


def _operand(a, shape, transposed):
    """numpy_compute.py:222-229: `.T` if flagged, then reshape if the stored shape differs."""
    a = upload(a)
    if transposed:
        a = _transpose_view(a)
    if tuple(a.shape) != tuple(shape):
        a = _reshape(a, shape)
    return a


def _broadcast_shape(s1, s2):
    """NumPy broadcasting of two shapes (pure Python: this sits on the per-block dispatch path)."""
    if s1 == s2:
        return s1
    n1, n2 = len(s1), len(s2)
    out = []
    for i in range(max(n1, n2)):
        d1 = s1[n1 - 1 - i] if i < n1 else 1
        d2 = s2[n2 - 1 - i] if i < n2 else 1
        if d1 != d2 and d1 != 1 and d2 != 1:
            raise ValueError("operands could not be broadcast together with shapes %s %s" % (s1, s2))
        out.append(d2 if d1 == 1 else d1)
    return tuple(reversed(out))


_REDUCE_TYPES = {}


def reduce_type(op_name, dt):
    key = (op_name, dt)
    hit = _REDUCE_TYPES.get(key)
    if hit is None:
        hit = getattr(np, op_name)(np.zeros((1,), dtype=dt)).dtype
        _REDUCE_TYPES[key] = hit
    return hit


def _np_dtype(name):
    if isinstance(name, str):
        return np.dtype({"int": np.int64, "float": np.float64, "bool": np.bool_}.get(name) or getattr(np, name))
    return np.dtype(name)


_BOP_TYPES = {}
_UOP_TYPES = {}


def _ufunc(name):
    fn = getattr(np, name, None)
    if fn is None:
        fn = getattr(scipy.special, name)  # numpy_compute.py:234-238
    return fn


def bop_types(name, dt_a, dt_b):
    """(loop dtype, output dtype) NumPy's type resolution picks for ufunc(name) on these inputs."""
    key = (name, dt_a, dt_b)
    hit = _BOP_TYPES.get(key)
    if hit is None:
        in_a, _in_b, out = _ufunc(name).resolve_dtypes((dt_a, dt_b, None))
        hit = (np.dtype(in_a), np.dtype(out))
        _BOP_TYPES[key] = hit
    return hit


def uop_types(name, dt):
    key = (name, dt)
    hit = _UOP_TYPES.get(key)
    if hit is None:
        in_a, out = _ufunc(name).resolve_dtypes((dt, None))
        hit = (np.dtype(in_a), np.dtype(out))
        _UOP_TYPES[key] = hit
    return hit


def _apply_sel(t, sel):
    """Basic (slice / int) indexing as a view; mirrors ndarray[sel] for the selections
    ArrayView generates (view.py:170-178, :358-365)."""
    if sel is None:
        return t
    if not isinstance(sel, tuple):
        sel = (sel,)
    for s in sel:
        if isinstance(s, slice) and s.step is not None and s.step < 0:
            raise NotImplementedError("negative-step slices are not supported by cuda_compute")
    return t[sel]


class Touched(object):
    """Result of ``touch``: resolves to a bool once the stream has been synchronised (by ``get``)."""
    __slots__ = ("ok",)

    def __init__(self, ok):
        self.ok = ok


class ComputeCls(_ComputeImp):
    # ------------------------------------------------------------------ I/O-ish
    def touch(self, arr):
        """"Touch" a block (numpy_compute.py:88-89).  Returns a token; waiting happens when the token
        is fetched with ``system.get`` -- BlockArray.touch collects one token per block and gets them
        all at once (blockarray.py:117-126), which costs one stream synchronisation, not one per block."""
        return Touched(isinstance(arr, torch.Tensor))

    def empty(self, grid_entry, grid_meta):
        grid = ArrayGrid.from_meta(grid_meta)
        return _empty(grid.get_block_shape(grid_entry), grid.dtype)

    def new_block(self, op_name, grid_entry, grid_meta):
        grid = ArrayGrid.from_meta(grid_meta)
        shape = grid.get_block_shape(grid_entry)
        out = _empty(shape, grid.dtype)
        if op_name == "eye":
            assert np.all(np.diff(grid_entry) == 0)
            if out.numel():
                LIB.check(LIB.dll.nums_eye(describe(out), _stream()))
        elif op_name in ("zeros", "ones"):
            if out.numel():
                LIB.check(LIB.dll.nums_fill(describe(out), 1.0 if op_name == "ones" else 0.0, _stream()))
        elif op_name != "empty":
            raise NotImplementedError("new_block(%s)" % op_name)
        return out

    def random_block(self, rng_params, rfunc_name, rfunc_args, shape, dtype):
        rng = block_rng(*rng_params)
        result = getattr(rng, rfunc_name)(*rfunc_args).reshape(shape)
        if rfunc_name not in ("random", "integers"):
            result = result.astype(dtype)
        return upload(result)

    def permutation(self, rng_params, size):
        return upload(block_rng(*rng_params).permutation(size))

    def diag(self, arr):
        arr = upload(arr)
        if arr.dim() == 1:
            n = arr.shape[0]
            out = _empty((n, n), arr.dtype)
            if n:
                LIB.check(LIB.dll.nums_fill(describe(out), 0.0, _stream()))
                _copy_into(out.as_strided((n,), (n + 1,)), arr)
            return out
        if arr.dim() == 2:
            k = min(arr.shape)
            src = arr.as_strided((k,), (arr.stride(0) + arr.stride(1),), arr.storage_offset())
            return _materialize(src)
        raise ValueError("Input must be 1- or 2-d.")

    def arange(self, start, stop, step, dtype):
        ref = np.arange(0, 1, 1, dtype)  # dtype resolution as NumPy does it
        length = int(max(0, np.ceil((stop - start) / step)))
        out = _empty((length,), ref.dtype)
        if length:
            LIB.check(LIB.dll.nums_arange(describe(out), float(start), float(step), _stream()))
        return out

    # ------------------------------------------------------------------ data movement
    def create_block(self, *src_arrs, src_params, dst_params, dst_shape, dst_shape_bc):
        first = upload(src_arrs[0])
        result = _empty(dst_shape, first.dtype)
        assert len(src_params) == len(dst_params)
        target = result if dst_shape_bc is None else result.view(tuple(dst_shape_bc))
        for i in range(len(src_params)):
            src = upload(src_arrs[i])
            src_sel, src_t = src_params[i]
            if src_t:
                src = _transpose_view(src)
            dst_sel, _dst_t = dst_params[i]
            _copy_into(_apply_sel(target, dst_sel), _apply_sel(src, src_sel))
        return result

    def update_block(self, dst_arr, *src_arrs, src_params, dst_params):
        assert len(src_params) == len(dst_params)
        dst = _materialize(upload(dst_arr))  # inputs are immutable (numpy_compute.py:136-138)
        if dst_params[0][1]:
            dst = _transpose_view(dst)
        for i in range(len(src_params)):
            src = upload(src_arrs[i])
            src_sel, src_shape_bc, src_t = src_params[i]
            if src_t:
                src = _transpose_view(src)
            if src_shape_bc is not None:
                src = _reshape(src, src_shape_bc)
            dst_sel, _t = dst_params[i]
            _copy_into(_apply_sel(dst, dst_sel), _apply_sel(src, src_sel))
        return dst

    def update_block_by_index(self, dst_arr, src_arr, index_pairs):
        result = _materialize(upload(dst_arr))
        src = upload(src_arr)
        pairs = list(index_pairs)
        if not pairs:
            return result
        dst_lin = _ravel_indices([p[0] for p in pairs], tuple(result.shape))
        src_lin = _ravel_indices([p[1] for p in pairs], tuple(src.shape))
        _scatter(result, src, dst_lin, src_lin, 1, result.numel(), src.numel(), 1)
        return result

    def update_block_along_axis(self, dst_arr, src_arr, index_pairs, axis):
        result = _materialize(upload(dst_arr))
        src = upload(src_arr)
        pairs = list(index_pairs)
        if not pairs:
            return result
        axis = int(axis) % result.dim()
        other = tuple(result.shape[:axis]) + tuple(result.shape[axis + 1:])
        if src.dim() != result.dim() or other != tuple(src.shape[:axis]) + tuple(src.shape[axis + 1:]):
            # shapes that only agree through broadcasting: pair by pair with the broadcasting copy kernel
            for dst_index, src_index in pairs:
                _copy_into(result.select(axis, int(dst_index)), src.select(axis, int(src_index)))
            return result
        dst_idx = _ravel_indices([(p[0],) for p in pairs], (result.shape[axis],))
        src_idx = _ravel_indices([(p[1],) for p in pairs], (src.shape[axis],))
        outer = int(np.prod(result.shape[:axis], dtype=np.int64))
        inner = int(np.prod(result.shape[axis + 1:], dtype=np.int64))
        _scatter(result, src, dst_idx, src_idx, outer, result.shape[axis], src.shape[axis], inner)
        return result

    def transpose(self, arr):
        return _transpose_view(upload(arr))

    def reshape(self, arr, shape):
        return _reshape(upload(arr), shape if isinstance(shape, (tuple, list)) else (shape,))

    def split(self, arr, indices_or_sections, axis, transposed):
        arr = upload(arr)
        if transposed:
            arr = _transpose_view(arr)
        n = arr.shape[axis]
        if isinstance(indices_or_sections, (int, np.integer)):
            if n % indices_or_sections:
                raise ValueError("array split does not result in an equal division")
            step = n // int(indices_or_sections)
            bounds = [(i * step, (i + 1) * step) for i in range(int(indices_or_sections))]
        else:
            cuts = [0] + [int(i) for i in indices_or_sections] + [n]
            bounds = [(min(cuts[i], n), min(max(cuts[i + 1], cuts[i]), n)) for i in range(len(cuts) - 1)]
        return [_carry_tag(arr.narrow(axis, lo, max(hi - lo, 0)), arr) for lo, hi in bounds]

    def astype(self, arr, dtype_str):
        arr = upload(arr)
        return _materialize(arr, _lib.torch_dtype(_np_dtype(dtype_str)))

    # ------------------------------------------------------------------ elementwise
    def map_uop(self, op_name, arr, args, kwargs):
        if args or kwargs:
            raise NotImplementedError("map_uop(%s) with extra arguments" % op_name)
        arr = upload(arr)
        if op_name not in _lib.UOP_CODE:
            raise NotImplementedError("unary ufunc %s" % op_name)
        loop, out_dt = uop_types(op_name, _lib.numpy_dtype(arr.dtype))
        out = _empty(arr.shape, out_dt)
        if out.numel():
            LIB.check(LIB.dll.nums_uop(_lib.UOP_CODE[op_name], _lib.dtype_code(loop), describe(arr),
                                       describe(out), _stream()))
        return out

    def xlogy(self, arr_x, arr_y):
        return elementwise("xlogy", upload(arr_x), upload(arr_y))

    def bop(self, op, a1, a2, a1_shape, a2_shape, a1_T, a2_T, axes):
        a1 = _operand(a1, a1_shape, a1_T)
        a2 = _operand(a2, a2_shape, a2_T)
        if op == "tensordot":
            return tensordot(a1, a2, axes)
        return elementwise(_SHORT_OP_NAMES.get(op, op), a1, a2)

    # ------------------------------------------------------------------ reductions
    def reduce_axis(self, op_name, arr, axis, keepdims, transposed):
        arr = upload(arr)
        if op_name not in _lib.REDUCE_CODE:
            raise NotImplementedError("reduce_axis(%s)" % op_name)
        nd = arr.dim()
        if axis is not None:
            axis = int(axis)
            if axis < 0:
                axis += nd
        # (arr.T).op(axis=a) == (arr.op(axis=nd-1-a)).T: reduce the stored array, transpose the view
        flip = bool(transposed) and nd > 1
        if flip and axis is not None:
            axis = nd - 1 - axis
        if not arr.is_contiguous():
            arr = _materialize(arr)
        in_dt = _lib.numpy_dtype(arr.dtype)
        out_dt = reduce_type(op_name, in_dt)
        shape = tuple(arr.shape)
        if axis is None:
            outer, red, inner = 1, int(np.prod(shape, dtype=np.int64)), 1
            out_shape = (1,) * nd if keepdims else ()
        else:
            outer = int(np.prod(shape[:axis], dtype=np.int64))
            red = shape[axis]
            inner = int(np.prod(shape[axis + 1:], dtype=np.int64))
            out_shape = shape[:axis] + ((1,) if keepdims else ()) + shape[axis + 1:]
        out = _empty(out_shape, out_dt)
        if out.numel():
            if red == 0:
                if op_name not in ("sum", "prod", "product", "any", "all"):
                    raise ValueError("zero-size array to reduction operation %s which has no identity" % op_name)
                LIB.check(LIB.dll.nums_fill(describe(out), 1.0 if op_name in ("prod", "product", "all") else 0.0,
                                            _stream()))
            else:
                LIB.call_ws(LIB.dll.nums_reduce, arr.device,
                            ((_lib.REDUCE_CODE[op_name], arr.data_ptr(), _lib.dtype_code(arr.dtype), outer, red,
                              inner, out.data_ptr(), _lib.dtype_code(out.dtype)), (_stream(),)))
        return _transpose_view(out) if flip else out

    def sum_reduce(self, *arrs):
        arrs = [upload(a) for a in arrs]
        first = arrs[0]
        dt = np.result_type(*[_lib.numpy_dtype(a.dtype) for a in arrs])
        tdt = _lib.torch_dtype(dt)
        prepared = []
        for a in arrs:
            if tuple(a.shape) != tuple(first.shape):
                raise ValueError("sum_reduce needs same-shape blocks")
            if a.dtype != tdt or not a.is_contiguous():
                a = _materialize(a, tdt)
            prepared.append(a)
        out = _empty(first.shape, dt)
        if out.numel():
            ptrs = (_lib.ctypes.c_void_p * len(prepared))(*[a.data_ptr() for a in prepared])
            LIB.check(LIB.dll.nums_sum_reduce(len(prepared), ptrs, _lib.dtype_code(tdt), out.numel(),
                                              out.data_ptr(), _stream()))
        return out

    def arg_op(self, op_name, arr, block_slice, other_argoptima=None, other_optima=None):
        if op_name not in ("argmin", "argmax"):
            raise Exception("Unsupported arg op.")
        arr = upload(arr)
        if arr.dim() != 1:
            arr = _reshape(arr, (arr.numel(),))
        if not arr.is_contiguous():
            arr = _materialize(arr)
        out_index = _empty((), np.int64)
        out_value = _empty((), _lib.numpy_dtype(arr.dtype))
        ci = cv = None
        if other_optima is not None:
            ci = upload(other_argoptima, np.int64)
            if ci.dtype != torch.int64:
                ci = _materialize(ci, torch.int64)
            cv = upload(other_optima)
            if cv.dtype != arr.dtype:
                cv = _materialize(cv, arr.dtype)
        LIB.call_ws(LIB.dll.nums_arg_op, arr.device,
                    ((1 if op_name == "argmax" else 0, arr.data_ptr(), _lib.dtype_code(arr.dtype), arr.numel(),
                      int(block_slice.start or 0), ci.data_ptr() if ci is not None else None,
                      cv.data_ptr() if cv is not None else None, out_index.data_ptr(), out_value.data_ptr()),
                     (_stream(),)))
        return out_index, out_value

    def where(self, arr, x, y, block_slice_tuples):
        if x is not None or y is not None:
            raise NotImplementedError("three-argument where is not reachable from nums.numpy.where "
                                      "(nums/numpy/api.py:379-381) and is not implemented")
        arr = upload(arr)
        if not arr.is_contiguous():
            arr = _materialize(arr)
        shape = tuple(arr.shape) if arr.dim() else (1,)
        count = _empty((), np.int64)
        LIB.call_ws(LIB.dll.nums_nonzero_count, arr.device,
                    ((arr.data_ptr(), _lib.dtype_code(arr.dtype), arr.numel(), count.data_ptr()), (_stream(),)))
        n = int(count.cpu().item())  # one 8-byte read-back: the caller fetches the shape anyway (application.py:588)
        outs = [_empty((n,), np.int64) for _ in shape]
        if n:
            c = _lib.ctypes
            shape_arr = (c.c_int64 * len(shape))(*shape)
            offs = (c.c_int64 * len(shape))(*[int(s[0]) for s in block_slice_tuples][:len(shape)])
            ptrs = (c.c_void_p * len(shape))(*[o.data_ptr() for o in outs])
            LIB.call_ws(LIB.dll.nums_nonzero_fill, arr.device,
                        ((arr.data_ptr(), _lib.dtype_code(arr.dtype), len(shape), shape_arr, offs, ptrs),
                         (_stream(),)))
        return tuple(outs + [(n,)])

    def allclose(self, a, b, rtol, atol):
        a, b = upload(a), upload(b)
        dt = np.result_type(_lib.numpy_dtype(a.dtype), _lib.numpy_dtype(b.dtype))
        tdt = _lib.torch_dtype(dt)
        shape = _broadcast_shape(tuple(a.shape), tuple(b.shape))
        a = _materialize(a.expand(shape), tdt) if (a.dtype != tdt or tuple(a.shape) != tuple(shape) or not a.is_contiguous()) else a
        b = _materialize(b.expand(shape), tdt) if (b.dtype != tdt or tuple(b.shape) != tuple(shape) or not b.is_contiguous()) else b
        flag = _empty((), np.bool_)
        LIB.call_ws(LIB.dll.nums_allclose, a.device,
                    ((a.data_ptr(), b.data_ptr(), _lib.dtype_code(tdt), a.numel(), float(rtol), float(atol),
                      flag.data_ptr()), (_stream(),)))
        return flag

    def logical_and(self, *bool_list):
        flags = _empty((len(bool_list),), np.bool_)
        for i, b in enumerate(bool_list):
            _copy_into(flags[i], upload(b, np.bool_) if not isinstance(b, torch.Tensor) else b)
        out = _empty((), np.bool_)
        LIB.call_ws(LIB.dll.nums_reduce, flags.device,
                    ((_lib.REDUCE_CODE["all"], flags.data_ptr(), _lib.BOOL, 1, len(bool_list), 1, out.data_ptr(),
                      _lib.BOOL), (_stream(),)))
        return out

    # ------------------------------------------------------------------ dense linear algebra
    def qr(self, *arrays, mode="reduced", axis=None):
        if len(arrays) > 1:
            assert axis is not None
            if mode == "r" and int(axis) == 0 and all(a.__class__ is DeferredR and a.value is None for a in arrays) \
                    and len({a.shape for a in arrays}) == 1:
                fused = _fused_stacked_r(list(arrays))        # the stacked-R call of indirect_tsr (application.py:807-814)
                if fused is not None:
                    return fused
            arr = concatenate([a.materialize() if a.__class__ is DeferredR else upload(a) for a in arrays], axis)
        elif arrays[0].__class__ is DeferredR:
            arr = arrays[0].materialize()                     # R of a triangular factor: itself
            if mode == "r":
                return arr
        else:
            arr = upload(arrays[0])
        if mode == "r":
            if QR_DEFER_ENABLED and len(arrays) == 1 and arr.dim() == 2 and _gram_path_ok(arr):
                return DeferredR(_gram_of(arr), arr)
            return qr_r(arr)
        if mode == "reduced":
            return qr_reduced(arr)
        raise NotImplementedError("qr mode %r" % (mode,))

    def cholesky(self, arr):
        return _single_cta_factor(LIB.dll.nums_cholesky, upload(arr), "Matrix is not positive definite")

    def inv(self, arr):
        known = getattr(arr, "_nums_inverse_t", None) if arr.__class__ is torch.Tensor else None
        if known is not None:             # the R factor of a Gram-path qr: its inverse is (L^-1)^T, already computed
            return _materialize(_transpose_view(known))
        return _single_cta_factor(LIB.dll.nums_inv, upload(arr), "Singular matrix")

    def svd(self, arr):
        """u, sigma, vT of a square block (numpy_compute.py:251-254; only ever the n x n R factor,
        application.py:946).  One-sided Jacobi; singular vectors are unique up to signs only."""
        arr = upload(arr)
        if arr.dim() != 2 or arr.shape[0] != arr.shape[1]:
            raise NotImplementedError("svd of non-square blocks is not on the hot path")
        if arr.dtype not in (torch.float64, torch.float32):
            arr = _materialize(arr, torch.float64)
        if not arr.is_contiguous():
            arr = _materialize(arr)
        n = arr.shape[0]
        dt = _lib.numpy_dtype(arr.dtype)
        u, sigma, vt = _empty((n, n), dt), _empty((n,), dt), _empty((n, n), dt)
        LIB.call_ws(LIB.dll.nums_svd, arr.device,
                    ((_lib.dtype_code(arr.dtype), n, arr.data_ptr(), n, u.data_ptr(), sigma.data_ptr(), vt.data_ptr()),
                     (_stream(),)))
        return u, sigma, vt


# ---------------------------------------------------------------------------------------------
# helpers (module level: ComputeCls itself must expose nothing but the 28 interface methods,
# the reference's SerialSystem.init looks every method up in ComputeInterface, systems.py:76-89)
# ---------------------------------------------------------------------------------------------
_BOP_PLANS = {}      # (ufunc name, torch dtype a, torch dtype b) -> (op code, loop dtype code, output torch dtype)


def elementwise(name, a1, a2):
    """np.<name>(a1, a2) with NumPy type resolution and broadcasting (nums_bop)."""
    key = (name, a1.dtype, a2.dtype)
    plan = _BOP_PLANS.get(key)
    if plan is None:
        if name not in _lib.BOP_CODE:
            raise NotImplementedError("binary ufunc %s" % name)
        loop, out_dt = bop_types(name, _lib.numpy_dtype(a1.dtype), _lib.numpy_dtype(a2.dtype))
        plan = _BOP_PLANS[key] = (_lib.BOP_CODE[name], _lib.dtype_code(loop), _lib.torch_dtype(out_dt),
                                  _lib.dtype_code(a1.dtype), _lib.dtype_code(a2.dtype), _lib.dtype_code(out_dt),
                                  name == "power" and np.dtype(loop).kind in "iu")
    if plan[6] and a2.numel():
        _check_integer_exponent(a2)
    s1, s2 = a1.shape, a2.shape
    shape = s1 if s1 == s2 else _broadcast_shape(tuple(s1), tuple(s2))
    out = torch.empty(shape, dtype=plan[2], device=a1.device)
    n = out.numel()
    if n:
        n1, n2 = a1.numel(), a2.numel()
        if (n1 == n or n1 == 1) and (n2 == n or n2 == 1) and a1.is_contiguous() and a2.is_contiguous():
            # dense over the output (broadcasting stretched nothing) or a single value: no descriptors needed
            rc = LIB.dll.nums_bop_flat(plan[0], plan[1], a1.data_ptr(), plan[3], n1, a2.data_ptr(), plan[4], n2,
                                       out.data_ptr(), plan[5], n, _stream())
        else:
            rc = LIB.dll.nums_bop(plan[0], plan[1], describe(a1), describe(a2), describe(out), _stream())
        if rc:
            LIB.check(rc)
    return out


def _check_integer_exponent(exponent):
    """np.power on integers raises ValueError("Integers to negative integer powers are not allowed.") when any
    exponent is negative -- a property of the DATA.  The minimum of the exponent block is reduced on the device and
    looked at with the other deferred status words (check_pending_status) at the next host synchronisation."""
    if exponent.dtype == torch.bool:
        return
    src = exponent if exponent.is_contiguous() else _materialize(exponent)
    low = _empty((), _lib.numpy_dtype(src.dtype))
    code = _lib.dtype_code(src.dtype)
    LIB.call_ws(LIB.dll.nums_reduce, src.device,
                ((_lib.REDUCE_CODE["min"], src.data_ptr(), code, 1, src.numel(), 1, low.data_ptr(), code), (_stream(),)))
    _PENDING_STATUS.append((low, _negative, ValueError("Integers to negative integer powers are not allowed.")))
    if len(_PENDING_STATUS) > _PENDING_STATUS_MAX:
        check_pending_status()


def _as_matrix(t, rows, cols):
    """View a tensor as a (rows, cols) matrix stored row-major or column-major without copying
    when possible.  Returns (tensor, transposed_flag, pitch)."""
    if t.is_contiguous():
        return t, False, max(cols, 1)
    tt = _transpose_view(t)
    if tt.is_contiguous():  # fully reversed axes of a dense array: a (cols, rows) row-major matrix
        return tt, True, max(rows, 1)
    m = _materialize(t)
    return m, False, max(cols, 1)


def gemm_into(out, a, ta, lda, b, tb, ldb, m, n, k, accumulate=False):
    LIB.call_ws(LIB.dll.nums_gemm, out.device,
                ((_lib.dtype_code(out.dtype), int(ta), int(tb), m, n, k, a.data_ptr(), lda, b.data_ptr(), ldb,
                  out.data_ptr(), max(n, 1), int(accumulate)), (_stream(),)))


def tensordot(a1, a2, axes):
    """np.tensordot(a1, a2, axes=k) for integer k (numpy_compute.py:231-232, blockarray.py:410-414)."""
    if not isinstance(axes, (int, np.integer)):
        raise NotImplementedError("tensordot with explicit axis lists")
    axes = int(axes)
    dt = np.result_type(_lib.numpy_dtype(a1.dtype), _lib.numpy_dtype(a2.dtype))
    if dt == np.bool_:
        raise NotImplementedError("tensordot on bool blocks")
    tdt = _lib.torch_dtype(dt)
    if a1.dtype != tdt:
        a1 = _materialize(a1, tdt)
    if a2.dtype != tdt:
        a2 = _materialize(a2, tdt)
    free1 = tuple(a1.shape[:a1.dim() - axes])
    con1 = tuple(a1.shape[a1.dim() - axes:])
    con2 = tuple(a2.shape[:axes])
    free2 = tuple(a2.shape[axes:])
    if con1 != con2:
        raise ValueError("shape-mismatch for sum")
    m = int(np.prod(free1, dtype=np.int64))
    k = int(np.prod(con1, dtype=np.int64))
    n = int(np.prod(free2, dtype=np.int64))
    out = _empty(free1 + free2, dt)
    if out.numel() == 0:
        return out
    # N-D operands: a fully reversed (lazy .T) view of a dense array is a transposed matrix only
    # in the 2-D case; otherwise materialise.
    if a1.dim() > 2 and not a1.is_contiguous():
        a1 = _materialize(a1)
    if a2.dim() > 2 and not a2.is_contiguous():
        a2 = _materialize(a2)
    A, ta, lda = _as_matrix(a1, m, k)
    B, tb, ldb = _as_matrix(a2, k, n)
    gemm_into(out, A, ta, lda, B, tb, ldb, m, n, k)
    return out


def concatenate(arrs, axis):
    axis = int(axis)
    shape = list(arrs[0].shape)
    shape[axis] = sum(a.shape[axis] for a in arrs)
    out = _empty(shape, _lib.numpy_dtype(arrs[0].dtype))
    pos = 0
    for a in arrs:
        _copy_into(out.narrow(axis, pos, a.shape[axis]), a)
        pos += a.shape[axis]
    return out


# Tall-skinny fast path of qr_r: a block is factored through its Gram matrix (all DMMA GEMM work)
# when a rigorous bound on its condition number says that is as accurate as Householder.
QR_GRAM_MIN_ASPECT = 8        # m >= 8 n
QR_GRAM_MAX_COLS = 256
QR_GRAM_ACCEPT_KAPPA = 30.0   # one Cholesky pass: error ~ kappa^2 eps  (< 1e-12)
QR_GRAM_REFINE_KAPPA = 1.0e6  # two passes (CholeskyQR2) are as good as Householder below ~1e7
QR_SHIFT_PLAIN_KAPPA = 1.0e7  # inside the shifted iteration a pass without shift is taken up to this bound
QR_SHIFT_MAX_SHIFTS = 3       # shifted passes before the Householder kernel takes over
QR_STATS = {"gram": 0, "gram2": 0, "gram3": 0, "householder": 0, "gram_fused": 0}
# NUMS_QR_SHIFTED=0: ill-conditioned tall blocks go straight to the Householder kernel
QR_SHIFTED_ENABLED = __import__("os").environ.get("NUMS_QR_SHIFTED", "1") != "0"
# NUMS_QR_DEFER=0: qr(block, mode="r") factors at once instead of handing out a DeferredR
QR_DEFER_ENABLED = __import__("os").environ.get("NUMS_QR_DEFER", "1") != "0"
# NUMS_QR_GRAM=0 sends every block through the Householder kernel (measurements, paranoia)
QR_GRAM_ENABLED = __import__("os").environ.get("NUMS_QR_GRAM", "1") != "0"


def _householder_r(arr):
    m, n = arr.shape
    r = _empty((min(m, n), n), _lib.numpy_dtype(arr.dtype))
    LIB.call_ws(LIB.dll.nums_qr, arr.device,
                ((_lib.dtype_code(arr.dtype), m, n, arr.data_ptr(), n, None, 0, r.data_ptr(), n), (_stream(),)))
    QR_STATS["householder"] += 1
    return r


def sum_of_squares(t):
    """sum(t * t) as a 0-d device tensor (multiply + two-stage sum, both ours)."""
    if not t.is_contiguous():
        t = _materialize(t)
    n = t.numel()
    flat = t.view(n)
    sq = _empty((n,), _lib.numpy_dtype(t.dtype))
    code = _lib.dtype_code(t.dtype)
    LIB.check(LIB.dll.nums_bop(_lib.BOP_CODE["multiply"], code, describe(flat), describe(flat), describe(sq), _stream()))
    out = _empty((), _lib.numpy_dtype(t.dtype))
    LIB.call_ws(LIB.dll.nums_reduce, t.device,
                ((_lib.REDUCE_CODE["sum"], sq.data_ptr(), code, 1, n, 1, out.data_ptr(), code), (_stream(),)))
    return out


def _norm_1_inf(t):
    """(max column abs-sum, max row abs-sum) of a square matrix, as two 0-d device tensors."""
    n = t.shape[0]
    absolute = _empty((n, n), np.float64)
    LIB.check(LIB.dll.nums_uop(_lib.UOP_CODE["absolute"], _lib.F64, describe(t), describe(absolute), _stream()))
    out = []
    for outer, red, inner in ((1, n, n), (n, n, 1)):      # axis 0 sums (columns), axis 1 sums (rows)
        sums = _empty((n,), np.float64)
        LIB.call_ws(LIB.dll.nums_reduce, t.device,
                    ((_lib.REDUCE_CODE["sum"], absolute.data_ptr(), _lib.F64, outer, red, inner, sums.data_ptr(),
                      _lib.F64), (_stream(),)))
        top = _empty((), np.float64)
        LIB.call_ws(LIB.dll.nums_reduce, t.device,
                    ((_lib.REDUCE_CODE["max"], sums.data_ptr(), _lib.F64, 1, n, 1, top.data_ptr(), _lib.F64),
                     (_stream(),)))
        out.append(top)
    return out


GRAM_FACTOR_FUSED_MAX = 128   # nums_gram_factor: Cholesky + triangular inverse + norms in one launch


def _gram_factor(a):
    """Cholesky factor L (lower) of a^T a, its inverse, R = L^T and a rigorous bound on cond_2(a):
    cond_2(a) = cond_2(L) <= sqrt(|L|_1 |L|_inf |L^-1|_1 |L^-1|_inf)   (|M|_2^2 <= |M|_1 |M|_inf),
    which is tight for the nearly diagonal factors of well-conditioned blocks (a Frobenius bound
    would be off by a factor n).  Returns (L, Linv, R, kappa_bound); the bound is inf if the Gram
    matrix is not numerically positive definite.  One 40-byte read-back."""
    return _factor_gram(_gram_of(a))


def _gram_of(a):
    """a^T a of a tall float64 block on the DMMA GEMM (split-K)."""
    m, n = a.shape
    gram = _empty((n, n), np.float64)
    gemm_into(gram, a, True, n, a, False, n, n, n, m)
    return gram


def _factor_gram(gram):
    """(L, L^-1, R = L^T, bound on cond_2) from a Gram matrix; see _gram_factor."""
    n = gram.shape[0]
    a = gram
    low = _empty((n, n), np.float64)
    stats = _empty((5,), np.float64)
    if n <= GRAM_FACTOR_FUSED_MAX:
        low_inv = _empty((n, n), np.float64)
        upper = _empty((n, n), np.float64)
        LIB.check(LIB.dll.nums_gram_factor(n, gram.data_ptr(), n, low.data_ptr(), n, upper.data_ptr(), n,
                                           low_inv.data_ptr(), n, stats.data_ptr(), _stream()))
    else:
        info = _empty((), np.int32)
        LIB.call_ws(LIB.dll.nums_cholesky, a.device,
                    ((_lib.F64, n, gram.data_ptr(), n, low.data_ptr(), n, info.data_ptr()), (_stream(),)))
        low_inv = _inv_nocheck(low)
        upper = _materialize(_transpose_view(low))
        _copy_into(stats[0], info)
        for i, v in enumerate(_norm_1_inf(low) + _norm_1_inf(low_inv)):
            _copy_into(stats[1 + i], v)
    failed, l1, linf, i1, iinf = (float(v) for v in stats.cpu())   # 40-byte D2H, the one sync of this path
    bound = l1 * linf * i1 * iinf
    if failed != 0 or not np.isfinite(bound):
        return low, low_inv, upper, float("inf")
    # R = L^T came with its inverse for free ((L^-1)^T): remember it on the tensor object, so that the inv(R) that
    # indirect_tsqr asks for next (application.py:833, one 0.6 ms single-CTA Gauss-Jordan on every rank) is a transpose
    upper._nums_inverse_t = low_inv
    return low, low_inv, upper, float(np.sqrt(bound))


def _gram_path_ok(arr):
    m, n = arr.shape
    return (QR_GRAM_ENABLED and arr.dtype == torch.float64 and arr.is_contiguous() and n % 2 == 0
            and 2 <= n <= QR_GRAM_MAX_COLS and m >= QR_GRAM_MIN_ASPECT * n and m >= 1024 and arr.data_ptr() % 16 == 0)


class DeferredR(object):
    """The R factor of a tall block that so far exists only as the block's Gram matrix.

    ``indirect_tsr`` (application.py:784-814) asks for one ``qr(block, mode="r")`` per row block and then for the
    ``qr`` of the stacked R's.  On the Gram path both stages are Cholesky factorizations -- of G_i = X_i^T X_i and of
    sum_i R_i^T R_i = sum_i G_i -- so the per-block factorizations (a small-matrix kernel and a 40-byte read-back
    each) are pure overhead when the stacked call follows: ``qr`` hands out this handle instead, the stacked call
    adds the Gram matrices and factors ONCE, with the condition bound checked on the result.  Anything else that
    touches the handle (``get``, any other kernel) materialises R_i exactly as before (``qr_r_ex``)."""
    __slots__ = ("gram", "source", "value", "shape", "__weakref__")
    dtype = torch.float64

    def __init__(self, gram, source):
        self.gram, self.source, self.value = gram, source, None
        n = int(gram.shape[0])
        self.shape = (n, n)

    def materialize(self):
        if self.value is None:
            self.value = qr_r_ex(self.source, self.gram)[0]
            self.gram = self.source = None
        return self.value


def _fused_stacked_r(parts):
    """R of the stacked blocks behind ``parts`` (all DeferredR of one width) from the SUM of their Gram matrices;
    None if the condition bound of the result is not small enough for a single Cholesky pass."""
    n = parts[0].shape[0]
    grams = [p.gram for p in parts]
    if len(grams) == 1:
        total = grams[0]
    else:
        total = _empty((n, n), np.float64)
        ptrs = (_lib.ctypes.c_void_p * len(grams))(*[g.data_ptr() for g in grams])
        LIB.check(LIB.dll.nums_sum_reduce(len(grams), ptrs, _lib.F64, n * n, total.data_ptr(), _stream()))
    _low, _low_inv, upper, kappa = _factor_gram(total)
    if kappa <= QR_GRAM_ACCEPT_KAPPA:
        QR_STATS["gram_fused"] += 1
        return upper
    return None


def qr_r_ex(arr, gram=None):
    """(R, kappa_bound).  R is the k x n (k = min(m, n)) triangular factor of a 2-D block.

    Tall float64 blocks go through the Gram matrix: R = chol(A^T A)^T when the condition bound is
    tiny, CholeskyQR2 when it is moderate, iterated shifted Cholesky passes when it is large or the Gram
    matrix is not numerically positive definite (_shifted_cholesky_r); everything else (wide, short, f32,
    exact zeros, more than 128 columns) uses the blocked Householder TSQR kernel (nums_qr), which is
    backward stable for any input."""
    if arr.dim() != 2:
        raise ValueError("qr needs a 2-D block")
    if arr.dtype not in (torch.float64, torch.float32):
        arr = _materialize(arr, torch.float64)
    if not arr.is_contiguous():
        arr = _materialize(arr)
    m, n = arr.shape
    if arr.dtype == torch.float32 and gram is None and QR_GRAM_ENABLED and n % 2 == 0 and 2 <= n <= QR_GRAM_MAX_COLS \
            and m >= QR_GRAM_MIN_ASPECT * n and m >= 1024:
        # tall float32 block: factor a float64 copy on the Gram path (one extra pass over the block instead of the
        # Householder kernel's 128 dependent reflector steps per chunk) and round the result; the arithmetic is
        # more accurate than the float32 LAPACK routine the reference calls
        r64, kappa = qr_r_ex(_materialize(arr, torch.float64))
        return _materialize(r64, torch.float32), kappa
    if gram is None and not _gram_path_ok(arr):
        return _householder_r(arr), None
    g0 = gram if gram is not None else _gram_of(arr)
    low, low_inv, upper, kappa = _factor_gram(g0)
    if kappa <= QR_GRAM_ACCEPT_KAPPA:
        QR_STATS["gram"] += 1
        return upper, kappa
    if kappa <= QR_GRAM_REFINE_KAPPA:
        # CholeskyQR2: Q1 = A L^-T, second Gram factor, R = (L1 L2)^T
        q1 = _empty((m, n), np.float64)
        gemm_into(q1, arr, False, n, _materialize(_transpose_view(low_inv)), False, n, m, n, n)
        low2, _low2_inv, _upper2, kappa2 = _gram_factor(q1)
        if kappa2 <= QR_GRAM_ACCEPT_KAPPA:
            prod = _empty((n, n), np.float64)
            gemm_into(prod, low, False, n, low2, False, n, n, n, n)
            QR_STATS["gram2"] += 1
            return _materialize(_transpose_view(prod)), kappa
    if QR_SHIFTED_ENABLED and n <= GRAM_FACTOR_FUSED_MAX:
        r = _shifted_cholesky_r(arr, g0)
        if r is not None:
            QR_STATS["gram3"] += 1
            return r, kappa
    return _householder_r(arr), kappa


def _shifted_cholesky_r(arr, gram):
    """R of an ill-conditioned tall float64 block by repeated (shifted) Cholesky passes on the streaming kernels
    -- shifted CholeskyQR3 (Fukaya, Kannan, Nakatsukasa, Yamamoto, Yanagisawa, SIAM J. Sci. Comput. 2020),
    iterated: with Q_0 = A and R = I, every pass factors the Gram matrix G of the current Q_i; while the bound on
    cond_2(Q_i) is above QR_SHIFT_PLAIN_KAPPA (or G is not numerically positive definite) the factor is taken from
    G + s I, s a small multiple of u trace(G) (see below), which is positive definite and lowers the condition number
    by ~sqrt(s / |Q_i|^2) ~ 1e5 per pass; then Q_{i+1} = Q_i R_i^-1, R <- R_i R, until one plain
    pass with a bound <= QR_GRAM_ACCEPT_KAPPA closes the iteration.  The result satisfies R^T R = A^T A and agrees
    with LAPACK's Householder R to ~1e-14 for condition numbers up to 1e16 (oracle/shifted_cholqr_study.py); every
    pass is one streaming SYRK, one 128-column triangular solve as a GEMM, one small factorization and one 40-byte
    read-back.  Returns None when QR_SHIFT_MAX_SHIFTS shifted passes were not enough (exact zeros: the Householder
    kernel takes over)."""
    m, n = arr.shape
    u = 2.0 ** -53
    shift_rel = 11.0 * (float(m) * n + n * (n + 1.0)) * u
    cur, r_acc, shifts = arr, None, 0
    for _ in range(QR_SHIFT_MAX_SHIFTS + 3):
        g = gram if gram is not None else _gram_of(cur)
        gram = None
        low, low_inv, upper, kappa = _factor_gram(g)
        if kappa <= QR_GRAM_ACCEPT_KAPPA:
            return upper if r_acc is None else _small_matmul(upper, r_acc)
        if kappa > QR_SHIFT_PLAIN_KAPPA:
            if shifts >= QR_SHIFT_MAX_SHIFTS:
                return None
            shifts += 1
            dg = _materialize(g.as_strided((n,), (n + 1,)))             # trace(G): diagonal + sum, one 8-byte read-back
            tr = _empty((), np.float64)
            LIB.call_ws(LIB.dll.nums_reduce, g.device,
                        ((_lib.REDUCE_CODE["sum"], dg.data_ptr(), _lib.F64, 1, n, 1, tr.data_ptr(), _lib.F64), (_stream(),)))
            trace = float(tr.cpu().item())
            if not np.isfinite(trace) or trace <= 0.0:
                return None
            # The theoretical shift 11 (m n + n (n + 1)) u |A|^2 guarantees that Cholesky runs to completion whatever
            # the rounding errors of the Gram matrix were, but it grows with m and lowers the condition number by
            # only sqrt(1 / shift) per pass (1.7e3 at m = 2 M).  Start at 11 n (n + 1) u trace(G) -- enough in every
            # case of the study -- and escalate by factors of 100 up to the theoretical value when the factorization
            # still breaks down (a retry costs one 0.3 ms small-matrix kernel, not a pass over the block).
            kappa_s = float("inf")
            c = min(11.0 * n * (n + 1.0) * u, shift_rel)
            while True:
                shift = _empty((n, n), np.float64)
                LIB.check(LIB.dll.nums_fill(describe(shift), 0.0, _stream()))
                LIB.check(LIB.dll.nums_fill(describe(shift.as_strided((n,), (n + 1,))), c * trace, _stream()))
                low, low_inv, upper, kappa_s = _factor_gram(elementwise("add", g, shift))
                if np.isfinite(kappa_s) or c >= shift_rel:
                    break
                c = min(c * 100.0, shift_rel)
            if not np.isfinite(kappa_s):
                return None
        r_inv = _materialize(_transpose_view(low_inv))                  # R_i^-1 = (L^-1)^T, dense for the streaming GEMM
        nxt = _empty((m, n), np.float64)
        gemm_into(nxt, cur, False, n, r_inv, False, n, m, n, n)
        cur = nxt
        r_acc = upper if r_acc is None else _small_matmul(upper, r_acc)
    return None


def _small_matmul(a, b):
    n = a.shape[0]
    out = _empty((n, b.shape[1]), np.float64)
    gemm_into(out, a, False, a.shape[1], b, False, b.shape[1], n, b.shape[1], a.shape[1])
    return out


def qr_r(arr):
    return qr_r_ex(arr)[0]


def _inv_nocheck(a):
    n = a.shape[0]
    out = _empty((n, n), _lib.numpy_dtype(a.dtype))
    LIB.call_ws(LIB.dll.nums_inv, a.device,
                ((_lib.dtype_code(a.dtype), n, a.data_ptr(), n, out.data_ptr(), n, None), (_stream(),)))
    return out


def qr_reduced(arr):
    """(Q, R) with Q (m x k) orthonormal.  Q is formed as A R^-1 (what the reference itself does one
    level up, application.py:833-845); unless R came with a tiny condition bound, one
    re-orthogonalisation pass (R2 = qr_r(Q); Q <- Q R2^-1; R <- R2 R) restores orthogonality to
    working precision for any block whose condition number is below ~1/sqrt(eps).
    """
    if arr.dtype not in (torch.float64, torch.float32):
        arr = _materialize(arr, torch.float64)
    if not arr.is_contiguous():
        arr = _materialize(arr)
    m, n = arr.shape
    k = min(m, n)
    r, kappa = qr_r_ex(arr)
    lead = arr if k == n else _materialize(arr[:, :k])
    r_sq = r if k == n else _materialize(r[:, :k])
    dt = _lib.numpy_dtype(arr.dtype)
    q = _empty((m, k), dt)
    gemm_into(q, lead, False, k, _inv_nocheck(r_sq), False, k, m, k, k)
    if kappa is not None and kappa <= QR_GRAM_ACCEPT_KAPPA:
        return q, r
    r2 = qr_r(q)
    q2 = _empty((m, k), dt)
    gemm_into(q2, q, False, k, _inv_nocheck(r2), False, k, m, k, k)
    r_out = _empty((k, n), dt)
    gemm_into(r_out, r2, False, k, r, False, n, k, n, k)
    return q2, r_out


def _single_cta_factor(fn, arr, message):
    if arr.dim() != 2 or arr.shape[0] != arr.shape[1]:
        raise np.linalg.LinAlgError("Last 2 dimensions of the array must be square")
    if arr.dtype not in (torch.float64, torch.float32):
        arr = _materialize(arr, torch.float64)
    if not arr.is_contiguous():
        arr = _materialize(arr)
    n = arr.shape[0]
    out = _empty((n, n), _lib.numpy_dtype(arr.dtype))
    mode = CHECK_FACTORIZATION_STATUS
    if mode is True:
        mode = "eager"
    info = _empty((), np.int32) if mode and mode != "off" else None
    LIB.call_ws(fn, arr.device,
                ((_lib.dtype_code(arr.dtype), n, arr.data_ptr(), n, out.data_ptr(), n,
                  info.data_ptr() if info is not None else None), (_stream(),)))
    if info is not None:
        if mode == "eager":
            if int(info.cpu().item()) != 0:
                raise np.linalg.LinAlgError(message)
        else:
            _PENDING_STATUS.append((info, _nonzero, np.linalg.LinAlgError(message)))
            if len(_PENDING_STATUS) > _PENDING_STATUS_MAX:
                check_pending_status()
    return out


def _lr_operands(X, y, beta):
    """The fused kernels read raw float64 pointers: X (n, d) with unit column stride, y (n,) and beta (d,)
    dense.  Anything else is converted (y of another dtype, e.g. bool / int labels) or rejected."""
    if X.dim() != 2 or X.dtype != torch.float64 or X.stride(1) != 1:
        raise ValueError("lr_grad_hess: X must be a 2-D float64 block with unit column stride")
    if y.dtype != torch.float64 or not y.is_contiguous():
        y = _materialize(y, torch.float64)
    if beta.dtype != torch.float64 or not beta.is_contiguous():
        beta = _materialize(beta, torch.float64)
    if y.numel() != X.shape[0] or beta.numel() != X.shape[1]:
        raise ValueError("lr_grad_hess: shapes X %s, y %s, beta %s do not agree"
                         % (tuple(X.shape), tuple(y.shape), tuple(beta.shape)))
    return X, y, beta


def lr_grad_hess(X, y, beta):
    """Fused g = X^T (mu - y), H = X^T diag(mu (1 - mu)) X for one row block (nums_lr_grad_hess).

    Returns a 1-D tensor of d + d*d doubles (g followed by row-major H)."""
    X, y, beta = _lr_operands(X, y, beta)
    n, d = X.shape
    out = _empty((d + d * d,), np.float64)
    LIB.call_ws(LIB.dll.nums_lr_grad_hess, X.device,
                ((n, d, X.data_ptr(), X.stride(0), y.data_ptr(), beta.data_ptr(), out.data_ptr()), (_stream(),)))
    return out


def lr_grad_hess_blocks(x_blocks, y_blocks, beta):
    """g | H summed over a list of row blocks.  Dense blocks of a supported width go through
    nums_lr_grad_hess_blocks (16 blocks per launch); anything else block by block."""
    checked = [_lr_operands(x, y, beta) for x, y in zip(x_blocks, y_blocks)]
    x_blocks, y_blocks, beta = [c[0] for c in checked], [c[1] for c in checked], checked[0][2]
    d = x_blocks[0].shape[1]
    dense = ((d % 16 in (4, 12)) and d <= 48
             and all(x.is_contiguous() and x.dtype == torch.float64 and x.data_ptr() % 16 == 0 for x in x_blocks))
    total = None
    if dense:
        ct = _lib.ctypes
        for lo in range(0, len(x_blocks), 16):
            xs, ys = x_blocks[lo:lo + 16], y_blocks[lo:lo + 16]
            out = _empty((d + d * d,), np.float64)
            xp = (ct.c_void_p * len(xs))(*[x.data_ptr() for x in xs])
            yp = (ct.c_void_p * len(xs))(*[y.data_ptr() for y in ys])
            rows = (ct.c_int64 * len(xs))(*[x.shape[0] for x in xs])
            LIB.call_ws(LIB.dll.nums_lr_grad_hess_blocks, xs[0].device,
                        ((len(xs), xp, yp, rows, d, beta.data_ptr(), out.data_ptr()), (_stream(),)))
            total = out if total is None else elementwise("add", total, out)
        return total
    for x, y in zip(x_blocks, y_blocks):
        part = lr_grad_hess(x, y, beta)
        total = part if total is None else elementwise("add", total, part)
    return total


def newton_step(gh, beta):
    """One Newton update from the summed ``g | H`` buffer (glms.py:362-372): returns
    ``(beta - inv(H) g, status)`` where ``status`` is a 2-element device tensor {max |g|, info}.  The
    caller reads ``status`` back once per iteration (convergence test + singularity check)."""
    d = beta.shape[0]
    if gh.shape[0] != d + d * d or gh.dtype != torch.float64 or beta.dtype != torch.float64:
        raise ValueError("newton_step: expected float64 g | H of %d entries" % (d + d * d))
    gh = gh if gh.is_contiguous() else _materialize(gh)
    beta = beta if beta.is_contiguous() else _materialize(beta)
    out = _empty((d,), np.float64)
    status = _empty((2,), np.float64)
    LIB.check(LIB.dll.nums_newton_step(d, gh.data_ptr(), beta.data_ptr(), out.data_ptr(), status.data_ptr(), _stream()))
    return out, status


# ---------------------------------------------------------------------------------------------
# delimited text ingest
# ---------------------------------------------------------------------------------------------
_CSV_STATUS = {1: "could not convert field to %s", 2: "field is valid for the reference but not supported by cuda_compute "
               "(hexadecimal float, lone carriage return, or a literal of more than 19 digits that cannot be decided) for %s",
               3: "rows have different numbers of fields (%s block)"}


def _csv_dtype(dtype):
    """The reference's converter choice (filesystem.py:169-190) -> (output dtype, kernel dtype code)."""
    if dtype is float:
        return np.dtype(np.float64)
    if dtype is int or dtype is bool:
        # int -> int(float(x)) into int64, bool -> Python bool(): neither is what the typed kernels implement
        raise NotImplementedError("read_csv_block: pass a NumPy dtype (float64, float32, int64, int32, bool_)")
    dt = np.dtype(dtype)
    if dt not in (np.dtype(np.float64), np.dtype(np.float32), np.dtype(np.int64), np.dtype(np.int32), np.dtype(np.bool_)):
        raise NotImplementedError("read_csv_block: dtype %s" % dt)
    return dt


def _find_byte(view, value, start):
    """Index of the first `value` in the uint8 array `view` at or after `start`, or -1."""
    step = 1 << 16
    n = view.shape[0]
    while start < n:
        hits = np.flatnonzero(view[start:start + step] == value)
        if hits.size:
            return start + int(hits[0])
        start += step
    return -1


_READ_CHUNK = 8 << 20
_read_pool = None


def _read_range(fd, offset, target, filename):
    """Fill the writable buffer `target` with the bytes of file `fd` starting at `offset`.  Large ranges
    are read by a few threads with positional reads (the page-cache copy runs outside the GIL): a
    single `readinto` of page-locked memory was the largest part of an end-to-end CSV ingest."""
    import os
    total = len(target)

    def fill(lo, hi):
        pos = lo
        while pos < hi:
            got = os.preadv(fd, [target[pos:hi]], offset + pos)
            if got <= 0:
                raise IOError("short read from %s" % filename)
            pos += got

    if total <= _READ_CHUNK:
        fill(0, total)
        return
    global _read_pool
    if _read_pool is None:
        from concurrent.futures import ThreadPoolExecutor
        _read_pool = ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1), thread_name_prefix="nums-csv-read")
    jobs = [_read_pool.submit(fill, lo, min(total, lo + _READ_CHUNK)) for lo in range(0, total, _READ_CHUNK)]
    for job in jobs:
        job.result()


def read_csv_block(filename, file_start, file_end, dtype, delimiter, has_header):
    """``read_csv_block`` of the reference (filesystem.py:157-212) with the parsing on the device.

    The chunk's lines are chosen exactly as the reference's text-mode loop does: if ``file_start``
    is not 0, everything up to and including the first newline at or after it is skipped (:198-201);
    then every line that *starts* before ``file_end`` is taken (:203-204), the header being dropped in
    the first chunk (:205-207).  The bytes of those lines are read straight into page-locked memory,
    uploaded, and split / converted by ``nums_csv_index`` + ``nums_csv_parse``.  Returns
    ``(block, shape)`` like the reference (:211-212); malformed input raises ``ValueError``.
    """
    import os
    dt = _csv_dtype(dtype)
    if not isinstance(delimiter, str) or len(delimiter.encode()) != 1:
        raise NotImplementedError("read_csv_block: single-byte delimiters only")
    delim = delimiter.encode()[0]
    size = os.path.getsize(filename)
    file_start, file_end = int(file_start), min(int(file_end), size)
    empty = (upload(np.array([], dtype=dt)), (0,))
    if file_start >= file_end:
        return empty
    slack = 1 << 16
    with open(filename, "rb") as fh:
        while True:
            want = min(size, file_end + slack) - file_start
            pinned = torch.empty(((want + 63) // 32 * 32,), dtype=torch.uint8, pin_memory=True)
            view = pinned.numpy()
            _read_range(fh.fileno(), file_start, memoryview(view)[:want], filename)
            view[want:] = 0
            # the last line of the chunk is the one holding byte file_end - 1; it ends at its newline
            last_nl = _find_byte(view[:want], 10, file_end - 1 - file_start)
            if last_nl >= 0 or file_start + want == size:
                break
            slack *= 8      # a very long line: read further
    stop = last_nl + 1 if last_nl >= 0 else want
    first = 0
    if file_start != 0:
        nl = _find_byte(view[:stop], 10, 0)
        if nl < 0:
            return empty
        first = nl + 1
    if has_header and file_start == 0:
        nl = _find_byte(view[:stop], 10, 0)
        first = nl + 1 if nl >= 0 else stop
    if first >= stop:
        return empty
    line_end = _find_byte(view[:stop], 10, first)
    line_end = stop if line_end < 0 else line_end
    cols = int(np.count_nonzero(view[first:line_end] == delim)) + 1

    text = _upload_tensor(pinned)
    try:
        return parse_csv_text(text, first, stop, delim, dt, cols)
    except (ValueError, NotImplementedError) as exc:
        where = getattr(exc, "offset", None)
        at = "" if where is None else " at byte %d of %s" % (file_start + where, filename)
        raise type(exc)(str(exc) + at) from None


def parse_csv_text(text, first, stop, delimiter, dtype, cols):
    """Bytes [first, stop) of the device uint8 tensor ``text`` (whole lines) -> ``(block, shape)``.
    ``text`` must be readable up to ``stop`` rounded up to a multiple of 32."""
    dt = np.dtype(dtype)
    summary = _empty((4,), np.int64)
    ws = LIB.workspace(text.device, 2 * ((stop - first) // 8192 + 4) * 8)
    LIB.check(LIB.dll.nums_csv_index(text.data_ptr(), first, stop, delimiter, summary.data_ptr(), ws.data_ptr(),
                                     ws.numel(), _stream()))
    rows, fields, status, where = (int(v) for v in summary.cpu())      # 32-byte read-back: the block's shape
    if status == 0 and fields != rows * cols:
        status = 3
    out = _empty((rows, cols), dt)
    if status == 0:
        LIB.check(LIB.dll.nums_csv_parse(text.data_ptr(), first, stop, delimiter, _lib.dtype_code(dt), rows, cols,
                                         out.data_ptr(), summary.data_ptr(), ws.data_ptr(), _stream()))
        _rows, _fields, status, where = (int(v) for v in summary.cpu())
    if status != 0:
        exc = (NotImplementedError if status == 2 else ValueError)("read_csv_block: " + _CSV_STATUS[status] % dt)
        exc.offset = None if where >= (1 << 62) else where
        raise exc
    return out, (rows, cols)


# ---------------------------------------------------------------------------------------------
# block persistence (the reference's on-disk array format, used as checkpoint format)
# ---------------------------------------------------------------------------------------------
def _block_path(filename, grid_entry):
    """<filename>/<i>_<j>...pkl -- the reference's layout (filesystem.py:63,111-129)."""
    import os
    os.makedirs(filename, exist_ok=True)
    return os.path.join(filename, "_".join(str(int(i)) for i in grid_entry) + ".pkl")


def write_block_fs(block, filename, grid_entry):
    """``write_block_fs`` (filesystem.py:111-119): the block is downloaded and pickled as the same NumPy
    array the reference would have written, so either side can read the other's files."""
    import pickle
    arr = download(upload(block))
    with open(_block_path(filename, grid_entry), "wb") as fh:
        pickle.dump(np.ascontiguousarray(arr), fh)
    return None


def read_block_fs(filename, grid_entry):
    """``read_block_fs`` (filesystem.py:121-129): unpickle and upload."""
    import pickle
    with open(_block_path(filename, grid_entry), "rb") as fh:
        return upload(np.asarray(pickle.load(fh)))


def delete_block_fs(filename, grid_entry):
    """``delete_block_fs`` (filesystem.py:131-139)."""
    import os
    os.remove(_block_path(filename, grid_entry))
    return None


# ---------------------------------------------------------------------------------------------
# optional kernels beyond ComputeInterface: offered to the host layers as
# ``system.lr_grad_hess`` / ``system.newton_step`` (CudaSystem registers EXTRA_KERNELS at init) and used
# by nums_b200.glms_fused when present.  Same calling convention as the 28 interface methods: blocks in,
# blocks out, ``syskwargs`` stripped by the system.
# ---------------------------------------------------------------------------------------------
def lr_grad_hess_block(X, y, beta):
    """g | H of one (n_b, d) row block: what glms.forward / gradient / hessian (glms.py:140-143,213-240)
    compute with ~12 kernel calls and six passes over X, in ONE pass (nums_lr_grad_hess_blocks when the
    block is dense and d = 4 or 12 mod 16, nums_lr_grad_hess otherwise)."""
    X, y, beta = upload(X), upload(y), upload(beta)
    if y.dim() != 1:
        y = _reshape(y, (y.numel(),))
    return lr_grad_hess_blocks([X], [y], beta)


def lr_grad_hess_multi(*blocks):
    """g | H summed over several row blocks that live on one device: ``(X_0, y_0, X_1, y_1, ..., beta)``.  One
    nums_lr_grad_hess_blocks launch per 16 blocks instead of one launch (and one partial to sum) per block."""
    if len(blocks) < 3 or len(blocks) % 2 == 0:
        raise ValueError("lr_grad_hess_multi: expected X_0, y_0, ..., X_k, y_k, beta")
    beta = upload(blocks[-1])
    xs = [upload(b) for b in blocks[0:-1:2]]
    ys = []
    for yb in blocks[1:-1:2]:
        yb = upload(yb)
        ys.append(yb if yb.dim() == 1 else _reshape(yb, (yb.numel(),)))
    return lr_grad_hess_blocks(xs, ys, beta)


def newton_step_block(gh, beta):
    """(beta - inv(H) g, status = {max |g|, info}) from the summed g | H buffer (glms.py:368-370)."""
    return newton_step(upload(gh), upload(beta))


EXTRA_KERNELS = {
    "lr_grad_hess": lr_grad_hess_block,
    "lr_grad_hess_multi": lr_grad_hess_multi,
    "newton_step": newton_step_block,
}


def loadtxt_block(fname, dtype, comments, delimiter, converters, skiprows, usecols, unpack, ndmin, encoding,
                  max_rows):
    """``loadtxt_block`` (filesystem.py:144-155): NumPy's text reader on the host (it is a general
    tokenizer with comments / converters / usecols, i.e. I/O, not arithmetic), result uploaded."""
    return upload(np.loadtxt(fname, dtype=dtype, comments=comments, delimiter=delimiter, converters=converters,
                             skiprows=skiprows, usecols=usecols, unpack=unpack, ndmin=ndmin, encoding=encoding,
                             max_rows=max_rows))


# block-level I/O functions the reference's FileSystem registers (filesystem.py:224-231) for which this
# module has device-aware versions; CudaSystem.register substitutes them by name.
DEVICE_FUNCTIONS = {
    "write_block_fs": write_block_fs,
    "read_block_fs": read_block_fs,
    "delete_block_fs": delete_block_fs,
    "read_csv_block": read_csv_block,
    "loadtxt_block": loadtxt_block,
}
