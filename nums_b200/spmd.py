"""``SpmdSystem`` -- the block grid partitioned over the GPUs of one box, behind the unchanged kernel interface.

The reference scales by handing every per-block kernel call to Ray together with
``syskwargs={"grid_entry", "grid_shape"}`` and letting the scheduler pick a node
(/root/reference/nums/core/systems/schedulers.py:170-246); operands travel through the object store.
Here one process drives one GPU (``torchrun``) and EVERY rank runs the same, unmodified host program
(``BlockArray`` / ``ArrayApplication`` / ``glms.newton``): each kernel call reaches this system on all ranks
with identical arguments, the ranks agree -- without talking -- on who executes it, and only that rank
launches the kernel.  What the host layers hold as ``oid`` is a ``Handle``: the same object on every rank
(id, home rank, shape, dtype), with the device tensor attached only where the block lives.

* **Placement** (SURVEY.md 8e): a call runs where most of its operand bytes already are ("compute where the
  operands live", the fork's rule, gpu_systems.py:156-161); ties and calls without device-resident operands
  go to ``owner(grid_entry, grid_shape)`` -- row-major flattened entry ``mod world`` for grids with one long
  axis (gpu_systems.py:163-164), ``(i mod pr, j mod pc)`` on the ``pr x pc`` device grid for true 2-D grids
  (BlockCyclicScheduler.get_cluster_entry, schedulers.py:170-191).  Blocks made by ``put`` and results of
  single-block arrays (scalars, beta, the d x d Hessian, R) are *replicated*: every rank holds them and
  computes on them redundantly, which costs nothing and removes all small-message traffic.
* **Operand movement**: a remote operand is sent point to point (NCCL over NVLink) to the executing rank and
  cached there (blocks are immutable, numpy_compute.py:136-138).
* **Reductions across ranks** (the reference gathers to one node: ``sum_reduce`` in ``_matvec``,
  blockarray.py:574-578; the ``add`` chains of ``_tensordot`` :468-471 and of ``reduce_axis`` :402-407; the
  stacked ``qr(*R_oids)``, application.py:807-814): ``sum_reduce`` over blocks living on several ranks is a
  local partial sum + ONE all-reduce; ``tensordot`` / ``add`` chains are recorded lazily (``_Lazy``) and
  materialised together -- operands that must move are exchanged in one batched transfer, every rank
  contracts its own terms in one grouped DMMA launch (CudaSystem's deferred contractions), and results whose
  terms live on several ranks are all-reduced; the stacked-R ``qr`` is a local QR + binary tree over the
  ranks + broadcast.  Results of cross-rank reductions are replicated.
* **Shuffles / permutations** (SURVEY.md 8f.4): ``BlockArray._advanced_single_array_subscript``
  (blockarray.py:229-316) builds every destination block as a CHAIN of ``update_block_along_axis(dst, src_j,
  index_pairs_j, axis)`` calls, one per source block.  Executed call by call that ships every source block WHOLE
  to every destination owner (G x the array over the links) and copies the destination G times.  The chains are
  recorded instead (``_LazyScatter``) and materialised together: every source owner gathers just the rows a
  destination asks for into a packed buffer (``nums_scatter_axis`` as a gather), ONE batched all-to-all moves
  the packed rows (each row crosses the links at most once), and every destination owner copies its base block
  once and scatters all its pieces into it.
* ``get`` broadcasts from the home rank, so every rank's host program sees the same values (it may branch
  on them: ``if max(abs(g)) <= tol``, glms.py:370).

Shapes / dtypes of results must be known on ranks that do not execute a call (receive buffers, later
placement decisions): they are inferred for the kernels on the hot path (``infer_result``) and, for the rest
(dynamic sizes such as ``where``, registered I/O functions), broadcast as a small pickled description from
the executing rank over a gloo side channel.  ``NUMS_SPMD_CHECK=1`` asserts inference == actual.

The class is backend-agnostic: ``local`` is a ``CudaSystem`` in production and the oracle system (NumPy
blocks, gloo) in the CPU tests (tests/test_spmd_cpu.py), which is how the host logic is covered without a GPU.
"""
import os
import weakref

import numpy as np

REPLICATED = -1
_F64 = np.dtype(np.float64)
_NOT_HANDLED = object()
_BOP_PARAMS = ("op", "a1", "a2", "a1_shape", "a2_shape", "a1_T", "a2_T", "axes")
LAZY_MIN_EXTENT = 64            # same eligibility as CudaSystem's deferred contractions (deferred.py)
PROMOTE_BYTES = 1 << 20         # fetched blocks up to this size stay on every rank
COALESCE_BYTES = 1 << 20        # all-reduces of partials up to this size share one buffer


class Handle(object):
    """One block of the distributed array store; identical metadata on every rank."""
    __slots__ = ("hid", "home", "value", "shape", "dtype", "copies", "lazy", "__weakref__")

    def __init__(self, hid, home, value, shape, dtype):
        self.hid = hid
        self.home = home                  # owning rank, or REPLICATED
        self.value = value                # local block where available, else None
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.copies = ()                  # ranks (besides home) that hold a cached copy
        self.lazy = None                  # _Lazy while the value is still an unevaluated sum

    @classmethod
    def _make(cls, hid, home, value, shape, dtype):
        """Constructor for an already normalised ``shape`` (tuple of int) / ``dtype`` (np.dtype)."""
        h = cls.__new__(cls)
        h.hid, h.home, h.value, h.shape, h.dtype, h.copies, h.lazy = hid, home, value, shape, dtype, (), None
        return h

    def available_on(self, rank):
        return self.home == REPLICATED or self.home == rank or rank in self.copies

    @property
    def nbytes(self):
        return int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize

    def __repr__(self):
        return "Handle(#%d home=%s %s %s%s)" % (self.hid, "*" if self.home == REPLICATED else self.home, self.shape,
                                                self.dtype, " lazy" if self.lazy is not None else "")


class _Lazy(object):
    """value = sum of ``items`` in order; an item is ("dot", a1, a2, a1_shape, a2_shape, a1_T, a2_T, rank) --
    np.tensordot(a1, a2, 1) evaluated on ``rank`` -- or ("blk", handle, rank)."""
    __slots__ = ("items", "hint")

    def __init__(self, items, hint):
        self.items = items
        self.hint = hint

    def ranks(self):
        return sorted({it[-1] for it in self.items})


class _LazyScatter(object):
    """value = ``base`` with, in order, ``dst[.., dst_idx, ..] = src[.., src_idx, ..]`` along ``axis`` for every
    item (src handle, dst_idx, src_idx) -- the chain of update_block_along_axis calls that builds one block of a
    shuffled array; evaluated on rank ``rank``."""
    __slots__ = ("base", "items", "axis", "rank")

    def __init__(self, base, items, axis, rank):
        self.base, self.items, self.axis, self.rank = base, items, axis, rank


# -----------------------------------------------------------------------------------------------------
# local back ends: what SpmdSystem needs besides the kernel interface
# -----------------------------------------------------------------------------------------------------
class _TorchBackend(object):
    def __init__(self, system):
        self.system = system

    def is_block(self, v):
        import torch
        from nums_b200.cuda_compute import DeferredR
        from nums_b200.deferred import DeferredContraction
        return isinstance(v, (torch.Tensor, DeferredContraction, DeferredR))

    def concrete(self, v):
        """A dense tensor holding the block (launches whatever the local system still defers)."""
        from nums_b200 import cuda_compute
        v = self.system.contractions.resolve(v)
        return v if v.is_contiguous() else cuda_compute._materialize(v)

    def settle(self, v):
        """The tensor behind a (by now launched) deferred contraction; other values unchanged."""
        return self.system.contractions.resolve(v)

    def meta(self, v):
        from nums_b200 import _lib
        return tuple(int(s) for s in v.shape), _lib.numpy_dtype(v.dtype)

    def empty(self, shape, dtype):
        from nums_b200 import cuda_compute
        return cuda_compute._empty(shape, dtype)

    def zeros(self, shape, dtype):
        from nums_b200 import cuda_compute
        from nums_b200._lib import LIB, describe
        out = cuda_compute._empty(shape, dtype)
        if out.numel():
            LIB.check(LIB.dll.nums_fill(describe(out), 0.0, cuda_compute._stream()))
        return out

    def clone(self, v):
        from nums_b200 import cuda_compute
        return cuda_compute._materialize(self.system.contractions.resolve(v))

    def copy_into(self, dst, src):
        from nums_b200 import cuda_compute
        cuda_compute._copy_into(dst, src)

    def flat_view(self, buf, offset, shape):
        n = int(np.prod(shape, dtype=np.int64))
        return buf[offset:offset + n].view(tuple(shape))

    def take_along(self, v, index, axis):
        """v[.., index, ..] along ``axis`` as a dense block (one nums_scatter_axis launch used as a gather)."""
        import numpy as np_
        from nums_b200 import cuda_compute as cc
        v = self.system.contractions.resolve(v)
        shape = list(v.shape)
        outer = int(np_.prod(shape[:axis], dtype=np_.int64))
        inner = int(np_.prod(shape[axis + 1:], dtype=np_.int64))
        n = int(len(index))
        out = cc._empty(shape[:axis] + [n] + shape[axis + 1:], cc._lib.numpy_dtype(v.dtype))
        if n and out.numel():
            cc._scatter(out, v, np_.arange(n, dtype=np_.int64), index, outer, n, shape[axis], inner)
        return out

    def scatter_chain(self, base, pieces, axis):
        """A private copy of ``base`` with every (src, dst_index, src_index) of ``pieces`` applied in order."""
        import numpy as np_
        from nums_b200 import cuda_compute as cc
        out = cc._materialize(self.system.contractions.resolve(base))
        shape = tuple(out.shape)
        outer = int(np_.prod(shape[:axis], dtype=np_.int64))
        inner = int(np_.prod(shape[axis + 1:], dtype=np_.int64))
        for src, dst_index, src_index in pieces:
            src = self.system.contractions.resolve(src)
            if len(dst_index) and out.numel():
                cc._scatter(out, src, dst_index, src_index, outer, shape[axis], src.shape[axis], inner)
        return out

    def flush(self):
        self.system.flush()

    def synchronize(self):
        self.system.synchronize()


class _NumpyBackend(object):
    def __init__(self, system):
        self.system = system

    def is_block(self, v):
        return isinstance(v, (np.ndarray, np.generic))

    def concrete(self, v):
        return np.ascontiguousarray(v)

    def settle(self, v):
        return v

    def meta(self, v):
        v = np.asarray(v)
        return tuple(v.shape), v.dtype

    def empty(self, shape, dtype):
        return np.empty(shape, dtype=dtype)

    def zeros(self, shape, dtype):
        return np.zeros(shape, dtype=dtype)

    def clone(self, v):
        return np.array(v, copy=True)

    def copy_into(self, dst, src):
        dst[...] = src

    def flat_view(self, buf, offset, shape):
        n = int(np.prod(shape, dtype=np.int64))
        return buf[offset:offset + n].reshape(tuple(shape))

    def take_along(self, v, index, axis):
        return np.ascontiguousarray(np.take(np.asarray(v), np.asarray(index, dtype=np.int64), axis=axis))

    def scatter_chain(self, base, pieces, axis):
        out = np.array(base, copy=True)
        for src, dst_index, src_index in pieces:
            dst_sel = [slice(None)] * out.ndim
            dst_sel[axis] = np.asarray(dst_index, dtype=np.int64)
            out[tuple(dst_sel)] = np.take(np.asarray(src), np.asarray(src_index, dtype=np.int64), axis=axis)
        return out

    def flush(self):
        pass

    def synchronize(self):
        pass


# -----------------------------------------------------------------------------------------------------
# result shapes / dtypes without executing (hot-path kernels only; None = ask the executing rank)
# -----------------------------------------------------------------------------------------------------
def _broadcast(s1, s2):
    return tuple(int(x) for x in np.broadcast_shapes(tuple(s1), tuple(s2)))


def _float_like(dt):
    return np.dtype(np.float32) if np.dtype(dt) == np.float32 else np.dtype(np.float64)


def infer_result(name, args, kwargs, meta_of):
    """Description of what kernel ``name`` returns -- ("b", shape, dtype str) per block, ("t", [...]) for
    tuples -- from the arguments alone, or None when it cannot be told cheaply.  ``meta_of(x)`` gives
    (shape, dtype) of a block argument."""
    from nums_b200 import cuda_compute as cc
    from nums_b200.grid import ArrayGrid

    def blk(shape, dtype):
        return ("b", tuple(int(s) for s in shape), np.dtype(dtype).str)

    try:
        if name == "bop":
            b = dict(zip(_BOP_PARAMS, args))
            b.update(kwargs)
            op = b["op"]
            (_s1, d1), (_s2, d2) = meta_of(b["a1"]), meta_of(b["a2"])
            s1, s2 = tuple(b["a1_shape"]), tuple(b["a2_shape"])
            if op == "tensordot":
                axes = int(b["axes"])
                dt = np.result_type(d1, d2)
                if dt == np.bool_:
                    return None
                return blk(s1[:len(s1) - axes] + s2[axes:], dt)
            ufunc = cc._SHORT_OP_NAMES.get(op, op)
            return blk(_broadcast(s1, s2), cc.bop_types(ufunc, np.dtype(d1), np.dtype(d2))[1])
        if name == "map_uop":
            op_name, arr = args[0], args[1]
            if (len(args) > 2 and args[2]) or (len(args) > 3 and args[3]):
                return None
            shape, dt = meta_of(arr)
            return blk(shape, cc.uop_types(op_name, np.dtype(dt))[1])
        if name == "reduce_axis":
            b = dict(zip(("op_name", "arr", "axis", "keepdims", "transposed"), args))
            b.update(kwargs)
            shape, dt = meta_of(b["arr"])
            if b["transposed"]:
                shape = tuple(reversed(shape))
            nd, axis = len(shape), b["axis"]
            if axis is None:
                out = (1,) * nd if b["keepdims"] else ()
            else:
                axis = int(axis) % nd
                out = shape[:axis] + ((1,) if b["keepdims"] else ()) + shape[axis + 1:]
            return blk(out, cc.reduce_type(b["op_name"], np.dtype(dt)))
        if name == "sum_reduce":
            metas = [meta_of(a) for a in args]
            return blk(metas[0][0], np.result_type(*[m[1] for m in metas]))
        if name in ("new_block", "empty"):
            entry, grid_meta = args[-2], args[-1]
            grid = ArrayGrid.from_meta(grid_meta)
            return blk(grid.get_block_shape(entry), np.dtype(grid.dtype))
        if name == "transpose":
            shape, dt = meta_of(args[0])
            return blk(tuple(reversed(shape)), dt)
        if name == "reshape":
            shape = args[1] if len(args) > 1 else kwargs["shape"]
            shape = tuple(shape) if isinstance(shape, (tuple, list)) else (shape,)
            if any(int(s) < 0 for s in shape):
                return None
            return blk(shape, meta_of(args[0])[1])
        if name == "astype":
            return blk(meta_of(args[0])[0], cc._np_dtype(args[1] if len(args) > 1 else kwargs["dtype_str"]))
        if name in ("inv", "cholesky"):
            shape, dt = meta_of(args[0])
            return blk(shape, _float_like(dt))
        if name == "xlogy":
            (s1, _d1), (s2, _d2) = meta_of(args[0]), meta_of(args[1])
            return blk(_broadcast(s1, s2), np.float64)
        if name == "qr":
            metas = [meta_of(a) for a in args]
            mode, axis = kwargs.get("mode", "reduced"), kwargs.get("axis")
            if any(len(m[0]) != 2 for m in metas):
                return None
            if len(metas) > 1:
                axis = int(axis)
                rows = sum(m[0][0] for m in metas) if axis == 0 else metas[0][0][0]
                cols = metas[0][0][1] if axis == 0 else sum(m[0][1] for m in metas)
            else:
                rows, cols = metas[0][0]
            dt = _float_like(np.result_type(*[m[1] for m in metas]))
            k = min(rows, cols)
            if mode == "r":
                return blk((k, cols), dt)
            if mode == "reduced":
                return ("t", [blk((rows, k), dt), blk((k, cols), dt)])
            return None
        if name == "create_block":
            return blk(kwargs["dst_shape"], meta_of(args[0])[1])
        if name in ("update_block_by_index", "update_block_along_axis"):
            shape, dt = meta_of(args[0])
            return blk(shape, dt)
        if name == "lr_grad_hess":
            d = int(meta_of(args[2])[0][0])
            return blk((d + d * d,), np.float64)
        if name == "lr_grad_hess_multi":
            d = int(meta_of(args[-1])[0][0])
            return blk((d + d * d,), np.float64)
        if name == "newton_step":
            d = int(meta_of(args[1])[0][0])
            return ("t", [blk((d,), np.float64), blk((2,), np.float64)])
    except Exception:  # noqa: BLE001 -- anything unusual: let the executing rank describe the result
        return None
    return None


class PeerExchange(object):
    """Bulk operand movement over NVLink peer memory, driven by the copy engines (GPU back end only).

    NCCL point-to-point transfers occupy SMs next to the DMMA kernel and run as one un-overlapped phase; here
    the exchange of a flush is a set of *pulls* instead: every rank stages the float64 blocks it has to serve
    into a symmetric-memory arena (``torch.distributed._symmetric_memory``: the same allocation on every rank,
    mapped into every peer's address space; the staging copy is device-to-device and costs ~10 us per 32 MB
    block), two device-side barriers bracket the staging (nobody still reads the previous contents / every
    arena is filled), and each destination then copies its operands straight out of the owners' arenas on the
    upload stream.  A pulled block looks exactly like an asynchronously uploaded one (``_nums_ready`` tag,
    cuda_compute._Transfers), so CudaSystem's deferred-contraction flush starts the grouped GEMM of the first
    result blocks as soon as THEIR operands have landed and overlaps the rest of the exchange with it.
    No SM is taken from the GEMM and no rank waits for another beyond the two barriers."""

    MIN_BYTES = 1 << 20

    def __init__(self, comm):
        self.comm = comm
        self.capacity = 0          # float64 elements
        self.buf = None
        self.handle = None
        self.failed = False
        self.pulled = False        # set by run(): pulls of this exchange may still be in flight
        self.stream = None         # high-priority stream for barriers + staging (so they do not queue behind a GEMM)
        self.slots = {}            # hid -> (weakref to the handle, offset in its home rank's arena): what is staged where
        self.fill = {}             # rank -> elements used in that rank's arena

    def usable(self, moves):
        if self.failed or not moves:
            return False
        f64 = np.dtype(np.float64)
        return all(h.dtype == f64 for h, _dst in moves) and sum(h.nbytes for h, _dst in moves) >= self.MIN_BYTES

    def _ensure(self, elements):
        """Collective: every rank calls with the same ``elements``."""
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        if elements <= self.capacity:
            return True
        ok = 1
        if self.buf is not None:
            # growing replaces the arena: nobody may still be pulling from the old one when it is released
            torch.cuda.synchronize()
            dist.barrier()
        try:
            cap = max(int(elements * 1.25), 1 << 20)
            buf = symm.empty(cap, dtype=torch.float64, device=torch.device("cuda", torch.cuda.current_device()))
            handle = symm.rendezvous(buf, dist.group.WORLD)
        except Exception as exc:  # noqa: BLE001 -- any failure selects the NCCL transfers
            import sys
            sys.stderr.write("[nums_b200] symmetric-memory arena unavailable (%s: %s); using NCCL point-to-point\n"
                             % (type(exc).__name__, exc))
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)          # all ranks take the same path
        if int(flag.item()) != 1:
            self.failed = True
            return False
        self.buf, self.handle, self.capacity = buf, handle, cap
        return True

    def run(self, moves, rank, backend):
        """``moves`` = [(handle, dst)] (identical on all ranks, none available on its dst yet).  Returns False
        if the exchange has to go through NCCL instead.

        Staged blocks STAY in the arena while their handle lives (blocks are immutable), so serving the same
        operand again -- every product of a loop over the same matrices -- costs no staging copy and no
        barrier: the destinations simply pull.  The arena is an append-only log per rank, compacted (reset +
        re-stage, bracketed by the "nobody is still pulling" barrier) when it fills up.  Every rank tracks
        every rank's log, from the same global move lists, so offsets never have to be communicated."""
        import torch
        from nums_b200 import cuda_compute as cc
        from nums_b200 import trace

        def elements(h):
            return (-(-int(np.prod(h.shape, dtype=np.int64)) // 16)) * 16        # 128-byte aligned slots

        needed = {}
        for h, _dst in moves:
            needed.setdefault(h.hid, h)
        for hid in [hid for hid, (ref, _off) in self.slots.items() if ref() is None]:
            del self.slots[hid]                                                   # (space comes back at the next reset)
        fresh = [h for hid, h in needed.items() if hid not in self.slots]
        reset = False
        if fresh:
            fill = dict(self.fill)
            for h in fresh:
                fill[h.home] = fill.get(h.home, 0) + elements(h)
            if max(fill.values()) > self.capacity:
                reset, fresh, fill = True, list(needed.values()), {}
                for h in fresh:
                    fill[h.home] = fill.get(h.home, 0) + elements(h)
                if not self._ensure(max(fill.values())):
                    return False
                self.slots, self.fill = {}, {}
            for h in fresh:
                off = self.fill.get(h.home, 0)
                self.slots[h.hid] = (weakref.ref(h), off)
                self.fill[h.home] = off + elements(h)
        home = torch.cuda.current_stream()
        up = cc._upload_stream()
        if fresh:
            if self.stream is None:
                self.stream = torch.cuda.Stream(priority=-1)
            side = self.stream
            # Barriers and staging run on a side stream ordered after everything the compute stream has been
            # given so far (the producers of the blocks to stage; the launches that consumed earlier pulls), so
            # the compute stream itself is free to start this flush's first launch on local operands at once.
            begin = torch.cuda.Event()
            begin.record(home)
            side.wait_event(begin)
            cc.await_uploads(side)
            with torch.cuda.stream(side):
                trace.mark("exchange: begin (staging %d blocks%s)" % (len(fresh), ", arena reset" if reset else ""), side)
                if reset:
                    self.handle.barrier(channel=0)           # nobody is still pulling the contents about to be overwritten
                for h in fresh:
                    if h.home == rank:
                        n = int(np.prod(h.shape, dtype=np.int64))
                        off = self.slots[h.hid][1]
                        src = backend.settle(h.value)
                        src.record_stream(side)
                        self.buf[off:off + n].view(h.shape).copy_(src)          # D2D placement copy (plumbing)
                trace.mark("exchange: staged", side)
                self.handle.barrier(channel=1)               # every rank's new blocks are in place
                ready = torch.cuda.Event()
                ready.record(side)
            up.wait_event(ready)
        trace.mark("exchange: pulls begin", up)
        device = torch.device("cuda", torch.cuda.current_device())
        for h, dst in moves:
            if dst != rank:
                continue
            remote = self.handle.get_buffer(h.home, h.shape, torch.float64, self.slots[h.hid][1])
            with torch.cuda.stream(up):
                dev = torch.empty(h.shape, dtype=torch.float64, device=device)
                dev.copy_(remote, non_blocking=True)
            dev.record_stream(home)
            event = torch.cuda.Event()
            event.record(up)
            cc._Transfers.seq += 1
            cc._Transfers.last_event = event
            dev._nums_ready = (cc._Transfers.seq, event)
            h.value = dev
        trace.mark("exchange: pulls done", up)
        self.pulled = True
        return True


class SpmdSystem(object):
    """See the module docstring.  ``local`` executes kernels on this rank; ``comm`` is a ``multi_gpu.Comm``."""

    def __init__(self, local, comm=None, placement="auto", check=None):
        from nums_b200 import multi_gpu
        self.methods = {}         # first: the reference's System.__getattribute__ (systems.py:62-66) reads it on every access
        self.remote_functions = {}
        self.local = local
        self.comm = comm if comm is not None else multi_gpu.Comm()
        self.rank, self.world_size = self.comm.rank, self.comm.world
        self.placement = placement
        self.device_grid = multi_gpu.device_grid(self.world_size)
        self.check = bool(int(os.environ.get("NUMS_SPMD_CHECK", "0"))) if check is None else bool(check)
        self.backend = _TorchBackend(local) if hasattr(local, "contractions") else _NumpyBackend(local)
        self.compute_module = getattr(local, "compute_module", None)
        self.compute_imp = getattr(local, "compute_imp", None)
        self.rng_cls = getattr(local, "rng_cls", None)
        self._registered = set()
        self._next = 0
        self._bop_meta = {}
        self._infer_memo = {}
        self._specials = {n[3:]: getattr(self, n) for n in dir(type(self)) if n.startswith("_k_")}   # kernel name -> handler
        self._owners = {}
        self._lazies = []
        self._scatters = []                   # weakrefs to handles whose value is a _LazyScatter
        self._copied = weakref.WeakSet()      # handles with cached copies away from home
        self.stats = {"moves": 0, "moved_bytes": 0, "all_reduces": 0, "broadcasts": 0, "meta_broadcasts": 0,
                      "executed": 0, "skipped": 0, "replicated": 0, "flushes": 0, "peer_exchanges": 0,
                      "scatter_chains": 0, "scatter_exchanges": 0, "scatter_moved_bytes": 0}
        self._peer = None
        if (isinstance(self.backend, _TorchBackend) and self.world_size > 1
                and os.environ.get("NUMS_SPMD_PEER", "1") != "0"):
            self._peer = PeerExchange(self.comm)
            local.contractions.group_policy = "single"   # flush() below cuts launches along k itself

    # -- lifecycle ------------------------------------------------------------------------------------
    def init(self):
        if hasattr(self.local, "init") and not getattr(self.local, "remote_functions", None):
            self.local.init()
        names = getattr(self.local, "remote_functions", None)
        if names is None:      # the oracle system of the CPU tests: kernel names from its implementation
            names = [n for n in dir(self.local.imp) if not n.startswith("_") and callable(getattr(self.local.imp, n))]
        for name in names:
            self.remote_functions[name] = name
            self._publish(name, self._make_callable(name))
        self.comm.setup_side_channel()

    def shutdown(self):
        self.flush()

    def _make_callable(self, name):
        def kernel(*args, **kwargs):
            return self.call(name, *args, **kwargs)
        kernel.__name__ = name
        return kernel

    def _publish(self, name, fn):
        """``system.<name>`` -> ``fn`` through ``methods`` and as an instance attribute (see CudaSystem._publish)."""
        self.methods[name] = fn
        owner = next((k for k in type(self).__mro__ if name in k.__dict__), None)
        if owner is None or owner.__name__ == "ComputeInterface":      # never shadow the system's own API
            self.__dict__[name] = fn

    def __getattr__(self, name):
        methods = self.__dict__.get("methods", {})
        if name in methods:
            return methods[name]
        raise AttributeError(name)

    def remote(self, function, remote_params):
        return function

    def register(self, name, func, remote_params=None):
        """Registered (I/O) functions run on exactly ONE rank -- they have side effects."""
        if name in self.remote_functions:
            return
        self.local.register(name, func, remote_params)
        self.remote_functions[name] = name
        self._publish(name, self._make_callable(name))
        self._registered.add(name)

    def nodes(self):
        return [{"Resources": {"node:%d" % r: 1.0}} for r in range(self.world_size)]

    def get_rng(self, seed):
        if seed is None:        # every rank must hand out the same (seed, jump_index) pairs
            import random
            seed = self.comm.broadcast_object(random.getrandbits(128) if self.rank == 0 else None, 0)
        return self.rng_cls(seed)

    def get_options(self, cluster_entry, cluster_shape):
        return {"resources": {"node:%d" % self.owner(tuple(cluster_entry), tuple(cluster_shape)): 1.0 / 10 ** 4}}

    def get_block_addresses(self, grid):
        return {entry: "node:%d" % self.owner(entry, grid.grid_shape) for entry in grid.get_entry_iterator()}

    def call_with_options(self, name, args, kwargs, options):
        rank = None
        for key in (options or {}).get("resources", {}):
            if isinstance(key, str) and key.startswith("node:"):
                rank = int(key.split(":")[1])
        return self.call(name, *args, _rank=rank, **kwargs)

    # -- placement ------------------------------------------------------------------------------------
    def owner(self, grid_entry, grid_shape):
        """Rank that executes calls scheduled on ``grid_entry`` of a grid of ``grid_shape`` when the operands
        do not decide it."""
        if self.world_size == 1 or len(grid_entry) == 0:
            return 0
        key = (grid_entry, grid_shape)
        hit = self._owners.get(key)
        if hit is not None:
            return hit
        long_axes = [i for i, g in enumerate(grid_shape) if g > 1]
        if self.placement == "flat" or len(long_axes) <= 1:
            out = int(np.ravel_multi_index(tuple(grid_entry), tuple(grid_shape))) % self.world_size
        else:
            pr, pc = self.device_grid
            a0, a1 = long_axes[0], long_axes[1]
            out = int((grid_entry[a0] % pr) * pc + (grid_entry[a1] % pc))
        if len(self._owners) < (1 << 16):
            self._owners[key] = out
        return out

    def _hint(self, sysk, rank=None):
        if rank is not None:
            return rank % self.world_size
        if not sysk or "grid_entry" not in sysk:
            return None
        entry, shape = sysk["grid_entry"], sysk["grid_shape"]
        if entry.__class__ is not tuple or shape.__class__ is not tuple:
            entry, shape = tuple(int(e) for e in entry), tuple(int(g) for g in shape)
        return self.owner(entry, shape)

    def _exec_rank(self, handles, sysk, single, rank=None):
        owned = [h for h in handles if h.home != REPLICATED]
        hint = self._hint(sysk, rank)
        if owned:
            score = {}
            for h in owned:
                for r in (h.home,) + tuple(h.copies):
                    score[r] = score.get(r, 0) + h.nbytes
            best = max(score.values())
            ranks = sorted(r for r, s in score.items() if s == best)
            return hint if hint in ranks else ranks[0]
        if single:
            return hint if hint is not None else 0
        if hint is None or sysk is None or int(np.prod(sysk.get("grid_shape", ()), dtype=np.int64)) <= 1:
            return REPLICATED
        return hint

    # -- handles --------------------------------------------------------------------------------------
    def _new_handle(self, home, value, shape, dtype):
        self._next += 1
        return Handle(self._next, home, value, shape, dtype)

    def _meta_of(self, x):
        if isinstance(x, Handle):
            return x.shape, x.dtype
        return self.backend.meta(x)

    def _describe(self, r):
        if self.backend.is_block(r):
            shape, dt = self.backend.meta(r)
            return ("b", tuple(shape), np.dtype(dt).str)
        if isinstance(r, tuple):
            return ("t", [self._describe(x) for x in r])
        if isinstance(r, list):
            return ("l", [self._describe(x) for x in r])
        return ("v", r)

    def _wrap(self, result, desc, home):
        """Handles for a kernel result: ``result`` is the actual value on ranks that executed, else None."""
        kind = desc[0]
        if kind == "b":
            return self._new_handle(home, result, desc[1], np.dtype(desc[2]))
        if kind in ("t", "l"):
            items = [self._wrap(None if result is None else result[i], d, home) for i, d in enumerate(desc[1])]
            return tuple(items) if kind == "t" else items
        return desc[1]

    def _unwrap(self, x):
        if isinstance(x, Handle):
            return x.value
        if isinstance(x, (list, tuple)) and any(isinstance(v, Handle) for v in x):
            return type(x)(self._unwrap(v) for v in x)
        return x

    @staticmethod
    def _handles_in(args, kwargs):
        out = []
        for v in list(args) + list(kwargs.values()):
            if isinstance(v, Handle):
                out.append(v)
            elif isinstance(v, (list, tuple)):
                out.extend(h for h in v if isinstance(h, Handle))
        return out

    # -- object store ---------------------------------------------------------------------------------
    def put(self, value):
        block = self.local.put(value)
        shape, dt = self.backend.meta(block)
        return self._new_handle(REPLICATED, block, shape, dt)

    def put_at(self, value, grid_entry, grid_shape, shape=None, dtype=None):
        """``put`` for callers that know where the block belongs.  The reference's ``put`` carries no placement
        information (systems.py:94-95), so plain ``put`` replicates; here only ``owner(grid_entry, grid_shape)``
        uploads ``value`` (a host array or an existing device block).  The other ranks pass ``value=None`` with
        ``shape`` and ``dtype`` (or the same array, which is then ignored)."""
        home = self.owner(tuple(grid_entry), tuple(grid_shape))
        if value is not None:
            vshape, vdtype = self.backend.meta(value) if self.backend.is_block(value) else (np.shape(value), np.asarray(value).dtype)
            shape = vshape if shape is None else shape
            dtype = vdtype if dtype is None else dtype
        elif shape is None or dtype is None:
            raise ValueError("put_at: ranks that do not own the block must give shape and dtype")
        if self.rank == home and value is None:
            raise ValueError("put_at: rank %d owns block %s and must pass its data" % (home, (grid_entry,)))
        block = self.local.put(value) if self.rank == home else None
        return self._new_handle(home, block, shape, np.dtype(dtype))

    def get_owned(self, oids):
        """{position in ``oids``: host array} for the blocks that live on THIS rank (no exchange): how a
        distributed result is drained to the hosts that own it, one PCIe link per GPU."""
        self.flush()
        mine = [(i, h) for i, h in enumerate(oids) if isinstance(h, Handle) and h.home in (self.rank, REPLICATED)]
        values = self.local.get([h.value for _i, h in mine])
        return {i: v for (i, _h), v in zip(mine, values)}

    def evict_copies(self):
        """Forget every cached copy of a remote block (the blocks themselves stay at home).  Operand copies
        are kept by default because blocks are immutable; a benchmark that wants every product to pay for
        its exchange calls this between steps."""
        for h in list(self._copied):
            if h.home != REPLICATED and h.home != self.rank:
                h.value = None
            h.copies = ()
        self._copied = weakref.WeakSet()

    def get(self, oids):
        self.flush()
        seen = {}

        def walk(o):
            if isinstance(o, Handle):
                if o.hid not in seen:
                    seen[o.hid] = self._fetch(o)
                return seen[o.hid]
            if isinstance(o, list):
                return [walk(v) for v in o]
            if isinstance(o, tuple):
                return tuple(walk(v) for v in o)
            return o
        local_values = walk(oids)
        return self.local.get(local_values)

    def get_async(self, oid):
        """``CudaSystem.get_async`` for a replicated block (every rank reads its own copy); other blocks are
        fetched synchronously."""
        self.flush()
        if isinstance(oid, Handle) and oid.home == REPLICATED and hasattr(self.local, "get_async"):
            return self.local.get_async(oid.value)
        value = self.get(oid)
        return lambda: value

    def _fetch(self, h):
        """The block of ``h`` on THIS rank (broadcast from its home when it is not replicated)."""
        if h.home == REPLICATED:
            return h.value
        src = h.home
        if self.rank == src:
            buf = self.backend.concrete(h.value)
        else:
            buf = self.backend.empty(h.shape, h.dtype)
        self.comm.broadcast(buf, src)
        self.stats["broadcasts"] += 1
        if h.nbytes <= PROMOTE_BYTES:
            if self.rank != src:
                h.value = buf
            h.home = REPLICATED
        return buf

    def flush_and_sync(self):
        self.flush()
        self.backend.synchronize()

    synchronize = flush_and_sync

    # -- transfers ------------------------------------------------------------------------------------
    def _move_many(self, moves):
        """[(handle, dst)]: make each block available on ``dst`` (cached there).  One batched exchange."""
        moves = [(h, dst) for h, dst in moves if not h.available_on(dst)]
        if not moves:
            return
        if self._peer is not None and self._peer.usable(moves) and self._peer.run(moves, self.rank, self.backend):
            for h, dst in moves:
                h.copies = tuple(h.copies) + (dst,)
                self._copied.add(h)
                self.stats["moves"] += 1
                self.stats["moved_bytes"] += h.nbytes
            self.stats["peer_exchanges"] += 1
            return
        sends, recvs = [], []
        for h, dst in moves:
            src = h.home
            if self.rank == src:
                sends.append((self.backend.concrete(h.value), dst))
            elif self.rank == dst:
                buf = self.backend.empty(h.shape, h.dtype)
                recvs.append((buf, src))
                h.value = buf
            h.copies = tuple(h.copies) + (dst,)
            self._copied.add(h)
            self.stats["moves"] += 1
            self.stats["moved_bytes"] += h.nbytes
        self.comm.exchange(sends, recvs, order=[(h.home, dst) for h, dst in moves])

    # -- dispatch -------------------------------------------------------------------------------------
    def call(self, name, *args, **kwargs):
        sysk = kwargs.pop("syskwargs", None)          # (**kwargs is a fresh dict per call)
        rank = kwargs.pop("_rank", None)
        if name == "touch":
            self.flush()
            return True
        special = self._specials.get(name)
        if special is not None:
            out = special(args, kwargs, sysk)
            if out is not _NOT_HANDLED:
                return out
        return self._generic(name, args, kwargs, sysk, rank)

    def _colocated(self, name, args, kwargs):
        """Fast path of ``_generic`` for the common case of the per-block kernels (map_uop, reduce_axis, GEMV-shaped
        tensordot, lr_grad_hess, ...): positional block arguments only, nothing lazy, no cached copies, and every
        operand that is not replicated lives on ONE rank.  That rank executes (the same answer ``_exec_rank`` gives:
        it holds all the operand bytes), nothing has to move, and the result description comes from a memo of
        ``infer_result`` keyed by the call's shapes / dtypes / scalar arguments.  ``_NOT_HANDLED`` for anything else."""
        e = REPLICATED
        key = [name]
        for a in args:
            if a.__class__ is Handle:
                if a.lazy is not None or a.copies:
                    return _NOT_HANDLED
                if a.home != REPLICATED:
                    if e == REPLICATED:
                        e = a.home
                    elif a.home != e:
                        return _NOT_HANDLED
                key.append((a.shape, a.dtype))
            elif a is None or a.__class__ in (str, int, bool, float):
                key.append(a)
            elif a.__class__ is tuple and len(a) <= 8 and all(x.__class__ in (int, str, bool, float) for x in a):
                key.append(a)
            elif (a.__class__ is tuple or a.__class__ is dict) and not a:
                key.append(())
            else:
                return _NOT_HANDLED
        if e == REPLICATED:
            return _NOT_HANDLED                      # all replicated: the generic path decides (may replicate the call)
        for k, v in kwargs.items():
            if v is None or v.__class__ in (str, int, bool, float):
                key.append((k, v))
            else:
                return _NOT_HANDLED
        key = tuple(key)
        desc = self._infer_memo.get(key)
        if desc is None:
            desc = infer_result(name, args, kwargs, self._meta_of)
            if desc is None:
                return _NOT_HANDLED
            if len(self._infer_memo) < 8192:
                self._infer_memo[key] = desc
        result = None
        if self.rank == e:
            self.stats["executed"] += 1
            result = self.local.call(name, *[a.value if a.__class__ is Handle else a for a in args], **kwargs)
            if self.check:
                actual = self._describe(result)
                if actual != desc:
                    raise AssertionError("SPMD result inference mismatch for %s: inferred %r, actual %r" % (name, desc, actual))
        else:
            self.stats["skipped"] += 1
        if desc[0] == "b":
            self._next += 1
            return Handle._make(self._next, e, result, desc[1], np.dtype(desc[2]))
        return self._wrap(result, desc, e)

    def _generic(self, name, args, kwargs, sysk, rank=None):
        if rank is None and name not in self._registered:
            out = self._colocated(name, args, kwargs)
            if out is not _NOT_HANDLED:
                return out
        handles = self._handles_in(args, kwargs)
        if any(h.lazy is not None for h in handles):
            self.flush()
        single = name in self._registered
        e = self._exec_rank(handles, sysk, single, rank)
        if e == REPLICATED:
            self.stats["replicated"] += 1
            result = self.local.call(name, *[self._unwrap(a) for a in args], **{k: self._unwrap(v) for k, v in kwargs.items()})
            return self._wrap(result, self._describe(result), REPLICATED)
        desc = None if single else infer_result(name, args, kwargs, self._meta_of)
        self._move_many([(h, e) for h in handles])
        if self.rank == e:
            self.stats["executed"] += 1
            error = None
            try:
                result = self.local.call(name, *[self._unwrap(a) for a in args],
                                         **{k: self._unwrap(v) for k, v in kwargs.items()})
            except Exception as exc:  # noqa: BLE001
                if desc is not None:
                    raise
                error, result = exc, None
            if desc is None:
                self.stats["meta_broadcasts"] += 1
                payload = ("error", repr(error)) if error is not None else ("ok", self._describe(result))
                self.comm.broadcast_object(payload, e)
                if error is not None:
                    raise error
                desc = payload[1]
            elif self.check:
                actual = self._describe(result)
                if actual != desc:
                    raise AssertionError("SPMD result inference mismatch for %s: inferred %r, actual %r" % (name, desc, actual))
            return self._wrap(result, desc, e)
        self.stats["skipped"] += 1
        if desc is None:
            self.stats["meta_broadcasts"] += 1
            status, desc = self.comm.broadcast_object(None, e)
            if status == "error":
                raise RuntimeError("kernel %s failed on rank %d: %s" % (name, e, desc))
        return self._wrap(None, desc, e)

    # -- lazily summed contractions ---------------------------------------------------------------------
    def _lazy_handle(self, items, hint, shape, dtype):
        lazy = _Lazy(items, hint)
        first = items[0][-1]
        single = all(it[-1] == first for it in items)
        self._next += 1
        h = Handle._make(self._next, first if single else REPLICATED, None, shape, dtype)
        h.lazy = lazy
        self._lazies.append(weakref.ref(h))
        return h

    def _term_rank(self, a1, a2, hint):
        if a1.home == REPLICATED and a2.home == REPLICATED:
            return hint
        for cand in ([hint] if hint is not None else []) + [a1.home, a2.home]:
            if cand != REPLICATED and a1.available_on(cand) and a2.available_on(cand):
                return cand
        return hint if hint is not None else (a1.home if a1.home != REPLICATED else a2.home)

    def _k_bop(self, args, kwargs, sysk):
        if len(args) == 7 and len(kwargs) == 1 and "axes" in kwargs:      # how Block.bop calls (base.py:220-231)
            op, a1, a2, s1, s2, t1, t2 = args
            axes = kwargs["axes"]
        else:
            bound = dict(zip(_BOP_PARAMS, args))
            bound.update(kwargs)
            if len(bound) != len(_BOP_PARAMS):
                return _NOT_HANDLED
            op, a1, a2, s1, s2, t1, t2, axes = (bound[k] for k in _BOP_PARAMS)
        if a1.__class__ is not Handle or a2.__class__ is not Handle:
            return _NOT_HANDLED
        if op == "tensordot":
            if (axes == 1 and len(s1) == 2 and len(s2) == 2 and a1.lazy is None and a2.lazy is None
                    and a1.dtype == _F64 and a2.dtype == _F64 and s1[1] == s2[0]
                    and s1[0] >= LAZY_MIN_EXTENT and s2[1] >= LAZY_MIN_EXTENT and s1[1] >= 1
                    and not (s1[1] == 128 and s2[1] == 128 and s1[0] >= 16384 and not t1 and not t2)):
                hint = self._hint(sysk)
                hint = 0 if hint is None else hint
                item = ("dot", a1, a2, tuple(s1), tuple(s2), bool(t1), bool(t2), self._term_rank(a1, a2, hint))
                return self._lazy_handle([item], hint, (int(s1[0]), int(s2[1])), _F64)
            return _NOT_HANDLED
        if op == "add" and not t1 and not t2 and a1.dtype == _F64 and a2.dtype == _F64 \
                and tuple(s1) == tuple(s2) == a1.shape == a2.shape:
            l1, l2 = a1.lazy, a2.lazy
            if (l1 is not None and l1.__class__ is not _Lazy) or (l2 is not None and l2.__class__ is not _Lazy):
                return _NOT_HANDLED                       # a pending shuffle: _generic materialises it first
            if l1 is not None or l2 is not None:
                lazy_hint = (l1 or l2).hint
                items = []
                for h, lz in ((a1, l1), (a2, l2)):
                    if lz is not None:
                        items.extend(lz.items)
                    else:
                        items.append(("blk", h, h.home if h.home != REPLICATED else lazy_hint))
                return self._lazy_handle(items, lazy_hint, a1.shape, _F64)
            if (a1.home != REPLICATED and a2.home != REPLICATED and a1.home != a2.home
                    and not a1.available_on(a2.home) and not a2.available_on(a1.home)):
                # partial results living on different ranks (the add chains of _tensordot / reduce_axis):
                # summed where they are and all-reduced, instead of being shipped to one rank one by one
                hint = self._hint(sysk)
                items = [("blk", a1, a1.home), ("blk", a2, a2.home)]
                return self._lazy_handle(items, 0 if hint is None else hint, a1.shape, _F64)
        # elementwise on operands that already sit together (config 1, the Newton iteration's vector algebra):
        # the executing rank and the result's shape / dtype follow without the generic machinery
        if a1.lazy is None and a2.lazy is None and not a1.copies and not a2.copies:
            h1, h2 = a1.home, a2.home
            if h1 == h2:
                e = h1
            elif h1 == REPLICATED:
                e = h2
            elif h2 == REPLICATED:
                e = h1
            else:
                return _NOT_HANDLED
            if e == REPLICATED:
                return _NOT_HANDLED
            key = (op, a1.dtype, a2.dtype, tuple(s1), tuple(s2))
            meta = self._bop_meta.get(key)
            if meta is None:
                desc = infer_result("bop", (op, a1, a2, s1, s2, t1, t2, axes), {}, self._meta_of)
                if desc is None or desc[0] != "b":
                    return _NOT_HANDLED
                meta = (desc[1], np.dtype(desc[2]))
                if len(self._bop_meta) < 4096:
                    self._bop_meta[key] = meta
            value = None
            if e == self.rank:
                self.stats["executed"] += 1
                value = self.local.call("bop", op, a1.value, a2.value, s1, s2, t1, t2, axes=axes)
            else:
                self.stats["skipped"] += 1
            self._next += 1
            return Handle._make(self._next, e, value, meta[0], meta[1])
        return _NOT_HANDLED

    # -- shuffles: chains of update_block_along_axis recorded, then one gather / all-to-all / scatter ---------
    def _k_update_block_along_axis(self, args, kwargs, sysk):
        if kwargs or len(args) != 4 or self.world_size == 1:
            return _NOT_HANDLED
        dst, src, index_pairs, axis = args
        if dst.__class__ is not Handle or src.__class__ is not Handle:
            return _NOT_HANDLED
        if src.lazy is not None or (dst.lazy is not None and dst.lazy.__class__ is not _LazyScatter):
            return _NOT_HANDLED
        nd = len(dst.shape)
        if nd == 0 or len(src.shape) != nd:
            return _NOT_HANDLED                           # shapes that only agree through broadcasting
        axis = int(axis) % nd
        if dst.shape[:axis] + dst.shape[axis + 1:] != src.shape[:axis] + src.shape[axis + 1:]:
            return _NOT_HANDLED
        pairs = list(index_pairs)
        dst_idx = np.fromiter((int(p[0]) for p in pairs), dtype=np.int64, count=len(pairs))
        src_idx = np.fromiter((int(p[1]) for p in pairs), dtype=np.int64, count=len(pairs))
        if len(pairs) and (dst_idx.min() < 0 or src_idx.min() < 0 or dst_idx.max() >= dst.shape[axis]
                           or src_idx.max() >= src.shape[axis]):
            return _NOT_HANDLED                           # negative / out-of-range indices: the kernel's own checks
        if dst.lazy is not None:
            chain = dst.lazy
            if chain.axis != axis:
                return _NOT_HANDLED
            base, items, rank = chain.base, chain.items + [(src, dst_idx, src_idx)], chain.rank
        else:
            hint = self._hint(sysk)
            rank = dst.home if dst.home != REPLICATED else (0 if hint is None else hint)
            base, items = dst, [(src, dst_idx, src_idx)]
        self._next += 1
        h = Handle._make(self._next, rank, None, dst.shape, dst.dtype)
        h.lazy = _LazyScatter(base, items, axis, rank)
        self._scatters.append(weakref.ref(h))
        return h

    def _flush_scatters(self):
        """Materialise every pending shuffle chain that is still referenced (identical plan on every rank).
        Intermediate links of a chain are dead by now (each call replaced ``dst_block.oid``), so only the last
        handle of every chain does any work."""
        alive = []
        for ref in self._scatters:
            h = ref()
            if h is not None and h.lazy is not None and h.lazy.__class__ is _LazyScatter:
                alive.append(h)
        self._scatters = []
        if not alive:
            return
        backend = self.backend
        self.stats["scatter_chains"] += len(alive)
        # A. rows that must travel: the source's owner packs them, one batched exchange moves them
        sends, recvs, order = [], [], []
        pieces = {}                                      # hid -> [(src block or packed rows, dst_idx, src_idx)]
        for h in alive:
            chain = h.lazy
            mine = self.rank == chain.rank
            plan = pieces.setdefault(h.hid, [])
            for src, dst_idx, src_idx in chain.items:
                if src.available_on(chain.rank):
                    if mine:
                        plan.append((src.value, dst_idx, src_idx))
                    continue
                n = int(len(dst_idx))
                if n == 0:
                    continue
                packed_shape = src.shape[:chain.axis] + (n,) + src.shape[chain.axis + 1:]
                nbytes = int(np.prod(packed_shape, dtype=np.int64)) * src.dtype.itemsize
                order.append((src.home, chain.rank))
                self.stats["scatter_moved_bytes"] += nbytes
                if self.rank == src.home:
                    sends.append((backend.take_along(src.value, src_idx, chain.axis), chain.rank))
                elif mine:
                    buf = backend.empty(packed_shape, src.dtype)
                    recvs.append((buf, src.home))
                    plan.append((buf, dst_idx, np.arange(n, dtype=np.int64)))
        if order:
            # a base block that lives elsewhere (not the case for BlockArray.empty results) travels whole, first
            self.stats["scatter_exchanges"] += 1
            self.comm.exchange(sends, recvs, order=order)
        moves = [(h.lazy.base, h.lazy.rank) for h in alive if not h.lazy.base.available_on(h.lazy.rank)]
        if moves:
            self._move_many(moves)
        # B. every destination owner: one copy of the base block, all pieces scattered into it
        for h in alive:
            chain = h.lazy
            if self.rank == chain.rank:
                h.value = backend.scatter_chain(chain.base.value, pieces[h.hid], chain.axis)
            h.home = chain.rank
            h.lazy = None

    def flush(self):
        """Materialise every lazy sum that is still referenced (identical plan on every rank)."""
        if self._scatters:
            self._flush_sums()
            self._flush_scatters()
            return
        self._flush_sums()

    def _flush_sums(self):
        alive = []
        for ref in self._lazies:
            h = ref()
            if h is not None and h.lazy is not None:
                alive.append(h)
        self._lazies = []
        if not alive:
            self.backend.flush()
            return
        self.stats["flushes"] += 1
        # A. operands that must travel: one batched exchange.  Every block gets a HEAD term -- the one with most
        #    operand bytes already on its executing rank (often fully local) -- whose missing operands go first;
        #    the other terms follow in k order rotated past the head, so that at any moment different ranks pull
        #    from different owners instead of all hitting the owners of k = 0.
        moves, seen, heads = [], set(), {}
        rounds = {}
        for h in alive:
            dots = [it for it in h.lazy.items if it[0] == "dot"]
            if not dots:
                continue
            local_bytes = [sum(op.nbytes for op in (it[1], it[2]) if op.available_on(it[-1])) for it in dots]
            head = max(range(len(dots)), key=lambda t: (local_bytes[t], -t))
            heads[h.hid] = dots[head]
            for t, it in enumerate(dots):
                rounds.setdefault((t - head) % len(dots), []).append(it)
        for rnd in sorted(rounds):
            for it in rounds[rnd]:
                for operand in (it[1], it[2]):
                    key = (operand.hid, it[-1])
                    if key not in seen and not operand.available_on(it[-1]):
                        seen.add(key)
                        moves.append((operand, it[-1]))
        if self._peer is not None:
            self._peer.pulled = False
        self._move_many(moves)
        overlap = self._peer is not None and self._peer.pulled        # pulls are in flight on the upload stream
        # B. every rank evaluates its own terms.  CudaSystem defers them into grouped launches whose CTAs keep a
        #    block's k-chain in registers.  While pulls are in flight the work is cut ALONG k, not by block:
        #    launch 1 = the first term of every block (waits for those operands only, and -- covering all
        #    blocks -- fills the machine as evenly as the whole product), launch 2 = all remaining terms
        #    accumulated onto launch 1's result; by then the rest of the exchange has landed.
        partial, borrowed, tails = {}, {}, []
        build = getattr(getattr(self.local, "contractions", None), "build", None)
        for h in alive:
            acc, count = None, 0
            mine = [it for it in h.lazy.items if it[-1] == self.rank]
            if build is not None and len(mine) > 1 and all(it[0] == "dot" for it in mine):
                # the whole k-chain of this block as ONE deferred contraction (no per-term dot / add replay)
                terms = [(it[1].value, it[2].value, it[3], it[4], it[5], it[6]) for it in mine]
                if overlap and len(terms) >= 3 and heads.get(h.hid) in mine:
                    at = mine.index(heads[h.hid])
                    head = build(terms[at:at + 1], h.shape)
                    if head is not None:
                        partial[h.hid], borrowed[h.hid] = head, False
                        tails.append((h, head, terms[:at] + terms[at + 1:]))
                        continue
                acc = build(terms, h.shape)
                if acc is not None:
                    partial[h.hid], borrowed[h.hid] = acc, False
                    continue
            for it in h.lazy.items:
                if it[-1] != self.rank:
                    continue
                if it[0] == "blk":
                    v = it[1].value
                else:
                    _k, a1, a2, s1, s2, t1, t2, _r = it
                    v = self.local.call("bop", "tensordot", a1.value, a2.value, s1, s2, t1, t2, axes=1)
                acc = v if acc is None else self.local.call("bop", "add", acc, v, h.shape, h.shape, False, False, axes=None)
                count += 1
            partial[h.hid] = acc
            borrowed[h.hid] = count == 1 and any(it[0] == "blk" and it[-1] == self.rank for it in h.lazy.items)
        # drop the loop's last handles before the local flush: a deferred contraction that something besides
        # `partial` still references would be launched a second time on its own (see deferred.py)
        acc = v = head = None
        self.backend.flush()
        if tails:
            for h, head, terms in tails:
                rest = build(terms, h.shape)
                done = self.backend.settle(head)
                if rest is None:      # (cannot happen for terms that built once; keep the chain correct anyway)
                    rest = done
                    for a1v, a2v, s1, s2, t1, t2 in terms:
                        dot = self.local.call("bop", "tensordot", a1v, a2v, s1, s2, t1, t2, axes=1)
                        rest = self.local.call("bop", "add", rest, dot, h.shape, h.shape, False, False, axes=None)
                    partial[h.hid] = rest
                else:
                    partial[h.hid] = self.local.call("bop", "add", rest, done, h.shape, h.shape, False, False, axes=None)
            rest = done = head = dot = None
            tails = None
            self.backend.flush()
        # C. sums whose terms live on several ranks: all-reduce (small ones share a buffer)
        shared, small = [], []
        for h in alive:
            ranks = h.lazy.ranks()
            if len(ranks) == 1:
                h.value = self.backend.settle(partial[h.hid]) if self.rank == ranks[0] else None
            elif h.nbytes <= COALESCE_BYTES:
                small.append(h)
            else:
                shared.append(h)
        if len(small) == 1:
            shared.append(small.pop())
        if small:
            total = sum(int(np.prod(h.shape, dtype=np.int64)) for h in small)
            buf = self.backend.zeros((total,), np.float64)
            views, off = [], 0
            for h in small:
                view = self.backend.flat_view(buf, off, h.shape)
                if partial[h.hid] is not None:
                    self.backend.copy_into(view, self.backend.concrete(partial[h.hid]))
                views.append(view)
                off += int(np.prod(h.shape, dtype=np.int64))
            self.comm.all_reduce_sum(buf)
            self.stats["all_reduces"] += 1
            for h, view in zip(small, views):
                h.value = view
        for h in shared:
            acc = partial[h.hid]
            if acc is None:
                buf = self.backend.zeros(h.shape, np.float64)
            elif borrowed[h.hid]:
                buf = self.backend.clone(acc)           # never reduce in place into somebody's block
            else:
                buf = self.backend.concrete(acc)
            self.comm.all_reduce_sum(buf)
            self.stats["all_reduces"] += 1
            h.value = buf
        for h in alive:
            h.lazy = None

    # -- n-ary sum over blocks living on several ranks: partial sums + one all-reduce ----------------------
    def _k_sum_reduce(self, args, kwargs, sysk):
        if kwargs or not args or not all(isinstance(a, Handle) for a in args):
            return _NOT_HANDLED
        if any(h.lazy is not None for h in args):
            self.flush()
        owned = [h for h in args if h.home != REPLICATED]
        if not owned:
            return _NOT_HANDLED
        for r in sorted({h.home for h in owned}):
            if all(h.available_on(r) for h in args):
                return _NOT_HANDLED                      # everything already sits on one rank
        first = args[0]
        if any(h.shape != first.shape or h.dtype != first.dtype for h in args) or first.dtype.kind not in "fiu":
            return _NOT_HANDLED
        anchor = owned[0].home                           # replicated addends are counted once, there
        mine = [h.value for h in args if (h.home if h.home != REPLICATED else anchor) == self.rank]
        part = self.local.call("sum_reduce", *mine) if mine else self.backend.zeros(first.shape, first.dtype)
        part = self.backend.concrete(part)
        self.comm.all_reduce_sum(part)
        self.stats["all_reduces"] += 1
        return self._new_handle(REPLICATED, part, first.shape, first.dtype)

    def _fused_gram_qr(self, args, where, n, dt):
        """Stacked-R ``qr`` when every per-block R is still a ``cuda_compute.DeferredR`` (a Gram matrix): the
        Gram matrices are summed where they are, all-reduced (n*n + 2 doubles, the last two counting blocks so
        that every rank learns whether ALL blocks were deferred) and factored once on every rank -- one
        collective and one small-matrix kernel instead of per-block factorizations plus a tree of stacked QRs.
        None (no collective issued) on the CPU back end; None after the all-reduce when a block was not deferred
        or the condition bound of the sum is too large, in which case the tree below takes over."""
        if not isinstance(self.backend, _TorchBackend) or dt != _F64:
            return None
        import torch
        from nums_b200 import cuda_compute as cc
        from nums_b200._lib import LIB
        mine = [h.value for h, w in zip(args, where) if w == self.rank]
        deferred = [v for v in mine if v.__class__ is cc.DeferredR and v.value is None and v.shape == (n, n)]
        buf = self.backend.zeros((n * n + 2,), np.float64)
        if mine:
            if len(deferred) == len(mine):
                ptrs = (cc._lib.ctypes.c_void_p * len(deferred))(*[v.gram.data_ptr() for v in deferred])
                LIB.check(LIB.dll.nums_sum_reduce(len(deferred), ptrs, cc._lib.F64, n * n, buf.data_ptr(), cc._stream()))
            counts = torch.tensor([float(len(deferred)), float(len(mine))], dtype=torch.float64)
            buf[n * n:].copy_(counts.to(buf.device, non_blocking=True))
        self.comm.all_reduce_sum(buf)
        self.stats["all_reduces"] += 1
        got, total = (float(v) for v in buf[n * n:].cpu())
        if got != total:
            return None
        _low, _low_inv, upper, kappa = cc._factor_gram(buf[:n * n].view(n, n))
        if kappa > cc.QR_GRAM_ACCEPT_KAPPA:
            return None
        cc.QR_STATS["gram_fused"] += 1
        return self._new_handle(REPLICATED, upper, (n, n), dt)

    # -- stacked-R qr over blocks living on several ranks: local QR, binary tree, broadcast ---------------
    def _k_qr(self, args, kwargs, sysk):
        mode, axis = kwargs.get("mode", "reduced"), kwargs.get("axis")
        if mode != "r" or len(args) < 2 or axis is None or int(axis) != 0 or not all(isinstance(a, Handle) for a in args):
            return _NOT_HANDLED
        if any(h.lazy is not None for h in args):
            self.flush()
        owned = [h for h in args if h.home != REPLICATED]
        if not owned or len({h.home for h in owned}) < 2:
            return _NOT_HANDLED
        if any(len(h.shape) != 2 or h.shape[1] != args[0].shape[1] or h.dtype != args[0].dtype for h in args) \
                or args[0].dtype.kind != "f":
            return _NOT_HANDLED
        n, dt = args[0].shape[1], args[0].dtype
        anchor = owned[0].home
        where = [h.home if h.home != REPLICATED else anchor for h in args]
        ranks = sorted(set(where))
        rows = {r: min(sum(h.shape[0] for h, w in zip(args, where) if w == r), n) for r in ranks}
        fused = self._fused_gram_qr(args, where, n, dt)
        if fused is not None:
            return fused
        r_local = None
        if self.rank in ranks:
            mine = [h.value for h, w in zip(args, where) if w == self.rank]
            r_local = self.local.call("qr", *mine, mode="r", axis=0) if len(mine) > 1 else self.local.call("qr", mine[0], mode="r")
        step = 1
        while step < len(ranks):
            for idx in range(0, len(ranks), 2 * step):
                if idx + step >= len(ranks):
                    continue
                dst, src = ranks[idx], ranks[idx + step]
                if self.rank == src:
                    self.comm.send(self.backend.concrete(r_local), dst)
                elif self.rank == dst:
                    other = self.backend.empty((rows[src], n), dt)
                    self.comm.recv(other, src)
                    r_local = self.local.call("qr", r_local, other, mode="r", axis=0)
                rows[dst] = min(rows[dst] + rows[src], n)
            step *= 2
        root = ranks[0]
        out = self.backend.concrete(r_local) if self.rank == root else self.backend.empty((rows[root], n), dt)
        self.comm.broadcast(out, root)
        self.stats["broadcasts"] += 1
        return self._new_handle(REPLICATED, out, (rows[root], n), dt)
