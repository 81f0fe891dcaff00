"""Deferred block contractions: turning ``dot`` + ``add`` chains into one grouped DMMA launch.

``BlockArray._tensordot`` (blockarray.py:460-472) computes every result block as
``dot(A[i,0], B[0,j]); for k > 0: result_block += dot(A[i,k], B[k,j])`` -- 512 ``tensordot`` and
448 ``add`` kernel calls for an 8 x 8 grid, each ``add`` a full read-read-write pass over a result
block and each 2048^2 ``tensordot`` 1.73 waves on 148 SMs.  Blocks are immutable in NumS, so
``CudaSystem`` may postpone the arithmetic: a float64 ``tensordot`` returns a
``DeferredContraction`` (a list of (A, B) terms), ``add`` of deferred contractions concatenates the
term lists, and the first call that needs actual data (``touch``, ``get``, any other kernel)
flushes *all* outstanding contractions through ``nums_gemm_grouped``: one launch whose CTAs walk
the tiles of every result block and accumulate each block's whole k-chain in registers.  The
unchanged host code therefore runs the blocked matmul as a single large GEMM.
"""
import weakref

import numpy as np
import torch

from nums_b200 import _lib, cuda_compute
from nums_b200._lib import LIB

MIN_EXTENT = 64   # smaller outputs / vector forms run eagerly (GEMV / split-K kernels)
STREAM_MIN_ROWS = 16384   # (m x 128) . (128 x 128) with m at least this runs on dgemm_tall128_stream_kernel (gemm.cu)
MAX_GROUPS = 16   # launches a flush may be cut into to overlap in-flight uploads (see _launch_groups)


GemmTerm, GemmProblem = _lib.GemmTerm, _lib.GemmProblem


class DeferredContraction(object):
    """value = addend + sum_t op(A_t) . op(B_t); materialised by ``ContractionQueue.flush``."""
    __slots__ = ("terms", "addend", "shape", "flags", "value", "__weakref__")

    def __init__(self, terms, shape, flags, addend=None):
        self.terms = terms        # tuple of (A tensor, lda, B tensor, ldb, k, source block of A, source block of B)
        self.shape = shape        # (m, n)
        self.flags = flags        # (trans_a, trans_b), shared by all terms
        self.addend = addend      # concrete (m, n) float64 tensor or None
        self.value = None

    @property
    def dtype(self):
        return torch.float64


def _dmma_ok(t, ld):
    return t.dtype == torch.float64 and t.data_ptr() % 16 == 0 and ld % 2 == 0


def plan_launch_groups(needs, done, max_groups, policy="even"):
    """Pure planning step of ``ContractionQueue._launch_groups``.

    ``policy`` = "even": neighbouring runs are merged evenly (operands streaming in over PCIe at a rate
    comparable to the GEMM's); "head": the first run alone, everything else in ONE launch -- for operands
    pulled over NVLink (nums_b200.spmd.PeerExchange), which all arrive within the first launch's run time,
    so a second split would only add a partial wave; "single": one launch that waits for everything
    (the caller cuts the work itself, nums_b200.spmd.SpmdSystem.flush).

    ``needs[i]`` is ``None`` or ``(sequence number, event)`` of the latest upload contraction i reads;
    uploads up to sequence number ``done`` have already been waited for.  Returns a list of
    ``(indices, upto)``: contractions that become ready together form a run, runs are ordered by
    readiness and neighbouring runs merged until at most ``max_groups`` launches remain; ``upto`` is
    the latest upload the launch has to wait for (``None``: nothing)."""
    if all(n is None or n[0] <= done for n in needs):
        return [(list(range(len(needs))), None)]
    seq_of = [n[0] if n is not None and n[0] > done else 0 for n in needs]
    order = sorted(range(len(needs)), key=lambda i: seq_of[i])     # stable: creation order inside a run
    runs = []
    for i in order:
        if runs and seq_of[runs[-1][-1]] == seq_of[i]:
            runs[-1].append(i)
        else:
            runs.append([i])
    if policy == "single":
        chunks = [runs]
    elif policy == "head" and len(runs) > 2:
        first = 1 if seq_of[runs[0][0]] else 2      # run 0 may be "needs nothing": take the first waiting run too
        chunks = [runs[:first], runs[first:]] if len(runs) > first else [runs]
    else:
        merge = max(1, -(-len(runs) // max_groups))
        chunks = [runs[lo:lo + merge] for lo in range(0, len(runs), merge)]
    out = []
    for chunk in chunks:
        idx = [i for run in chunk for i in run]
        tags = [needs[i] for i in idx if seq_of[i]]
        out.append((idx, max(tags, key=lambda t: t[0]) if tags else None))
    return out


class ContractionQueue(object):

    def __init__(self):
        self._pending = []   # weak references to unmaterialised contractions, in creation order
        self.flushes = 0
        self.enabled = True
        self.group_policy = "even"       # how a flush is cut into launches, see plan_launch_groups
        self.last_flush = None           # {"contractions", "tile_terms", "redundant_tile_terms"} of the latest flush
        self.redundant_tile_terms = 0
        self._warned = False
        self.materialized_seen = False   # deferred handles exist somewhere: keep resolving arguments

    # -- building -------------------------------------------------------------------------------
    def tensordot(self, a1, a2, a1_shape, a2_shape, a1_T, a2_T, axes):
        """Deferred np.tensordot(a1, a2, axes=1) for 2-D float64 operands, or None if ineligible."""
        if not self.enabled or axes != 1 or len(a1_shape) != 2 or len(a2_shape) != 2:
            return None
        if not (isinstance(a1, torch.Tensor) and isinstance(a2, torch.Tensor)):
            return None
        if a1.dtype != torch.float64 or a2.dtype != torch.float64:
            return None
        m, k = int(a1_shape[0]), int(a1_shape[1])
        k2, n = int(a2_shape[0]), int(a2_shape[1])
        if k != k2 or m < MIN_EXTENT or n < MIN_EXTENT or k < 1:
            return None
        if k == 128 and n == 128 and m >= STREAM_MIN_ROWS and not a1_T and not a2_T and a1.is_contiguous() \
                and a2.is_contiguous():
            return None       # tall block times small square (the Q of TSQR): the eager streaming kernel is faster
        x = cuda_compute._operand(a1, a1_shape, a1_T)     # metadata only: no kernel, no wait for uploads
        y = cuda_compute._operand(a2, a2_shape, a2_T)
        A, ta, lda = cuda_compute._as_matrix(x, m, k)
        B, tb, ldb = cuda_compute._as_matrix(y, k, n)
        if not (_dmma_ok(A, lda) and _dmma_ok(B, ldb)):
            return None
        return self._register(DeferredContraction(((A, lda, B, ldb, k, a1, a2),), (m, n), (bool(ta), bool(tb))))

    def build(self, terms, shape):
        """One deferred contraction sum_t np.tensordot(a1_t, a2_t, 1) from a whole term list
        [(a1, a2, a1_shape, a2_shape, a1_T, a2_T)] (what a dot / add chain would have built call by call);
        None if any term is ineligible or the terms do not share one (trans_a, trans_b) class."""
        if not self.enabled or not terms:
            return None
        out, flags = [], None
        m, n = shape
        for a1, a2, a1_shape, a2_shape, a1_T, a2_T in terms:
            if not (isinstance(a1, torch.Tensor) and isinstance(a2, torch.Tensor)) \
                    or a1.dtype != torch.float64 or a2.dtype != torch.float64 or len(a1_shape) != 2 or len(a2_shape) != 2:
                return None
            k = int(a1_shape[1])
            if (int(a1_shape[0]), int(a2_shape[1])) != (m, n) or k != int(a2_shape[0]) or m < MIN_EXTENT or n < MIN_EXTENT or k < 1:
                return None
            x = cuda_compute._operand(a1, a1_shape, a1_T)
            y = cuda_compute._operand(a2, a2_shape, a2_T)
            A, ta, lda = cuda_compute._as_matrix(x, m, k)
            B, tb, ldb = cuda_compute._as_matrix(y, k, n)
            if not (_dmma_ok(A, lda) and _dmma_ok(B, ldb)):
                return None
            if flags is None:
                flags = (bool(ta), bool(tb))
            elif flags != (bool(ta), bool(tb)):
                return None
            out.append((A, lda, B, ldb, k, a1, a2))
        return self._register(DeferredContraction(tuple(out), (m, n), flags))

    def add(self, x, y, x_shape, y_shape, x_T, y_T):
        """Lazy x + y when at least one side is an unmaterialised contraction of the same shape."""
        dx = isinstance(x, DeferredContraction) and x.value is None
        dy = isinstance(y, DeferredContraction) and y.value is None
        if not (dx or dy) or x_T or y_T or tuple(x_shape) != tuple(y_shape):
            return None
        if dx and dy:
            if x.flags != y.flags or (x.addend is not None and y.addend is not None):
                return None
            addend = x.addend if x.addend is not None else y.addend
            return self._register(DeferredContraction(x.terms + y.terms, x.shape, x.flags, addend))
        lazy, other = (x, y) if dx else (y, x)
        if isinstance(other, DeferredContraction):
            other = other.value
        if (not isinstance(other, torch.Tensor) or other.dtype != torch.float64 or lazy.addend is not None
                or tuple(other.shape) != tuple(lazy.shape) or not other.is_contiguous()):
            return None
        return self._register(DeferredContraction(lazy.terms, lazy.shape, lazy.flags, other))

    def _register(self, d):
        self._pending.append(weakref.ref(d))
        self.materialized_seen = True
        return d

    # -- materialising ----------------------------------------------------------------------------
    def resolve(self, obj):
        """Replace deferred contractions (also inside lists / tuples) by concrete tensors."""
        if isinstance(obj, DeferredContraction):
            if obj.value is None:
                self.flush()
            return obj.value
        if obj.__class__ is cuda_compute.DeferredR:          # an R factor still held as a Gram matrix
            return obj.materialize()
        if isinstance(obj, list):
            return [self.resolve(o) for o in obj]
        if isinstance(obj, tuple) and any(isinstance(o, (DeferredContraction, cuda_compute.DeferredR)) for o in obj):
            return tuple(self.resolve(o) for o in obj)
        return obj

    def flush(self):
        """One grouped launch per (trans_a, trans_b) class for everything still referenced."""
        alive = []
        for ref in self._pending:
            d = ref()
            if d is not None and d.value is None:
                alive.append(d)
        self._pending = []
        if not alive:
            return
        self.flushes += 1
        self._audit(alive)
        by_flags = {}
        for d in alive:
            by_flags.setdefault(d.flags, []).append(d)
        for (ta, tb), group in by_flags.items():
            launches = self._launch_groups(group)
            for chunk, upto in launches:
                self._launch(ta, tb, chunk, upto, len(launches) > 1)

    def _audit(self, alive):
        """Tile accounting of a flush, and the dangling-chain check.

        ``x += dot`` builds a new contraction from the previous partial chain; the partial chain normally dies
        at once (CPython reference counting).  If something -- typically a loop variable of the caller --
        still references it when the flush happens, it is materialised on its own: a whole extra block of GEMM
        work whose result nobody reads (this cost 3.3 ms per product once).  A live
        contraction whose term list is a proper prefix of another live one is exactly that case: it is counted
        in ``last_flush["redundant_tiles"]``, reported once with a warning, and raises under
        ``NUMS_DEFERRED_STRICT=1`` (set by the tests)."""
        tiles = redundant = 0
        by_first = {}
        for d in alive:
            m, n = d.shape
            tiles += (-(-m // 128)) * (-(-n // 128)) * max(len(d.terms), 1)
            if d.terms:
                by_first.setdefault(id(d.terms[0]), []).append(d)
        dangling = []
        for group in by_first.values():
            if len(group) < 2:
                continue
            group.sort(key=lambda d: len(d.terms))
            keys = [tuple(map(id, d.terms)) for d in group]
            for i, d in enumerate(group):
                if any(len(keys[j]) > len(keys[i]) and keys[j][:len(keys[i])] == keys[i] for j in range(i + 1, len(group))):
                    m, n = d.shape
                    redundant += (-(-m // 128)) * (-(-n // 128)) * len(d.terms)
                    dangling.append(d)
        self.last_flush = {"contractions": len(alive), "tile_terms": tiles, "redundant_tile_terms": redundant}
        self.redundant_tile_terms += redundant
        if dangling:
            import os
            import warnings
            msg = ("%d partial dot/add chain(s) were still referenced at flush time and are materialised separately "
                   "(%d of %d tile-terms of this flush are wasted): drop the intermediate handles (e.g. loop "
                   "variables) before the flush" % (len(dangling), redundant, tiles))
            if os.environ.get("NUMS_DEFERRED_STRICT") == "1":
                raise AssertionError(msg)
            if not self._warned:
                self._warned = True
                warnings.warn(msg, RuntimeWarning, stacklevel=3)

    @staticmethod
    def _need(d):
        """Latest in-flight upload ((sequence, event) or None) among the operands of a contraction."""
        latest = None
        for term in d.terms:
            for src in term[5:]:
                tag = cuda_compute.upload_tag(src)
                if tag is not None and (latest is None or tag[0] > latest[0]):
                    latest = tag
        return latest

    def _launch_groups(self, group):
        """Split one (trans_a, trans_b) class into launch groups by operand readiness.

        Operands uploaded asynchronously (cuda_compute.upload) carry their upload's sequence number;
        contractions are ordered by the latest upload they read and cut into at most MAX_GROUPS
        launches (one per set of contractions that become ready together), each waiting only for what
        it reads, so the GEMM overlaps the rest of the H2D traffic.  With nothing in flight this is a
        single launch."""
        done = cuda_compute._Transfers.awaited.get(cuda_compute._stream_unordered(), 0)
        needs = [self._need(d) for d in group]
        return [([group[i] for i in idx], upto)
                for idx, upto in plan_launch_groups(needs, done, MAX_GROUPS, self.group_policy)]

    def _launch(self, ta, tb, group, upto, mark_done):
        if upto is not None:
            cuda_compute.await_uploads(upto=upto)
        stream = cuda_compute._stream_unordered()
        nterms = sum(len(d.terms) for d in group)
        problems = (GemmProblem * len(group))()
        terms = (GemmTerm * nterms)()
        cursor = 0
        for i, d in enumerate(group):
            m, n = d.shape
            out = cuda_compute._empty((m, n), np.float64)
            d.value = out
            pr = problems[i]
            pr.C = out.data_ptr()
            pr.Cin = d.addend.data_ptr() if d.addend is not None else None
            pr.ldc, pr.ldcin, pr.m, pr.n = n, n, m, n
            pr.term_begin, pr.term_count = cursor, len(d.terms)
            for term in d.terms:
                A, lda, B, ldb, k = term[:5]
                tm = terms[cursor]
                tm.A, tm.B, tm.lda, tm.ldb, tm.k = A.data_ptr(), B.data_ptr(), lda, ldb, k
                cursor += 1
        device = group[0].value.device
        from nums_b200 import trace
        trace.mark("grouped gemm: start (%d blocks, %d terms)" % (len(group), nterms))
        LIB.call_ws(LIB.dll.nums_gemm_grouped, device,
                    ((_lib.F64, int(ta), int(tb), len(group), problems, nterms, terms), (stream,)))
        trace.mark("grouped gemm: end")
        if mark_done:
            # lets get_assembled start the D2H of these blocks while later groups still compute
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream())
            for d in group:
                d.value._nums_done = done
        for d in group:      # operands may be released now; the stream keeps the ordering
            d.terms = ()
            d.addend = None
