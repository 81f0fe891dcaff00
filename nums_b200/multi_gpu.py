"""Single-box multi-GPU drivers: one process per GPU, ``torch.distributed`` for the exchanges.

The reference places blocks with Ray (``BlockCyclicScheduler``, schedulers.py:170-246) and moves
them through the object store; its only "collectives" are gathers to one node (``sum_reduce`` in
``_matvec``, blockarray.py:574-578; the final ``qr(*R_oids)``, application.py:807-814).  Here the
block grid is partitioned over the ranks of one NVSwitch box instead:

* ``bop`` / elementwise      : blocks are dealt round-robin, nothing is exchanged;
* blocked matmul             : SUMMA on a ``pr x pc`` device grid -- C(i,j) lives on rank
                               ``(i mod pr, j mod pc)``; step k needs the A(:,k) panel of the owner in
                               my device row and the B(k,:) panel of the owner in my device column.
                               On GPUs the resident panels live in symmetric (peer-mapped) memory and
                               every rank PULLS the panels it needs with copy-engine transfers over
                               NVLink -- no SM is taken from the DMMA kernel and no rank waits for
                               another; NCCL broadcasts (prefetched one step ahead) are the fallback;
* TSQR                       : local Householder R per rank, then a binary tree of
                               ``qr([R_a; R_b])`` over send/recv pairs, R broadcast back;
* Newton logistic regression : fused local gradient/Hessian, one all-reduce of d + d*d doubles per
                               iteration, replicated d x d solve.

The drivers take a ``system`` (``CudaSystem`` in production; the CPU oracle system in the gloo
tests) and only use its kernel interface plus the ``Comm`` wrapper below, so the host logic is
testable on CPU with ``world_size = 2`` (tests/test_multi_gpu_cpu.py).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist


def device_grid(world_size):
    """pr x pc factorisation used for SUMMA: 1x1, 1x2, 2x2, 2x4, ... (pr <= pc)."""
    pr = int(np.floor(np.sqrt(world_size)))
    while world_size % pr:
        pr -= 1
    return pr, world_size // pr


def nccl_options():
    """NCCL kernels on a HIGH-PRIORITY stream.  A long GEMM grid keeps thousands of CTAs pending; the
    block scheduler only starts CTAs of a later, equal-priority kernel once those are all dispatched,
    which would push every panel broadcast to the tail of the GEMM it is supposed to overlap with
    (measured: 1.3 ms exposed per SUMMA step).  With priority the broadcast CTAs are picked as soon as
    any GEMM tile retires."""
    if not (dist.is_available() and dist.is_nccl_available() and torch.cuda.is_available()):
        return None
    if dist.is_initialized() and dist.get_backend() != "nccl":
        return None
    opts = dist.ProcessGroupNCCL.Options()
    opts.is_high_priority_stream = True
    return opts


class Comm(object):
    """Thin layer over ``torch.distributed`` that moves either torch tensors (NCCL, device memory)
    or NumPy arrays (gloo, used by the CPU tests)."""

    def __init__(self):
        self.enabled = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank() if self.enabled else 0
        self.world = dist.get_world_size() if self.enabled else 1
        self._groups = {}
        self._side = None

    def group(self, ranks):
        """Sub-communicator for `ranks` (every rank must call this in the same order)."""
        ranks = tuple(sorted(ranks))
        if len(ranks) == self.world:
            return None
        if ranks not in self._groups:
            self._groups[ranks] = dist.new_group(list(ranks), pg_options=nccl_options())
        return self._groups[ranks]

    @staticmethod
    def _as_tensor(x):
        if isinstance(x, np.ndarray):
            return torch.from_numpy(x), True
        if isinstance(x, torch.Tensor) and x.is_cuda:
            # blocks put() asynchronously may still be on the upload stream: NCCL orders itself after
            # the current stream only
            from nums_b200 import cuda_compute
            cuda_compute.await_uploads()
        return x, False

    def broadcast(self, buf, src, group=None, async_op=False):
        if not self.enabled or self.world == 1:
            return None
        t, _ = self._as_tensor(buf)
        return dist.broadcast(t, src=src, group=group, async_op=async_op)

    def all_reduce_sum(self, buf):
        if not self.enabled or self.world == 1:
            return buf
        t, _ = self._as_tensor(buf)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return buf

    def send(self, buf, dst):
        t, _ = self._as_tensor(buf)
        dist.send(t, dst=dst)

    def recv(self, buf, src):
        t, _ = self._as_tensor(buf)
        dist.recv(t, src=src)

    def barrier(self):
        if self.enabled and self.world > 1:
            dist.barrier()

    # -- used by nums_b200.spmd -----------------------------------------------------------------------
    def setup_side_channel(self):
        """A gloo group for small pickled messages (result descriptions, seeds): a host-side exchange that
        does not touch the GPU streams.  Collective: every rank calls it once."""
        if self.enabled and self.world > 1 and self._side is None:
            self._side = dist.group.WORLD if dist.get_backend() == "gloo" else dist.new_group(backend="gloo")

    def broadcast_object(self, obj, src):
        if not self.enabled or self.world == 1:
            return obj
        if self._side is None:
            self.setup_side_channel()
        box = [obj]
        dist.broadcast_object_list(box, src=src, group=self._side)
        return box[0]

    def exchange(self, sends, recvs, order):
        """One batched point-to-point exchange.  ``order`` lists (src, dst) of every transfer of the plan in
        the same order on all ranks; ``sends`` = [(buffer, dst)] and ``recvs`` = [(buffer, src)] are this
        rank's part of it, in plan order."""
        if not self.enabled or self.world == 1:
            return
        ops, si, ri = [], 0, 0
        for src, dst in order:
            if self.rank == src:
                t, _ = self._as_tensor(sends[si][0])
                ops.append(dist.P2POp(dist.isend, t, dst))
                si += 1
            elif self.rank == dst:
                t, _ = self._as_tensor(recvs[ri][0])
                ops.append(dist.P2POp(dist.irecv, t, src))
                ri += 1
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()


def init_distributed():
    """torchrun environment -> (rank, world): one process per GPU, NCCL over NVLink (RANK / LOCAL_RANK /
    WORLD_SIZE / MASTER_* from the environment).  No-op when the process group exists or world is 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not (dist.is_available() and dist.is_initialized()):
        if torch.cuda.is_available():
            local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
            torch.cuda.set_device(local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), pg_options=nccl_options())
        else:
            dist.init_process_group("gloo")
    return rank, world


def _empty_like_block(system, shape, like):
    """Receive buffer of `shape` in the same memory space / dtype as block `like`."""
    if isinstance(like, np.ndarray):
        return np.empty(shape, dtype=like.dtype)
    return torch.empty(shape, dtype=like.dtype, device=like.device)


# ---------------------------------------------------------------------------------------------
# elementwise: shard-local
# ---------------------------------------------------------------------------------------------
def local_entries(grid_shape, rank, world):
    """Round-robin deal of a flattened grid (the fork's placement, gpu_systems.py:163-164)."""
    count = int(np.prod(grid_shape))
    return [np.unravel_index(i, grid_shape) for i in range(count) if i % world == rank]


# ---------------------------------------------------------------------------------------------
# SUMMA
# ---------------------------------------------------------------------------------------------
class SummaMatmul(object):
    """C = A @ B for square block grids, blocks distributed 2-D block-cyclically.

    ``a_blocks[(i, k)]`` / ``b_blocks[(k, j)]`` hold this rank's blocks (others absent).  All
    blocks are (bs x bs) here -- the bench shapes; ragged grids go through the single-GPU path.
    """

    def __init__(self, system, comm, grid, block_size, dtype_like):
        self.system, self.comm = system, comm
        self.g = grid
        self.bs = block_size
        self.pr, self.pc = device_grid(comm.world)
        self.r, self.c = divmod(comm.rank, self.pc)
        self.like = dtype_like
        # sub-communicators: my device row (same r) and my device column (same c)
        self.row_groups = [comm.group([rr * self.pc + cc for cc in range(self.pc)]) for rr in range(self.pr)]
        self.col_groups = [comm.group([rr * self.pc + cc for rr in range(self.pr)]) for cc in range(self.pc)]
        self.my_i = [i for i in range(grid) if i % self.pr == self.r]
        self.my_j = [j for j in range(grid) if j % self.pc == self.c]
        self.trace = None      # set to a list to record (label, k, cuda event) marks (development aid)
        self._peer = None      # _PeerPanels once set up, False when unavailable

    def owner_a(self, i, k):
        return (i % self.pr) * self.pc + (k % self.pc)

    def owner_b(self, k, j):
        return (k % self.pr) * self.pc + (j % self.pc)

    def owner_c(self, i, j):
        return (i % self.pr) * self.pc + (j % self.pc)

    def pack(self, a_blocks, b_blocks):
        """Stack this rank's blocks into one contiguous panel per k, so that a SUMMA step needs ONE
        transfer per operand instead of one per block: panels[0][k] holds A(i, k) for i in my_i
        (owned when k mod pc == c), panels[1][k] holds B(k, j) for j in my_j (owned when k mod pr == r).
        On GPUs the panels are written into symmetric memory that the peers map (see _PeerPanels)."""
        first = next(iter(a_blocks.values())) if a_blocks else next(iter(b_blocks.values()))
        ka = [k for k in range(self.g) if k % self.pc == self.c]
        kb = [k for k in range(self.g) if k % self.pr == self.r]
        peer = self._peer_panels(first)
        if peer is not None:
            return peer.publish({k: [a_blocks[(i, k)] for i in self.my_i] for k in ka},
                                {k: [b_blocks[(k, j)] for j in self.my_j] for k in kb})

        if isinstance(first, torch.Tensor) and first.is_cuda:
            # blocks put() from page-locked memory may still be crossing PCIe on the upload stream; the stack
            # copies below run on the current stream
            from nums_b200 import cuda_compute
            cuda_compute.await_uploads()

        def stack(blocks):
            first = blocks[0]
            if isinstance(first, np.ndarray):
                return np.stack(blocks)
            out = torch.empty((len(blocks),) + tuple(first.shape), dtype=first.dtype, device=first.device)
            for idx, blk in enumerate(blocks):
                out[idx].copy_(blk)      # device-to-device placement copy (plumbing, not arithmetic)
            return out
        pa = {k: stack([a_blocks[(i, k)] for i in self.my_i]) for k in ka}
        pb = {k: stack([b_blocks[(k, j)] for j in self.my_j]) for k in kb}
        return pa, pb

    def _peer_panels(self, like):
        """The symmetric-memory exchange, set up on first use; None when it does not apply (CPU blocks,
        one rank, ragged device grid, NUMS_SUMMA_PEER=0, or the set-up failed on any rank)."""
        if self._peer is None:
            self._peer = False
            usable = (isinstance(like, torch.Tensor) and like.is_cuda and self.comm.world > 1
                      and self.g % self.pr == 0 and self.g % self.pc == 0
                      and os.environ.get("NUMS_SUMMA_PEER", "1") != "0")
            if usable:
                peer = None
                try:
                    peer = _PeerPanels(self, like)
                except Exception as exc:  # noqa: BLE001 -- any failure selects the NCCL broadcasts
                    sys.stderr.write("[nums_b200] symmetric-memory panels unavailable (%s: %s); using NCCL broadcasts\n"
                                     % (type(exc).__name__, exc))
                ok = torch.tensor([1 if peer is not None else 0], dtype=torch.int32, device=like.device)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)     # all ranks take the same path
                if int(ok.item()) == 1:
                    self._peer = peer
        return self._peer or None

    def _panels(self, k, packed):
        """Start the (at most two) broadcasts of step k; returns (A panel, B panel, pending works)."""
        pa, pb = packed
        works = []
        src_a = self.r * self.pc + (k % self.pc)
        shape_a = (len(self.my_i), self.bs, self.bs)
        buf_a = pa[k] if src_a == self.comm.rank else _empty_like_block(self.system, shape_a, self.like)
        if self.pc > 1:
            works.append(self.comm.broadcast(buf_a, src_a, self.row_groups[self.r], async_op=True))
        src_b = (k % self.pr) * self.pc + self.c
        shape_b = (len(self.my_j), self.bs, self.bs)
        buf_b = pb[k] if src_b == self.comm.rank else _empty_like_block(self.system, shape_b, self.like)
        if self.pr > 1:
            works.append(self.comm.broadcast(buf_b, src_b, self.col_groups[self.c], async_op=True))
        a_panel = {i: buf_a[idx] for idx, i in enumerate(self.my_i)}
        b_panel = {j: buf_b[idx] for idx, j in enumerate(self.my_j)}
        return a_panel, b_panel, [w for w in works if w is not None]

    def run(self, a_blocks, b_blocks=None, flush_every=None):
        """Returns {(i, j): block} for the C blocks this rank owns.

        `a_blocks` is either the result of ``pack`` or a dict of this rank's A blocks (then `b_blocks`
        is the dict of B blocks and they are packed here).  The local updates of consecutive k-steps
        are handed to the system as one deferred chain, i.e. one grouped GEMM launch whose CTAs
        accumulate the whole chain in registers.  `flush_every` = number of k-steps per launch; the
        default depends on the exchange: with peer pulls every transfer is queued up front, so step 0
        is launched alone (it starts as soon as the first panels have landed) and steps 1 .. g-1 as ONE
        launch -- by then their panels have arrived, and one long launch pays the per-tile prologue
        (pipeline fill, C read and write) once instead of g - 1 times; with NCCL broadcasts, which are
        prefetched one step ahead, every step is its own launch."""
        packed = a_blocks if b_blocks is None else self.pack(a_blocks, b_blocks)
        peer = packed if isinstance(packed, _PublishedPanels) else None
        if flush_every is None or flush_every <= 0:
            flush_at = {0, self.g - 1} if peer is not None else set(range(self.g))
        else:
            flush_at = {k for k in range(self.g) if (k + 1) % flush_every == 0} | {self.g - 1}
        shape = (self.bs, self.bs)
        c_blocks = {}
        trace = self.trace
        if peer is not None:
            pulls = peer.owner.pull_all(peer)       # every transfer of this product, queued on the copy streams
            nxt = pulls[0]
        else:
            nxt = self._panels(0, packed)
        for k in range(self.g):
            a_panel, b_panel, works = nxt
            for w in works:
                w.wait()
            if trace is not None:
                trace.append(("panels_ready", k, self._mark()))
            if k + 1 < self.g:
                # NCCL path: prefetch while the GEMMs below run
                nxt = pulls[k + 1] if peer is not None else self._panels(k + 1, packed)
            for i in self.my_i:
                for j in self.my_j:
                    sysk = {"grid_entry": (i, j), "grid_shape": (self.g, self.g)}
                    dot = self.system.bop("tensordot", a_panel[i], b_panel[j], shape, shape, False, False,
                                          axes=1, syskwargs=sysk)
                    prev = c_blocks.get((i, j))
                    c_blocks[(i, j)] = dot if prev is None else self.system.bop(
                        "add", prev, dot, shape, shape, False, False, axes=None, syskwargs=sysk)
            # Drop the loop's last handles before flushing: a deferred contraction that is still referenced
            # is materialised, and `prev` is the complete k-chain of the last block so far -- one whole extra
            # block of GEMM work per launch (measured: a constant 3.3 ms per product at every N).
            prev = dot = None
            if hasattr(self.system, "flush") and k in flush_at:
                self.system.flush()   # one grouped launch for the C += A(:,k) B(k,:) updates so far
                if trace is not None:
                    trace.append(("gemm_done", k, self._mark()))
        if peer is not None:
            peer.owner.product_done()
        return c_blocks

    def _mark(self):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        return ev


class _EventWork(object):
    """`wait()` orders the current stream after a CUDA event (same call shape as a c10d Work)."""
    __slots__ = ("event",)

    def __init__(self, event):
        self.event = event

    def wait(self):
        torch.cuda.current_stream().wait_event(self.event)


class _PublishedPanels(object):
    """Result of ``SummaMatmul.pack`` on the symmetric-memory path."""
    __slots__ = ("owner", "pa", "pb", "ready")

    def __init__(self, owner, pa, pb, ready):
        self.owner, self.pa, self.pb, self.ready = owner, pa, pb, ready


class _PeerPanels(object):
    """Panel exchange of SUMMA over NVLink peer memory, driven by the copy engines.

    Every rank keeps its resident panels -- A(my_i, k) for the k it owns, B(k, my_j) likewise -- in one
    buffer of symmetric memory (``torch.distributed._symmetric_memory``: the same allocation on every
    rank, mapped into every peer's address space).  A product then needs no collective at all: each
    rank copies the panels of the owners in its device row / column straight out of their memory
    with plain device-to-device copies on two copy streams (one per operand, so the row and the
    column transfer use different engines), all of them queued when the product starts, and the
    grouped DMMA launch of step k waits for the events of panels k only.  Compared with NCCL
    broadcasts this takes no SM away from the GEMM (a broadcast channel is a resident CTA whose
    registers do not fit next to a 288-thread DMMA CTA, so every channel evicts one GEMM CTA), needs
    no rendezvous between ranks per step, and the data crosses each link once.

    The published panels are immutable while products run.  ``publish`` brackets the rewrite of the
    buffer with two device-side barriers of the symmetric-memory handle: the first guarantees that no
    peer is still pulling the previous contents, the second that every rank's new panels are in place.
    Receive buffers are double-buffered across products, so the pulls of product s + 1 may start while
    product s still computes.
    """

    def __init__(self, summa, like):
        import torch.distributed._symmetric_memory as symm
        self.s = summa
        g, bs = summa.g, summa.bs
        self.device = like.device
        self.dtype = like.dtype
        self.na, self.nb = g // summa.pc, g // summa.pr            # panels owned per operand
        self.shape_a = (len(summa.my_i), bs, bs)
        self.shape_b = (len(summa.my_j), bs, bs)
        self.numel_a = int(np.prod(self.shape_a))
        self.numel_b = int(np.prod(self.shape_b))
        total = self.na * self.numel_a + self.nb * self.numel_b
        self.buf = symm.empty(total, dtype=self.dtype, device=self.device)
        self.handle = symm.rendezvous(self.buf, dist.group.WORLD)
        self.streams = (torch.cuda.Stream(), torch.cuda.Stream())
        # Both receive sets are allocated now, on the current stream, before any product runs: allocating a set
        # lazily when its parity is first used could hand the copy streams memory the caching allocator had just
        # freed on the compute stream (e.g. the previous product's C blocks) while the GEMM still writes it.
        self._recv = [self._new_recv_set(), self._new_recv_set()]     # per parity: ({k: A panel}, {k: B panel})
        allocated = torch.cuda.Event()
        allocated.record()
        for stream in self.streams:
            stream.wait_event(allocated)
        self._done = [None, None]     # per parity: event after the last GEMM that read the set
        self._product = 0
        self._remote = {}

    def _new_recv_set(self):
        s = self.s
        return ({k: torch.empty(self.shape_a, dtype=self.dtype, device=self.device)
                 for k in range(s.g) if k % s.pc != s.c},
                {k: torch.empty(self.shape_b, dtype=self.dtype, device=self.device)
                 for k in range(s.g) if k % s.pr != s.r})

    def _mine_a(self, k):
        off = (k // self.s.pc) * self.numel_a
        return self.buf[off:off + self.numel_a].view(self.shape_a)

    def _mine_b(self, k):
        off = self.na * self.numel_a + (k // self.s.pr) * self.numel_b
        return self.buf[off:off + self.numel_b].view(self.shape_b)

    def _remote_view(self, rank, operand, k):
        key = (rank, operand, k)
        view = self._remote.get(key)
        if view is None:
            if operand == 0:
                view = self.handle.get_buffer(rank, self.shape_a, self.dtype, (k // self.s.pc) * self.numel_a)
            else:
                view = self.handle.get_buffer(rank, self.shape_b, self.dtype,
                                              self.na * self.numel_a + (k // self.s.pr) * self.numel_b)
            self._remote[key] = view
        return view

    def publish(self, a_lists, b_lists):
        """Write this rank's panels (lists of blocks per owned k) into the symmetric buffer."""
        from nums_b200 import cuda_compute
        cuda_compute.await_uploads()
        self.handle.barrier(channel=0)          # nobody is still reading the previous panels
        pa, pb = {}, {}
        for k, blocks in a_lists.items():
            pa[k] = self._mine_a(k)
            for idx, blk in enumerate(blocks):
                pa[k][idx].copy_(blk)           # device-to-device placement copy (plumbing, not arithmetic)
        for k, blocks in b_lists.items():
            pb[k] = self._mine_b(k)
            for idx, blk in enumerate(blocks):
                pb[k][idx].copy_(blk)
        self.handle.barrier(channel=1)          # every rank's panels are in place
        ready = torch.cuda.Event()
        ready.record()
        return _PublishedPanels(self, pa, pb, ready)

    def pull_all(self, published):
        """Queue every panel transfer of one product; returns per k (A panel, B panel, works)."""
        s = self.s
        parity = self._product & 1
        recv_a, recv_b = self._recv[parity]
        for stream in self.streams:
            stream.wait_event(published.ready)
            if self._done[parity] is not None:
                stream.wait_event(self._done[parity])     # the product that last read this receive set
        out = []
        for k in range(s.g):
            works = []
            src_a = s.r * s.pc + (k % s.pc)
            if src_a == s.comm.rank:
                buf_a = published.pa[k]
            else:
                buf_a = recv_a[k]
                with torch.cuda.stream(self.streams[0]):
                    buf_a.copy_(self._remote_view(src_a, 0, k), non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record()
                works.append(_EventWork(ev))
            src_b = (k % s.pr) * s.pc + s.c
            if src_b == s.comm.rank:
                buf_b = published.pb[k]
            else:
                buf_b = recv_b[k]
                with torch.cuda.stream(self.streams[1]):
                    buf_b.copy_(self._remote_view(src_b, 1, k), non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record()
                works.append(_EventWork(ev))
            out.append(({i: buf_a[idx] for idx, i in enumerate(s.my_i)},
                        {j: buf_b[idx] for idx, j in enumerate(s.my_j)}, works))
        return out

    def product_done(self):
        parity = self._product & 1
        ev = torch.cuda.Event()
        ev.record()
        self._done[parity] = ev
        self._product += 1


# ---------------------------------------------------------------------------------------------
# TSQR: tree reduction of R
# ---------------------------------------------------------------------------------------------
def tsqr_r_tree(system, comm, local_row_blocks, n):
    """R factor of the row-blocked matrix whose local row blocks are given; R ends on every rank.

    Local stage: one ``qr(mode='r')`` per block and one over the stacked local R's (what
    indirect_tsr does globally, application.py:784-814).  Cross-rank stage: binary tree."""
    rs = [system.qr(b, mode="r", axis=1, syskwargs={"grid_entry": (i, 0), "grid_shape": (len(local_row_blocks), 1)})
          for i, b in enumerate(local_row_blocks)]
    r = rs[0] if len(rs) == 1 else system.qr(*rs, mode="r", axis=0, syskwargs={"grid_entry": (0, 0), "grid_shape": (1, 1)})
    if hasattr(r, "materialize"):         # cuda_compute.DeferredR: a block's R that so far exists only as its Gram matrix
        r = r.materialize()
    step = 1
    while step < comm.world:
        if comm.rank % (2 * step) == step:
            comm.send(_contig(r), comm.rank - step)
        elif comm.rank % (2 * step) == 0 and comm.rank + step < comm.world:
            other = _empty_like_block(system, (n, n), r)
            comm.recv(other, comm.rank + step)
            r = system.qr(r, other, mode="r", axis=0, syskwargs={"grid_entry": (0, 0), "grid_shape": (1, 1)})
        step *= 2
    r = _contig(r)
    comm.broadcast(r, 0)
    return r


def _contig(x):
    if isinstance(x, np.ndarray):
        return np.ascontiguousarray(x)
    return x if x.is_contiguous() else x.contiguous()


def tsqr_q(system, x_block, r_inv, entry, grid_shape):
    m, n = x_block.shape
    return system.bop("tensordot", x_block, r_inv, (m, n), (n, n), False, False, axes=1,
                      syskwargs={"grid_entry": entry, "grid_shape": grid_shape})


# ---------------------------------------------------------------------------------------------
# Newton logistic regression
# ---------------------------------------------------------------------------------------------
def newton_lr(system, comm, x_blocks, y_blocks, d, tol, max_iter, grad_hess, step=None):
    """Newton iterations (glms.py:362-372) on row-sharded data; returns (beta, iterations).

    ``grad_hess(x_blocks, y_blocks, beta) -> 1-D buffer of d + d*d`` (summed over the blocks) is the
    fused kernel (cuda_compute.lr_grad_hess_blocks) or, in the CPU tests, its NumPy statement.  The all-reduce of that
    buffer replaces the reference's gathers (``sum_reduce`` of G gradients, the (d, d) add chain).
    ``step(gh, beta) -> (new beta, status)`` optionally fuses the update itself (cuda_compute.newton_step:
    solve, subtract, max |g| and the singularity flag in one launch, one 16-byte read-back per iteration);
    without it the update runs through the kernel interface like the reference (inv, tensordot, sub, abs, max).
    """
    beta = system.new_block("zeros", (0,), {"shape": (d,), "block_shape": (d,), "dtype": "float64"},
                            syskwargs={"grid_entry": (0,), "grid_shape": (1,)})
    sk = {"grid_entry": (0,), "grid_shape": (1,)}
    iters = 0
    get_async = getattr(system, "get_async", None)
    pending = None           # (beta after that iteration, its iteration number, waiter for {max |g|, info})

    def converged(entry):
        gmax, info = (float(v) for v in np.asarray(entry[2]()))
        if info != 0:
            raise np.linalg.LinAlgError("Singular matrix")
        return gmax <= tol

    for _ in range(max_iter):
        iters += 1
        acc = grad_hess(x_blocks, y_blocks, beta)    # g | H summed over this rank's row blocks
        comm.all_reduce_sum(acc)
        if step is not None:
            beta, status = step(acc, beta)
            # the status of iteration i is read after iteration i + 1 has been enqueued (no device idle time on the
            # 16-byte read-back); a converged iteration's beta is returned and the speculative step dropped
            waiter = get_async(status) if get_async is not None else (lambda v: (lambda: v))(system.get(status))
            if pending is not None and converged(pending):
                beta, iters = pending[0], pending[1]
                pending = None
                break
            pending = (beta, iters, waiter)
            continue
        g = acc[:d]
        h = acc[d:].reshape(d, d) if isinstance(acc, np.ndarray) else acc[d:].view(d, d)
        h_inv = system.inv(h, syskwargs={"grid_entry": (0, 0), "grid_shape": (1, 1)})
        step_vec = system.bop("tensordot", h_inv, g, (d, d), (d,), False, False, axes=1, syskwargs=sk)
        beta = system.bop("sub", beta, step_vec, (d,), (d,), False, False, axes=None, syskwargs=sk)
        gmax = system.reduce_axis("max", system.map_uop("abs", g, (), {}, syskwargs=sk), None, False, False, syskwargs=sk)
        if float(np.asarray(system.get(gmax))) <= tol:     # the one host sync per iteration
            break
    if pending is not None:
        converged(pending)
    return beta, iters
