"""Single-box multi-GPU drivers: one process per GPU, ``torch.distributed`` for the exchanges.

The reference places blocks with Ray (``BlockCyclicScheduler``, schedulers.py:170-246) and moves
them through the object store; its only "collectives" are gathers to one node (``sum_reduce`` in
``_matvec``, blockarray.py:574-578; the final ``qr(*R_oids)``, application.py:807-814).  Here the
block grid is partitioned over the ranks of one NVSwitch box instead:

* ``bop`` / elementwise      : blocks are dealt round-robin, nothing is exchanged;
* blocked matmul             : SUMMA on a ``pr x pc`` device grid -- C(i,j) lives on rank
                               ``(i mod pr, j mod pc)``; at step k the owners broadcast the A(:,k)
                               blocks along device rows and the B(k,:) blocks along device columns
                               (NCCL broadcasts, prefetched one step ahead of the local GEMMs);
* TSQR                       : local Householder R per rank, then a binary tree of
                               ``qr([R_a; R_b])`` over send/recv pairs, R broadcast back;
* Newton logistic regression : fused local gradient/Hessian, one all-reduce of d + d*d doubles per
                               iteration, replicated d x d solve.

The drivers take a ``system`` (``CudaSystem`` in production; the CPU oracle system in the gloo
tests) and only use its kernel interface plus the ``Comm`` wrapper below, so the host logic is
testable on CPU with ``world_size = 2`` (tests/test_multi_gpu_cpu.py).
"""
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

    def owner_a(self, i, k):
        return (i % self.pr) * self.pc + (k % self.pc)

    def owner_b(self, k, j):
        return (k % self.pr) * self.pc + (j % self.pc)

    def owner_c(self, i, j):
        return (i % self.pr) * self.pc + (j % self.pc)

    def pack(self, a_blocks, b_blocks):
        """Stack this rank's blocks into one contiguous panel per k, so that a SUMMA step needs ONE
        broadcast per operand instead of one per block: panels[0][k] holds A(i, k) for i in my_i
        (owned when k mod pc == c), panels[1][k] holds B(k, j) for j in my_j (owned when k mod pr == r)."""
        def stack(blocks):
            first = blocks[0]
            if isinstance(first, np.ndarray):
                return np.stack(blocks)
            out = torch.empty((len(blocks),) + tuple(first.shape), dtype=first.dtype, device=first.device)
            for idx, blk in enumerate(blocks):
                out[idx].copy_(blk)      # device-to-device placement copy (plumbing, not arithmetic)
            return out
        pa = {k: stack([a_blocks[(i, k)] for i in self.my_i]) for k in range(self.g) if k % self.pc == self.c}
        pb = {k: stack([b_blocks[(k, j)] for j in self.my_j]) for k in range(self.g) if k % self.pr == self.r}
        return pa, pb

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

    def run(self, a_blocks, b_blocks=None, flush_every=1):
        """Returns {(i, j): block} for the C blocks this rank owns.

        `a_blocks` is either the result of ``pack`` or a dict of this rank's A blocks (then `b_blocks`
        is the dict of B blocks and they are packed here).  The local updates of `flush_every`
        consecutive k-steps are handed to the system as one deferred chain: one grouped GEMM launch
        that overlaps with the broadcasts of the following step."""
        packed = a_blocks if b_blocks is None else self.pack(a_blocks, b_blocks)
        shape = (self.bs, self.bs)
        c_blocks = {}
        trace = self.trace
        nxt = self._panels(0, packed)
        for k in range(self.g):
            a_panel, b_panel, works = nxt
            for w in works:
                w.wait()
            if trace is not None:
                trace.append(("panels_ready", k, self._mark()))
            if k + 1 < self.g:
                nxt = self._panels(k + 1, packed)   # prefetch while the GEMMs below run
            for i in self.my_i:
                for j in self.my_j:
                    sysk = {"grid_entry": (i, j), "grid_shape": (self.g, self.g)}
                    dot = self.system.bop("tensordot", a_panel[i], b_panel[j], shape, shape, False, False,
                                          axes=1, syskwargs=sysk)
                    prev = c_blocks.get((i, j))
                    c_blocks[(i, j)] = dot if prev is None else self.system.bop(
                        "add", prev, dot, shape, shape, False, False, axes=None, syskwargs=sysk)
            if hasattr(self.system, "flush") and ((k + 1) % flush_every == 0 or k + 1 == self.g):
                self.system.flush()   # one grouped launch for the C += A(:,k) B(k,:) updates so far
                if trace is not None:
                    trace.append(("gemm_done", k, self._mark()))
        return c_blocks

    def _mark(self):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        return ev


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
def newton_lr(system, comm, x_blocks, y_blocks, d, tol, max_iter, grad_hess):
    """Newton iterations (glms.py:362-372) on row-sharded data; returns (beta, iterations).

    ``grad_hess(x_blocks, y_blocks, beta) -> 1-D buffer of d + d*d`` (summed over the blocks) is the
    fused kernel (cuda_compute.lr_grad_hess_blocks) or, in the CPU tests, its NumPy statement.  The all-reduce of that
    buffer replaces the reference's gathers (``sum_reduce`` of G gradients, the (d, d) add chain).
    """
    beta = system.new_block("zeros", (0,), {"shape": (d,), "block_shape": (d,), "dtype": "float64"},
                            syskwargs={"grid_entry": (0,), "grid_shape": (1,)})
    sk = {"grid_entry": (0,), "grid_shape": (1,)}
    iters = 0
    for _ in range(max_iter):
        iters += 1
        acc = grad_hess(x_blocks, y_blocks, beta)    # g | H summed over this rank's row blocks
        comm.all_reduce_sum(acc)
        g = acc[:d]
        h = acc[d:].reshape(d, d) if isinstance(acc, np.ndarray) else acc[d:].view(d, d)
        h_inv = system.inv(h, syskwargs={"grid_entry": (0, 0), "grid_shape": (1, 1)})
        step = system.bop("tensordot", h_inv, g, (d, d), (d,), False, False, axes=1, syskwargs=sk)
        beta = system.bop("sub", beta, step, (d,), (d,), False, False, axes=None, syskwargs=sk)
        gmax = system.reduce_axis("max", system.map_uop("abs", g, (), {}, syskwargs=sk), None, False, False, syskwargs=sk)
        if float(np.asarray(system.get(gmax))) <= tol:     # the one host sync per iteration
            break
    return beta, iters
