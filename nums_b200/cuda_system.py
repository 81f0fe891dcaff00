"""``CudaSystem`` -- the system object that owns block placement for ``cuda_compute``.

Mirror of the reference's ``System`` / ``SerialSystem`` (/root/reference/nums/core/systems/
systems.py:31-142): it loads ``compute_module.ComputeCls``, exposes every kernel both through
``system.call(name, ...)`` and as an attribute (``system.bop(...)``), strips ``syskwargs`` and
runs the kernel in-process.  Unlike ``SerialSystem`` (identity put/get, systems.py:94-98) ``put``
uploads a host array to HBM and ``get`` synchronises and downloads, as the fork's
``CupySerialSystem`` does (gpu_systems.py:88-100).

One process drives one GPU (``torch.distributed`` rank == device).  ``owner(grid_entry,
grid_shape)`` implements the reference's placement rules so that SPMD drivers
(``nums_b200.blocks``) know which rank computes a block:

* ``"cyclic"``  -- ``grid_entry[axis] mod device_grid[axis]`` (BlockCyclicScheduler.get_cluster_entry,
  schedulers.py:170-191);
* ``"flat"``    -- row-major flattened entry ``mod world`` (the fork's CupyParallelSystem,
  gpu_systems.py:163-164), which spreads tall-skinny grids over all devices.
"""
import inspect

import numpy as np
import torch

from nums_b200 import cuda_compute
from nums_b200.deferred import ContractionQueue, DeferredContraction

_BOP_PARAMS = ("op", "a1", "a2", "a1_shape", "a2_shape", "a1_T", "a2_T", "axes")


def _host_function(func):
    """Adapter for registered host-side (I/O) functions: device blocks in -> ndarrays, ndarray results
    of a supported dtype -> device blocks."""
    def to_host(v):
        return cuda_compute.download(v) if isinstance(v, torch.Tensor) else v

    def to_device(v):
        if isinstance(v, np.ndarray) and v.dtype != object and cuda_compute._lib.supported_dtype(v.dtype):
            return cuda_compute.upload(v)
        if isinstance(v, tuple):
            return tuple(to_device(x) for x in v)
        return v

    def wrapper(*args, **kwargs):
        out = func(*[to_host(a) for a in args], **{k: to_host(v) for k, v in kwargs.items()})
        return to_device(out)
    wrapper.__name__ = getattr(func, "__name__", "host_function")
    return wrapper


class CudaSystem(object):

    def __init__(self, compute_module=cuda_compute, device=None, rank=0, world_size=1,
                 device_grid=None, placement="flat"):
        self.compute_module = compute_module
        self.compute_imp = compute_module.ComputeCls
        if getattr(compute_module, "RNG", None) is None:
            raise Exception("No random number generator implemented for compute module %s" % compute_module)
        self.rng_cls = compute_module.RNG
        self.methods = {}
        self.remote_functions = {}
        self.rank = int(rank)
        self.world_size = int(world_size)
        self.device_grid = tuple(device_grid) if device_grid is not None else (self.world_size, 1)
        self.placement = placement
        self._device = device
        self._imp = None
        # float64 tensordot / add chains are deferred and run as one grouped launch (deferred.py)
        self.contractions = ContractionQueue()

    # -- lifecycle ---------------------------------------------------------------------------
    def init(self):
        if self._device is None and torch.cuda.is_available():
            self._device = torch.device("cuda", torch.cuda.current_device())
        if self._device is not None and self._device.type == "cuda":
            torch.cuda.set_device(self._device)
        self._imp = self.compute_imp()
        for name, fn in inspect.getmembers(self._imp, predicate=inspect.ismethod):
            if name.startswith("_"):
                continue
            self.remote_functions[name] = fn
            self._publish(name, self._make_callable(name))
        if "bop" in self.remote_functions:
            self._bop_kernel = self.remote_functions["bop"]
            self._publish("bop", self._call_bop)          # the per-block hot call: a leaner copy of ``call``
        # optional kernels beyond the 28 interface methods (fused logistic-regression step)
        for name, fn in getattr(self.compute_module, "EXTRA_KERNELS", {}).items():
            self.remote_functions[name] = fn
            self._publish(name, self._make_callable(name))

    def shutdown(self):
        # like SerialSystem.shutdown (systems.py:91-92) this leaves the system usable: the reference keeps
        # module-level objects (e.g. the nums.numpy.random state) that outlive an application
        self.contractions.flush()

    def _make_callable(self, name):
        def kernel(*args, **kwargs):
            return self.call(name, *args, **kwargs)
        kernel.__name__ = name
        return kernel

    def _publish(self, name, fn):
        """Make ``system.<name>`` resolve to ``fn``: in ``methods`` (where the reference's System looks,
        systems.py:62-66) AND as an instance attribute, which shadows the abstract stub of the same name that
        the reference's ``System`` inherits from ``ComputeInterface`` -- so that reference_compat's subclasses can
        use the plain C attribute lookup instead of the reference's Python-level ``__getattribute__``."""
        self.methods[name] = fn
        owner = next((k for k in type(self).__mro__ if name in k.__dict__), None)
        if owner is None or owner.__name__ == "ComputeInterface":      # never shadow the system's own API
            self.__dict__[name] = fn

    def __getattr__(self, name):
        # only reached when normal lookup fails: kernel names are routed to ``call``
        methods = self.__dict__.get("methods", {})
        if name in methods:
            return methods[name]
        raise AttributeError(name)

    # -- object store ----------------------------------------------------------------------------
    def put(self, value):
        return cuda_compute.upload(value)

    def get(self, object_ids):
        object_ids = self.contractions.resolve(object_ids)
        if isinstance(object_ids, list):
            if all(isinstance(o, torch.Tensor) for o in object_ids):
                return cuda_compute.download_many(object_ids)
            if object_ids and all(isinstance(o, cuda_compute.Touched) for o in object_ids):
                torch.cuda.current_stream().synchronize()     # one wait for a whole BlockArray.touch()
                cuda_compute.check_pending_status()
                return [o.ok for o in object_ids]
            return [self.get(o) for o in object_ids]
        if isinstance(object_ids, tuple):
            return tuple(self.get(o) for o in object_ids)
        return cuda_compute.download(object_ids)

    def get_async(self, oid):
        """Start fetching a (small) block without blocking: returns a callable that waits for the copy and
        gives the NumPy array.  Lets a driver loop read a convergence flag one iteration late, with the next
        iteration already enqueued (nums_b200.glms_fused.newton)."""
        t = self.contractions.resolve(oid)
        if not isinstance(t, torch.Tensor):
            return lambda: t
        host = cuda_compute._to_pinned(t)
        done = torch.cuda.Event()
        done.record()

        def wait():
            done.synchronize()
            cuda_compute.check_pending_status()
            return host.numpy()
        return wait

    def get_assembled(self, grid, oids_by_entry):
        """Whole array of a block grid as one NumPy array (BlockArrayBase.get, base.py:348-360).

        The blocks are placed into one device buffer with the strided-copy kernel and come back as
        page-locked D2H transfers of whole block rows (contiguous in the C-ordered result) instead of
        one transfer per block plus a host-side re-assembly.  Assembly and transfers run on the
        download stream, each block row as soon as the launch group that produced its blocks has
        finished (``_nums_done`` events set by the deferred-contraction flush), so the D2H of early
        rows overlaps the GEMM of later ones."""
        cc = cuda_compute
        dtype = np.dtype(grid.dtype)
        tdtype = cuda_compute._lib.torch_dtype(dtype)
        shape = tuple(int(x) for x in grid.shape)
        if len(shape) == 0 or int(np.prod(shape)) * dtype.itemsize < (1 << 20):
            full = cc._empty(shape, dtype)
            for entry, oid in oids_by_entry:
                block = self.contractions.resolve(oid)
                view = full[grid.get_slice(entry)]
                cc._copy_into(view, block.reshape(view.shape) if tuple(block.shape) != tuple(view.shape) else block)
            return cc.download(full)
        blocks = [(entry, self.contractions.resolve(oid)) for entry, oid in oids_by_entry]
        home = torch.cuda.current_stream()
        cc.await_uploads(home)
        tail = None                      # producers without their own event: everything queued so far
        down = cc._download_stream()
        host = torch.empty(shape, dtype=tdtype, pin_memory=True)
        rows = {}
        for entry, block in blocks:
            rows.setdefault(entry[0], []).append((entry, block))
        with torch.cuda.stream(down):
            full = torch.empty(shape, dtype=tdtype, device=cc._device())
            for r in sorted(rows):
                lo = hi = None
                for entry, block in rows[r]:
                    done = getattr(block, "_nums_done", None)
                    if done is None:
                        if tail is None:
                            tail = torch.cuda.Event()
                            tail.record(home)
                        done = tail
                    down.wait_event(done)
                    block.record_stream(down)
                    sl = grid.get_slice(entry)
                    view = full[sl]
                    cc._copy_into(view, block.reshape(view.shape) if tuple(block.shape) != tuple(view.shape) else block)
                    lo, hi = sl[0].start, sl[0].stop
                host[lo:hi].copy_(full[lo:hi], non_blocking=True)
        down.synchronize()
        return host.numpy()

    def remote(self, function, remote_params):
        return function

    def register(self, name, func, remote_params=None):
        """``System.register`` (systems.py:100-105).  The reference's ``FileSystem`` registers its
        block-level I/O functions here (filesystem.py:224-231); those written against NumPy blocks are
        replaced by the device-aware versions the compute module provides (``DEVICE_FUNCTIONS``:
        block persistence and the CSV parser), everything else (metadata, S3, ``loadtxt_block``) runs
        on the host with blocks downloaded on the way in and array results uploaded on the way out."""
        if name in self.remote_functions:
            return
        device_fn = getattr(self.compute_module, "DEVICE_FUNCTIONS", {}).get(name)
        self.remote_functions[name] = device_fn if device_fn is not None else _host_function(func)
        self._publish(name, self._make_callable(name))

    def nodes(self):
        return [{"Resources": {"node:%d" % r: 1.0}} for r in range(self.world_size)]

    def get_rng(self, seed):
        return self.rng_cls(seed)

    # -- dispatch -----------------------------------------------------------------------------------
    def call(self, name, *args, **kwargs):
        kwargs.pop("syskwargs", None)        # (**kwargs is a fresh dict per call)
        q = self.contractions
        if name == "bop" and q.enabled:
            op = args[0] if args else kwargs.get("op")
            if op == "tensordot" or (op == "add" and q._pending):
                lazy = self._try_defer(args, kwargs)
                if lazy is not None:
                    return lazy
        if q._pending or q.materialized_seen:
            if name == "qr":         # the stacked-R call consumes DeferredR handles as they are (cuda_compute.DeferredR)
                args = tuple(a if a.__class__ is cuda_compute.DeferredR else q.resolve(a) for a in args)
            else:
                args = tuple(q.resolve(a) for a in args)
            kwargs = {k: q.resolve(v) for k, v in kwargs.items()}
        out = self.remote_functions[name](*args, **kwargs)
        if out.__class__ is cuda_compute.DeferredR:
            q.materialized_seen = True
        return out

    def _call_bop(self, op, a1, a2, a1_shape, a2_shape, a1_T, a2_T, axes=None, syskwargs=None):
        """``call("bop", ...)`` without the generic argument handling (Block.bop, base.py:220-231, calls with
        exactly these seven positionals + ``axes`` + ``syskwargs``)."""
        q = self.contractions
        if q.enabled:
            if op == "tensordot":
                if a1.__class__ is not DeferredContraction and a2.__class__ is not DeferredContraction:
                    lazy = q.tensordot(a1, a2, a1_shape, a2_shape, a1_T, a2_T, axes)
                    if lazy is not None:
                        return lazy
            elif op == "add" and q._pending and (a1.__class__ is DeferredContraction or a2.__class__ is DeferredContraction):
                lazy = q.add(a1, a2, a1_shape, a2_shape, a1_T, a2_T)
                if lazy is not None:
                    return lazy
        if a1.__class__ is not torch.Tensor:          # deferred contraction / deferred R factor / host value
            a1 = q.resolve(a1)
        if a2.__class__ is not torch.Tensor:
            a2 = q.resolve(a2)
        return self._bop_kernel(op, a1, a2, a1_shape, a2_shape, a1_T, a2_T, axes)

    def _try_defer(self, args, kwargs):
        """tensordot -> DeferredContraction; add of deferred contractions -> longer term list."""
        bound = dict(zip(_BOP_PARAMS, args))
        bound.update(kwargs)
        if len(bound) != len(_BOP_PARAMS):
            return None
        op, a1, a2 = bound["op"], bound["a1"], bound["a2"]
        q = self.contractions
        if op == "tensordot":
            if isinstance(a1, DeferredContraction) or isinstance(a2, DeferredContraction):
                return None
            return q.tensordot(a1, a2, bound["a1_shape"], bound["a2_shape"], bound["a1_T"], bound["a2_T"], bound["axes"])
        if op == "add" and (isinstance(a1, DeferredContraction) or isinstance(a2, DeferredContraction)):
            return q.add(a1, a2, bound["a1_shape"], bound["a2_shape"], bound["a1_T"], bound["a2_T"])
        return None

    def flush(self):
        """Launch every deferred contraction now (used between SUMMA steps)."""
        self.contractions.flush()

    def call_with_options(self, name, args, kwargs, options):
        return self.call(name, *args, **kwargs)

    def get_options(self, cluster_entry, cluster_shape):
        return {"resources": {"node:%d" % self.rank: 1.0 / 10 ** 4}}

    def get_block_addresses(self, grid):
        return {entry: "node:%d" % self.owner(entry, grid.grid_shape) for entry in grid.get_entry_iterator()}

    # -- placement ------------------------------------------------------------------------------------
    def owner(self, grid_entry, grid_shape):
        """Rank that owns / computes the block at ``grid_entry`` of a grid of ``grid_shape``."""
        if self.world_size == 1 or len(grid_entry) == 0:
            return 0
        if self.placement == "flat":
            flat = int(np.ravel_multi_index(tuple(grid_entry), tuple(grid_shape)))
            return flat % self.world_size
        dg = self.device_grid
        coords = [grid_entry[i] % dg[i] if i < len(grid_entry) else 0 for i in range(len(dg))]
        return int(np.ravel_multi_index(tuple(coords), dg))

    def synchronize(self):
        self.contractions.flush()
        torch.cuda.current_stream().synchronize()
        cuda_compute.check_pending_status()
