"""Which host layers drive the kernels: the reference's own, or the built-in mirror.

The product is the per-block compute back end; the layers above it (``BlockArray``,
``ArrayApplication``, ``nums.models.glms``) are the reference's and run unchanged
(nums_b200.reference_compat).  ``HostLayers`` hands bench.py / smoke() those classes when a NumS
installation is present (``kind == "reference"``: /root/reference, baseline/_ref, or an installed
``nums``) and otherwise falls back to ``nums_b200.blocks`` (``kind == "mirror"``), the call-for-call
restatement of the hot-path operations that tests/test_call_trace.py pins against the reference.
"""
import numpy as np

from nums_b200 import reference_compat


class HostLayers(object):

    def __init__(self, system=None, prefer_reference=True, **system_kwargs):
        self.kind = "mirror"
        if prefer_reference and reference_compat.available():
            try:
                reference_compat.load_reference()
                self.kind = "reference"
            except Exception:  # noqa: BLE001 -- an unusable installation selects the mirror
                self.kind = "mirror"
        if self.kind == "reference":
            from nums.core.array.application import ArrayApplication
            from nums.core.array.blockarray import BlockArray
            from nums.core.storage.storage import ArrayGrid
            from nums.core.systems.filesystem import FileSystem
            if system is None:
                system = reference_compat.cuda_system(**system_kwargs)    # SpmdSystem under torchrun
            self.system = system
            self.app = ArrayApplication(system=system, filesystem=FileSystem(system))
            self.BlockArray, self.ArrayGrid = BlockArray, ArrayGrid
        else:
            from nums_b200 import blocks
            from nums_b200.grid import ArrayGrid
            if system is None:
                from nums_b200.cuda_system import CudaSystem
                system = CudaSystem(**system_kwargs)
                system.init()
            self.system = system
            self.app = blocks.ArrayApp(system)
            self.BlockArray, self.ArrayGrid = blocks.BlockArray, ArrayGrid

    @property
    def description(self):
        if self.kind == "reference":
            return "the reference's unmodified BlockArray / ArrayApplication / glms (%s)" % reference_compat.reference_root()
        return "nums_b200.blocks (call-for-call mirror of the reference's host layers; no NumS installation found)"

    # -- construction ---------------------------------------------------------------------------
    def blockarray(self, shape, block_shape, dtype="float64"):
        """A BlockArray whose blocks have no oid yet (the caller assigns ``ba.blocks[entry].oid``)."""
        return self.BlockArray(self.ArrayGrid(tuple(shape), tuple(block_shape), dtype), self.system)

    def from_blocks(self, shape, block_shape, fill, dtype="float64"):
        ba = self.blockarray(shape, block_shape, dtype)
        for entry in ba.grid.get_entry_iterator():
            ba.blocks[entry].oid = fill(entry, ba.grid.get_block_shape(entry))
        return ba

    def array(self, arr, block_shape):
        return self.app.array(arr, block_shape)

    # -- materialisation --------------------------------------------------------------------------
    def sync(self, ba=None):
        """Launch whatever ``ba`` still defers and wait for the device (this fork's ``BlockArray.touch``
        does not wait: its ``system.get(oids)`` is commented out, blockarray.py:117-126)."""
        if ba is not None:
            ba.touch()
        self.system.synchronize()

    def launch(self, ba):
        """Enqueue everything ``ba`` depends on without waiting."""
        ba.touch()
        self.system.flush()

    def get(self, ba):
        """Whole array on the host through ``CudaSystem.get_assembled`` (device-side assembly, block rows
        drained while later ones compute) -- same values as ``ba.get()``."""
        entries = list(ba.grid.get_entry_iterator())
        if len(entries) > 1 and hasattr(self.system, "get_assembled"):
            return self.system.get_assembled(ba.grid, [(e, ba.blocks[e].oid) for e in entries])
        return ba.get()

    # -- workloads -------------------------------------------------------------------------------
    def logistic_model(self):
        if self.kind == "reference":
            from nums.core import application_manager
            from nums.models.glms import LogisticRegression
            if not application_manager.is_initialized():
                application_manager.set_instance(self.app)     # GLM.__init__ asks for the global application
            model = LogisticRegression(solver="newton", penalty="none")
            model._app = self.app
            return model
        from nums_b200 import blocks
        return blocks.LogisticRegression(self.app)

    def newton(self, model, X, y, tol, max_iter):
        """glms.newton (glms.py:362-372) from beta = 0; returns beta (BlockArray)."""
        d = X.shape[1]
        beta = self.app.zeros((d,), (d,), np.float64)
        tol = self.app.scalar(tol)
        if self.kind == "reference":
            from nums.models.glms import newton
            return newton(self.app, model, beta, X, y, tol, max_iter)
        from nums_b200 import blocks
        return blocks.newton(self.app, model, beta, X, y, tol, max_iter)[0]
