"""Block-grid geometry: which slice of an array each grid entry covers.

Host-side mirror of the reference's ``ArrayGrid`` / ``Batch``
(/root/reference/nums/core/storage/storage.py:29-86, storage/utils.py:23-62); the kernels
``empty`` / ``new_block`` receive ``grid.to_meta()`` dicts exactly like the reference's do
(numpy_compute.py:91-104).
"""
import itertools

import numpy as np


def axis_batches(dim, block_dim):
    """[start, stop) pairs along one axis (storage/utils.py:45-62)."""
    if dim < block_dim:
        return [(0, dim)]
    edges = list(range(0, dim, block_dim)) + [dim]
    return [(lo, hi) for lo, hi in zip(edges[:-1], edges[1:]) if lo < hi]


class ArrayGrid(object):

    @classmethod
    def from_meta(cls, d):
        return cls(**d)

    def __init__(self, shape, block_shape, dtype):
        self.shape = tuple(int(s) for s in shape)
        if len(self.shape) != len(block_shape):
            raise ValueError("shape %s and block_shape %s differ in rank" % (shape, block_shape))
        self.block_shape = tuple(int(min(s, b)) for s, b in zip(self.shape, block_shape))
        if isinstance(dtype, str):
            dtype = {"int": np.int64, "float": np.float64, "bool": np.bool_}.get(dtype) or getattr(np, dtype)
        self.dtype = np.dtype(dtype).type
        self.grid_slices = []
        for dim, bdim in zip(self.shape, block_shape):
            self.grid_slices.append([] if dim == 0 else axis_batches(dim, int(bdim)))
        self.grid_shape = tuple(len(s) for s in self.grid_slices)

    def to_meta(self):
        return {"shape": self.shape, "block_shape": self.block_shape, "dtype": self.dtype.__name__}

    def copy(self):
        return self.from_meta(self.to_meta())

    def get_entry_iterator(self):
        if 0 in self.shape:
            return iter(())
        return itertools.product(*map(range, self.grid_shape))

    def get_slice_tuples(self, grid_entry):
        return [tuple(self.grid_slices[axis][i]) for axis, i in enumerate(grid_entry)]

    def get_slice(self, grid_entry):
        return tuple(slice(lo, hi) for lo, hi in self.get_slice_tuples(grid_entry))

    def get_block_shape(self, grid_entry):
        return tuple(hi - lo for lo, hi in self.get_slice_tuples(grid_entry))
