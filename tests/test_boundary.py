"""CPU checks of the drop-in boundary: the C ABI library, the interface surface, no CPU fallback."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import ref_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "nums_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nums_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from nums_b200._lib import LIB, LIB_PATH
    assert os.path.exists(LIB_PATH), "build it first: python -c 'import __graft_entry__ as g; g.build()'"
    dll = ctypes.CDLL(LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 22
    for name in names:
        assert hasattr(dll, name), "%s declared in include/nums_cuda.h but not exported" % name
    assert sorted(LIB.EXPORTS) == names          # the ctypes binding covers the whole header
    assert LIB.dll.nums_abi_version() == 1       # loading binds argtypes for every entry point
    assert LIB.dll.nums_launch_count() == 0      # nothing ran: no GPU here


def test_compute_cls_has_the_28_interface_methods():
    from nums_b200 import cuda_compute
    names = ("touch empty new_block random_block permutation diag arange sum_reduce transpose create_block "
             "update_block update_block_by_index update_block_along_axis bop split qr cholesky svd inv allclose "
             "map_uop where reduce_axis xlogy logical_and astype arg_op reshape").split()
    imp = cuda_compute.ComputeCls()          # instantiable with no arguments (systems/utils.py:50)
    for name in names:
        assert callable(getattr(imp, name)), name
    assert len(names) == 28
    rng = cuda_compute.RNG(7)
    assert rng.new_block_rng_params() == (7, 0) and rng.new_block_rng_params() == (7, 1)


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present on this box")
def test_reference_accepts_cuda_compute_as_a_compute_module():
    ref_loader.load()
    from nums.core.systems.interfaces import ComputeImp, ComputeInterface
    from nums.core.systems.utils import check_implementation, extract_functions
    import importlib
    import nums_b200.cuda_compute as cc
    cc = importlib.reload(cc)                 # pick up the real ComputeImp base now that nums imports
    check_implementation(ComputeInterface, cc.ComputeCls)        # systems/utils.py:59-72
    assert issubclass(cc.ComputeCls, ComputeImp)
    assert set(extract_functions(cc.ComputeCls)) >= {"bop", "qr", "reduce_axis"}
    # ... and the reference's own System base class wires it up (systems.py:34-49)
    from nums.core.systems.systems import SerialSystem
    system = SerialSystem(compute_module=cc)
    system.init()
    assert "bop" in system.methods and system.get_rng(3).seed == 3


def test_no_cpu_fallback_and_no_oracle_in_the_product():
    from nums_b200 import _lib
    from nums_b200.cuda_system import CudaSystem
    system = CudaSystem()
    system.init()
    if not torch.cuda.is_available():
        with pytest.raises(_lib.NumsCudaError):
            system.put(np.ones(4))
        with pytest.raises(_lib.NumsCudaError):
            system.new_block("zeros", (0,), {"shape": (4,), "block_shape": (4,), "dtype": "float64"}, syskwargs={})
    pkg = os.path.join(ROOT, "nums_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn
    with pytest.raises(NotImplementedError):
        _lib.dtype_code(np.float16)


def test_host_side_rules():
    from nums_b200 import cuda_compute as cc
    from nums_b200._lib import describe
    from nums_b200.grid import ArrayGrid
    assert cc._broadcast_shape((6, 1), (8,)) == (6, 8)
    assert cc._broadcast_shape((), (3, 1)) == (3, 1)
    with pytest.raises(ValueError):
        cc._broadcast_shape((3,), (4,))
    f8, i8, f4, b1 = (np.dtype(x) for x in ("f8", "i8", "f4", "?"))
    assert cc.bop_types("add", f8, i8) == (f8, f8)
    assert cc.bop_types("true_divide", i8, i8) == (f8, f8)
    assert cc.bop_types("less", f8, f4) == (f8, b1)
    assert cc.bop_types("xlogy", f8, f8) == (f8, f8)            # scipy.special fallback (numpy_compute.py:234-238)
    assert cc.uop_types("sqrt", i8) == (f8, f8)
    assert cc.reduce_type("sum", b1) == i8
    t = torch.zeros((3, 5), dtype=torch.float64).t()
    d = describe(t)
    assert (d.ndim, d.shape[0], d.shape[1], d.stride[0], d.stride[1]) == (2, 5, 3, 1, 5)
    g = ArrayGrid((2345, 9), (123, 9), "float64")
    assert g.grid_shape == (20, 1) and g.get_block_shape((19, 0)) == (8, 9)
    assert g.to_meta() == {"shape": (2345, 9), "block_shape": (123, 9), "dtype": "float64"}
    if ref_loader.available():
        ref_loader.load()
        from nums.core.storage.storage import ArrayGrid as RefGrid
        for shape, bs in [((2345, 9), (123, 9)), ((10,), (3,)), ((7, 7, 7), (2, 7, 3)), ((5,), (9,))]:
            a, b = ArrayGrid(shape, bs, "int64"), RefGrid(shape, bs, "int64")
            assert a.grid_shape == b.grid_shape
            for e in a.get_entry_iterator():
                assert a.get_block_shape(e) == tuple(b.get_block_shape(e))
                assert a.get_slice(e) == b.get_slice(e)
