import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# a flush that materialises a dangling partial dot/add chain is an error in the tests (nums_b200/deferred.py)
os.environ.setdefault("NUMS_DEFERRED_STRICT", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the tests exercise the built artefact: (re)build libnumscuda.so if it is missing (nvcc needed)
    from nums_b200 import _build
    if not os.path.exists(_build.LIB):
        _build.build()


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def cuda_system():
    from nums_b200.cuda_system import CudaSystem
    system = CudaSystem()
    system.init()
    return system


@pytest.fixture(scope="session")
def oracle():
    from oracle.np_oracle import OracleCompute
    return OracleCompute()
