"""nums_b200 -- a B200-native (sm_100a) compute backend for NumS.

``cuda_compute``  : drop-in for ``nums.core.systems.numpy_compute`` (ComputeCls + RNG)
``cuda_system``   : the System object that places blocks on the GPU (put / get / call)
``blocks``        : the minimal block-array host layer used by tests and ``bench.py``
``libnumscuda.so``: hand-written CUDA kernels behind a C ABI (``include/nums_cuda.h``)
"""
__version__ = "0.1.0"
