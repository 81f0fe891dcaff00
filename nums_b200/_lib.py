"""ctypes binding of ``libnumscuda.so`` (C ABI in ``include/nums_cuda.h``).

There is deliberately no fallback: if the shared library is missing or a call fails, an
exception is raised.  PyTorch is used only to own device memory and streams; the pointers
handed to the library are ``tensor.data_ptr()`` values.
"""
import ctypes
import os

import numpy as np
import torch

MAX_DIMS = 8
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libnumscuda.so")

# nums_dtype_t
BOOL, I32, I64, F32, F64 = 0, 1, 2, 3, 4
_TORCH_TO_CODE = {torch.bool: BOOL, torch.int32: I32, torch.int64: I64,
                  torch.float32: F32, torch.float64: F64}
_NP_TO_CODE = {np.dtype(np.bool_): BOOL, np.dtype(np.int32): I32, np.dtype(np.int64): I64,
               np.dtype(np.float32): F32, np.dtype(np.float64): F64}
_CODE_TO_TORCH = {v: k for k, v in _TORCH_TO_CODE.items()}
_NP_TO_TORCH = {k: _CODE_TO_TORCH[v] for k, v in _NP_TO_CODE.items()}
_TORCH_TO_NP = {v: k for k, v in _NP_TO_TORCH.items()}

BOPS = ("add subtract multiply true_divide floor_divide remainder fmod power float_power maximum "
        "minimum fmax fmin arctan2 hypot copysign nextafter heaviside logaddexp logaddexp2 ldexp "
        "xlogy less less_equal greater greater_equal equal not_equal logical_and logical_or "
        "logical_xor bitwise_and bitwise_or bitwise_xor left_shift right_shift gcd lcm").split()
BOP_CODE = {name: i for i, name in enumerate(BOPS)}
BOP_CODE["divide"] = BOP_CODE["true_divide"]
BOP_CODE["mod"] = BOP_CODE["remainder"]

UOPS = ("copy absolute negative positive sign sqrt cbrt square reciprocal exp exp2 expm1 log log2 "
        "log10 log1p sin cos tan arcsin arccos arctan sinh cosh tanh arcsinh arccosh arctanh floor "
        "ceil trunc rint deg2rad rad2deg spacing isnan isinf isfinite signbit logical_not "
        "invert").split()
UOP_CODE = {name: i for i, name in enumerate(UOPS)}
UOP_CODE.update({"abs": UOP_CODE["absolute"], "fabs": UOP_CODE["absolute"],
                 "radians": UOP_CODE["deg2rad"], "degrees": UOP_CODE["rad2deg"],
                 "bitwise_not": UOP_CODE["invert"], "conjugate": UOP_CODE["positive"],
                 "conj": UOP_CODE["positive"]})

REDUCE_CODE = {"sum": 0, "prod": 1, "product": 1, "min": 2, "amin": 2, "max": 3, "amax": 3,
               "any": 4, "all": 5}

ERR_WORKSPACE = -4


class NumsCudaError(RuntimeError):
    pass


class NumsArray(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("dtype", ctypes.c_int32), ("ndim", ctypes.c_int32),
                ("shape", ctypes.c_int64 * MAX_DIMS), ("stride", ctypes.c_int64 * MAX_DIMS)]


class GemmTerm(ctypes.Structure):       # nums_gemm_term_t
    _fields_ = [("A", ctypes.c_void_p), ("B", ctypes.c_void_p), ("lda", ctypes.c_int64),
                ("ldb", ctypes.c_int64), ("k", ctypes.c_int64)]


class GemmProblem(ctypes.Structure):    # nums_gemm_problem_t
    _fields_ = [("C", ctypes.c_void_p), ("Cin", ctypes.c_void_p), ("ldc", ctypes.c_int64),
                ("ldcin", ctypes.c_int64), ("m", ctypes.c_int64), ("n", ctypes.c_int64),
                ("term_begin", ctypes.c_int32), ("term_count", ctypes.c_int32)]


def dtype_code(dt):
    """torch / numpy dtype -> nums_dtype_t; raises for dtypes the library has no loops for."""
    if isinstance(dt, torch.dtype):
        code = _TORCH_TO_CODE.get(dt)
    else:
        code = _NP_TO_CODE.get(np.dtype(dt))
    if code is None:
        raise NotImplementedError("libnumscuda has no kernels for dtype %s "
                                  "(supported: bool, int32, int64, float32, float64)" % (dt,))
    return code


def supported_dtype(dt):
    try:
        return np.dtype(dt) in _NP_TO_CODE
    except TypeError:
        return False


def torch_dtype(dt):
    if isinstance(dt, torch.dtype):
        return dt
    t = _NP_TO_TORCH.get(np.dtype(dt))
    if t is None:
        raise NotImplementedError("unsupported dtype %s" % (dt,))
    return t


def numpy_dtype(dt):
    if isinstance(dt, torch.dtype):
        return _TORCH_TO_NP[dt]
    return np.dtype(dt)


def describe(t, shape=None, strides=None):
    """Build a ``nums_array_t`` for a torch tensor (optionally viewed with other shape/strides)."""
    shape = tuple(t.shape) if shape is None else tuple(shape)
    strides = tuple(t.stride()) if strides is None else tuple(strides)
    if len(shape) > MAX_DIMS:
        raise NotImplementedError("arrays with more than %d axes are not supported" % MAX_DIMS)
    a = NumsArray()
    a.data = t.data_ptr()
    a.dtype = _TORCH_TO_CODE[t.dtype] if t.dtype in _TORCH_TO_CODE else dtype_code(t.dtype)
    a.ndim = len(shape)
    for i, (s, st) in enumerate(zip(shape, strides)):
        a.shape[i] = s
        a.stride[i] = st
    return a


class _Lib(object):
    """Lazy singleton around the shared library."""

    def __init__(self):
        self._dll = None
        self._ws = {}

    @property
    def dll(self):
        if self._dll is None:
            if not os.path.exists(LIB_PATH):
                raise NumsCudaError(
                    "%s is missing: build it with `python -m nums_b200._build` (nvcc, sm_100a). "
                    "There is no CPU fallback." % LIB_PATH)
            dll = ctypes.CDLL(LIB_PATH)
            self._declare(dll)
            if dll.nums_abi_version() != 1:
                raise NumsCudaError("libnumscuda ABI version mismatch")
            self._dll = dll
        return self._dll

    @staticmethod
    def _declare(dll):
        c = ctypes
        P, I, L, D, Z = c.c_void_p, c.c_int, c.c_int64, c.c_double, c.c_size_t
        A = c.POINTER(NumsArray)
        sig = {
            "nums_abi_version": ([], c.c_int),
            "nums_last_error": ([], c.c_char_p),
            "nums_last_workspace_request": ([], Z),
            "nums_sm_count": ([], c.c_int),
            "nums_launch_count": ([], c.c_uint64),
            "nums_bop": ([I, I, A, A, A, P], I),
            "nums_bop_flat": ([I, I, P, I, L, P, I, L, P, I, L, P], I),
            "nums_uop": ([I, I, A, A, P], I),
            "nums_sum_reduce": ([I, c.POINTER(P), I, L, P, P], I),
            "nums_fill": ([A, D, P], I),
            "nums_arange": ([A, D, D, P], I),
            "nums_eye": ([A, P], I),
            "nums_reduce": ([I, P, I, L, L, L, P, I, P, Z, P], I),
            "nums_arg_op": ([I, P, I, L, L, P, P, P, P, P, Z, P], I),
            "nums_allclose": ([P, P, I, L, D, D, P, P, Z, P], I),
            "nums_nonzero_count": ([P, I, L, P, P, Z, P], I),
            "nums_nonzero_fill": ([P, I, I, c.POINTER(L), c.POINTER(L), c.POINTER(P), P, Z, P], I),
            "nums_gemm": ([I, I, I, L, L, L, P, L, P, L, P, L, I, P, Z, P], I),
            "nums_gemm_grouped": ([I, I, I, I, c.POINTER(GemmProblem), I, c.POINTER(GemmTerm), P, Z, P], I),
            "nums_qr": ([I, L, L, P, L, P, L, P, L, P, Z, P], I),
            "nums_inv": ([I, L, P, L, P, L, P, P, Z, P], I),
            "nums_cholesky": ([I, L, P, L, P, L, P, P, Z, P], I),
            "nums_gram_factor": ([L, P, L, P, L, P, L, P, L, P, P], I),
            "nums_newton_step": ([L, P, P, P, P, P], I),
            "nums_scatter_axis": ([I, L, L, L, L, L, P, P, P, P, P], I),
            "nums_csv_index": ([P, L, L, I, P, P, Z, P], I),
            "nums_csv_parse": ([P, L, L, I, I, L, L, P, P, P, P], I),
            "nums_svd": ([I, L, P, L, P, P, P, P, Z, P], I),
            "nums_lr_grad_hess": ([L, L, P, L, P, P, P, P, Z, P], I),
            "nums_lr_grad_hess_blocks": ([I, c.POINTER(P), c.POINTER(P), c.POINTER(L), L, P, P, P, Z, P], I),
        }
        for name, (argtypes, restype) in sig.items():
            fn = getattr(dll, name)  # AttributeError if the symbol is not exported
            fn.argtypes = argtypes
            fn.restype = restype

    EXPORTS = ("nums_abi_version nums_last_error nums_last_workspace_request nums_sm_count nums_launch_count nums_bop "
               "nums_bop_flat nums_uop nums_sum_reduce nums_fill nums_arange nums_eye nums_reduce nums_arg_op "
               "nums_allclose nums_nonzero_count nums_nonzero_fill nums_gemm nums_gemm_grouped nums_qr nums_inv "
               "nums_cholesky nums_gram_factor nums_svd nums_lr_grad_hess nums_lr_grad_hess_blocks nums_newton_step nums_scatter_axis nums_csv_index "
               "nums_csv_parse").split()

    # -- error handling ---------------------------------------------------------------------
    def check(self, rc):
        if rc == 0:
            return
        msg = self.dll.nums_last_error().decode("utf-8", "replace")
        if rc == -2:
            raise NotImplementedError("libnumscuda: " + msg)
        if rc == -1:
            raise ValueError("libnumscuda: " + msg)
        raise NumsCudaError("libnumscuda error %d: %s" % (rc, msg))

    # -- workspace ------------------------------------------------------------------------------
    def workspace(self, device, min_bytes=0):
        """Per-device scratch buffer (uint8 torch tensor), grown on demand."""
        key = (device.type, device.index)
        ws = self._ws.get(key)
        if ws is None or ws.numel() < min_bytes:
            size = max(int(min_bytes), 64 << 20)
            ws = torch.empty(size, dtype=torch.uint8, device=device)
            self._ws[key] = ws
        return ws

    def call_ws(self, fn, device, args_before_ws_and_after):
        """Call ``fn(*before, ws, ws_bytes, *after)``; retry once with a larger workspace."""
        before, after = args_before_ws_and_after
        ws = self.workspace(device)
        rc = fn(*before, ws.data_ptr(), ws.numel(), *after)
        if rc == ERR_WORKSPACE:
            need = int(self.dll.nums_last_workspace_request())
            ws = self.workspace(device, need)
            rc = fn(*before, ws.data_ptr(), ws.numel(), *after)
        self.check(rc)


LIB = _Lib()


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream
