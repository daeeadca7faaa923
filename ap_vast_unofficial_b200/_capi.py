"""ctypes binding of the C-ABI in ``include/apvast_b200.h`` (libapvast_b200.so, sm_100a).

There is no CPU fallback: if the shared library is missing or no CUDA device is present the
engine fails loudly (``ImportError`` / ``RuntimeError``).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libapvast_b200.so")

# status codes (enum apv_status)
OK, EINVAL, ENOTPD, ECUDA, ENOMEM, ENOCONV, ENCCL = range(7)

# tensor ids (enum apv_tensor)
(T_W, T_LAMBDA, T_U, T_R, T_RVEC, T_WEIGHT, T_RESP, T_RESP_T, T_OLA, T_OLA_T, T_STATS, T_STATS_T,
 T_OUT_OLA, T_OUT_OLA_T, T_INPUT, T_TARGET_FRAME) = range(16)


class Config(C.Structure):
    _fields_ = [
        ("block_size", C.c_int32), ("hop_size", C.c_int32), ("rir_length", C.c_int32),
        ("n_srcs", C.c_int32), ("n_mics", C.c_int32), ("filter_length", C.c_int32),
        ("stats_length", C.c_int32), ("n_eig", C.c_int32), ("modeling_delay", C.c_int32),
        ("ref_A", C.c_int32), ("ref_B", C.c_int32), ("run_A", C.c_int32), ("run_B", C.c_int32),
        ("perceptual", C.c_int32), ("normalize_gains", C.c_int32), ("eig_mode", C.c_int32),
        ("stats_mode", C.c_int32), ("device", C.c_int32),
        ("mu", C.c_double), ("reg", C.c_double), ("sampling_rate", C.c_double),
        ("toeplitz_clean", C.c_int32), ("normalize_stats", C.c_int32), ("loading_mode", C.c_int32),
        ("target_ref_per_zone", C.c_int32), ("bright_load", C.c_double), ("dark_load", C.c_double),
        ("active_mics_A", C.c_int32), ("reg_relative", C.c_int32),
    ]


_dp = C.POINTER(C.c_double)
_lib = None

# name -> (restype, argtypes): every entry point include/apvast_b200.h declares
SIGNATURES = {
    "apv_tensor_size": (C.c_size_t, [C.c_void_p, C.c_int]),
    "apv_create": (C.c_int, [C.POINTER(Config), _dp, _dp, _dp, C.POINTER(C.c_void_p)]),
    "apv_destroy": (None, [C.c_void_p]),
    "apv_process_block": (C.c_int, [C.c_void_p, _dp, _dp, _dp, _dp, _dp, _dp]),
    "apv_process_blocks": (C.c_int, [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp]),
    "apv_process_block_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "apv_begin_block": (C.c_int, [C.c_void_p, _dp, _dp]),
    "apv_finish_block": (C.c_int, [C.c_void_p, _dp, _dp, _dp, _dp]),
    "apv_advance_state": (C.c_int, [C.c_void_p, _dp, _dp]),
    "apv_set_pipeline": (C.c_int, [C.c_void_p, C.c_int]),
    "apv_set_depth": (C.c_int, [C.c_void_p, C.c_int]),
    "apv_debug_timeline": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float)]),
    "apv_set_reg_mode": (C.c_int, [C.c_void_p, C.c_int]),
    "apv_copy_weights": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "apv_comm_unique_id": (C.c_int, [C.c_void_p]),
    "apv_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "apv_comm_destroy": (C.c_int, [C.c_void_p]),
    "apv_nccl_version": (C.c_int, [C.POINTER(C.c_int)]),
    "apv_range_reserve": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "apv_range_run": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]),
    "apv_range_exchange_halo": (C.c_int, [C.c_void_p]),
    "apv_range_gather": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_void_p]),
    "apv_range_device_ptrs": (C.c_int, [C.c_void_p] + [C.POINTER(C.c_void_p)] * 4),
    "apv_range_tail_get": (C.c_int, [C.c_void_p, _dp]),
    "apv_range_tail_add": (C.c_int, [C.c_void_p, _dp]),
    "apv_alloc_pinned": (C.c_void_p, [C.c_size_t]),
    "apv_free_pinned": (None, [C.c_void_p]),
    "apv_get": (C.c_int, [C.c_void_p, C.c_int, _dp, C.c_size_t]),
    "apv_set": (C.c_int, [C.c_void_p, C.c_int, _dp, C.c_size_t]),
    "apv_set_mu": (C.c_int, [C.c_void_p, C.c_double]),
    "apv_set_gain_table": (C.c_int, [C.c_void_p, C.c_int, _dp, C.c_double, C.c_double, C.c_double]),
    "apv_sweep": (C.c_int, [C.c_void_p, C.c_int, _dp, _dp]),
    "apv_sweep_device": (C.c_int, [C.c_void_p, C.c_int, _dp, C.c_void_p, _dp]),
    "apv_eval_zone": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _dp, _dp, _dp]),
    "apv_device_ptr": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "apv_synchronize": (C.c_int, [C.c_void_p]),
    "apv_stage_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "apv_jdiag_phase_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "apv_kernel_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "apv_timer_start": (C.c_int, [C.c_void_p]),
    "apv_timer_stop": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "apv_launch_count": (C.c_int, [C.c_void_p]),
    "apv_jdiag": (C.c_int, [C.c_int, C.c_int, _dp, _dp, C.c_double, C.c_int, _dp, _dp, C.POINTER(C.c_int)]),
    "apv_util_gemm": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _dp, _dp, C.c_double, _dp]),
    "apv_util_fft": (C.c_int, [C.c_int, C.c_int, _dp, _dp]),
    "apv_bench_dmma_peak": (C.c_int, [C.c_int, _dp]),
    "apv_bench_dfma": (C.c_int, [C.c_int, _dp]),
    "apv_bench_gemm": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "apv_bench_gemm_shape": (C.c_int, [C.c_int] * 8 + [C.c_double, C.c_int, C.POINTER(C.c_float)]),
    "apv_last_error": (C.c_char_p, []),
    "apv_version": (C.c_char_p, []),
}


def lib():
    """Load libapvast_b200.so (once).  Raises ImportError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C ap_vast_unofficial_b200/csrc` (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def ptr(a):
    """double* of a C-contiguous float64 array (or NULL for None)."""
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def pinned_array(shape):
    """float64 NumPy array in page-locked host memory (cudaMallocHost): the destination of asynchronous D2H copies."""
    n = int(np.prod(shape))
    p = lib().apv_alloc_pinned(max(n, 1) * 8)
    if not p:
        raise MemoryError(last_error())
    buf = (C.c_double * max(n, 1)).from_address(p)
    a = np.frombuffer(buf, dtype=np.float64, count=n).reshape(shape)
    _PINNED[a.__array_interface__["data"][0]] = p
    return a


def free_pinned(a):
    p = _PINNED.pop(a.__array_interface__["data"][0], None)
    if p:
        lib().apv_free_pinned(p)


_PINNED = {}


def last_error() -> str:
    return lib().apv_last_error().decode("utf-8", "replace")


def check(rc: int):
    """Map a status to the exception the reference would raise at the same point."""
    if rc == OK:
        return
    msg = last_error()
    if rc == ENOTPD:
        raise np.linalg.LinAlgError(msg or "Matrix is not positive definite")   # apvast.py:21-24
    if rc == EINVAL:
        raise RuntimeError(msg or "invalid argument")                           # apvast.py:86-90,154-155
    raise RuntimeError(f"apvast_b200 error {rc}: {msg}")
