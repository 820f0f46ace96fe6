"""ctypes binding of include/zquatev_b200.h + the host-side mirror of ts::zquatev."""
from __future__ import annotations

import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libzquatev_b200.so")

_lib = None


class ZqOptions(ctypes.Structure):
    """struct zq_options of include/zquatev_b200.h"""
    _fields_ = [("jobz", ctypes.c_int), ("device_ptrs", ctypes.c_int), ("nb", ctypes.c_int),
                ("stream", ctypes.c_void_p), ("sync", ctypes.c_int), ("col0", ctypes.c_int), ("ncols", ctypes.c_int), ("dist", ctypes.c_int), ("host_result", ctypes.c_int)]


# every symbol include/zquatev_b200.h declares: name -> (restype, argtypes)
_P, _I, _LL, _D = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_double
SYMBOLS = {
    "zquatev_b200": (_I, [_I, _P, _I, _P]),
    "zquatev_b200_ex": (_I, [_I, _P, _I, _P, ctypes.POINTER(ZqOptions)]),
    "zquatev_b200_create": (_I, [ctypes.POINTER(_P)]),
    "zquatev_b200_destroy": (_I, [_P]),
    "zquatev_b200_solve": (_I, [_P, _I, _P, _I, _P, ctypes.POINTER(ZqOptions)]),
    "zquatev_b200_reserve": (_I, [_P, _I, ctypes.POINTER(ZqOptions)]),
    "zquatev_b200_workspace_query": (_I, [_I, ctypes.POINTER(ZqOptions), ctypes.POINTER(ctypes.c_ulonglong)]),
    "zquatev_b200_handle_phases": (_I, [_P, _P]),
    "zquatev_b200_qgemm": (_I, [_I, _I, _I, _I, _I, _D, _P, _I, _P, _I, _D, _P, _I, _P]),
    "zquatev_b200_congruence": (_I, [_I, _I, _P, _I, _P, _I, _P, _I, _P, _P]),
    "zquatev_b200_fill_pairing": (_I, [_I, _I, _P, _I, _P]),
    "zquatev_b200_batched": (_I, [_I, _I, _P, _I, _LL, _P, _LL, _P]),
    "zquatev_b200_batched_stats": (None, [_P, _P]),
    "zquatev_b200_release": (None, []),
    "zquatev_b200_dist_unique_id": (_I, [_P]),
    "zquatev_b200_dist_init": (_I, [_I, _I, _P]),
    "zquatev_b200_dist_finalize": (None, []),
    "zquatev_b200_dist_transport": (_I, []),
    "zquatev_b200_last_phases": (_I, [_P]),
    "zquatev_b200_set_profiling": (None, [_I]),
    "zquatev_b200_last_trailing_ms": (_D, []),
    "zquatev_b200_last_gather_ms": (_D, []),
    "zquatev_b200_version": (ctypes.c_char_p, []),
    "zq_test_matvec": (_I, [_I, _I, _P, _LL, _P, _P, _I, _P]),
    "zq_test_zgemm": (_I, [_I, _I, _I, _I, _I, _P, _P, _LL, _P, _LL, _P, _P, _LL, _I, _I, _P]),
    "zq_test_set_gemm_3m": (None, [_I]),
    "zq_test_qgemm": (_I, [_I, _I, _I, _I, _I, _D, _P, _LL, _LL, _P, _LL, _LL, _D, _P, _LL, _LL, _I, _I, _P]),
    "zq_test_stedc": (_I, [_I, _P, _P, _P, _P]),
    "zq_test_bisect": (_I, [_I, _P, _P, _P]),
    "zq_test_tridiag": (_I, [_I, _I, _P, _LL, _P, _P, _P, _P]),
    "zq_test_dist_chunks": (_I, [_I, _I, _P]),
    "zq_test_upload_bounds": (_I, [_I, _I, _P]),
}


def lib():
    """Loads libzquatev_b200.so (fails loudly when it has not been built: there is no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not built: run `python -m zquatev_b200.build` (no CPU fallback exists)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def version() -> str:
    return lib().zquatev_b200_version().decode()


def _check(info: int, what: str) -> int:
    if info <= -1000:
        raise RuntimeError(f"{what}: CUDA runtime error {-info - 1000} (a B200/CUDA device is required; no CPU fallback)")
    if info < 0:
        raise ValueError(f"{what}: illegal argument {-info}")
    return info


def zquatev(n2: int, D: np.ndarray, ld2: int, eig: np.ndarray, jobz: int = 1, nb: int = 0) -> int:
    """ts::zquatev(n2, D, ld2, eig) -- reference zquatev.h:54.

    D   : complex128 array holding the column-major ld2 x n2 matrix, i.e. a Fortran-ordered
          (ld2, n2) array (or any C-contiguous buffer with the same memory).  Overwritten.
    eig : float64 array with >= n2/2 entries.  Returns info (0 = success, > 0 solver failure).
    """
    if D.dtype != np.complex128 or eig.dtype != np.float64:
        raise TypeError("D must be complex128 and eig float64")
    if not (D.flags.f_contiguous or D.flags.c_contiguous):
        raise ValueError("D must be contiguous")
    if D.size < ld2 * n2 or eig.size < n2 // 2:
        raise ValueError("D or eig too small")
    opt = ZqOptions(jobz, 0, nb, None, 1, 0, 0, 0, 0)
    info = lib().zquatev_b200_ex(n2, D.ctypes.data, ld2, eig.ctypes.data, ctypes.byref(opt))
    return _check(info, "zquatev")


def zquatev_device(n2: int, D_ptr: int, ld2: int, eig_ptr: int, jobz: int = 1, nb: int = 0, stream: int = 0,
                   sync: bool = True, col0: int = 0, ncols: int = 0, dist: bool = False) -> int:
    """Device-resident variant: D_ptr / eig_ptr are CUDA device addresses (e.g. torch
    ``tensor.data_ptr()``) of a column-major ld2 x n2 complex128 array and n doubles."""
    opt = ZqOptions(jobz, 1, nb, ctypes.c_void_p(stream) if stream else None, 1 if sync else 0, col0, ncols, 1 if dist else 0, 0)
    info = lib().zquatev_b200_ex(n2, ctypes.c_void_p(D_ptr), ld2, ctypes.c_void_p(eig_ptr), ctypes.byref(opt))
    return _check(info, "zquatev_device")


def zquatev_batched(D: np.ndarray, eig: np.ndarray) -> np.ndarray:
    """D: (batch, n2, n2) array where D[b] is the column-major matrix of problem b stored
    Fortran-style (i.e. D[b].T is the C-ordered view); eig: (batch, n) float64."""
    batch, n2 = D.shape[0], D.shape[-1]
    info = np.zeros(batch, dtype=np.int32)
    rc = lib().zquatev_b200_batched(batch, n2, D.ctypes.data, n2, n2 * n2, eig.ctypes.data, eig.shape[-1], info.ctypes.data)
    _check(rc if rc < 0 else 0, "zquatev_batched")
    return info


def batched_stats():
    """(graph_launches, eager_solves) of zquatev_batched since the library was loaded."""
    g, e = ctypes.c_int(0), ctypes.c_int(0)
    lib().zquatev_b200_batched_stats(ctypes.byref(g), ctypes.byref(e))
    return g.value, e.value


def last_phases():
    """dict of device milliseconds of the last solve (see zquatev_b200_last_phases)."""
    ms = (ctypes.c_double * 8)()
    if not lib().zquatev_b200_last_phases(ms):
        return None
    keys = ["h2d", "tridiag", "tridiag_eig", "backtransform", "d2h", "device_total", "k1_matvec", "launches"]
    return dict(zip(keys, list(ms)))


def last_trailing_ms() -> float:
    """device milliseconds of the trailing-update GEMMs of the last profiled single-GPU solve"""
    return float(lib().zquatev_b200_last_trailing_ms())


def last_gather_ms() -> float:
    """device milliseconds of the NCCL gather of the eigenvector shards in the last collective solve (0: single GPU)"""
    return float(lib().zquatev_b200_last_gather_ms())


def set_profiling(on: bool):
    lib().zquatev_b200_set_profiling(1 if on else 0)


def release():
    lib().zquatev_b200_release()
