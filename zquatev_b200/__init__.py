"""zquatev_b200 -- B200-native quaternionic Hermitian eigensolver (drop-in for ts::zquatev).

Python mirror of the reference's one public function (reference zquatev.h:54):

    info = zquatev(n2, D, ld2, eig)

with the same argument meaning (D: column-major 2n x 2n complex128 array of which the left
half is read and which is overwritten by ( U -V* ; V U* ); eig: float64 array receiving the n
eigenvalues in ascending order).  The work is done by hand-written sm_100a CUDA kernels in
``lib/libzquatev_b200.so`` reached through the C ABI of ``include/zquatev_b200.h``.  There is
no CPU path: importing works anywhere, calling without the library or a CUDA device raises.
"""
from .api import (  # noqa: F401
    LIB_PATH,
    ZqOptions,
    batched_stats,
    last_phases,
    last_gather_ms,
    last_trailing_ms,
    lib,
    release,
    set_profiling,
    version,
    zquatev,
    zquatev_batched,
    zquatev_device,
)

__all__ = ["zquatev", "zquatev_device", "zquatev_batched", "batched_stats", "last_phases", "last_gather_ms", "last_trailing_ms", "set_profiling", "release", "version",
           "lib", "LIB_PATH", "ZqOptions"]
