"""dctz_b200 -- B200-native (sm_100a) implementation of DCTZ's data-parallel hot path.

The product is two native libraries built in-tree by `__graft_entry__.build()`:

  dctz_b200/libdctz_gpu.so   CUDA kernels + the C-ABI of include/dctz_gpu.h
  dctz_b200/libdctz.so       host C library with the reference's public API (dctz.h:121-128:
                             dctz_compress / dctz_decompress / calc_data_stat / gen_bins / dct_* ...)
                             that calls the GPU through that C-ABI and keeps zlib on the host

This Python package is only a thin ctypes binding over the C-ABI, used by tests/ and bench.py
(torch supplies device memory, streams and torch.distributed; it is plumbing, not the product).
There is no CPU fallback: without the built library or without a CUDA device every call raises.
"""
from .binding import (  # noqa: F401
    DOUBLE,
    FLOAT,
    Context,
    DctzGpuError,
    GpuInfo,
    LIB_PATH,
    load_library,
)

__all__ = ["DOUBLE", "FLOAT", "Context", "DctzGpuError", "GpuInfo", "LIB_PATH", "load_library"]
