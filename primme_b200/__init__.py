"""primme_b200 -- B200-native (sm_100a) implementation of PRIMME's block Davidson inner loop.

The product is the shared library ``libprimme_b200.so`` (host control code in C + hand-written
CUDA kernels) exporting PRIMME's own C API (``dprimme``, ``cublas_dprimme``, ``primme_initialize``,
...) and the kernel-level C-ABI of ``include/primme_b200.h``.  This package only adds a ctypes
binding (``primme_b200.api``) and deterministic matrix generators (``primme_b200.matrices``).
"""
from . import api, matrices  # noqa: F401
