"""qblas_b200 — B200-native IEEE binary128 BLAS hot path (qgemm / qgemv / qdot / qnrm2 / qaxpy)
behind the API surface of SwayamInSync/QBLAS.  The product is the CUDA shared library
libqblas_b200.so (C ABI in include/qblas_b200.h); this package is its ctypes binding plus host
helpers for quad bit patterns.  No CPU compute path exists.
"""
from . import quad  # noqa: F401
from ._lib import LIB_PATH, QblasError, lib  # noqa: F401
from .api import *  # noqa: F401,F403
