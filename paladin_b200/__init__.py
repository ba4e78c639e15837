"""
paladin_b200 -- B200 (sm_100a) replacement for paladin's data-parallel hot path: many independent
FP64 dgeev eigendecompositions read from MatrixMarket COO (reference /root/reference:
src/paladin.hpp:216-343 compute(), src/lapack-wrapper.hpp:47-60 dgeev_, src/square-matrix.hpp:101-117
parse + scatter).

The product is the C-ABI shared library paladin_b200/lib/libpaladin_gpu.so (include/paladin_gpu.h).
This package is the thin Python host side used by tests/ and bench.py: ctypes bindings plus the
host-side mirror of the reference's listing / measure / balancer / writer logic (paladin_b200.host).
There is no CPU fallback anywhere in this package: if the library is missing or no B200 is present
every compute entry raises.
"""
from .api import (  # noqa: F401
    Batch,
    PaladinGpuError,
    device_count,
    geev_batch,
    dgeev,
    lib,
    lib_path,
    version,
)
