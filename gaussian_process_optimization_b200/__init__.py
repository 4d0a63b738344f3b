"""gaussian_process_optimization_b200 -- B200-native (sm_100a) exact-GP inner loop behind the GPy / GPyOpt API surface.

The numerical path is hand-written CUDA in libgpb200.so (C ABI: include/gpb200.h); this package is the host-side mirror
of the reference's plug-in interfaces.  There is no CPU fallback.
"""
from . import _lib, native  # noqa: F401

__version__ = "0.1.0"
