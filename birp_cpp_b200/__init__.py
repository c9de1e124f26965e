"""birp_cpp_b200: B200-native line-of-sight integrator behind the BIRP Camera/Space API.

The product is `libbirp_b200.so` (CUDA kernels + C ABI, include/birp_b200.h) and the C++
classes in birp_cpp_b200/include + birp_cpp_b200/src that mirror the reference's
Space / DataCube / CMEM / Camera interface.  This Python package is a thin ctypes
binding used by the tests and bench.py.
"""
from . import capi, synthetic  # noqa: F401

__all__ = ["capi", "synthetic"]
