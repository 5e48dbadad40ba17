"""nbldpc_b200: B200-native (sm_100a) non-binary LDPC decode path behind a C ABI.

csrc/     CUDA kernels + extern "C" layer (libnbldpc_b200.so, declared in include/nbldpc_b200.h)
host/     C++ harness keeping the reference's CSimulation / CComm / CNBLDPC call surface on top of the C ABI
capi.py   ctypes binding used by the tests and bench.py
"""
from . import capi  # noqa: F401
