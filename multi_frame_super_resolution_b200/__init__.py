"""B200-native burst multi-frame super-resolution hot path (CUDA sm_100a behind a C ABI).

The compute lives in libmfsr_b200.so (csrc/*.cu, include/mfsr.h); this package is the thin
host-side mirror used by tests and benches.  Importing it never falls back to a CPU path.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
