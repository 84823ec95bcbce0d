"""zig-lz4_b200 — B200-native LZ4 block/frame codec behind the API surface of jedisct1/zig-lz4.

The product is the C-ABI shared library ``libb2lz4.so`` (include/b2lz4.h; hand-written sm_100a CUDA
kernels).  This package is the Python-side mirror of the reference's module layout
(/root/reference/src/root.zig:3-57): namespaces ``lz4``, ``lz4hc``, ``lz4f`` plus the flat re-exports,
all forwarding to the C-ABI through ctypes.  There is no CPU fallback: without the built library or
without a CUDA device every compute call raises.

Import name: the directory is called ``zig-lz4_b200`` (not a valid identifier); ``zig_lz4_b200.py`` at
the repo root aliases it, so ``import zig_lz4_b200`` works.
"""
from . import _native
from ._native import (B2Error, lib, build, library_path, kernel_launch_count, debug_tune, Context, Prefs, status_name)
from . import lz4, lz4hc, lz4f

# ---- flat re-exports, reference src/root.zig:7-57 ----
Error = lz4.Error
compressDefault = lz4.compressDefault
compressFast = lz4.compressFast
compressBound = lz4.compressBound
decompressSafe = lz4.decompressSafe
decompressSafeUsingDict = lz4.decompressSafeUsingDict
MINMATCH = lz4.MINMATCH
LZ4_MAX_INPUT_SIZE = lz4.LZ4_MAX_INPUT_SIZE
LZ4_DISTANCE_MAX = lz4.LZ4_DISTANCE_MAX
compressHC = lz4hc.compressHC
LZ4HC_CLEVEL_MIN = lz4hc.LZ4HC_CLEVEL_MIN
LZ4HC_CLEVEL_DEFAULT = lz4hc.LZ4HC_CLEVEL_DEFAULT
LZ4HC_CLEVEL_MAX = lz4hc.LZ4HC_CLEVEL_MAX

__all__ = ["lz4", "lz4hc", "lz4f", "B2Error", "Context", "Prefs", "lib", "build", "library_path",
           "kernel_launch_count", "debug_tune", "status_name"]
