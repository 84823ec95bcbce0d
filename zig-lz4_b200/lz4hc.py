"""`lz4hc` namespace — mirrors /root/reference/src/lz4hc.zig's one-shot API over the C-ABI."""
from ._native import lib
from . import lz4 as _lz4

LZ4HC_CLEVEL_MIN = 2        # reference src/lz4hc.zig:28-31
LZ4HC_CLEVEL_DEFAULT = 9
LZ4HC_CLEVEL_OPT_MIN = 10
LZ4HC_CLEVEL_MAX = 12

Error = _lz4.Error


def compressBound(inputSize):
    """reference src/lz4hc.zig:1430-1432"""
    return _lz4.compressBound(inputSize)


def compressHC(src, compressionLevel=LZ4HC_CLEVEL_DEFAULT, dst_capacity=None):
    """reference src/lz4hc.zig:1440-1453.  Levels 3..9 (and <2 -> 9) run the hash-chain kernel; level 2
    (LZ4MID) and 10..12 (optimal parser) raise b2lz4.UnsupportedLevel — they are outside the accelerated
    path and are never silently rerouted."""
    n = len(src) if isinstance(src, (bytes, bytearray)) else memoryview(src).nbytes
    cap = compressBound(n) if dst_capacity is None else dst_capacity
    return _lz4._call_out(lib().b2lz4_compress_hc, src, cap, compressionLevel)
