"""Synthetic corpus generator (SURVEY.md §8d classes) — test/bench input only, not part of the codec."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libb2datagen.so")
TEXT, BINARY, REDUNDANT, RANDOM, MIXED = 0, 1, 2, 3, 4
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise RuntimeError("libb2datagen.so is not built (run __graft_entry__.build())")
        _lib = C.CDLL(_SO)
        _lib.b2gen_fill.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64, C.c_int]
    return _lib


def fill(out, seed=0x4C5A3442, mode=MIXED, span=65536, threads=None):
    """Fill a uint8 numpy array (or anything with .ctypes.data / nbytes) in place."""
    threads = threads or min(64, os.cpu_count() or 1)
    _load().b2gen_fill(out.ctypes.data, out.nbytes, seed, mode, span, threads)
    return out


def fill_ptr(ptr, nbytes, seed=0x4C5A3442, mode=MIXED, span=65536, threads=None):
    threads = threads or min(64, os.cpu_count() or 1)
    _load().b2gen_fill(ptr, nbytes, seed, mode, span, threads)


def generate(nbytes, seed=0x4C5A3442, mode=MIXED, span=65536, threads=None):
    return fill(np.empty(nbytes, dtype=np.uint8), seed, mode, span, threads)
