"""`lz4` namespace — mirrors /root/reference/src/lz4.zig's public block API over the C-ABI.

Zig slices become bytes-like objects; `dst: []u8` becomes a capacity (default: the tight bound) and the
written prefix is returned as bytes.  Zig error unions become `B2Error` (``.name == "lz4.CorruptedData"``).
"""
import ctypes as C
from ._native import lib, check, as_buffer, B2Error

# constants, reference src/lz4.zig:12-44
MINMATCH = 4
WILDCOPYLENGTH = 8
LASTLITERALS = 5
MFLIMIT = 12
ML_BITS = 4
ML_MASK = 15
RUN_BITS = 4
RUN_MASK = 15
LZ4_MAX_INPUT_SIZE = 0x7E000000
LZ4_DISTANCE_ABSOLUTE_MAX = 65535
LZ4_DISTANCE_MAX = 65535
LZ4_MEMORY_USAGE = 14
LZ4_HASHLOG = 12
ACCELERATION_DEFAULT = 1
ACCELERATION_MAX = 65537

Error = B2Error


def compressBound(inputSize):
    """reference src/lz4.zig:80-83"""
    return lib().b2lz4_compress_bound(inputSize)


def _call_out(fn, src, cap, *extra_before_out):
    p, n, keep = as_buffer(src)
    buf = bytearray(cap)
    dp, dn, dkeep = as_buffer(buf) if cap else (0, 0, None)
    out = C.c_size_t(0)
    check(fn(p, n, dp, cap, *extra_before_out, C.byref(out)))
    del dkeep
    return bytes(buf[:out.value])


def compressFast(src, acceleration=1, dst_capacity=None):
    """reference src/lz4.zig:292-447 (byte-identical output)"""
    n = len(memoryview(src).cast("B")) if not isinstance(src, bytes) else len(src)
    cap = compressBound(n) if dst_capacity is None else dst_capacity
    return _call_out(lib().b2lz4_compress_fast, src, cap, acceleration)


def compressDefault(src, dst_capacity=None):
    """reference src/lz4.zig:283-285"""
    return compressFast(src, ACCELERATION_DEFAULT, dst_capacity)


def compressFastUsingDict(src, dict, acceleration=1, dst_capacity=None):
    """compressFast with the table primed by `dict` (what Stream.loadDict promises, reference src/lz4.zig:798-836, but
    never delivers — SURVEY F6).  Decode with decompressSafeUsingDict(…, dict); an empty dict gives compressFast's bytes."""
    n = len(memoryview(src).cast("B")) if not isinstance(src, bytes) else len(src)
    cap = compressBound(n) if dst_capacity is None else dst_capacity
    dp, dn, dkeep = as_buffer(dict)
    return _call_out(lib().b2lz4_compress_fast_using_dict, src, cap, dp if dn else 0, dn, acceleration)


def compressDestSize(src, dst_capacity, src_size=None):
    """reference src/lz4.zig:551-616: compress the longest prefix of src[:src_size] the reference's bisection finds to
    fit dst_capacity bytes.  Returns (compressed bytes, consumed) — the Zig call returns the size and updates
    srcSizePtr.*."""
    p, n, keep = as_buffer(src)
    used = C.c_size_t(n if src_size is None else src_size)
    dst = (C.c_uint8 * max(1, dst_capacity))()
    out = C.c_size_t(0)
    check(lib().b2lz4_compress_dest_size(p, dst, dst_capacity, C.byref(used), C.byref(out)))
    return bytes(dst[:out.value]), used.value


def decompressSafe(src, dst_capacity):
    """reference src/lz4.zig:257-259"""
    return _call_out(lib().b2lz4_decompress_safe, src, dst_capacity)


def decompressSafeUsingDict(src, dst_capacity, dict):
    """reference src/lz4.zig:960-964"""
    dp, dn, dkeep = as_buffer(dict)
    if dn == 0:
        dp = C.cast(C.c_char_p(b"\0"), C.c_void_p).value
    return _call_out(lib().b2lz4_decompress_safe_using_dict, src, dst_capacity, dp, dn)


def xxh32(data, seed=0):
    """std.hash.XxHash32.hash(seed, data) as the reference uses it (src/lz4f.zig:139,424)"""
    p, n, keep = as_buffer(data)
    out = C.c_uint32(0)
    check(lib().b2lz4_xxh32(p, n, seed, C.byref(out)))
    return out.value
