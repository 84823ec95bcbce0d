"""ctypes binding of the CPU oracle (oracle/libb2oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module.  The product package (zig-lz4_b200/) never does.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libb2oracle.so")

OK = 0
OutputTooSmall, InputTooLarge, CorruptedData, DecompressionFailed, InvalidState, AllocationFailed = range(1, 7)
F_BASE = 100
UnsupportedLevel = 201

LZ4_ERRORS = ["OutputTooSmall", "InputTooLarge", "CorruptedData", "DecompressionFailed", "InvalidState",
              "AllocationFailed"]
LZ4F_ERRORS = ["Generic", "MaxBlockSizeInvalid", "BlockModeInvalid", "ParameterInvalid",
               "CompressionLevelInvalid", "HeaderVersionWrong", "BlockChecksumInvalid", "ReservedFlagSet",
               "AllocationFailed", "SrcSizeTooLarge", "DstMaxSizeTooSmall", "FrameHeaderIncomplete",
               "FrameTypeUnknown", "FrameSizeWrong", "SrcPtrWrong", "DecompressionFailed",
               "HeaderChecksumInvalid", "ContentChecksumInvalid", "FrameDecodingAlreadyStarted",
               "CompressionStateUninitialized", "ParameterNull", "MaxCode", "OutOfMemory"]


def status_name(code):
    if code == 0:
        return "ok"
    if 1 <= code <= 6:
        return "lz4." + LZ4_ERRORS[code - 1]
    if 100 <= code < 100 + len(LZ4F_ERRORS):
        return "lz4f." + LZ4F_ERRORS[code - 100]
    return "status%d" % code


class Prefs(C.Structure):
    _fields_ = [("block_size_id", C.c_uint32), ("block_mode", C.c_uint32), ("content_checksum", C.c_uint32),
                ("frame_type", C.c_uint32), ("content_size", C.c_uint64), ("dict_id", C.c_uint32),
                ("block_checksum", C.c_uint32), ("compression_level", C.c_int32), ("auto_flush", C.c_uint32),
                ("favor_dec_speed", C.c_uint32)]


class XxhState(C.Structure):
    """b2o_xxh32_state (b2o.h)"""
    _fields_ = [("v", C.c_uint32 * 4), ("buf", C.c_uint8 * 16), ("buf_len", C.c_uint32), ("total", C.c_uint64),
                ("seed", C.c_uint32)]


def build(force=False):
    if force or not os.path.exists(_SO) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_SO)
            for f in os.listdir(_HERE) if f.endswith((".c", ".h"))):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libb2oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None
_native = False


def prefer_native():
    """bench.py's CPU legs: rebuild the oracle with -O3 -march=native ON THE BOX THAT RUNS IT (oracle/_native/, not
    shipped, not tracked) and load that copy; the portable -O3 build stays the one tests use.  Returns the flags used."""
    global _SO, _lib, _native
    if _native:
        return "-O3 -march=native"
    out_dir = os.path.join(_HERE, "_native")
    so = os.path.join(out_dir, "libb2oracle.so")
    try:
        os.makedirs(out_dir, exist_ok=True)
        srcs = [os.path.join(_HERE, f) for f in ("oracle_lz4.c", "oracle_lz4hc.c", "oracle_lz4f.c", "oracle_xxh32.c")]
        subprocess.check_call(["gcc", "-O3", "-march=native", "-fPIC", "-std=c11", "-D_GNU_SOURCE", "-pthread", "-shared",
                               "-o", so] + srcs + ["-lpthread"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    except Exception:
        return "-O3 (portable build; native rebuild failed)"
    _SO, _lib, _native = so, None, True
    return "-O3 -march=native"


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p, sz, szp = C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)
        L.b2o_compress_bound.restype = sz
        L.b2o_compress_bound.argtypes = [sz]
        L.b2o_compress_fast.argtypes = [u8p, sz, u8p, sz, C.c_uint32, szp]
        L.b2o_decompress_safe.argtypes = [u8p, sz, u8p, sz, szp]
        L.b2o_decompress_safe_using_dict.argtypes = [u8p, sz, u8p, sz, u8p, sz, szp]
        L.b2o_compress_dest_size.argtypes = [u8p, u8p, sz, szp, szp]
        L.b2o_compress_hc.argtypes = [u8p, sz, u8p, sz, C.c_int, szp]
        L.b2o_hc_f8_guard_hits.restype = C.c_uint64
        L.b2o_xxh32.restype = C.c_uint32
        L.b2o_xxh32.argtypes = [u8p, sz, C.c_uint32]
        L.b2o_xxh32_init.restype = None
        L.b2o_xxh32_init.argtypes = [C.POINTER(XxhState), C.c_uint32]
        L.b2o_xxh32_update.restype = None
        L.b2o_xxh32_update.argtypes = [C.POINTER(XxhState), u8p, sz]
        L.b2o_xxh32_final.restype = C.c_uint32
        L.b2o_xxh32_final.argtypes = [C.POINTER(XxhState)]
        L.b2o_compress_frame_bound.restype = sz
        L.b2o_compress_frame_bound.argtypes = [sz, C.POINTER(Prefs)]
        L.b2o_compress_frame.argtypes = [u8p, sz, u8p, sz, C.POINTER(Prefs), szp]
        L.b2o_compress_frame_mt.argtypes = [u8p, sz, u8p, sz, C.POINTER(Prefs), szp, C.c_int]
        L.b2o_decompress_frame.argtypes = [u8p, sz, u8p, sz, szp]
        L.b2o_decompress_frame_mt.argtypes = [u8p, sz, u8p, sz, szp, C.c_int]
        L.b2o_header_size.argtypes = [u8p, sz, szp]
        L.b2o_write_frame_header.argtypes = [u8p, sz, C.POINTER(Prefs), szp]
        L.b2o_parse_frame_header.argtypes = [u8p, sz, C.POINTER(Prefs), szp]
        L.b2o_batch.argtypes = [C.c_int, C.c_int, u8p, u8p, u8p, u8p, u8p, u8p, u8p, u8p, sz, C.c_int]
        L.b2o_hardware_threads.restype = C.c_int
        _lib = L
    return _lib


class OracleError(Exception):
    def __init__(self, code):
        super().__init__(status_name(code))
        self.code = code


def _buf(b):
    """bytes-like -> (ctypes pointer value, length, keepalive)"""
    if isinstance(b, (bytes, bytearray, memoryview)):
        mv = memoryview(b)
        if mv.readonly:
            arr = (C.c_uint8 * max(1, len(mv))).from_buffer_copy(bytes(mv) if len(mv) else b"\0")
        else:
            arr = (C.c_uint8 * max(1, len(mv))).from_buffer(mv) if len(mv) else (C.c_uint8 * 1)()
        return C.addressof(arr), len(mv), arr
    # numpy array
    return b.ctypes.data, b.nbytes, b


def compress_bound(n):
    return lib().b2o_compress_bound(n)


def compress_fast(src, accel=1, cap=None):
    p, n, keep = _buf(src)
    cap = compress_bound(n) if cap is None else cap
    dst = (C.c_uint8 * max(1, cap))()
    out = C.c_size_t(0)
    rc = lib().b2o_compress_fast(p, n, dst, cap, accel, C.byref(out))
    if rc:
        raise OracleError(rc)
    return bytes(dst[:out.value])


def decompress_safe(src, cap, dict=None):
    p, n, keep = _buf(src)
    dst = (C.c_uint8 * max(1, cap))()
    out = C.c_size_t(0)
    if dict is None:
        rc = lib().b2o_decompress_safe(p, n, dst, cap, C.byref(out))
    else:
        dp, dn, dk = _buf(dict)
        rc = lib().b2o_decompress_safe_using_dict(p, n, dst, cap, dp, dn, C.byref(out))
    if rc:
        raise OracleError(rc)
    return bytes(dst[:out.value])


def compress_dest_size(src, cap, src_size=None):
    """lz4.compressDestSize: returns (consumed, compressed size, dst bytes as the reference leaves them)."""
    p, n, keep = _buf(src)
    used = C.c_size_t(n if src_size is None else src_size)
    dst = (C.c_uint8 * max(1, cap))()
    out = C.c_size_t(0)
    rc = lib().b2o_compress_dest_size(p, dst, cap, C.byref(used), C.byref(out))
    if rc:
        raise OracleError(rc)
    return used.value, out.value, bytes(dst[:cap])


def compress_hc(src, level=9, cap=None):
    p, n, keep = _buf(src)
    cap = compress_bound(n) if cap is None else cap
    dst = (C.c_uint8 * max(1, cap))()
    out = C.c_size_t(0)
    rc = lib().b2o_compress_hc(p, n, dst, cap, level, C.byref(out))
    if rc:
        raise OracleError(rc)
    return bytes(dst[:out.value])


def xxh32(data, seed=0):
    p, n, keep = _buf(data)
    return lib().b2o_xxh32(p, n, seed)


def xxh32_state_init(seed=0):
    st = XxhState()
    lib().b2o_xxh32_init(C.byref(st), seed)
    return st


def xxh32_state_update(st, data):
    p, n, keep = _buf(data)
    if n:
        lib().b2o_xxh32_update(C.byref(st), p, n)
    return st


def xxh32_state_final(st):
    return lib().b2o_xxh32_final(C.byref(st))


def make_prefs(block_size_id=0, block_mode=0, content_checksum=0, content_size=0, dict_id=0, block_checksum=0,
               compression_level=0):
    return Prefs(block_size_id, block_mode, content_checksum, 0, content_size, dict_id, block_checksum,
                 compression_level, 0, 0)


def compress_frame_bound(n, prefs=None):
    return lib().b2o_compress_frame_bound(n, C.byref(prefs) if prefs is not None else None)


def compress_frame(src, prefs=None, cap=None, threads=1):
    p, n, keep = _buf(src)
    cap = compress_frame_bound(n, prefs) if cap is None else cap
    dst = (C.c_uint8 * max(1, cap))()
    out = C.c_size_t(0)
    pp = C.byref(prefs) if prefs is not None else None
    if threads > 1:
        rc = lib().b2o_compress_frame_mt(p, n, dst, cap, pp, C.byref(out), threads)
    else:
        rc = lib().b2o_compress_frame(p, n, dst, cap, pp, C.byref(out))
    if rc:
        raise OracleError(rc)
    return bytes(memoryview(dst)[:out.value])


def decompress_frame(src, cap, threads=1):
    p, n, keep = _buf(src)
    dst = (C.c_uint8 * max(1, cap))()
    out = C.c_size_t(0)
    if threads > 1:
        rc = lib().b2o_decompress_frame_mt(p, n, dst, cap, C.byref(out), threads)
    else:
        rc = lib().b2o_decompress_frame(p, n, dst, cap, C.byref(out))
    if rc:
        raise OracleError(rc)
    return bytes(memoryview(dst)[:out.value])


def header_size(src):
    p, n, keep = _buf(src)
    out = C.c_size_t(0)
    rc = lib().b2o_header_size(p, n, C.byref(out))
    if rc:
        raise OracleError(rc)
    return out.value


def hardware_threads():
    return lib().b2o_hardware_threads()
