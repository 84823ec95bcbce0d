"""ctypes binding of libb2lz4.so (include/b2lz4.h).  Fails loudly when the library is missing."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libb2lz4.so")
_CSRC = os.path.join(_HERE, "csrc")

LZ4_ERRORS = ["OutputTooSmall", "InputTooLarge", "CorruptedData", "DecompressionFailed", "InvalidState",
              "AllocationFailed"]                                   # reference src/lz4.zig:48-55
LZ4F_ERRORS = ["Generic", "MaxBlockSizeInvalid", "BlockModeInvalid", "ParameterInvalid",
               "CompressionLevelInvalid", "HeaderVersionWrong", "BlockChecksumInvalid", "ReservedFlagSet",
               "AllocationFailed", "SrcSizeTooLarge", "DstMaxSizeTooSmall", "FrameHeaderIncomplete",
               "FrameTypeUnknown", "FrameSizeWrong", "SrcPtrWrong", "DecompressionFailed",
               "HeaderChecksumInvalid", "ContentChecksumInvalid", "FrameDecodingAlreadyStarted",
               "CompressionStateUninitialized", "ParameterNull", "MaxCode", "OutOfMemory"]  # src/lz4f.zig:31-55
ERR_CUDA = 200
ERR_UNSUPPORTED_LEVEL = 201


class B2Error(Exception):
    """A non-zero status from the C-ABI; `.code` is the status, `.name` the reference error name."""

    def __init__(self, code, detail=""):
        self.code = code
        self.name = status_name(code)
        super().__init__(self.name + ((": " + detail) if detail else ""))


def status_name(code):
    if code == 0:
        return "ok"
    if 1 <= code <= 6:
        return "lz4." + LZ4_ERRORS[code - 1]
    if 100 <= code < 123:
        return "lz4f." + LZ4F_ERRORS[code - 100]
    if code == ERR_CUDA:
        return "b2lz4.CudaError"
    if code == ERR_UNSUPPORTED_LEVEL:
        return "b2lz4.UnsupportedLevel"
    return "b2lz4.status%d" % code


class Prefs(C.Structure):
    """b2lz4f_prefs == lz4f.Preferences + FrameInfo (reference src/lz4f.zig:106-122)."""
    _fields_ = [("block_size_id", C.c_uint32), ("block_mode", C.c_uint32), ("content_checksum", C.c_uint32),
                ("frame_type", C.c_uint32), ("content_size", C.c_uint64), ("dict_id", C.c_uint32),
                ("block_checksum", C.c_uint32), ("compression_level", C.c_int32), ("auto_flush", C.c_uint32),
                ("favor_dec_speed", C.c_uint32)]


class XxhState(C.Structure):
    _fields_ = [("v", C.c_uint32 * 4), ("tail", C.c_uint8 * 16), ("tail_len", C.c_uint32), ("seed", C.c_uint32),
                ("total", C.c_uint64)]


class FrameIndex(C.Structure):
    """struct b2lz4f_frame_index (include/b2lz4.h)"""
    _fields_ = [("nblocks", C.c_uint64), ("end_pos", C.c_uint64), ("content_size", C.c_uint64), ("terminal", C.c_uint32),
                ("header_size", C.c_uint32), ("block_size", C.c_uint32), ("block_checksum", C.c_uint32),
                ("content_checksum", C.c_uint32), ("max_stored", C.c_uint32)]


def library_path():
    return _SO


def build(force=False, verbose=False):
    """Compile libb2lz4.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    srcs.append(os.path.join(_HERE, "..", "include", "b2lz4.h"))
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        cmd = ["make", "-C", _CSRC, "-j8"] + (["-B"] if force else [])
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if verbose or r.returncode:
            print(r.stdout)
        if r.returncode:
            raise RuntimeError("building libb2lz4.so failed")
    return _SO


_lib = None

EXPORTS = [
    "b2lz4_status_name", "b2lz4_last_cuda_error", "b2lz4_kernel_launch_count", "b2lz4_version", "b2lz4_debug_tune",
    "b2lz4_ctx_create", "b2lz4_ctx_destroy", "b2lz4_ctx_device", "b2lz4_ctx_workspace_bytes",
    "b2lz4_compress_bound", "b2lz4_compress_default", "b2lz4_compress_fast", "b2lz4_decompress_safe",
    "b2lz4_decompress_safe_using_dict", "b2lz4_compress_hc", "b2lz4_xxh32",
    "b2lz4_compress_fast_batch_dev", "b2lz4_decompress_safe_batch_dev", "b2lz4_compress_hc_batch_dev",
    "b2lz4_compress_fast_batch", "b2lz4_decompress_safe_batch", "b2lz4_compress_hc_batch", "b2lz4_xxh32_dev",
    "b2lz4_compress_fast_using_dict", "b2lz4_compress_fast_dict_batch", "b2lz4_compress_fast_dict_batch_dev",
    "b2lz4_compress_dest_size", "b2lz4_compress_dest_size_batch", "b2lz4_compress_dest_size_batch_dev",
    "b2lz4f_index_frame_dev", "b2lz4f_compress_frame_mgpu", "b2lz4f_decompress_frame_mgpu",
    "b2lz4f_prefs_init", "b2lz4f_compress_frame_bound", "b2lz4f_compress_frame", "b2lz4f_decompress_frame",
    "b2lz4f_header_size", "b2lz4f_write_frame_header", "b2lz4f_parse_frame_header",
    "b2lz4f_compress_frame_ctx", "b2lz4f_decompress_frame_ctx", "b2lz4f_compress_frame_dev",
    "b2lz4f_decompress_frame_dev", "b2lz4_ctx_last_phase_ms", "b2lz4_ctx_set_timing",
    "b2lz4f_compress_blocks_dev", "b2lz4f_decompress_blocks_dev", "b2lz4_xxh32_state_init",
    "b2lz4_xxh32_state_update_dev", "b2lz4_xxh32_state_final", "b2lz4f_create_compression_context",
    "b2lz4f_free_compression_context", "b2lz4f_compress_begin", "b2lz4f_compress_bound",
    "b2lz4f_compress_update", "b2lz4f_compress_end",
]


def lib():
    """Load libb2lz4.so.  Raises if it has not been built — there is no fallback implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise RuntimeError("libb2lz4.so is not built (run __graft_entry__.build() or make -C %s); "
                           "there is no CPU fallback" % _CSRC)
    L = C.CDLL(_SO)
    vp, sz, szp, u32, i32 = C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_uint32, C.c_int
    pp = C.POINTER(Prefs)
    L.b2lz4_status_name.restype = C.c_char_p
    L.b2lz4_status_name.argtypes = [i32]
    L.b2lz4_last_cuda_error.restype = C.c_char_p
    L.b2lz4_kernel_launch_count.restype = C.c_uint64
    L.b2lz4_version.restype = C.c_char_p
    L.b2lz4_debug_tune.argtypes = [C.c_char_p, i32]
    L.b2lz4_ctx_create.argtypes = [i32, C.POINTER(vp)]
    L.b2lz4_ctx_destroy.argtypes = [vp]
    L.b2lz4_ctx_destroy.restype = None
    L.b2lz4_ctx_device.argtypes = [vp]
    L.b2lz4_ctx_workspace_bytes.argtypes = [vp]
    L.b2lz4_ctx_workspace_bytes.restype = sz
    L.b2lz4_ctx_set_timing.argtypes = [vp, i32]
    L.b2lz4_ctx_set_timing.restype = None
    L.b2lz4_ctx_last_phase_ms.argtypes = [vp, C.POINTER(C.c_float * 5)]
    L.b2lz4_compress_bound.restype = sz
    L.b2lz4_compress_bound.argtypes = [sz]
    L.b2lz4_compress_default.argtypes = [vp, sz, vp, sz, szp]
    L.b2lz4_compress_fast.argtypes = [vp, sz, vp, sz, u32, szp]
    L.b2lz4_decompress_safe.argtypes = [vp, sz, vp, sz, szp]
    L.b2lz4_decompress_safe_using_dict.argtypes = [vp, sz, vp, sz, vp, sz, szp]
    L.b2lz4_compress_hc.argtypes = [vp, sz, vp, sz, i32, szp]
    L.b2lz4_compress_fast_using_dict.argtypes = [vp, sz, vp, sz, vp, sz, u32, szp]
    L.b2lz4_xxh32.argtypes = [vp, sz, u32, C.POINTER(u32)]
    batch = [vp, vp, vp, vp, vp, vp, vp, vp, vp, sz]
    L.b2lz4_compress_fast_batch_dev.argtypes = batch + [u32, vp]
    L.b2lz4_decompress_safe_batch_dev.argtypes = batch + [vp, sz, vp]
    L.b2lz4_compress_hc_batch_dev.argtypes = batch + [i32, vp]
    L.b2lz4_compress_fast_batch.argtypes = batch + [u32]
    L.b2lz4_compress_fast_dict_batch.argtypes = batch + [vp, sz, u32]
    L.b2lz4_compress_fast_dict_batch_dev.argtypes = batch + [vp, sz, u32, vp]
    L.b2lz4_decompress_safe_batch.argtypes = batch + [vp, sz]
    L.b2lz4_compress_dest_size.argtypes = [vp, vp, sz, szp, szp]
    L.b2lz4_compress_dest_size_batch.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz]
    L.b2lz4_compress_dest_size_batch_dev.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, u32, vp]
    L.b2lz4_compress_hc_batch.argtypes = batch + [i32]
    L.b2lz4_xxh32_dev.argtypes = [vp, vp, sz, u32, vp, vp]
    L.b2lz4f_prefs_init.argtypes = [pp]
    L.b2lz4f_prefs_init.restype = None
    L.b2lz4f_compress_frame_bound.restype = sz
    L.b2lz4f_compress_frame_bound.argtypes = [sz, pp]
    L.b2lz4f_compress_frame.argtypes = [vp, sz, vp, sz, pp, szp]
    L.b2lz4f_decompress_frame.argtypes = [vp, sz, vp, sz, szp]
    L.b2lz4f_header_size.argtypes = [vp, sz, szp]
    L.b2lz4f_write_frame_header.argtypes = [vp, sz, pp, szp]
    L.b2lz4f_parse_frame_header.argtypes = [vp, sz, pp, szp]
    L.b2lz4f_compress_frame_ctx.argtypes = [vp, vp, sz, vp, sz, pp, szp]
    L.b2lz4f_decompress_frame_ctx.argtypes = [vp, vp, sz, vp, sz, szp]
    L.b2lz4f_compress_frame_dev.argtypes = [vp, vp, sz, vp, sz, pp, szp, vp]
    L.b2lz4f_decompress_frame_dev.argtypes = [vp, vp, sz, vp, sz, szp, vp]
    L.b2lz4f_compress_blocks_dev.argtypes = [vp, vp, sz, vp, sz, pp, szp, vp]
    L.b2lz4f_decompress_blocks_dev.argtypes = [vp, vp, sz, vp, sz, sz, i32, szp, vp]
    L.b2lz4f_compress_frame_mgpu.argtypes = [vp, sz, vp, sz, pp, i32, szp]
    L.b2lz4f_decompress_frame_mgpu.argtypes = [vp, sz, vp, sz, i32, szp]
    L.b2lz4f_index_frame_dev.argtypes = [vp, vp, sz, vp, vp, sz, C.POINTER(FrameIndex), vp]
    L.b2lz4_xxh32_state_init.argtypes = [C.POINTER(XxhState), u32]
    L.b2lz4_xxh32_state_init.restype = None
    L.b2lz4_xxh32_state_update_dev.argtypes = [vp, C.POINTER(XxhState), vp, sz, vp]
    L.b2lz4_xxh32_state_final.argtypes = [C.POINTER(XxhState)]
    L.b2lz4_xxh32_state_final.restype = u32
    L.b2lz4f_create_compression_context.argtypes = [C.POINTER(vp)]
    L.b2lz4f_free_compression_context.argtypes = [vp]
    L.b2lz4f_free_compression_context.restype = None
    L.b2lz4f_compress_begin.argtypes = [vp, vp, sz, pp, szp]
    L.b2lz4f_compress_bound.restype = sz
    L.b2lz4f_compress_bound.argtypes = [sz, pp]
    L.b2lz4f_compress_update.argtypes = [vp, vp, sz, vp, sz, szp]
    L.b2lz4f_compress_end.argtypes = [vp, vp, sz, szp]
    _lib = L
    return L


def check(rc):
    if rc:
        detail = lib().b2lz4_last_cuda_error().decode() if rc == ERR_CUDA else ""
        raise B2Error(rc, detail)


def kernel_launch_count():
    return lib().b2lz4_kernel_launch_count()


def debug_tune(key, value):
    """Diagnostic knob of the library (include/b2lz4.h b2lz4_debug_tune); returns the previous value."""
    old = lib().b2lz4_debug_tune(key.encode(), int(value))
    if old < 0:
        raise KeyError(key)
    return old


def as_buffer(b):
    """bytes-like / numpy -> (address, nbytes, keepalive).  Read-only bytes are used in place."""
    if isinstance(b, bytes):
        return (C.cast(C.c_char_p(b), C.c_void_p).value or 0), len(b), b
    if isinstance(b, (bytearray, memoryview)):
        mv = memoryview(b).cast("B")
        if len(mv) == 0:
            return 0, 0, mv
        if mv.readonly:
            bb = bytes(mv)
            return C.cast(C.c_char_p(bb), C.c_void_p).value, len(bb), bb
        arr = (C.c_uint8 * len(mv)).from_buffer(mv)
        return C.addressof(arr), len(mv), arr
    if hasattr(b, "ctypes") and hasattr(b, "nbytes"):  # numpy
        return b.ctypes.data, b.nbytes, b
    raise TypeError("expected a bytes-like object or numpy array, got %r" % type(b))


class Context:
    """b2lz4_ctx: one GPU + its workspace.  Device-pointer entry points live here."""

    def __init__(self, device=-1):
        self._h = C.c_void_p()
        check(lib().b2lz4_ctx_create(device, C.byref(self._h)))

    def close(self):
        if self._h:
            lib().b2lz4_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    @property
    def device(self):
        return lib().b2lz4_ctx_device(self._h)

    def workspace_bytes(self):
        return lib().b2lz4_ctx_workspace_bytes(self._h)

    def set_timing(self, on):
        lib().b2lz4_ctx_set_timing(self._h, 1 if on else 0)

    def last_phase_ms(self):
        arr = (C.c_float * 5)()
        check(lib().b2lz4_ctx_last_phase_ms(self._h, C.byref(arr)))
        return list(arr)

    # ---- device-pointer frame codec (ints are raw device addresses, e.g. tensor.data_ptr()) ----
    def compress_frame_dev(self, src_ptr, n, dst_ptr, cap, prefs=None, stream=0):
        out = C.c_size_t(0)
        check(lib().b2lz4f_compress_frame_dev(self._h, src_ptr, n, dst_ptr, cap,
                                               C.byref(prefs) if prefs is not None else None, C.byref(out), stream))
        return out.value

    def decompress_frame_dev(self, src_ptr, n, dst_ptr, cap, stream=0):
        out = C.c_size_t(0)
        check(lib().b2lz4f_decompress_frame_dev(self._h, src_ptr, n, dst_ptr, cap, C.byref(out), stream))
        return out.value

    def compress_blocks_dev(self, src_ptr, n, dst_ptr, cap, prefs=None, stream=0):
        out = C.c_size_t(0)
        check(lib().b2lz4f_compress_blocks_dev(self._h, src_ptr, n, dst_ptr, cap,
                                                C.byref(prefs) if prefs is not None else None, C.byref(out), stream))
        return out.value

    def decompress_blocks_dev(self, src_ptr, n, dst_ptr, cap, block_size, block_checksum=False, stream=0):
        out = C.c_size_t(0)
        check(lib().b2lz4f_decompress_blocks_dev(self._h, src_ptr, n, dst_ptr, cap, block_size,
                                                  1 if block_checksum else 0, C.byref(out), stream))
        return out.value

    def index_frame_dev(self, src_ptr, n, off_ptr=0, hdr_ptr=0, capacity=0, stream=0):
        """block index of a frame in device memory -> FrameIndex; off/hdr device arrays are optional"""
        info = FrameIndex()
        check(lib().b2lz4f_index_frame_dev(self._h, src_ptr, n, off_ptr, hdr_ptr, capacity, C.byref(info), stream))
        return info

    def xxh32_state_update_dev(self, state, src_ptr, n, stream=0):
        check(lib().b2lz4_xxh32_state_update_dev(self._h, C.byref(state), src_ptr, n, stream))
        return state

    # ---- host-pointer frame codec with this context ----
    def compress_frame(self, src, prefs=None, cap=None, dst=None):
        p, n, keep = as_buffer(src)
        pref_p = C.byref(prefs) if prefs is not None else None
        if dst is None:
            cap = lib().b2lz4f_compress_frame_bound(n, pref_p) if cap is None else cap
            buf = bytearray(cap)
            dp, dn, dkeep = as_buffer(buf) if cap else (0, 0, None)
        else:
            dp, dn, dkeep = as_buffer(dst)
            cap = dn if cap is None else cap
            buf = None
        out = C.c_size_t(0)
        check(lib().b2lz4f_compress_frame_ctx(self._h, p, n, dp, cap, pref_p, C.byref(out)))
        if buf is None:
            return out.value
        del dkeep
        return bytes(buf[:out.value])

    def decompress_frame(self, src, cap=None, dst=None):
        p, n, keep = as_buffer(src)
        if dst is None:
            buf = bytearray(cap)
            dp, dn, dkeep = as_buffer(buf) if cap else (0, 0, None)
        else:
            dp, dn, dkeep = as_buffer(dst)
            cap = dn if cap is None else cap
            buf = None
        out = C.c_size_t(0)
        check(lib().b2lz4f_decompress_frame_ctx(self._h, p, n, dp, cap, C.byref(out)))
        if buf is None:
            return out.value
        del dkeep
        return bytes(buf[:out.value])

    # ---- batch (device pointers) ----
    def compress_fast_batch_dev(self, src, src_off, src_len, dst, dst_off, dst_cap, out_len, status, nblocks, accel=1,
                                stream=0):
        check(lib().b2lz4_compress_fast_batch_dev(self._h, src, src_off, src_len, dst, dst_off, dst_cap, out_len, status,
                                                   nblocks, accel, stream))

    def decompress_safe_batch_dev(self, src, src_off, src_len, dst, dst_off, dst_cap, out_len, status, nblocks,
                                  dict_ptr=0, dict_len=0, stream=0):
        check(lib().b2lz4_decompress_safe_batch_dev(self._h, src, src_off, src_len, dst, dst_off, dst_cap, out_len, status,
                                                     nblocks, dict_ptr, dict_len, stream))

    def compress_hc_batch_dev(self, src, src_off, src_len, dst, dst_off, dst_cap, out_len, status, nblocks, level=9,
                              stream=0):
        check(lib().b2lz4_compress_hc_batch_dev(self._h, src, src_off, src_len, dst, dst_off, dst_cap, out_len, status,
                                                 nblocks, level, stream))

    # ---- batch (host numpy arrays) ----
    def _host_batch(self, fn, src, src_off, src_len, dst_cap_total, dst_off, dst_cap, extra):
        import numpy as np
        nb = len(src_off)
        sp, sn, skeep = as_buffer(src)
        so = np.ascontiguousarray(src_off, dtype=np.uint64)
        sl = np.ascontiguousarray(src_len, dtype=np.uint32)
        do = np.ascontiguousarray(dst_off, dtype=np.uint64)
        dc = np.ascontiguousarray(dst_cap, dtype=np.uint32)
        dst = np.zeros(max(1, dst_cap_total), dtype=np.uint8)
        ol = np.zeros(max(1, nb), dtype=np.uint32)
        st = np.zeros(max(1, nb), dtype=np.int32)
        check(fn(self._h, sp, so.ctypes.data, sl.ctypes.data, dst.ctypes.data, do.ctypes.data, dc.ctypes.data,
                 ol.ctypes.data, st.ctypes.data, nb, *extra))
        return dst, ol[:nb], st[:nb]

    def compress_fast_batch(self, src, src_off, src_len, dst_total, dst_off, dst_cap, accel=1):
        return self._host_batch(lib().b2lz4_compress_fast_batch, src, src_off, src_len, dst_total, dst_off, dst_cap,
                                (accel,))

    def compress_fast_dict_batch(self, src, src_off, src_len, dst_total, dst_off, dst_cap, dict, accel=1):
        """records against one shared dictionary; decode with decompress_safe_batch(..., dict=dict)"""
        dp, dn, keep = as_buffer(dict)
        extra = (dp if dn else 0, dn, accel)
        return self._host_batch(lib().b2lz4_compress_fast_dict_batch, src, src_off, src_len, dst_total, dst_off, dst_cap,
                                extra)

    def compress_dest_size_batch(self, src, src_off, src_len, dst_total, dst_off, dst_cap):
        """lz4.compressDestSize per block: returns (dst, consumed, out_len, status)"""
        import numpy as np
        nb = len(src_off)
        sp, sn, skeep = as_buffer(src)
        so = np.ascontiguousarray(src_off, dtype=np.uint64)
        sl = np.ascontiguousarray(src_len, dtype=np.uint32)
        do = np.ascontiguousarray(dst_off, dtype=np.uint64)
        dc = np.ascontiguousarray(dst_cap, dtype=np.uint32)
        dst = np.zeros(max(1, dst_total), dtype=np.uint8)
        used = np.zeros(max(1, nb), dtype=np.uint32)
        ol = np.zeros(max(1, nb), dtype=np.uint32)
        st = np.zeros(max(1, nb), dtype=np.int32)
        check(lib().b2lz4_compress_dest_size_batch(self._h, sp, so.ctypes.data, sl.ctypes.data, dst.ctypes.data,
                                                   do.ctypes.data, dc.ctypes.data, used.ctypes.data, ol.ctypes.data,
                                                   st.ctypes.data, nb))
        return dst, used[:nb], ol[:nb], st[:nb]

    def decompress_safe_batch(self, src, src_off, src_len, dst_total, dst_off, dst_cap, dict=None):
        if dict is None:
            extra = (0, 0)
            keep = None
        else:
            dp, dn, keep = as_buffer(dict)
            extra = (dp if dn else C.cast(C.c_char_p(b"\0"), C.c_void_p).value, dn)
        return self._host_batch(lib().b2lz4_decompress_safe_batch, src, src_off, src_len, dst_total, dst_off, dst_cap,
                                extra)

    def compress_hc_batch(self, src, src_off, src_len, dst_total, dst_off, dst_cap, level=9):
        return self._host_batch(lib().b2lz4_compress_hc_batch, src, src_off, src_len, dst_total, dst_off, dst_cap,
                                (level,))
