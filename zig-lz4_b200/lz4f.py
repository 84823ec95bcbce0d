"""`lz4f` namespace — mirrors /root/reference/src/lz4f.zig (+ the README streaming trio) over the C-ABI."""
import ctypes as C
from ._native import lib, check, as_buffer, B2Error, Prefs

# constants, reference src/lz4f.zig:12-27
MAGICNUMBER = 0x184D2204
MAGIC_SKIPPABLE_START = 0x184D2A50
HEADER_SIZE_MIN = 7
HEADER_SIZE_MAX = 19
BLOCK_HEADER_SIZE = 4
BLOCK_CHECKSUM_SIZE = 4
CONTENT_CHECKSUM_SIZE = 4
ENDMARK_SIZE = 4

Error = B2Error


class BlockSizeID:      # reference src/lz4f.zig:64-70
    default = 0
    max64KB = 4
    max256KB = 5
    max1MB = 6
    max4MB = 7

    @staticmethod
    def toBlockSize(v):
        return {0: 65536, 4: 65536, 5: 262144, 6: 1 << 20, 7: 4 << 20}[v]


class BlockMode:        # :82-85
    linked = 0
    independent = 1


class ContentChecksum:  # :88-91
    disabled = 0
    enabled = 1


class BlockChecksum:    # :94-97
    disabled = 0
    enabled = 1


def Preferences(blockSizeID=BlockSizeID.default, blockMode=BlockMode.linked,
                contentChecksumFlag=ContentChecksum.disabled, contentSize=0, dictID=0,
                blockChecksumFlag=BlockChecksum.disabled, compressionLevel=0, autoFlush=False, favorDecSpeed=False):
    """lz4f.Preferences{ .frameInfo = .{...}, .compressionLevel = ... } (reference src/lz4f.zig:106-122)."""
    return Prefs(blockSizeID, blockMode, contentChecksumFlag, 0, contentSize, dictID, blockChecksumFlag,
                 compressionLevel, 1 if autoFlush else 0, 1 if favorDecSpeed else 0)


def _pp(prefs):
    return C.byref(prefs) if prefs is not None else None


def compressFrameBound(srcSize, prefs=None):
    """reference src/lz4f.zig:274-301"""
    return lib().b2lz4f_compress_frame_bound(srcSize, _pp(prefs))


def compressFrame(src, prefs=None, dst_capacity=None):
    """reference src/lz4f.zig:354-446"""
    p, n, keep = as_buffer(src)
    cap = compressFrameBound(n, prefs) if dst_capacity is None else dst_capacity
    buf = bytearray(cap)
    dp, dn, dkeep = as_buffer(buf) if cap else (0, 0, None)
    out = C.c_size_t(0)
    check(lib().b2lz4f_compress_frame(p, n, dp, cap, _pp(prefs), C.byref(out)))
    del dkeep
    return bytes(buf[:out.value])


def decompressFrame(src, dst_capacity):
    """reference src/lz4f.zig:541-638"""
    p, n, keep = as_buffer(src)
    buf = bytearray(dst_capacity)
    dp, dn, dkeep = as_buffer(buf) if dst_capacity else (0, 0, None)
    out = C.c_size_t(0)
    check(lib().b2lz4f_decompress_frame(p, n, dp, dst_capacity, C.byref(out)))
    del dkeep
    return bytes(buf[:out.value])


def compressFrameMultiGPU(src, prefs=None, ngpus=8, dst=None):
    """lz4f.compressFrame with the frame sharded over `ngpus` devices inside the call (b2lz4f_compress_frame_mgpu);
    same bytes as compressFrame.  dst: optional writable buffer (e.g. pinned numpy array) -> returns the size."""
    p, n, keep = as_buffer(src)
    if dst is None:
        cap = compressFrameBound(n, prefs)
        buf = bytearray(cap)
        dp, dn, dkeep = as_buffer(buf)
    else:
        dp, cap, dkeep = as_buffer(dst)
        buf = None
    out = C.c_size_t(0)
    check(lib().b2lz4f_compress_frame_mgpu(p, n, dp, cap, _pp(prefs), ngpus, C.byref(out)))
    if buf is None:
        return out.value
    del dkeep
    return bytes(buf[:out.value])


def decompressFrameMultiGPU(src, dst_capacity=None, ngpus=8, dst=None):
    """lz4f.decompressFrame over `ngpus` devices (b2lz4f_decompress_frame_mgpu)"""
    p, n, keep = as_buffer(src)
    if dst is None:
        buf = bytearray(dst_capacity)
        dp, dn, dkeep = as_buffer(buf) if dst_capacity else (0, 0, None)
        cap = dst_capacity
    else:
        dp, cap, dkeep = as_buffer(dst)
        buf = None
    out = C.c_size_t(0)
    check(lib().b2lz4f_decompress_frame_mgpu(p, n, dp, cap, ngpus, C.byref(out)))
    if buf is None:
        return out.value
    del dkeep
    return bytes(buf[:out.value])


def headerSize(src):
    """reference src/lz4f.zig:451-480"""
    p, n, keep = as_buffer(src)
    out = C.c_size_t(0)
    check(lib().b2lz4f_header_size(p, n, C.byref(out)))
    return out.value


def writeFrameHeader(prefs):
    """reference src/lz4f.zig:304-351"""
    buf = bytearray(HEADER_SIZE_MAX)
    dp, dn, dkeep = as_buffer(buf)
    out = C.c_size_t(0)
    check(lib().b2lz4f_write_frame_header(dp, dn, _pp(prefs if prefs is not None else Prefs()), C.byref(out)))
    del dkeep
    return bytes(buf[:out.value])


def parseFrameHeader(src):
    """reference src/lz4f.zig:483-538 -> (Prefs-as-FrameInfo, header size)"""
    p, n, keep = as_buffer(src)
    info = Prefs()
    size = C.c_size_t(0)
    check(lib().b2lz4f_parse_frame_header(p, n, C.byref(info), C.byref(size)))
    return info, size.value


# ---- README streaming trio (reference README.md:98-122; not present in src/lz4f.zig — SURVEY F4) ----
class CompressionContext:
    def __init__(self):
        self._h = C.c_void_p()
        check(lib().b2lz4f_create_compression_context(C.byref(self._h)))
        self.prefs = None

    def free(self):
        if self._h:
            lib().b2lz4f_free_compression_context(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def createCompressionContext():
    return CompressionContext()


def freeCompressionContext(cctx):
    cctx.free()


def compressBound(srcSize, prefs=None):
    return lib().b2lz4f_compress_bound(srcSize, _pp(prefs))


def compressBegin(cctx, prefs=None, dst_capacity=HEADER_SIZE_MAX):
    buf = bytearray(dst_capacity)
    dp, dn, dkeep = as_buffer(buf) if dst_capacity else (0, 0, None)
    out = C.c_size_t(0)
    check(lib().b2lz4f_compress_begin(cctx._h, dp, dst_capacity, _pp(prefs), C.byref(out)))
    cctx.prefs = prefs
    del dkeep
    return bytes(buf[:out.value])


def compressUpdate(cctx, src, dst_capacity=None):
    p, n, keep = as_buffer(src)
    cap = compressBound(n, cctx.prefs) if dst_capacity is None else dst_capacity
    buf = bytearray(cap)
    dp, dn, dkeep = as_buffer(buf) if cap else (0, 0, None)
    out = C.c_size_t(0)
    check(lib().b2lz4f_compress_update(cctx._h, dp, cap, p, n, C.byref(out)))
    del dkeep
    return bytes(buf[:out.value])


def compressEnd(cctx, dst_capacity=None):
    cap = compressBound(0, cctx.prefs) if dst_capacity is None else dst_capacity
    buf = bytearray(cap)
    dp, dn, dkeep = as_buffer(buf) if cap else (0, 0, None)
    out = C.c_size_t(0)
    check(lib().b2lz4f_compress_end(cctx._h, dp, cap, C.byref(out)))
    del dkeep
    return bytes(buf[:out.value])
