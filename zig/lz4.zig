//! zig/lz4.zig — drop-in shim: the module `lz4` of jedisct1/zig-lz4 (src/root.zig:3-57) re-exported over
//! the C-ABI of libb2lz4.so (include/b2lz4.h).  A maintainer replaces `src/root.zig` by this file (or
//! adds it as module "lz4" in build.zig, linking `b2lz4`) and user code keeps compiling unchanged:
//!     const lz4 = @import("lz4");
//!     const n = try lz4.compressDefault(input, compressed);
//!     _ = try lz4.decompressSafe(compressed[0..n], out);
//! NOTE: there is no zig toolchain in the build image, so this file is NOT compiled or tested here;
//! tests/ drive the identical C symbols through ctypes instead (see INTEGRATION.md).
const std = @import("std");

// ---- extern "C" surface (include/b2lz4.h) ----
const c = struct {
    pub const Prefs = extern struct {
        block_size_id: u32 = 0,
        block_mode: u32 = 0,
        content_checksum: u32 = 0,
        frame_type: u32 = 0,
        content_size: u64 = 0,
        dict_id: u32 = 0,
        block_checksum: u32 = 0,
        compression_level: i32 = 0,
        auto_flush: u32 = 0,
        favor_dec_speed: u32 = 0,
    };
    pub extern "c" fn b2lz4_compress_bound(n: usize) usize;
    pub extern "c" fn b2lz4_compress_fast(src: [*]const u8, n: usize, dst: [*]u8, cap: usize, accel: u32, out: *usize) c_int;
    pub extern "c" fn b2lz4_decompress_safe(src: [*]const u8, n: usize, dst: [*]u8, cap: usize, out: *usize) c_int;
    pub extern "c" fn b2lz4_decompress_safe_using_dict(src: [*]const u8, n: usize, dst: [*]u8, cap: usize, dict: [*]const u8, dict_len: usize, out: *usize) c_int;
    pub extern "c" fn b2lz4_compress_fast_using_dict(src: [*]const u8, n: usize, dst: [*]u8, cap: usize, dict: [*]const u8, dict_len: usize, acceleration: u32, out: *usize) c_int;
    pub extern "c" fn b2lz4_compress_dest_size(src: [*]const u8, dst: [*]u8, cap: usize, src_size: *usize, out: *usize) c_int;
    pub extern "c" fn b2lz4_compress_hc(src: [*]const u8, n: usize, dst: [*]u8, cap: usize, level: c_int, out: *usize) c_int;
    pub extern "c" fn b2lz4f_compress_frame_bound(n: usize, prefs: ?*const Prefs) usize;
    pub extern "c" fn b2lz4f_compress_frame(src: [*]const u8, n: usize, dst: [*]u8, cap: usize, prefs: ?*const Prefs, out: *usize) c_int;
    pub extern "c" fn b2lz4f_decompress_frame(src: [*]const u8, n: usize, dst: [*]u8, cap: usize, out: *usize) c_int;
    pub extern "c" fn b2lz4f_compress_frame_mgpu(src: [*]const u8, n: usize, dst: [*]u8, cap: usize, prefs: ?*const Prefs, ngpus: c_int, out: *usize) c_int;
    pub extern "c" fn b2lz4f_decompress_frame_mgpu(src: [*]const u8, n: usize, dst: [*]u8, cap: usize, ngpus: c_int, out: *usize) c_int;
    pub extern "c" fn b2lz4f_header_size(src: [*]const u8, n: usize, out: *usize) c_int;
    pub extern "c" fn b2lz4f_create_compression_context(out: *?*anyopaque) c_int;
    pub extern "c" fn b2lz4f_free_compression_context(cctx: ?*anyopaque) void;
    pub extern "c" fn b2lz4f_compress_begin(cctx: ?*anyopaque, dst: [*]u8, cap: usize, prefs: ?*const Prefs, out: *usize) c_int;
    pub extern "c" fn b2lz4f_compress_update(cctx: ?*anyopaque, dst: [*]u8, cap: usize, src: [*]const u8, n: usize, out: *usize) c_int;
    pub extern "c" fn b2lz4f_compress_end(cctx: ?*anyopaque, dst: [*]u8, cap: usize, out: *usize) c_int;
};

// ---- lz4 namespace (reference src/lz4.zig) ----
pub const lz4 = struct {
    /// The reference's six members (src/lz4.zig:48-55) plus two the reference cannot produce: `GpuUnavailable`
    /// (status 200: no CUDA device / runtime failure — there is no CPU fallback) and `UnsupportedLevel` (status 201:
    /// lz4hc levels 2 and 10..12, see lz4hc below).  Code that switches over lz4.Error needs an `else` arm for them.
    pub const Error = error{ OutputTooSmall, InputTooLarge, CorruptedData, DecompressionFailed, InvalidState, AllocationFailed, GpuUnavailable, UnsupportedLevel };
    pub const MINMATCH = 4;
    pub const LZ4_MAX_INPUT_SIZE = 0x7E000000;
    pub const LZ4_DISTANCE_MAX = 65535;

    fn check(status: c_int) Error!void {
        return switch (status) {
            0 => {},
            1 => error.OutputTooSmall,
            2 => error.InputTooLarge,
            3 => error.CorruptedData,
            4 => error.DecompressionFailed,
            5 => error.InvalidState,
            6 => error.AllocationFailed,
            201 => error.UnsupportedLevel,
            else => error.GpuUnavailable, // 200 (B2LZ4_ERR_CUDA)
        };
    }
    pub fn compressBound(inputSize: usize) usize {
        return c.b2lz4_compress_bound(inputSize);
    }
    pub fn compressFast(src: []const u8, dst: []u8, acceleration: u32) Error!usize {
        var out: usize = 0;
        try check(c.b2lz4_compress_fast(src.ptr, src.len, dst.ptr, dst.len, acceleration, &out));
        return out;
    }
    pub fn compressDefault(src: []const u8, dst: []u8) Error!usize {
        return compressFast(src, dst, 1);
    }
    pub fn decompressSafe(src: []const u8, dst: []u8) Error!usize {
        var out: usize = 0;
        try check(c.b2lz4_decompress_safe(src.ptr, src.len, dst.ptr, dst.len, &out));
        return out;
    }
    pub fn decompressSafeUsingDict(src: []const u8, dst: []u8, dict: []const u8) Error!usize {
        var out: usize = 0;
        try check(c.b2lz4_decompress_safe_using_dict(src.ptr, src.len, dst.ptr, dst.len, dict.ptr, dict.len, &out));
        return out;
    }
    /// reference src/lz4.zig:551-616 (same consumed / compressed sizes; dst holds the stream of the consumed prefix)
    pub fn compressDestSize(src: []const u8, dst: []u8, srcSizePtr: *usize) Error!usize {
        var out: usize = 0;
        try check(c.b2lz4_compress_dest_size(src.ptr, dst.ptr, dst.len, srcSizePtr, &out));
        return out;
    }
    /// Not in the reference: compressFast with the table primed by `dict` — the encode side of
    /// decompressSafeUsingDict (Stream.loadDict promises it, src/lz4.zig:798-836, but never matches into the dictionary).
    pub fn compressFastUsingDict(src: []const u8, dst: []u8, dict: []const u8, acceleration: u32) Error!usize {
        var out: usize = 0;
        try check(c.b2lz4_compress_fast_using_dict(src.ptr, src.len, dst.ptr, dst.len, dict.ptr, dict.len, acceleration, &out));
        return out;
    }
};

// ---- lz4hc namespace (reference src/lz4hc.zig) ----
pub const lz4hc = struct {
    pub const Error = lz4.Error;
    /// The reference's level range (src/lz4hc.zig:28-31).  The accelerated path covers the hash-chain strategy,
    /// levels 3..9 (and everything below 2, which the reference maps to 9); level 2 (LZ4MID) and 10..12 (optimal
    /// parser) are other algorithms outside the hot-path scope: compressHC returns error.UnsupportedLevel for them
    /// instead of silently compressing with a different level.
    pub const LZ4HC_CLEVEL_MIN = 2;
    pub const LZ4HC_CLEVEL_DEFAULT = 9;
    pub const LZ4HC_CLEVEL_MAX = 12;
    pub const LZ4HC_CLEVEL_ACCELERATED_MIN = 3;
    pub const LZ4HC_CLEVEL_ACCELERATED_MAX = 9;
    pub fn compressBound(inputSize: usize) usize {
        return c.b2lz4_compress_bound(inputSize);
    }
    pub fn compressHC(src: []const u8, dst: []u8, compressionLevel: i32) Error!usize {
        var out: usize = 0;
        try lz4.check(c.b2lz4_compress_hc(src.ptr, src.len, dst.ptr, dst.len, compressionLevel, &out));
        return out;
    }
};

// ---- lz4f namespace (reference src/lz4f.zig + README streaming trio) ----
pub const lz4f = struct {
    pub const Error = error{
        Generic,                 MaxBlockSizeInvalid,       BlockModeInvalid,         ParameterInvalid,
        CompressionLevelInvalid, HeaderVersionWrong,        BlockChecksumInvalid,     ReservedFlagSet,
        AllocationFailed,        SrcSizeTooLarge,           DstMaxSizeTooSmall,       FrameHeaderIncomplete,
        FrameTypeUnknown,        FrameSizeWrong,            SrcPtrWrong,              DecompressionFailed,
        HeaderChecksumInvalid,   ContentChecksumInvalid,    FrameDecodingAlreadyStarted,
        CompressionStateUninitialized, ParameterNull,       MaxCode,                  OutOfMemory,
    };
    // size constants and isError, reference src/lz4f.zig:12-27,57-59
    pub const MAGICNUMBER: u32 = 0x184D2204;
    pub const MAGIC_SKIPPABLE_START: u32 = 0x184D2A50;
    pub const MAGIC_SKIPPABLE_MASK: u32 = 0xFFFFFFF0;
    pub const HEADER_SIZE_MIN: usize = 7;
    pub const HEADER_SIZE_MAX: usize = 19;
    pub const MIN_SIZE_TO_KNOW_HEADER_LENGTH: usize = 5;
    pub const BLOCK_HEADER_SIZE: usize = 4;
    pub const BLOCK_CHECKSUM_SIZE: usize = 4;
    pub const CONTENT_CHECKSUM_SIZE: usize = 4;
    pub const ENDMARK_SIZE: usize = 4;
    pub fn isError(code: usize) bool {
        return code > @as(usize, @bitCast(@as(isize, -65536)));
    }
    pub const BlockSizeID = enum(u3) { default = 0, max64KB = 4, max256KB = 5, max1MB = 6, max4MB = 7 };
    pub const BlockMode = enum(u1) { linked = 0, independent = 1 };
    pub const ContentChecksum = enum(u1) { disabled = 0, enabled = 1 };
    pub const BlockChecksum = enum(u1) { disabled = 0, enabled = 1 };
    pub const FrameType = enum(u1) { frame = 0, skippableFrame = 1 };
    pub const FrameInfo = struct {
        blockSizeID: BlockSizeID = .default,
        blockMode: BlockMode = .linked,
        contentChecksumFlag: ContentChecksum = .disabled,
        frameType: FrameType = .frame,
        contentSize: u64 = 0,
        dictID: u32 = 0,
        blockChecksumFlag: BlockChecksum = .disabled,
    };
    pub const Preferences = struct {
        frameInfo: FrameInfo = .{},
        compressionLevel: i32 = 0,
        autoFlush: bool = false,
        favorDecSpeed: bool = false,
        pub fn init() Preferences {
            return .{};
        }
    };

    fn toC(p: ?Preferences) c.Prefs {
        const q = p orelse Preferences{};
        return .{
            .block_size_id = @intFromEnum(q.frameInfo.blockSizeID),
            .block_mode = @intFromEnum(q.frameInfo.blockMode),
            .content_checksum = @intFromEnum(q.frameInfo.contentChecksumFlag),
            .frame_type = @intFromEnum(q.frameInfo.frameType),
            .content_size = q.frameInfo.contentSize,
            .dict_id = q.frameInfo.dictID,
            .block_checksum = @intFromEnum(q.frameInfo.blockChecksumFlag),
            .compression_level = q.compressionLevel,
            .auto_flush = @intFromBool(q.autoFlush),
            .favor_dec_speed = @intFromBool(q.favorDecSpeed),
        };
    }
    fn check(status: c_int) Error!void {
        if (status == 0) return;
        const members = [_]Error{
            error.Generic,                 error.MaxBlockSizeInvalid,    error.BlockModeInvalid,       error.ParameterInvalid,
            error.CompressionLevelInvalid, error.HeaderVersionWrong,     error.BlockChecksumInvalid,   error.ReservedFlagSet,
            error.AllocationFailed,        error.SrcSizeTooLarge,        error.DstMaxSizeTooSmall,     error.FrameHeaderIncomplete,
            error.FrameTypeUnknown,        error.FrameSizeWrong,         error.SrcPtrWrong,            error.DecompressionFailed,
            error.HeaderChecksumInvalid,   error.ContentChecksumInvalid, error.FrameDecodingAlreadyStarted,
            error.CompressionStateUninitialized, error.ParameterNull,    error.MaxCode,                error.OutOfMemory,
        };
        if (status >= 100 and status < 100 + @as(c_int, members.len)) return members[@intCast(status - 100)];
        if (status == 201) return error.CompressionLevelInvalid; // HC level 2 / 10..12: not on the accelerated path
        return error.Generic; // 200: no CUDA device / runtime failure
    }
    pub fn compressFrameBound(srcSize: usize, prefs: ?Preferences) usize {
        const p = toC(prefs);
        return c.b2lz4f_compress_frame_bound(srcSize, &p);
    }
    pub fn compressFrame(allocator: std.mem.Allocator, src: []const u8, dst: []u8, prefs: ?Preferences) Error!usize {
        _ = allocator; // unused by the reference too (src/lz4f.zig:443)
        const p = toC(prefs);
        var out: usize = 0;
        try check(c.b2lz4f_compress_frame(src.ptr, src.len, dst.ptr, dst.len, &p, &out));
        return out;
    }
    pub fn decompressFrame(allocator: std.mem.Allocator, src: []const u8, dst: []u8) Error!usize {
        _ = allocator;
        var out: usize = 0;
        try check(c.b2lz4f_decompress_frame(src.ptr, src.len, dst.ptr, dst.len, &out));
        return out;
    }
    /// compressFrame / decompressFrame with the frame sharded over `ngpus` devices inside the call (same bytes and errors)
    pub fn compressFrameMultiGpu(src: []const u8, dst: []u8, prefs: ?Preferences, ngpus: c_int) Error!usize {
        const p = toC(prefs);
        var out: usize = 0;
        try check(c.b2lz4f_compress_frame_mgpu(src.ptr, src.len, dst.ptr, dst.len, &p, ngpus, &out));
        return out;
    }
    pub fn decompressFrameMultiGpu(src: []const u8, dst: []u8, ngpus: c_int) Error!usize {
        var out: usize = 0;
        try check(c.b2lz4f_decompress_frame_mgpu(src.ptr, src.len, dst.ptr, dst.len, ngpus, &out));
        return out;
    }
    pub fn headerSize(src: []const u8) Error!usize {
        var out: usize = 0;
        try check(c.b2lz4f_header_size(src.ptr, src.len, &out));
        return out;
    }
    // README.md:98-122
    pub const CompressionContext = opaque {};
    pub fn createCompressionContext(allocator: std.mem.Allocator) Error!*CompressionContext {
        _ = allocator;
        var h: ?*anyopaque = null;
        try check(c.b2lz4f_create_compression_context(&h));
        return @ptrCast(h.?);
    }
    pub fn freeCompressionContext(cctx: *CompressionContext) void {
        c.b2lz4f_free_compression_context(cctx);
    }
    pub fn compressBegin(cctx: *CompressionContext, dst: []u8, prefs: ?*const Preferences) Error!usize {
        const p = toC(if (prefs) |q| q.* else null);
        var out: usize = 0;
        try check(c.b2lz4f_compress_begin(cctx, dst.ptr, dst.len, &p, &out));
        return out;
    }
    pub fn compressUpdate(cctx: *CompressionContext, dst: []u8, src: []const u8, options: ?*const anyopaque) Error!usize {
        _ = options;
        var out: usize = 0;
        try check(c.b2lz4f_compress_update(cctx, dst.ptr, dst.len, src.ptr, src.len, &out));
        return out;
    }
    pub fn compressEnd(cctx: *CompressionContext, dst: []u8, options: ?*const anyopaque) Error!usize {
        _ = options;
        var out: usize = 0;
        try check(c.b2lz4f_compress_end(cctx, dst.ptr, dst.len, &out));
        return out;
    }
};

// ---- flat re-exports, reference src/root.zig:7-57 ----
pub const Error = lz4.Error;
pub const compressDefault = lz4.compressDefault;
pub const compressFast = lz4.compressFast;
pub const compressBound = lz4.compressBound;
pub const compressDestSize = lz4.compressDestSize;
pub const decompressSafe = lz4.decompressSafe;
pub const decompressSafeUsingDict = lz4.decompressSafeUsingDict;
pub const MINMATCH = lz4.MINMATCH;
pub const LZ4_MAX_INPUT_SIZE = lz4.LZ4_MAX_INPUT_SIZE;
pub const LZ4_DISTANCE_MAX = lz4.LZ4_DISTANCE_MAX;
pub const compressHC = lz4hc.compressHC;
pub const LZ4HC_CLEVEL_MIN = lz4hc.LZ4HC_CLEVEL_MIN;
pub const LZ4HC_CLEVEL_DEFAULT = lz4hc.LZ4HC_CLEVEL_DEFAULT;
pub const LZ4HC_CLEVEL_MAX = lz4hc.LZ4HC_CLEVEL_MAX;
