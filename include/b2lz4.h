/*
 * b2lz4.h — C-ABI of the B200-native LZ4 block/frame codec (libb2lz4.so).
 *
 * This is the drop-in boundary for the hot path of jedisct1/zig-lz4: a Zig shim (zig/lz4.zig, see
 * INTEGRATION.md) re-exports the reference's names (src/root.zig:3-57) over these `extern "C"`
 * entry points.  Plain pointers and sizes only; no torch / C++ types.  Every compute entry point
 * runs hand-written sm_100a CUDA kernels; there is NO CPU fallback — a call made without a usable
 * CUDA device returns B2LZ4_ERR_CUDA.
 *
 * Status convention (replaces Zig error unions):
 *   0            ok
 *   1..6         lz4.Error members in declaration order        (reference src/lz4.zig:48-55)
 *   100+k        lz4f.Error member k in declaration order      (reference src/lz4f.zig:31-55)
 *   200..        conditions that do not exist in the reference (CUDA failure, out-of-scope level)
 * Byte counts are returned through `size_t* out`.
 *
 * Pointer kinds: functions without a `_dev` suffix take HOST pointers (the reference's semantics:
 * caller-owned slices, synchronous).  `_dev` functions take DEVICE pointers plus a CUDA stream
 * (`void* stream` is a cudaStream_t; NULL = the context's own stream) and are the throughput path.
 * Device buffers must be readable/writable up to the next 16-byte boundary past their length
 * (true of every cudaMalloc / torch allocation).
 */
#ifndef B2LZ4_H
#define B2LZ4_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define B2LZ4_API __attribute__((visibility("default")))
#else
#define B2LZ4_API
#endif

/* ------------------------------------------------------------------ status codes ---- */
enum b2lz4_status {
    B2LZ4_OK = 0,
    /* lz4.Error — reference src/lz4.zig:48-55 */
    B2LZ4_ERR_OUTPUT_TOO_SMALL = 1,
    B2LZ4_ERR_INPUT_TOO_LARGE = 2,
    B2LZ4_ERR_CORRUPTED_DATA = 3,
    B2LZ4_ERR_DECOMPRESSION_FAILED = 4,
    B2LZ4_ERR_INVALID_STATE = 5,
    B2LZ4_ERR_ALLOCATION_FAILED = 6,
    /* lz4f.Error — reference src/lz4f.zig:31-55 */
    B2LZ4F_ERR_GENERIC = 100,
    B2LZ4F_ERR_MAX_BLOCK_SIZE_INVALID = 101,
    B2LZ4F_ERR_BLOCK_MODE_INVALID = 102,
    B2LZ4F_ERR_PARAMETER_INVALID = 103,
    B2LZ4F_ERR_COMPRESSION_LEVEL_INVALID = 104,
    B2LZ4F_ERR_HEADER_VERSION_WRONG = 105,
    B2LZ4F_ERR_BLOCK_CHECKSUM_INVALID = 106,
    B2LZ4F_ERR_RESERVED_FLAG_SET = 107,
    B2LZ4F_ERR_ALLOCATION_FAILED = 108,
    B2LZ4F_ERR_SRC_SIZE_TOO_LARGE = 109,
    B2LZ4F_ERR_DST_MAX_SIZE_TOO_SMALL = 110,
    B2LZ4F_ERR_FRAME_HEADER_INCOMPLETE = 111,
    B2LZ4F_ERR_FRAME_TYPE_UNKNOWN = 112,
    B2LZ4F_ERR_FRAME_SIZE_WRONG = 113,
    B2LZ4F_ERR_SRC_PTR_WRONG = 114,
    B2LZ4F_ERR_DECOMPRESSION_FAILED = 115,
    B2LZ4F_ERR_HEADER_CHECKSUM_INVALID = 116,
    B2LZ4F_ERR_CONTENT_CHECKSUM_INVALID = 117,
    B2LZ4F_ERR_FRAME_DECODING_ALREADY_STARTED = 118,
    B2LZ4F_ERR_COMPRESSION_STATE_UNINITIALIZED = 119,
    B2LZ4F_ERR_PARAMETER_NULL = 120,
    B2LZ4F_ERR_MAX_CODE = 121,
    B2LZ4F_ERR_OUT_OF_MEMORY = 122,
    /* not in the reference */
    B2LZ4_ERR_CUDA = 200,              /* no device / CUDA runtime failure; b2lz4_last_cuda_error() */
    B2LZ4_ERR_UNSUPPORTED_LEVEL = 201  /* HC level 2 (LZ4MID) and 10-12 (optimal parser): outside the
                                          accelerated path; never silently rerouted */
};

B2LZ4_API const char* b2lz4_status_name(int status);
B2LZ4_API const char* b2lz4_last_cuda_error(void);
/* Number of CUDA kernels this library has launched in this process (monotone counter). */
B2LZ4_API uint64_t b2lz4_kernel_launch_count(void);
B2LZ4_API const char* b2lz4_version(void);
/* Diagnostic knob (process-wide; 0 restores the shipped default): "k1_ctas", "k2_occ", "k2_variant", "k3_variant",
 * "pipe_blocks", "no_pipeline", "serial_walk", "xxh_variant", "spare0".."spare7".  Returns the previous value, -1 for an
 * unknown key.  Used by the experiments in DESIGN.md and by tests that force a rare path; the library never reads
 * the environment. */
B2LZ4_API int b2lz4_debug_tune(const char* key, int value);

/* ------------------------------------------------------------------ context ---- */
/* One context = one GPU + its workspace (block slots, size/offset tables, pinned staging) and a
 * private stream.  The reference has no such object (it is single-threaded CPU code, no globals —
 * SURVEY §8b "Threading"); functions without a ctx argument use a lazily created per-process
 * default context guarded by a mutex, so they stay callable from several host threads.
 * Stream rule for the *_dev calls (asynchronous on the caller's stream): a context owns ONE set of scratch (work
 * ticket, HC tables, dictionary table, checksum state), so the library orders every call on a context after the
 * previous call on that context on the device — the new call's stream waits for an event recorded at the end of the
 * previous one — whatever streams were passed.  Calls that should overlap on the device need one context each. */
typedef struct b2lz4_ctx b2lz4_ctx;
B2LZ4_API int b2lz4_ctx_create(int device /* -1 = current */, b2lz4_ctx** out);
B2LZ4_API void b2lz4_ctx_destroy(b2lz4_ctx* ctx);
B2LZ4_API int b2lz4_ctx_device(const b2lz4_ctx* ctx);
/* device bytes currently held by the context's workspace */
B2LZ4_API size_t b2lz4_ctx_workspace_bytes(const b2lz4_ctx* ctx);

/* ------------------------------------------------------------------ block API (host pointers) ---- */
/* replaces lz4.compressBound — reference src/lz4.zig:80-83 */
B2LZ4_API size_t b2lz4_compress_bound(size_t input_size);
/* replaces lz4.compressDefault — reference src/lz4.zig:283-285 */
B2LZ4_API int b2lz4_compress_default(const void* src, size_t n, void* dst, size_t cap, size_t* out);
/* replaces lz4.compressFast — reference src/lz4.zig:292-447 (byte-identical output) */
B2LZ4_API int b2lz4_compress_fast(const void* src, size_t n, void* dst, size_t cap, uint32_t acceleration,
                                  size_t* out);
/* replaces lz4.decompressSafe — reference src/lz4.zig:257-259 (same error kinds) */
B2LZ4_API int b2lz4_decompress_safe(const void* src, size_t n, void* dst, size_t cap, size_t* out);
/* replaces lz4.decompressSafeUsingDict — reference src/lz4.zig:960-964 */
B2LZ4_API int b2lz4_decompress_safe_using_dict(const void* src, size_t n, void* dst, size_t cap,
                                               const void* dict, size_t dict_len, size_t* out);
/* Dictionary compression, the encode side of decompressSafeUsingDict.  The reference's Stream.loadDict +
 * compressFastContinue (src/lz4.zig:798-836) accept a dictionary but never match into it; this is
 * the same matcher with the table primed by the last 64 KiB of `dict`.  Output decodes with
 * lz4.decompressSafeUsingDict(…, dict); with dict_len == 0 it is byte-identical to lz4.compressFast. */
B2LZ4_API int b2lz4_compress_fast_using_dict(const void* src, size_t n, void* dst, size_t cap, const void* dict,
                                             size_t dict_len, uint32_t acceleration, size_t* out);
/* replaces lz4.compressDestSize — reference src/lz4.zig:551-616.  *src_size: in = bytes available, out = bytes
 * consumed (the prefix the reference's bisection settles on); *out = compressed size.  dst[0..*out) is
 * compressDefault(src[0..*src_size)): the reference leaves the bytes of its LAST probe in dst, which is a
 * different (undecodable) stream whenever that probe did not fit — the returned sizes are the reference's,
 * the bytes are the ones its contract (and its own test, src/test_dictionary.zig:78-103) expects. */
B2LZ4_API int b2lz4_compress_dest_size(const void* src, void* dst, size_t cap, size_t* src_size, size_t* out);
/* replaces lz4hc.compressHC — reference src/lz4hc.zig:1440-1453 (levels 3..9; <2 -> 9) */
B2LZ4_API int b2lz4_compress_hc(const void* src, size_t n, void* dst, size_t cap, int level, size_t* out);
/* XXH32 as used by lz4f through std.hash.XxHash32 — reference src/lz4f.zig:139,424,438 */
B2LZ4_API int b2lz4_xxh32(const void* src, size_t n, uint32_t seed, uint32_t* out);

/* ------------------------------------------------------------------ batch API ---- */
/* Many independent blocks per call: block i is src[src_off[i] .. +src_len[i]) and may write
 * dst[dst_off[i] .. +dst_cap[i]).  out_len[i] / status[i] carry what the per-block reference call
 * would have returned.  The reference has no batch call; this is `for blocks |b| compressFast(b)`
 * (how src/lz4f.zig:379-430 drives the codec) hoisted into one launch.
 *
 * `_dev`: every pointer (including the offset/len/cap/out arrays) is a device pointer; the call is
 * asynchronous on `stream`.  Without `_dev`: host pointers, synchronous. */
B2LZ4_API int b2lz4_compress_fast_batch_dev(b2lz4_ctx* ctx, const void* src, const uint64_t* src_off,
                                            const uint32_t* src_len, void* dst, const uint64_t* dst_off,
                                            const uint32_t* dst_cap, uint32_t* out_len, int32_t* status,
                                            size_t nblocks, uint32_t acceleration, void* stream);
B2LZ4_API int b2lz4_decompress_safe_batch_dev(b2lz4_ctx* ctx, const void* src, const uint64_t* src_off,
                                              const uint32_t* src_len, void* dst, const uint64_t* dst_off,
                                              const uint32_t* dst_cap, uint32_t* out_len, int32_t* status,
                                              size_t nblocks, const void* dict, size_t dict_len,
                                              void* stream);
B2LZ4_API int b2lz4_compress_hc_batch_dev(b2lz4_ctx* ctx, const void* src, const uint64_t* src_off,
                                          const uint32_t* src_len, void* dst, const uint64_t* dst_off,
                                          const uint32_t* dst_cap, uint32_t* out_len, int32_t* status,
                                          size_t nblocks, int level, void* stream);
/* many records against one shared dictionary (BASELINE configs[4] record case); device / host pointers */
B2LZ4_API int b2lz4_compress_fast_dict_batch_dev(b2lz4_ctx* ctx, const void* src, const uint64_t* src_off,
                                                 const uint32_t* src_len, void* dst, const uint64_t* dst_off,
                                                 const uint32_t* dst_cap, uint32_t* out_len, int32_t* status,
                                                 size_t nblocks, const void* dict, size_t dict_len,
                                                 uint32_t acceleration, void* stream);
B2LZ4_API int b2lz4_compress_fast_dict_batch(b2lz4_ctx* ctx, const void* src, const uint64_t* src_off,
                                             const uint32_t* src_len, void* dst, const uint64_t* dst_off,
                                             const uint32_t* dst_cap, uint32_t* out_len, int32_t* status,
                                             size_t nblocks, const void* dict, size_t dict_len,
                                             uint32_t acceleration);
B2LZ4_API int b2lz4_compress_fast_batch(b2lz4_ctx* ctx, const void* src, const uint64_t* src_off,
                                        const uint32_t* src_len, void* dst, const uint64_t* dst_off,
                                        const uint32_t* dst_cap, uint32_t* out_len, int32_t* status,
                                        size_t nblocks, uint32_t acceleration);
B2LZ4_API int b2lz4_decompress_safe_batch(b2lz4_ctx* ctx, const void* src, const uint64_t* src_off,
                                          const uint32_t* src_len, void* dst, const uint64_t* dst_off,
                                          const uint32_t* dst_cap, uint32_t* out_len, int32_t* status,
                                          size_t nblocks, const void* dict, size_t dict_len);
B2LZ4_API int b2lz4_compress_hc_batch(b2lz4_ctx* ctx, const void* src, const uint64_t* src_off,
                                      const uint32_t* src_len, void* dst, const uint64_t* dst_off,
                                      const uint32_t* dst_cap, uint32_t* out_len, int32_t* status,
                                      size_t nblocks, int level);
/* compressDestSize for many blocks: src_len[i] = bytes block i may consume, dst_cap[i] = room; consumed[i] /
 * out_len[i] as lz4.compressDestSize returns them.  Every probe of the reference's bisection is one launch of
 * the fast compressor over all blocks.  max_src_len: an upper bound of src_len[] (0 = unknown: 31 probes). */
B2LZ4_API int b2lz4_compress_dest_size_batch_dev(b2lz4_ctx* ctx, const void* src, const uint64_t* src_off,
                                                 const uint32_t* src_len, void* dst, const uint64_t* dst_off,
                                                 const uint32_t* dst_cap, uint32_t* consumed, uint32_t* out_len,
                                                 int32_t* status, size_t nblocks, uint32_t max_src_len,
                                                 void* stream);
B2LZ4_API int b2lz4_compress_dest_size_batch(b2lz4_ctx* ctx, const void* src, const uint64_t* src_off,
                                             const uint32_t* src_len, void* dst, const uint64_t* dst_off,
                                             const uint32_t* dst_cap, uint32_t* consumed, uint32_t* out_len,
                                             int32_t* status, size_t nblocks);
/* XXH32 of a device buffer; asynchronous, result written to *out_dev (device u32). */
B2LZ4_API int b2lz4_xxh32_dev(b2lz4_ctx* ctx, const void* src, size_t n, uint32_t seed, uint32_t* out_dev,
                              void* stream);

/* ------------------------------------------------------------------ frame API ---- */
/* mirrors lz4f.FrameInfo + lz4f.Preferences — reference src/lz4f.zig:106-122 */
typedef struct b2lz4f_prefs {
    uint32_t block_size_id;     /* BlockSizeID: 0 default, 4 max64KB, 5 max256KB, 6 max1MB, 7 max4MB */
    uint32_t block_mode;        /* BlockMode: 0 linked, 1 independent (blocks are independent either
                                   way, exactly like the reference — SURVEY F5) */
    uint32_t content_checksum;  /* ContentChecksum: 0 disabled, 1 enabled */
    uint32_t frame_type;        /* FrameType: 0 frame, 1 skippableFrame (unused by the reference) */
    uint64_t content_size;      /* 0 = unknown */
    uint32_t dict_id;           /* 0 = none; serialised only */
    uint32_t block_checksum;    /* BlockChecksum: 0 disabled, 1 enabled */
    int32_t compression_level;  /* 0 = fast mode (compressFast accel 1); >0 = compressHC(level) */
    uint32_t auto_flush;        /* declared by the reference, never read */
    uint32_t favor_dec_speed;   /* declared by the reference, never read */
} b2lz4f_prefs;

/* Preferences.init() of the README (reference README.md:107) == all-zero defaults (src/lz4f.zig:117) */
B2LZ4_API void b2lz4f_prefs_init(b2lz4f_prefs* prefs);
/* replaces lz4f.compressFrameBound — reference src/lz4f.zig:274-301 (prefs may be NULL) */
B2LZ4_API size_t b2lz4f_compress_frame_bound(size_t src_size, const b2lz4f_prefs* prefs);
/* replaces lz4f.compressFrame — reference src/lz4f.zig:354-446 (host pointers, byte-identical) */
B2LZ4_API int b2lz4f_compress_frame(const void* src, size_t n, void* dst, size_t cap,
                                    const b2lz4f_prefs* prefs, size_t* out);
/* replaces lz4f.decompressFrame — reference src/lz4f.zig:541-638 (host pointers) */
B2LZ4_API int b2lz4f_decompress_frame(const void* src, size_t n, void* dst, size_t cap, size_t* out);
/* replaces lz4f.headerSize — reference src/lz4f.zig:451-480 */
B2LZ4_API int b2lz4f_header_size(const void* src, size_t n, size_t* out);
/* header codec, reference src/lz4f.zig:304-351 / :483-538 (host, a few bytes) */
B2LZ4_API int b2lz4f_write_frame_header(void* dst, size_t cap, const b2lz4f_prefs* prefs, size_t* out);
B2LZ4_API int b2lz4f_parse_frame_header(const void* src, size_t n, b2lz4f_prefs* info, size_t* header_size);

/* same two calls with an explicit context (host pointers; H2D / kernels / D2H are pipelined) */
B2LZ4_API int b2lz4f_compress_frame_ctx(b2lz4_ctx* ctx, const void* src, size_t n, void* dst, size_t cap,
                                        const b2lz4f_prefs* prefs, size_t* out);
B2LZ4_API int b2lz4f_decompress_frame_ctx(b2lz4_ctx* ctx, const void* src, size_t n, void* dst, size_t cap,
                                          size_t* out);

/* Device-resident frame codec (the measured hot path).  src/dst are device pointers.  The call
 * enqueues all kernels on `stream` and synchronises that stream once, to return *out / the status. */
B2LZ4_API int b2lz4f_compress_frame_dev(b2lz4_ctx* ctx, const void* src, size_t n, void* dst, size_t cap,
                                        const b2lz4f_prefs* prefs, size_t* out, void* stream);
B2LZ4_API int b2lz4f_decompress_frame_dev(b2lz4_ctx* ctx, const void* src, size_t n, void* dst, size_t cap,
                                          size_t* out, void* stream);

/* Per-phase device times (ms, CUDA events on the launching stream) of the last *_frame_dev call on
 * this context: [0] block codec kernel, [1] block checksums, [2] scan + assembly (compress) or frame
 * index walk (decompress), [3] content checksum, [4] whole call.  For bench.py's roofline. */
B2LZ4_API int b2lz4_ctx_last_phase_ms(const b2lz4_ctx* ctx, float out_ms[5]);
/* enable (1) / disable (0) the event timing above; off by default */
B2LZ4_API void b2lz4_ctx_set_timing(b2lz4_ctx* ctx, int enabled);

/* ------------------------------------------------------------------ shard API (multi-GPU) ---- */
/* A frame shards by block range: rank k encodes blocks [k*B/G, (k+1)*B/G) into a
 * "body": the concatenated block records `u32 header | payload | [u32 xxh32]` exactly as they
 * appear in the frame (reference src/lz4f.zig:379-430), without frame header / end mark / content
 * checksum.  frame = header(rank 0) ++ body_0 ++ ... ++ body_{G-1} ++ endmark [++ content xxh32]. */
B2LZ4_API int b2lz4f_compress_blocks_dev(b2lz4_ctx* ctx, const void* src, size_t n, void* dst, size_t cap,
                                         const b2lz4f_prefs* prefs, size_t* out, void* stream);
/* Decode a body (device pointers); block_size in bytes; expects exactly the records, no end mark. */
B2LZ4_API int b2lz4f_decompress_blocks_dev(b2lz4_ctx* ctx, const void* src, size_t n, void* dst, size_t cap,
                                           size_t block_size, int block_checksum, size_t* out, void* stream);
/* Block index of a whole frame held in device memory — the header chain lz4f.decompressFrame walks serially
 * (reference src/lz4f.zig:563-589), built by the parallel index kernels.  Parses the frame header (same errors
 * as lz4f.parseFrameHeader), then writes for block i the position of its payload (off_dev[i], right after
 * the 4-byte block header) and the header word (hdr_dev[i], bit 31 = stored raw) for up to `capacity` blocks
 * (device arrays; may be NULL with capacity 0 to only count).  A multi-GPU decode splits the frame on these:
 * blocks [b0, b1) are the body bytes [off[b0] - 4, off[b1] - 4), decodable with b2lz4f_decompress_blocks_dev. */
typedef struct b2lz4f_frame_index {
    uint64_t nblocks;         /* blocks on the chain */
    uint64_t end_pos;         /* position after the end mark (where the content checksum sits, if any) */
    uint64_t content_size;    /* from the header, 0 = absent */
    uint32_t terminal;        /* 0 = end mark seen, 1 = ran off the end without one, 2 = truncated (FrameSizeWrong) */
    uint32_t header_size;
    uint32_t block_size;      /* bytes */
    uint32_t block_checksum;  /* 0 / 1 */
    uint32_t content_checksum;
    uint32_t max_stored;      /* largest block payload */
} b2lz4f_frame_index;
B2LZ4_API int b2lz4f_index_frame_dev(b2lz4_ctx* ctx, const void* src, size_t n, uint64_t* off_dev, uint32_t* hdr_dev,
                                     size_t capacity, b2lz4f_frame_index* info, void* stream);
/* One process, several GPUs (SURVEY section 8b "multi-GPU variants taking ngpus"): lz4f.compressFrame /
 * lz4f.decompressFrame on host slices with the frame sharded by contiguous block range over devices 0..ngpus-1
 * (clamped to the devices present), one host thread and one context per device inside the call, each range on its
 * own GPU's PCIe link.  Same bytes, sizes and errors as the one-GPU calls; frames with fewer than 4 blocks per GPU,
 * foreign layouts and every error case run on one GPU. */
B2LZ4_API int b2lz4f_compress_frame_mgpu(const void* src, size_t n, void* dst, size_t cap, const b2lz4f_prefs* prefs,
                                         int ngpus, size_t* out);
B2LZ4_API int b2lz4f_decompress_frame_mgpu(const void* src, size_t n, void* dst, size_t cap, int ngpus, size_t* out);
/* Running XXH32 state for the content checksum hand-off rank k -> k+1: 4 lanes, the
 * <16-byte tail, and the byte count — 40 bytes, plain data so it can be sent with any transport. */
typedef struct b2lz4_xxh32_state {
    uint32_t v[4];
    uint8_t tail[16];
    uint32_t tail_len;
    uint32_t seed;
    uint64_t total;
} b2lz4_xxh32_state;
B2LZ4_API void b2lz4_xxh32_state_init(b2lz4_xxh32_state* st, uint32_t seed);
/* consume n device bytes (kernel), synchronous on `stream`; state lives on the host */
B2LZ4_API int b2lz4_xxh32_state_update_dev(b2lz4_ctx* ctx, b2lz4_xxh32_state* st, const void* src, size_t n,
                                           void* stream);
B2LZ4_API uint32_t b2lz4_xxh32_state_final(const b2lz4_xxh32_state* st);

/* ------------------------------------------------------------------ streaming trio ---- */
/* README-only API of the reference (README.md:98-122; absent from src/lz4f.zig — SURVEY F4).
 * Contract: begin ++ update* ++ end output == compressFrame(all input) for the same prefs. */
typedef struct b2lz4f_cctx b2lz4f_cctx;
B2LZ4_API int b2lz4f_create_compression_context(b2lz4f_cctx** out);
B2LZ4_API void b2lz4f_free_compression_context(b2lz4f_cctx* cctx);
B2LZ4_API int b2lz4f_compress_begin(b2lz4f_cctx* cctx, void* dst, size_t cap, const b2lz4f_prefs* prefs,
                                    size_t* out);
/* worst-case bytes one update(src_size) or end() call may write */
B2LZ4_API size_t b2lz4f_compress_bound(size_t src_size, const b2lz4f_prefs* prefs);
B2LZ4_API int b2lz4f_compress_update(b2lz4f_cctx* cctx, void* dst, size_t cap, const void* src, size_t n,
                                     size_t* out);
B2LZ4_API int b2lz4f_compress_end(b2lz4f_cctx* cctx, void* dst, size_t cap, size_t* out);

#ifdef __cplusplus
}
#endif
#endif /* B2LZ4_H */
