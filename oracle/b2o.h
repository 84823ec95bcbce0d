/*
 * b2o.h — CPU ORACLE for the LZ4 block/frame hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the reference algorithms (jedisct1/zig-lz4, Zig sources under
 * /root/reference/src).  The reference itself cannot be built in this image (no `zig`), so this
 * restatement is the parity checker and the reported CPU baseline ("kind": "port").
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libb2lz4.so) never links, loads or calls anything in oracle/.
 *
 * PARITY PINNING: the reference holds no golden vectors.  This oracle is pinned by
 *   - the reference's own round-trip assertions restated in tests/ (test.zig, test_compat.zig,
 *     test_lz4f.zig, test_lz4hc.zig inputs),
 *   - stock-decoder acceptance (liblz4.so.1 LZ4_decompress_safe / LZ4F_decompress, pyarrow) in
 *     place of the `lz4` CLI the reference shells out to (src/test_compat.zig:141-254),
 *   - XXH32 against libxxhash.so.0 / python-xxhash (Zig std.hash.XxHash32 is standard XXH32),
 *   - the second-source vectors of SURVEY.md §8(c) (an independent Python restatement, fast path),
 *   - tests/second_source/zlz4_second.py: a second restatement of the WHOLE path (fast, decoder, HC 3..9, frames,
 *     XXH32) written independently from the .zig text, whose committed outputs (tests/golden/
 *     second_source_vectors.json: 324 blocks, 208 frames, 49 decoder exits) this oracle must reproduce
 *     (tests/test_second_source.py).
 * There is still neither a reference binary nor a reference-held vector: "parity unpinned" against
 * reference-binary output in that strict sense, pinned by two independent derivations agreeing, and DESIGN.md §2
 * says so.
 *
 * Status codes (shared numbering with include/b2lz4.h, but deliberately re-declared here so the
 * oracle stays independent of the product headers):
 *   0 ok; 1..6 = lz4.Error members in declaration order (src/lz4.zig:48-55);
 *   100+k = lz4f.Error member k in declaration order (src/lz4f.zig:31-55).
 */
#ifndef B2O_H
#define B2O_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    B2O_OK = 0,
    /* lz4.Error, src/lz4.zig:48-55 */
    B2O_OutputTooSmall = 1,
    B2O_InputTooLarge = 2,
    B2O_CorruptedData = 3,
    B2O_DecompressionFailed = 4,
    B2O_InvalidState = 5,
    B2O_AllocationFailed = 6,
    /* lz4f.Error, src/lz4f.zig:31-55 */
    B2O_F_Generic = 100,
    B2O_F_MaxBlockSizeInvalid = 101,
    B2O_F_BlockModeInvalid = 102,
    B2O_F_ParameterInvalid = 103,
    B2O_F_CompressionLevelInvalid = 104,
    B2O_F_HeaderVersionWrong = 105,
    B2O_F_BlockChecksumInvalid = 106,
    B2O_F_ReservedFlagSet = 107,
    B2O_F_AllocationFailed = 108,
    B2O_F_SrcSizeTooLarge = 109,
    B2O_F_DstMaxSizeTooSmall = 110,
    B2O_F_FrameHeaderIncomplete = 111,
    B2O_F_FrameTypeUnknown = 112,
    B2O_F_FrameSizeWrong = 113,
    B2O_F_SrcPtrWrong = 114,
    B2O_F_DecompressionFailed = 115,
    B2O_F_HeaderChecksumInvalid = 116,
    B2O_F_ContentChecksumInvalid = 117,
    B2O_F_FrameDecodingAlreadyStarted = 118,
    B2O_F_CompressionStateUninitialized = 119,
    B2O_F_ParameterNull = 120,
    B2O_F_MaxCode = 121,
    B2O_F_OutOfMemory = 122,
    /* oracle-only: the reference strategy exists but is outside the hot-path scope:
       HC level 2 (compressMID) and levels 10-12 (compressOptimal) are not restated. */
    B2O_UnsupportedLevel = 201
};

/* ---- block codec (src/lz4.zig) ---- */
size_t b2o_compress_bound(size_t n);                                   /* src/lz4.zig:80-83 */
int b2o_compress_fast(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, uint32_t accel,
                      size_t* out);                                    /* src/lz4.zig:292-447 */
int b2o_decompress_safe(const uint8_t* src, size_t n, uint8_t* dst, size_t cap,
                        size_t* out);                                  /* src/lz4.zig:257 */
int b2o_decompress_safe_using_dict(const uint8_t* src, size_t n, uint8_t* dst, size_t cap,
                                   const uint8_t* dict, size_t dict_len,
                                   size_t* out);                       /* src/lz4.zig:960-964 */

/* *src_size: in = bytes available, out = bytes consumed; dst is left as the reference leaves it (last probe) */
int b2o_compress_dest_size(const uint8_t* src, uint8_t* dst, size_t cap, size_t* src_size,
                           size_t* out);                               /* src/lz4.zig:551-616 */

/* ---- HC (src/lz4hc.zig), levels routed to compressHashChain only (3..9; <2 -> 9) ---- */
int b2o_compress_hc(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, int level,
                    size_t* out);                                      /* src/lz4hc.zig:1440-1489 */
/* number of times the F8 guard (u32 underflow at src/lz4hc.zig:636) fired in this thread */
uint64_t b2o_hc_f8_guard_hits(void);

/* ---- XXH32 (Zig std.hash.XxHash32 == standard XXH32) ---- */
typedef struct {
    uint32_t v[4];
    uint8_t buf[16];
    uint32_t buf_len;
    uint64_t total;
    uint32_t seed;
} b2o_xxh32_state;
uint32_t b2o_xxh32(const void* p, size_t n, uint32_t seed);
void b2o_xxh32_init(b2o_xxh32_state* s, uint32_t seed);
void b2o_xxh32_update(b2o_xxh32_state* s, const void* p, size_t n);
uint32_t b2o_xxh32_final(const b2o_xxh32_state* s);

/* ---- frame (src/lz4f.zig) ---- */
typedef struct {
    uint32_t block_size_id;     /* 0 default, 4 64K, 5 256K, 6 1M, 7 4M   (src/lz4f.zig:64-70) */
    uint32_t block_mode;        /* 0 linked, 1 independent               (src/lz4f.zig:82-85) */
    uint32_t content_checksum;  /* 0/1                                   (src/lz4f.zig:88-91) */
    uint32_t frame_type;        /* 0 frame, 1 skippable                  (src/lz4f.zig:100-103) */
    uint64_t content_size;      /* 0 = unknown                           (src/lz4f.zig:111) */
    uint32_t dict_id;           /*                                        (src/lz4f.zig:112) */
    uint32_t block_checksum;    /* 0/1                                   (src/lz4f.zig:113) */
    int32_t compression_level;  /* 0 = fast                              (src/lz4f.zig:119) */
    uint32_t auto_flush;        /* declared, never read                  (src/lz4f.zig:120) */
    uint32_t favor_dec_speed;   /* declared, never read                  (src/lz4f.zig:121) */
} b2o_prefs;

void b2o_prefs_default(b2o_prefs* p);                                   /* src/lz4f.zig:106-122 */
size_t b2o_compress_frame_bound(size_t n, const b2o_prefs* p);          /* src/lz4f.zig:274-301 */
int b2o_compress_frame(const uint8_t* src, size_t n, uint8_t* dst, size_t cap,
                       const b2o_prefs* p, size_t* out);                /* src/lz4f.zig:354-446 */
int b2o_decompress_frame(const uint8_t* src, size_t n, uint8_t* dst, size_t cap,
                         size_t* out);                                  /* src/lz4f.zig:541-638 */
int b2o_header_size(const uint8_t* src, size_t n, size_t* out);         /* src/lz4f.zig:451-480 */
int b2o_write_frame_header(uint8_t* dst, size_t cap, const b2o_prefs* p,
                           size_t* out);                                /* src/lz4f.zig:304-351 */
int b2o_parse_frame_header(const uint8_t* src, size_t n, b2o_prefs* info,
                           size_t* size);                               /* src/lz4f.zig:483-538 */

/* ---- threaded batch drivers: the reported CPU baseline (one block per task) ---- */
/* mode: 0 = compress_fast(accel=param), 1 = decompress_safe, 2 = compress_hc(level=param) */
int b2o_batch(int mode, int param, const uint8_t* src, const uint64_t* src_off,
              const uint32_t* src_len, uint8_t* dst, const uint64_t* dst_off,
              const uint32_t* dst_cap, uint32_t* out_len, int32_t* status, size_t nblocks,
              int nthreads);
/* frame round trip on many threads: compresses src as one frame with blocks distributed over
   threads (block bodies are independent; assembly is sequential), returns the same bytes as
   b2o_compress_frame. */
int b2o_compress_frame_mt(const uint8_t* src, size_t n, uint8_t* dst, size_t cap,
                          const b2o_prefs* p, size_t* out, int nthreads);
int b2o_decompress_frame_mt(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out,
                            int nthreads);
int b2o_hardware_threads(void);

#ifdef __cplusplus
}
#endif
#endif
