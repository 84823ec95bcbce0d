/*
 * oracle_lz4f.c — CPU ORACLE (test infrastructure only; see b2o.h).
 * Literal C restatement of the reference frame codec, /root/reference/src/lz4f.zig.
 * Quirks kept: blocks are always compressed independently whatever blockMode says;
 * default prefs emit FLG 0x40 (the "linked" flag) anyway; decompressFrame ignores block mode,
 * contentSize and dictID, and does not bound a decoded block by blockSize.
 */
#include "b2o.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define MAGICNUMBER 0x184D2204u           /* src/lz4f.zig:12 */
#define MAGIC_SKIPPABLE_START 0x184D2A50u /* src/lz4f.zig:15 */
#define MAGIC_SKIPPABLE_MASK 0xFFFFFFF0u  /* src/lz4f.zig:16 */
#define HEADER_SIZE_MIN 7                 /* src/lz4f.zig:19 */
#define HEADER_SIZE_MAX 19                /* src/lz4f.zig:20 */
#define MIN_SIZE_TO_KNOW_HEADER_LENGTH 5  /* src/lz4f.zig:21 */
#define BLOCK_HEADER_SIZE 4
#define BLOCK_CHECKSUM_SIZE 4
#define CONTENT_CHECKSUM_SIZE 4
#define ENDMARK_SIZE 4

static inline uint32_t rd32(const uint8_t* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static inline void wr32(uint8_t* p, uint32_t v) {
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}
static inline uint64_t rd64(const uint8_t* p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }
static inline void wr64(uint8_t* p, uint64_t v) { wr32(p, (uint32_t)v); wr32(p + 4, (uint32_t)(v >> 32)); }

void b2o_prefs_default(b2o_prefs* p) { memset(p, 0, sizeof *p); }  /* src/lz4f.zig:106-122 */

/* BlockSizeID.toBlockSize, src/lz4f.zig:71-78.  The Zig enum(u3) only admits 0,4,5,6,7. */
static int to_block_size(uint32_t id, size_t* bs) {
    switch (id) {
        case 0: case 4: *bs = 64 * 1024; return 0;
        case 5: *bs = 256 * 1024; return 0;
        case 6: *bs = 1024 * 1024; return 0;
        case 7: *bs = 4 * 1024 * 1024; return 0;
        default: return -1;
    }
}

static uint8_t header_checksum(const uint8_t* d, size_t n) {        /* src/lz4f.zig:138-141 */
    return (uint8_t)((b2o_xxh32(d, n, 0) >> 8) & 0xFF);
}

static uint8_t encodeFLG(const b2o_prefs* p) {                      /* src/lz4f.zig:152-184 */
    uint8_t flg = 0x40;
    if (p->block_mode == 1) flg |= 0x20;
    if (p->block_checksum == 1) flg |= 0x10;
    if (p->content_size != 0) flg |= 0x08;
    if (p->content_checksum == 1) flg |= 0x04;
    if (p->dict_id != 0) flg |= 0x01;
    return flg;
}

static uint8_t encodeBD(uint32_t id) {                              /* src/lz4f.zig:224-232 */
    uint8_t v = (id == 0 || id == 4) ? 4 : (uint8_t)id;
    return (uint8_t)(v << 4);
}

size_t b2o_compress_frame_bound(size_t srcSize, const b2o_prefs* prefs) {  /* src/lz4f.zig:274-301 */
    b2o_prefs d;
    if (!prefs) { b2o_prefs_default(&d); prefs = &d; }
    size_t blockSize;
    if (to_block_size(prefs->block_size_id, &blockSize)) blockSize = 65536;
    size_t result = HEADER_SIZE_MAX;
    size_t numBlocks = (srcSize + blockSize - 1) / blockSize;
    size_t per = BLOCK_HEADER_SIZE + b2o_compress_bound(blockSize) + (prefs->block_checksum == 1 ? BLOCK_CHECKSUM_SIZE : 0);
    result += numBlocks * per;
    result += ENDMARK_SIZE;
    if (prefs->content_checksum == 1) result += CONTENT_CHECKSUM_SIZE;
    return result;
}

int b2o_write_frame_header(uint8_t* dst, size_t cap, const b2o_prefs* p, size_t* out) {  /* :304-351 */
    if (cap < HEADER_SIZE_MIN) return B2O_F_DstMaxSizeTooSmall;
    size_t pos = 0;
    wr32(dst, MAGICNUMBER); pos += 4;
    dst[pos++] = encodeFLG(p);
    dst[pos++] = encodeBD(p->block_size_id);
    const size_t headerStart = 4;
    if (p->content_size != 0) {
        if (cap < pos + 8) return B2O_F_DstMaxSizeTooSmall;
        wr64(dst + pos, p->content_size); pos += 8;
    }
    if (p->dict_id != 0) {
        if (cap < pos + 4) return B2O_F_DstMaxSizeTooSmall;
        wr32(dst + pos, p->dict_id); pos += 4;
    }
    /* NOTE: the reference indexes dst[pos] for the HC byte without a capacity test (:347); every
       caller passes >= HEADER_SIZE_MAX bytes, so the difference is unobservable there. */
    if (cap < pos + 1) return B2O_F_DstMaxSizeTooSmall;
    dst[pos] = header_checksum(dst + headerStart, pos - headerStart);
    pos += 1;
    *out = pos;
    return B2O_OK;
}

static int map_compression_error(int e) {                           /* src/lz4f.zig:144-149 */
    if (e == B2O_UnsupportedLevel) return e;
    return e == B2O_OutputTooSmall ? B2O_F_DstMaxSizeTooSmall : B2O_F_Generic;
}

/* One block body: compress, decide raw/compressed (src/lz4f.zig:393-408).  Writes the payload of the
 * block (compressed bytes or raw copy) to `body` and returns header word + stored size. */
static int compress_block_body(const uint8_t* srcBlock, size_t blockLen, uint8_t* body, size_t bodyCap,
                               int level, uint32_t* headerWord, size_t* actualSize) {
    size_t csize = 0;
    int rc = level > 0 ? b2o_compress_hc(srcBlock, blockLen, body, bodyCap, level, &csize)
                       : b2o_compress_fast(srcBlock, blockLen, body, bodyCap, 1, &csize);
    if (rc != B2O_OK) return map_compression_error(rc);
    int storeUncompressed = csize >= blockLen;                      /* :407 */
    size_t actual = storeUncompressed ? blockLen : csize;
    uint32_t hw = (uint32_t)actual;
    if (storeUncompressed) {
        hw |= 0x80000000u;
        memcpy(body, srcBlock, blockLen);                           /* :416 */
    }
    *headerWord = hw;
    *actualSize = actual;
    return B2O_OK;
}

int b2o_compress_frame(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, const b2o_prefs* prefs,
                       size_t* out) {                               /* src/lz4f.zig:354-446 */
    b2o_prefs d;
    if (!prefs) { b2o_prefs_default(&d); prefs = &d; }
    *out = 0;
    size_t requiredSize = b2o_compress_frame_bound(n, prefs);
    if (cap < requiredSize) return B2O_F_DstMaxSizeTooSmall;        /* :364 */
    size_t dstPos = 0;
    int rc = b2o_write_frame_header(dst, cap, prefs, &dstPos);      /* :369 */
    if (rc) return rc;
    size_t blockSize;
    if (to_block_size(prefs->block_size_id, &blockSize)) return B2O_F_MaxBlockSizeInvalid;
    b2o_xxh32_state cs;
    b2o_xxh32_init(&cs, 0);                                         /* :375 */
    size_t srcPos = 0;
    while (srcPos < n) {                                            /* :379 */
        size_t blockLen = n - srcPos < blockSize ? n - srcPos : blockSize;
        const uint8_t* srcBlock = src + srcPos;
        if (prefs->content_checksum == 1) b2o_xxh32_update(&cs, srcBlock, blockLen);
        size_t blockStart = dstPos + BLOCK_HEADER_SIZE;
        uint32_t hw; size_t actual;
        rc = compress_block_body(srcBlock, blockLen, dst + blockStart, cap - blockStart,
                                 prefs->compression_level, &hw, &actual);
        if (rc) return rc;
        wr32(dst + dstPos, hw);                                     /* :418 */
        dstPos = blockStart + actual;
        if (prefs->block_checksum == 1) {                           /* :422 */
            wr32(dst + dstPos, b2o_xxh32(dst + blockStart, actual, 0));
            dstPos += BLOCK_CHECKSUM_SIZE;
        }
        srcPos += blockLen;
    }
    wr32(dst + dstPos, 0); dstPos += ENDMARK_SIZE;                  /* :433 */
    if (prefs->content_checksum == 1) {                             /* :437 */
        wr32(dst + dstPos, b2o_xxh32_final(&cs));
        dstPos += CONTENT_CHECKSUM_SIZE;
    }
    *out = dstPos;
    return B2O_OK;
}

int b2o_header_size(const uint8_t* src, size_t n, size_t* out) {    /* src/lz4f.zig:451-480 */
    if (n < MIN_SIZE_TO_KNOW_HEADER_LENGTH) return B2O_F_FrameHeaderIncomplete;
    uint32_t magic = rd32(src);
    if (magic != MAGICNUMBER) {
        if ((magic & MAGIC_SKIPPABLE_MASK) == MAGIC_SKIPPABLE_START) { *out = 8; return B2O_OK; }
        return B2O_F_FrameTypeUnknown;
    }
    uint8_t flg = src[4];
    size_t size = 7;
    if (flg & 0x08) size += 8;
    if (flg & 0x01) size += 4;
    *out = size;
    return B2O_OK;
}

int b2o_parse_frame_header(const uint8_t* src, size_t n, b2o_prefs* info, size_t* size) {  /* :483-538 */
    if (n < HEADER_SIZE_MIN) return B2O_F_FrameHeaderIncomplete;
    if (rd32(src) != MAGICNUMBER) return B2O_F_FrameTypeUnknown;
    size_t pos = 4;
    uint8_t flg = src[pos];
    b2o_prefs_default(info);
    /* decodeFLG :187-221 */
    if (((flg >> 6) & 3) != 1) return B2O_F_HeaderVersionWrong;
    if (flg & 0x02) return B2O_F_ReservedFlagSet;
    info->block_mode = (flg & 0x20) ? 1 : 0;
    info->block_checksum = (flg & 0x10) ? 1 : 0;
    info->content_checksum = (flg & 0x04) ? 1 : 0;
    pos += 1;
    /* decodeBD :235-249 */
    uint8_t bd = src[pos];
    if (bd & 0x8F) return B2O_F_ReservedFlagSet;
    switch ((bd >> 4) & 7) {
        case 0: case 4: info->block_size_id = 4; break;
        case 5: info->block_size_id = 5; break;
        case 6: info->block_size_id = 6; break;
        case 7: info->block_size_id = 7; break;
        default: return B2O_F_MaxBlockSizeInvalid;
    }
    pos += 1;
    const size_t headerStart = 4;
    if (flg & 0x08) {
        if (n < pos + 8) return B2O_F_FrameHeaderIncomplete;
        info->content_size = rd64(src + pos); pos += 8;
    }
    if (flg & 0x01) {
        if (n < pos + 4) return B2O_F_FrameHeaderIncomplete;
        info->dict_id = rd32(src + pos); pos += 4;
    }
    if (n < pos + 1) return B2O_F_FrameHeaderIncomplete;
    if (src[pos] != header_checksum(src + headerStart, pos - headerStart)) return B2O_F_HeaderChecksumInvalid;
    pos += 1;
    *size = pos;
    return B2O_OK;
}

int b2o_decompress_frame(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out) {  /* :541-638 */
    *out = 0;
    b2o_prefs info; size_t srcPos = 0;
    int rc = b2o_parse_frame_header(src, n, &info, &srcPos);
    if (rc) return rc;
    size_t dstPos = 0;
    b2o_xxh32_state cs;
    b2o_xxh32_init(&cs, 0);
    while (srcPos < n) {                                            /* :563 */
        if (srcPos + BLOCK_HEADER_SIZE > n) return B2O_F_FrameSizeWrong;
        uint32_t blockHeader = rd32(src + srcPos);
        srcPos += BLOCK_HEADER_SIZE;
        if (blockHeader == 0) break;                                /* :573 */
        int isUncompressed = (blockHeader & 0x80000000u) != 0;
        size_t blockDataSize = blockHeader & 0x7FFFFFFFu;
        if (srcPos + blockDataSize > n) return B2O_F_FrameSizeWrong; /* :582 */
        const uint8_t* blockData = src + srcPos;
        srcPos += blockDataSize;
        if (info.block_checksum == 1) {                             /* :590 */
            if (srcPos + BLOCK_CHECKSUM_SIZE > n) return B2O_F_FrameSizeWrong;
            if (rd32(src + srcPos) != b2o_xxh32(blockData, blockDataSize, 0)) return B2O_F_BlockChecksumInvalid;
            srcPos += BLOCK_CHECKSUM_SIZE;
        }
        size_t decompressedSize;
        if (isUncompressed) {                                       /* :603 */
            if (dstPos + blockDataSize > cap) return B2O_F_DstMaxSizeTooSmall;
            memcpy(dst + dstPos, blockData, blockDataSize);
            decompressedSize = blockDataSize;
        } else {
            if (b2o_decompress_safe(blockData, blockDataSize, dst + dstPos, cap - dstPos, &decompressedSize))
                return B2O_F_DecompressionFailed;                   /* :610-612 */
        }
        if (info.content_checksum == 1) b2o_xxh32_update(&cs, dst + dstPos, decompressedSize);
        dstPos += decompressedSize;
    }
    if (info.content_checksum == 1) {                               /* :625 */
        if (srcPos + CONTENT_CHECKSUM_SIZE > n) return B2O_F_FrameSizeWrong;
        if (rd32(src + srcPos) != b2o_xxh32_final(&cs)) return B2O_F_ContentChecksumInvalid;
        srcPos += CONTENT_CHECKSUM_SIZE;
    }
    *out = dstPos;
    return B2O_OK;
}

/* ============================ threaded drivers (CPU baseline) ============================ */
/* "One thread per block across all cores" (BASELINE.json north_star): tasks are pulled from a
 * shared atomic counter; each task is one independent block.  Not part of the reference. */

int b2o_hardware_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n < 1 ? 1 : (int)n;
}

typedef struct {
    int mode, param;
    const uint8_t* src; const uint64_t* src_off; const uint32_t* src_len;
    uint8_t* dst; const uint64_t* dst_off; const uint32_t* dst_cap;
    uint32_t* out_len; int32_t* status; size_t nblocks;
    size_t next;
} batch_job;

static void* batch_worker(void* arg) {
    batch_job* j = (batch_job*)arg;
    for (;;) {
        size_t i = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (i >= j->nblocks) break;
        size_t out = 0; int rc;
        const uint8_t* s = j->src + j->src_off[i];
        uint8_t* d = j->dst + j->dst_off[i];
        if (j->mode == 0) rc = b2o_compress_fast(s, j->src_len[i], d, j->dst_cap[i], (uint32_t)j->param, &out);
        else if (j->mode == 1) rc = b2o_decompress_safe(s, j->src_len[i], d, j->dst_cap[i], &out);
        else rc = b2o_compress_hc(s, j->src_len[i], d, j->dst_cap[i], j->param, &out);
        j->out_len[i] = (uint32_t)out;
        j->status[i] = rc;
    }
    return NULL;
}

static void run_threads(void* (*fn)(void*), void* arg, int nthreads) {
    if (nthreads <= 1) { fn(arg); return; }
    pthread_t* t = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
    int started = 0;
    for (int i = 0; i < nthreads; i++) if (pthread_create(&t[started], NULL, fn, arg) == 0) started++;
    if (started == 0) fn(arg);
    for (int i = 0; i < started; i++) pthread_join(t[i], NULL);
    free(t);
}

int b2o_batch(int mode, int param, const uint8_t* src, const uint64_t* src_off, const uint32_t* src_len,
              uint8_t* dst, const uint64_t* dst_off, const uint32_t* dst_cap, uint32_t* out_len,
              int32_t* status, size_t nblocks, int nthreads) {
    batch_job j = {mode, param, src, src_off, src_len, dst, dst_off, dst_cap, out_len, status, nblocks, 0};
    run_threads(batch_worker, &j, nthreads);
    return B2O_OK;
}

/* Multi-threaded frame compress: block bodies are computed in parallel into bound-sized slots, then
 * assembled sequentially in block order — byte-identical to b2o_compress_frame by construction
 * (blocks are independent); the content checksum stays one serial chain (F11). */
typedef struct {
    const uint8_t* src; size_t n; size_t blockSize; int level;
    uint8_t* slots; size_t slotStride; uint32_t* hw; size_t* actual; int* rc; size_t nblocks; size_t next;
} fc_job;

static void* fc_worker(void* arg) {
    fc_job* j = (fc_job*)arg;
    for (;;) {
        size_t i = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (i >= j->nblocks) break;
        size_t off = i * j->blockSize;
        size_t len = j->n - off < j->blockSize ? j->n - off : j->blockSize;
        j->rc[i] = compress_block_body(j->src + off, len, j->slots + i * j->slotStride, j->slotStride, j->level,
                                       &j->hw[i], &j->actual[i]);
    }
    return NULL;
}

/* assembly of one wave, in parallel: record i goes to its prefix-summed position (header word, stored bytes, block
 * checksum) — the bytes are those of the serial loop at src/lz4f.zig:406-427, only the order of the copies differs */
typedef struct {
    const fc_job* w; uint8_t* dst; const size_t* pos; int blockChecksum; size_t next;
} fa_job;

static void* fa_worker(void* arg) {
    fa_job* a = (fa_job*)arg;
    for (;;) {
        size_t i = __atomic_fetch_add(&a->next, 1, __ATOMIC_RELAXED);
        if (i >= a->w->nblocks) break;
        uint8_t* d = a->dst + a->pos[i];
        wr32(d, a->w->hw[i]);
        memcpy(d + 4, a->w->slots + i * a->w->slotStride, a->w->actual[i]);
        if (a->blockChecksum) wr32(d + 4 + a->w->actual[i], b2o_xxh32(d + 4, a->w->actual[i], 0));
    }
    return NULL;
}

/* the content checksum is one serial chain over the raw input; it does not depend on the compressed
 * bytes, so it runs on its own thread next to the block workers */
typedef struct { const uint8_t* src; size_t n; uint32_t sum; } cc_job;
static void* cc_worker(void* arg) {
    cc_job* c = (cc_job*)arg;
    c->sum = b2o_xxh32(c->src, c->n, 0);
    return NULL;
}

/* scratch of the multi-threaded writer, kept between calls (a fresh 256 MiB malloc costs its page faults every call) */
static pthread_mutex_t g_scratch_mu = PTHREAD_MUTEX_INITIALIZER;
static uint8_t* g_scratch = NULL;
static size_t g_scratch_cap = 0;

int b2o_compress_frame_mt(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, const b2o_prefs* prefs,
                          size_t* out, int nthreads) {
    b2o_prefs d;
    if (!prefs) { b2o_prefs_default(&d); prefs = &d; }
    *out = 0;
    if (cap < b2o_compress_frame_bound(n, prefs)) return B2O_F_DstMaxSizeTooSmall;
    size_t dstPos = 0;
    int rc = b2o_write_frame_header(dst, cap, prefs, &dstPos);
    if (rc) return rc;
    size_t blockSize;
    if (to_block_size(prefs->block_size_id, &blockSize)) return B2O_F_MaxBlockSizeInvalid;
    size_t nblocks = (n + blockSize - 1) / blockSize;
    fc_job j;
    memset(&j, 0, sizeof j);
    j.src = src; j.n = n; j.blockSize = blockSize; j.level = prefs->compression_level;
    j.slotStride = b2o_compress_bound(blockSize);
    j.nblocks = nblocks;
    /* process in waves so the scratch stays bounded (<= 256 MiB, or one slot per thread for 4 MiB blocks) */
    size_t wave = (256u << 20) / j.slotStride;
    if (wave < (size_t)nthreads * 2) wave = (size_t)nthreads * 2;
    if (wave < 1) wave = 1;
    if (wave > nblocks) wave = nblocks ? nblocks : 1;
    pthread_mutex_lock(&g_scratch_mu);
    if (g_scratch_cap < wave * j.slotStride) {
        free(g_scratch);
        g_scratch = (uint8_t*)malloc(wave * j.slotStride);
        g_scratch_cap = g_scratch ? wave * j.slotStride : 0;
    }
    j.slots = g_scratch;
    j.hw = (uint32_t*)malloc(wave * sizeof(uint32_t));
    j.actual = (size_t*)malloc(wave * sizeof(size_t));
    j.rc = (int*)malloc(wave * sizeof(int));
    size_t* pos = (size_t*)malloc(wave * sizeof(size_t));
    if (!j.slots || !j.hw || !j.actual || !j.rc || !pos) {
        free(j.hw); free(j.actual); free(j.rc); free(pos);
        pthread_mutex_unlock(&g_scratch_mu);
        return B2O_F_AllocationFailed;
    }
    cc_job cc = {src, n, 0};
    pthread_t cct;
    int cc_threaded = 0;
    if (prefs->content_checksum == 1) {
        if (nthreads > 1 && pthread_create(&cct, NULL, cc_worker, &cc) == 0) cc_threaded = 1;
        else cc_worker(&cc);
    }
    rc = B2O_OK;
    for (size_t base = 0; base < nblocks && rc == B2O_OK; base += wave) {
        size_t cnt = nblocks - base < wave ? nblocks - base : wave;
        fc_job w = j;
        w.src = src + base * blockSize; w.n = n - base * blockSize; w.nblocks = cnt; w.next = 0;
        run_threads(fc_worker, &w, nthreads);
        for (size_t i = 0; i < cnt; i++) {
            if (w.rc[i]) { rc = w.rc[i]; break; }
            pos[i] = dstPos;
            dstPos += 4 + w.actual[i] + (prefs->block_checksum == 1 ? 4 : 0);
        }
        if (rc) break;
        fa_job a = {&w, dst, pos, prefs->block_checksum == 1, 0};
        run_threads(fa_worker, &a, nthreads);
    }
    if (cc_threaded) pthread_join(cct, NULL);
    free(j.hw); free(j.actual); free(j.rc); free(pos);
    pthread_mutex_unlock(&g_scratch_mu);
    if (rc) return rc;
    wr32(dst + dstPos, 0); dstPos += 4;
    if (prefs->content_checksum == 1) { wr32(dst + dstPos, cc.sum); dstPos += 4; }
    *out = dstPos;
    return B2O_OK;
}

/* Multi-threaded frame decode for well-formed frames whose non-final blocks decode to exactly
 * blockSize (everything our writer, the reference and the stock CLI produce): walk the header chain
 * (serial), decode blocks in parallel at i*blockSize, verify.  Falls back to the
 * sequential restatement whenever anything looks unusual, so results always equal
 * b2o_decompress_frame. */
typedef struct {
    const uint8_t* src; uint8_t* dst; size_t cap; size_t blockSize; int blockChecksum;
    const uint64_t* off; const uint32_t* hdr; size_t nblocks; size_t next; int bad;
    uint32_t* outLen;
} fd_job;

static void* fd_worker(void* arg) {
    fd_job* j = (fd_job*)arg;
    for (;;) {
        size_t i = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (i >= j->nblocks) break;
        uint32_t h = j->hdr[i];
        size_t sz = h & 0x7FFFFFFFu;
        const uint8_t* data = j->src + j->off[i];
        size_t dpos = i * j->blockSize;
        size_t room = dpos < j->cap ? j->cap - dpos : 0;
        if (room > j->blockSize) room = j->blockSize;
        if (j->blockChecksum && rd32(data + sz) != b2o_xxh32(data, sz, 0)) { __atomic_store_n(&j->bad, 1, __ATOMIC_RELAXED); continue; }
        size_t out = 0;
        if (h & 0x80000000u) {
            if (sz > room) { __atomic_store_n(&j->bad, 1, __ATOMIC_RELAXED); continue; }
            memcpy(j->dst + dpos, data, sz);
            out = sz;
        } else if (room == 0 || b2o_decompress_safe(data, sz, j->dst + dpos, room, &out)) {
            __atomic_store_n(&j->bad, 1, __ATOMIC_RELAXED);
            continue;
        }
        j->outLen[i] = (uint32_t)out;
    }
    return NULL;
}

int b2o_decompress_frame_mt(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out, int nthreads) {
    b2o_prefs info; size_t srcPos = 0;
    if (b2o_parse_frame_header(src, n, &info, &srcPos)) return b2o_decompress_frame(src, n, dst, cap, out);
    size_t blockSize;
    if (to_block_size(info.block_size_id, &blockSize)) return b2o_decompress_frame(src, n, dst, cap, out);
    size_t capBlocks = 1024, nb = 0;
    uint64_t* off = (uint64_t*)malloc(capBlocks * sizeof(uint64_t));
    uint32_t* hdr = (uint32_t*)malloc(capBlocks * sizeof(uint32_t));
    int clean = 0;
    while (srcPos < n) {
        if (srcPos + 4 > n) break;
        uint32_t h = rd32(src + srcPos); srcPos += 4;
        if (h == 0) { clean = 1; break; }
        size_t sz = h & 0x7FFFFFFFu;
        if (srcPos + sz + (info.block_checksum ? 4 : 0) > n) break;
        if (nb == capBlocks) {
            capBlocks *= 2;
            off = (uint64_t*)realloc(off, capBlocks * sizeof(uint64_t));
            hdr = (uint32_t*)realloc(hdr, capBlocks * sizeof(uint32_t));
        }
        off[nb] = srcPos; hdr[nb] = h; nb++;
        srcPos += sz + (info.block_checksum ? 4 : 0);
    }
    int ok = clean;
    uint32_t* outLen = (uint32_t*)calloc(nb ? nb : 1, sizeof(uint32_t));
    if (ok) {
        fd_job j = {src, dst, cap, blockSize, (int)info.block_checksum, off, hdr, nb, 0, 0, outLen};
        run_threads(fd_worker, &j, nthreads);
        if (j.bad) ok = 0;
        for (size_t i = 0; ok && i + 1 < nb; i++) if (outLen[i] != blockSize) ok = 0;
    }
    size_t total = 0;
    if (ok) {
        total = nb ? (nb - 1) * blockSize + outLen[nb - 1] : 0;
        if (info.content_checksum == 1) {
            if (srcPos + 4 > n || rd32(src + srcPos) != b2o_xxh32(dst, total, 0)) ok = 0;
        }
    }
    free(off); free(hdr); free(outLen);
    if (!ok) return b2o_decompress_frame(src, n, dst, cap, out);
    *out = total;
    return B2O_OK;
}
