/*
 * oracle_lz4.c — CPU ORACLE (test infrastructure only; see b2o.h header note).
 * Literal C restatement of the reference block codec, /root/reference/src/lz4.zig.
 * Every function cites the Zig lines it follows.  Deliberately keeps the reference's quirks:
 *   - position 0 is never matchable (table value 0 means "empty"), src/lz4.zig:345
 *   - the step schedule probes ip0+1 sixty-four times, src/lz4.zig:321-355
 *   - no look-ahead hash, no backward catch-up, byte-at-a-time extension, src/lz4.zig:401-413
 */
#include "b2o.h"
#include <string.h>

#define MINMATCH 4                      /* src/lz4.zig:12 */
#define LASTLITERALS 5                  /* src/lz4.zig:14 */
#define MFLIMIT 12                      /* src/lz4.zig:15 */
#define ML_BITS 4                       /* src/lz4.zig:18 */
#define ML_MASK 15u                     /* src/lz4.zig:19 */
#define RUN_MASK 15u                    /* src/lz4.zig:21 */
#define LZ4_MAX_INPUT_SIZE 0x7E000000u  /* src/lz4.zig:23 */
#define LZ4_DISTANCE_ABSOLUTE_MAX 65535u /* src/lz4.zig:24 */
#define LZ4_HASHLOG 12                  /* src/lz4.zig:31 */
#define LZ4_HASH_SIZE_U32 4096          /* src/lz4.zig:33 */
#define ACCELERATION_MAX 65537u         /* src/lz4.zig:36 */
#define HASH_MULTIPLIER 2654435761u     /* src/lz4.zig:44 */

static inline uint32_t rd32(const uint8_t* p) {  /* src/lz4.zig:65-67 */
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static inline uint32_t hash4(uint32_t seq) {     /* src/lz4.zig:75-77 */
    return (uint32_t)(seq * HASH_MULTIPLIER) >> ((MINMATCH * 8) - LZ4_HASHLOG);
}

size_t b2o_compress_bound(size_t n) {            /* src/lz4.zig:80-83 */
    if (n > LZ4_MAX_INPUT_SIZE) return 0;
    return n + (n / 255) + 16;
}

/* src/lz4.zig:449-482 */
static int compress_as_literals(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out) {
    size_t literalLength = n, op = 0;
    if (cap < 1) return B2O_OutputTooSmall;
    if (literalLength >= RUN_MASK) {
        dst[op++] = (uint8_t)(RUN_MASK << ML_BITS);
        size_t len = literalLength - RUN_MASK;
        while (len >= 255) {
            if (op >= cap) return B2O_OutputTooSmall;
            dst[op++] = 255;
            len -= 255;
        }
        if (op >= cap) return B2O_OutputTooSmall;
        dst[op++] = (uint8_t)len;
    } else {
        dst[op++] = (uint8_t)(literalLength << ML_BITS);
    }
    if (op + literalLength > cap) return B2O_OutputTooSmall;
    memcpy(dst + op, src, literalLength);
    op += literalLength;
    *out = op;
    return B2O_OK;
}

/* src/lz4.zig:484-519 */
static int finish_compression(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t anchor,
                              size_t op, size_t* out) {
    size_t literalLength = n - anchor;
    size_t outPos = op;
    if (literalLength == 0) { *out = outPos; return B2O_OK; }
    if (outPos >= cap) return B2O_OutputTooSmall;
    if (literalLength >= RUN_MASK) {
        dst[outPos++] = (uint8_t)(RUN_MASK << ML_BITS);
        size_t len = literalLength - RUN_MASK;
        while (len >= 255) {
            if (outPos >= cap) return B2O_OutputTooSmall;
            dst[outPos++] = 255;
            len -= 255;
        }
        if (outPos >= cap) return B2O_OutputTooSmall;
        dst[outPos++] = (uint8_t)len;
    } else {
        dst[outPos++] = (uint8_t)(literalLength << ML_BITS);
    }
    if (outPos + literalLength > cap) return B2O_OutputTooSmall;
    memcpy(dst + outPos, src + anchor, literalLength);
    outPos += literalLength;
    *out = outPos;
    return B2O_OK;
}

/* src/lz4.zig:292-447 (compressFast); compressDefault :283 is accel = 1 */
int b2o_compress_fast(const uint8_t* src, size_t srcSize, uint8_t* dst, size_t cap, uint32_t acceleration,
                      size_t* out) {
    *out = 0;
    if (srcSize > LZ4_MAX_INPUT_SIZE) return B2O_InputTooLarge;       /* :296 */
    if (srcSize == 0) return B2O_OK;                                   /* :299 */
    if (srcSize < MFLIMIT + 1) return compress_as_literals(src, srcSize, dst, cap, out); /* :302 */

    uint32_t table[LZ4_HASH_SIZE_U32];                                 /* :307 HashTable.init() */
    memset(table, 0, sizeof table);

    size_t ip = 0, op = 0, anchor = 0;
    const size_t mflimitPlusOne = srcSize - MFLIMIT;                   /* :313 */
    const size_t matchLimit = srcSize - LASTLITERALS;                  /* :314 */
    ip += 1;                                                           /* :317 */

    while (ip < mflimitPlusOne) {                                      /* :320 */
        uint32_t accel = acceleration < 1 ? 1 : (acceleration > ACCELERATION_MAX ? ACCELERATION_MAX : acceleration);
        size_t step = accel;
        size_t searchMatchNb = accel;
        size_t match;
        size_t forwardIp = ip;
        for (;;) {                                                     /* :329 */
            ip = forwardIp;
            forwardIp += step;
            step = searchMatchNb >> 6;
            searchMatchNb += 1;
            if (forwardIp > mflimitPlusOne)                            /* :335 */
                return finish_compression(src, srcSize, dst, cap, anchor, op, out);
            uint32_t h = hash4(rd32(src + ip));
            match = table[h];
            int is_valid_match = match > 0 && match < ip && match + LZ4_DISTANCE_ABSOLUTE_MAX >= ip &&
                                 rd32(src + match) == rd32(src + ip);  /* :345-348 */
            table[h] = (uint32_t)ip;                                   /* :350 */
            if (is_valid_match) break;
        }

        size_t literalLength = ip - anchor;                            /* :360 */
        size_t tokenPos = op;
        op += 1;
        if (op >= cap) return B2O_OutputTooSmall;                      /* :365 */
        if (literalLength >= RUN_MASK) {
            dst[tokenPos] = (uint8_t)(RUN_MASK << ML_BITS);
            size_t len = literalLength - RUN_MASK;
            while (len >= 255) {
                if (op >= cap) return B2O_OutputTooSmall;
                dst[op++] = 255;
                len -= 255;
            }
            if (op >= cap) return B2O_OutputTooSmall;
            dst[op++] = (uint8_t)len;
        } else {
            dst[tokenPos] = (uint8_t)(literalLength << ML_BITS);
        }
        if (op + literalLength > cap) return B2O_OutputTooSmall;       /* :388 */
        if (literalLength > 0) {
            memcpy(dst + op, src + anchor, literalLength);
            op += literalLength;
        }
        uint16_t offset = (uint16_t)(ip - match);                      /* :395 */
        if (op + 2 > cap) return B2O_OutputTooSmall;
        dst[op] = (uint8_t)(offset & 0xFF);
        dst[op + 1] = (uint8_t)(offset >> 8);
        op += 2;

        ip += MINMATCH;                                                /* :401 */
        match += MINMATCH;
        size_t matchLength = 0;
        while (ip < matchLimit) {                                      /* :405 */
            if (src[ip] == src[match]) { ip++; match++; matchLength++; }
            else break;
        }
        if (matchLength >= ML_MASK) {                                  /* :416 */
            dst[tokenPos] |= ML_MASK;
            size_t len = matchLength - ML_MASK;
            while (len >= 255) {
                if (op >= cap) return B2O_OutputTooSmall;
                dst[op++] = 255;
                len -= 255;
            }
            if (op >= cap) return B2O_OutputTooSmall;
            dst[op++] = (uint8_t)len;
        } else {
            dst[tokenPos] |= (uint8_t)matchLength;
        }
        anchor = ip;                                                   /* :435 */
        if (ip < mflimitPlusOne) {                                     /* :438 */
            uint32_t h = hash4(rd32(src + ip));
            table[h] = (uint32_t)ip;
            ip += 1;
        }
    }
    return finish_compression(src, srcSize, dst, cap, anchor, op, out); /* :446 */
}

/* src/lz4.zig:89-251 decompressGeneric with targetOutputSize == dst.len.
 * has_dict == 0 reproduces decompressSafe (:257, lowPrefix = dst, no dict);
 * has_dict == 1 reproduces decompressSafeUsingDict (:960, lowPrefix = dst, dict given). */
static int decompress_generic(const uint8_t* src, size_t srcLen, uint8_t* dst, size_t dstLen,
                              int has_dict, const uint8_t* dict, size_t dictSize, size_t* out) {
    *out = 0;
    if (srcLen == 0) return B2O_OK;                                    /* :97 */
    if (dstLen == 0) return B2O_OK;                                    /* :98 */
    size_t ip = 0, op = 0;
    const size_t iend = srcLen, oend = dstLen;
    for (;;) {
        if (ip >= iend) break;                                         /* :113 */
        uint8_t token = src[ip++];
        size_t literalLength = token >> ML_BITS;
        if (literalLength == RUN_MASK) {                               /* :123 */
            for (;;) {
                if (ip >= iend) return B2O_CorruptedData;
                uint8_t s = src[ip++];
                literalLength += s;
                if (s != 255) break;
            }
        }
        if (literalLength > 0) {                                       /* :134 */
            if (ip + literalLength > iend) return B2O_CorruptedData;
            if (op + literalLength > oend) return B2O_OutputTooSmall;
            memcpy(dst + op, src + ip, literalLength);
            ip += literalLength;
            op += literalLength;
        }
        if (ip >= iend) break;                                         /* :146 */
        if (ip + 2 > iend) return B2O_CorruptedData;                   /* :149 */
        size_t offset = (size_t)src[ip] | ((size_t)src[ip + 1] << 8);
        ip += 2;
        if (offset == 0) return B2O_CorruptedData;                     /* :154 */
        size_t matchLength = token & ML_MASK;
        if (matchLength == ML_MASK) {                                  /* :160 */
            for (;;) {
                if (ip >= iend) return B2O_CorruptedData;
                uint8_t s = src[ip++];
                matchLength += s;
                if (s != 255) break;
            }
        }
        matchLength += MINMATCH;                                       /* :171 */
        if (op + matchLength > oend) return B2O_OutputTooSmall;        /* :174 */

        if (offset > op) {                                             /* :181  matchPtr < lowPrefix */
            if (!has_dict) return B2O_CorruptedData;                   /* :183-186 */
            size_t prefixOffset = op;                                  /* :189 (lowPrefix == dst) */
            if (offset > prefixOffset + dictSize) return B2O_CorruptedData; /* :190 */
            size_t lowPrefixOffset = offset - op;                      /* :195 */
            const uint8_t* dictMatchPtr = dict + dictSize - lowPrefixOffset; /* :196 */
            if (matchLength <= lowPrefixOffset) {                      /* :199 */
                memcpy(dst + op, dictMatchPtr, matchLength);
                op += matchLength;
            } else {
                size_t copySize = lowPrefixOffset;
                size_t restSize = matchLength - copySize;
                memcpy(dst + op, dictMatchPtr, copySize);
                op += copySize;
                size_t restStart = 0;                                  /* :213 */
                /* :216-227: byte loop when overlapping, memcpy otherwise — same bytes either way */
                for (size_t i = 0; i < restSize; i++) dst[op + i] = dst[restStart + i];
                op += restSize;
            }
        } else {
            size_t matchPos = op - offset;                             /* :232 */
            /* :235-246: forward byte copy (overlap) or memcpy — a forward byte loop gives both */
            for (size_t i = 0; i < matchLength; i++) dst[op + i] = dst[matchPos + i];
            op += matchLength;
        }
    }
    *out = op;
    return B2O_OK;
}

int b2o_decompress_safe(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out) {
    return decompress_generic(src, n, dst, cap, 0, NULL, 0, out);
}

int b2o_decompress_safe_using_dict(const uint8_t* src, size_t n, uint8_t* dst, size_t cap,
                                   const uint8_t* dict, size_t dict_len, size_t* out) {
    return decompress_generic(src, n, dst, cap, 1, dict, dict_len, out);
}

/* src/lz4.zig:551-616 — literal restatement, including what it leaves in dst: every probe compresses into the
   same dst, so after the call dst holds the bytes of the LAST probe (complete or cut off where it stopped
   fitting), not necessarily those of the prefix whose sizes are returned. */
int b2o_compress_dest_size(const uint8_t* src, uint8_t* dst, size_t cap, size_t* srcSizePtr, size_t* out) {
    const size_t maxSrcSize = *srcSizePtr;
    *out = 0;
    if (maxSrcSize == 0) {                                              /* :553-556 */
        *srcSizePtr = 0;
        return B2O_OK;
    }
    const size_t maxCompressed = b2o_compress_bound(maxSrcSize);        /* :559 */
    if (cap >= maxCompressed) {                                         /* :560-564 */
        size_t result = 0;
        int rc = b2o_compress_fast(src, maxSrcSize, dst, cap, 1, &result);
        if (rc) return rc;
        *srcSizePtr = maxSrcSize;
        *out = result;
        return B2O_OK;
    }
    size_t low = 1, high = maxSrcSize, bestSize = 0, bestCompressedSize = 0;  /* :567-570 */
    if (cap <= maxSrcSize) {                                            /* :573-586 */
        const size_t estimate = cap < maxSrcSize ? cap : maxSrcSize;
        size_t size = 0;
        if (b2o_compress_fast(src, estimate, dst, cap, 1, &size) == B2O_OK) {
            if (size <= cap) {
                bestSize = estimate;
                bestCompressedSize = size;
                low = estimate + 1;
            } else {
                high = estimate - 1;
            }
        } else {
            high = estimate - 1;
        }
    }
    while (low <= high) {                                               /* :589-612 */
        const size_t mid = low + (high - low) / 2;
        if (mid == 0 || mid > maxSrcSize) break;
        size_t size = 0;
        if (b2o_compress_fast(src, mid, dst, cap, 1, &size) == B2O_OK) {
            if (size <= cap) {
                bestSize = mid;
                bestCompressedSize = size;
                if (mid == maxSrcSize) break;
                low = mid + 1;
            } else {
                high = mid - 1;
            }
        } else {
            high = mid - 1;
        }
        if (low > maxSrcSize) break;
    }
    *srcSizePtr = bestSize;                                             /* :614-615 */
    *out = bestCompressedSize;
    return B2O_OK;
}
