/*
 * oracle_lz4hc.c — CPU ORACLE (test infrastructure only; see b2o.h).
 * Literal C restatement of the reference HC hash-chain path, /root/reference/src/lz4hc.zig:
 *   compressHC :1440-1453 -> compressHCExtState :1457-1489 -> compressHashChain :976-1064
 *   -> insertAndFindBestMatch :514-535 -> insertHC :491-510 + insertAndGetWiderMatch :538-681.
 * In the one-shot path prefixStart = src, dictLimit = lowLimit = 0 (:1001-1006), so every "index"
 * below is simply a byte position in src.
 *
 * Quirks kept on purpose: greedy (no lazy evaluation), chainSwap is a stub (:548),
 * countBack never runs because iLowLimit == ip (:528,:596).
 *
 * ONE DOCUMENTED DEVIATION — the F8 guard: at :636 the reference computes `matchIndex - 1` on a
 * u32; when the chain walk ended with matchIndex == 0 and chainTable[0] == 1 (only reachable for
 * blocks > 64 KiB, where index 65536 aliases slot 0) that underflows: a panic in Debug/ReleaseSafe,
 * a wild read in ReleaseFast.  Here that case is treated as "no pattern candidate" and counted in
 * b2o_hc_f8_guard_hits().  Second deviation, unreachable with a compressBound-sized dst: the final
 * literal run's head-room test at :1037 ignores the length-extension bytes, so the reference can
 * write past dst; this restatement returns OutputTooSmall instead of overflowing.
 *
 * Levels 2 (compressMID :687-971) and 10-12 (compressOptimal :1068-1391) are outside the hot-path
 * scope and are not restated: B2O_UnsupportedLevel.
 */
#include "b2o.h"
#include <stdlib.h>
#include <string.h>

#define MINMATCH 4
#define LASTLITERALS 5
#define MFLIMIT 12
#define ML_BITS 4
#define ML_MASK 15u
#define RUN_MASK 15u
#define LZ4_MAX_INPUT_SIZE 0x7E000000u
#define LZ4_DISTANCE_MAX 65535u
#define LZ4HC_MAXD 65536u                 /* src/lz4hc.zig:34 */
#define LZ4HC_MAXD_MASK (LZ4HC_MAXD - 1)
#define LZ4HC_HASH_LOG 15                 /* src/lz4hc.zig:37 */
#define LZ4HC_HASHTABLESIZE (1u << LZ4HC_HASH_LOG)
#define HASH_MULTIPLIER 2654435761u

typedef struct {                          /* src/lz4hc.zig:391-403 (tables + nextToUpdate only) */
    uint32_t hashTable[LZ4HC_HASHTABLESIZE];
    uint16_t chainTable[LZ4HC_MAXD];
    uint32_t nextToUpdate;
} hc_ctx;

typedef struct { int32_t off, len; } hc_match;

static __thread uint64_t g_f8_hits = 0;
uint64_t b2o_hc_f8_guard_hits(void) { return g_f8_hits; }

static inline uint32_t rd32(const uint8_t* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static inline uint64_t rd64(const uint8_t* p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }
static inline uint32_t hashHC(uint32_t seq) {                      /* :129-131 */
    return (uint32_t)(seq * HASH_MULTIPLIER) >> ((MINMATCH * 8) - LZ4HC_HASH_LOG);
}

/* src/lz4hc.zig:170-199 */
static size_t countPattern(const uint8_t* ip, const uint8_t* iEnd, uint32_t pattern32) {
    const uint8_t* iStart = ip;
    const uint8_t* ptr = ip;
    uint64_t pattern64 = (uint64_t)pattern32 | ((uint64_t)pattern32 << 32);
    while (ptr + 7 < iEnd) {
        uint64_t diff = rd64(ptr) ^ pattern64;
        if (diff == 0) ptr += 8;
        else return (size_t)(ptr - iStart) + ((size_t)__builtin_ctzll(diff) >> 3);
    }
    uint32_t patternByte = pattern32;
    while (ptr < iEnd) {
        if (ptr[0] != (uint8_t)patternByte) break;
        ptr += 1;
        patternByte >>= 8;
        if (patternByte == 0) patternByte = pattern32;
    }
    return (size_t)(ptr - iStart);
}

/* src/lz4hc.zig:202-222 */
static size_t reverseCountPattern(const uint8_t* ip, const uint8_t* iLow, uint32_t pattern) {
    const uint8_t* iStart = ip;
    const uint8_t* ptr = ip;
    while (ptr >= iLow + 4) {
        if (rd32(ptr - 4) != pattern) break;
        ptr -= 4;
    }
    uint8_t patternBytes[4] = {(uint8_t)pattern, (uint8_t)(pattern >> 8), (uint8_t)(pattern >> 16),
                               (uint8_t)(pattern >> 24)};
    size_t byteIdx = 3;
    while (ptr > iLow) {
        if (ptr[-1] != patternBytes[byteIdx]) break;
        ptr -= 1;
        if (byteIdx == 0) byteIdx = 3; else byteIdx -= 1;
    }
    return (size_t)(iStart - ptr);
}

/* src/lz4hc.zig:225-228 */
static inline int isRepetitivePattern(uint32_t pattern) {
    return ((pattern & 0xFFFF) == (pattern >> 16)) && ((pattern & 0xFF) == (pattern >> 24));
}

/* src/lz4hc.zig:234-264 */
static size_t lz4Count(const uint8_t* ip, const uint8_t* match, const uint8_t* iLimit) {
    size_t counted = 0;
    while (ip + 8 <= iLimit) {
        uint64_t diff = rd64(ip) ^ rd64(match);
        if (diff == 0) { ip += 8; match += 8; counted += 8; }
        else return counted + ((size_t)__builtin_ctzll(diff) >> 3);
    }
    while (ip < iLimit) {
        if (ip[0] != match[0]) break;
        ip++; match++; counted++;
    }
    return counted;
}

/* src/lz4hc.zig:491-510 */
static void insertHC(hc_ctx* ctx, const uint8_t* src, uint32_t target) {
    uint32_t idx = ctx->nextToUpdate;
    while (idx < target) {
        uint32_t h = hashHC(rd32(src + idx));
        uint32_t prevIdx = ctx->hashTable[h];
        uint32_t delta = (prevIdx > idx) ? LZ4_DISTANCE_MAX + 1 : idx - prevIdx;
        uint16_t deltaClamped = (delta > LZ4_DISTANCE_MAX) ? (uint16_t)LZ4_DISTANCE_MAX : (uint16_t)delta;
        ctx->chainTable[idx & LZ4HC_MAXD_MASK] = deltaClamped;
        ctx->hashTable[h] = idx;
        idx += 1;
    }
    ctx->nextToUpdate = target;
}

/* src/lz4hc.zig:514-535 + :538-681 with iLowLimit == ip, longest == MINMATCH-1, chainSwap == false */
static hc_match insertAndFindBestMatch(hc_ctx* ctx, const uint8_t* src, uint32_t ipIndex,
                                       const uint8_t* iHighLimit, int32_t maxNbAttempts,
                                       int patternAnalysis) {
    insertHC(ctx, src, ipIndex);                                     /* :522 */
    const uint8_t* ip = src + ipIndex;
    const uint32_t lowLimit = 0, dictIdx = 0;
    const int withinStartDistance = (lowLimit + (LZ4_DISTANCE_MAX + 1) > ipIndex);          /* :553 */
    const uint32_t lowestMatchIndex = withinStartDistance ? lowLimit : ipIndex - LZ4_DISTANCE_MAX;
    int32_t nbAttempts = maxNbAttempts;
    const uint32_t pattern = rd32(ip);
    hc_match result = {0, MINMATCH - 1};

    uint32_t matchIndex = ctx->hashTable[hashHC(pattern)];           /* :563 */
    if (matchIndex == 0) return result;                              /* :566 */

    while (matchIndex > 0 && nbAttempts > 0) {                       /* :571 */
        if (matchIndex > ipIndex || (ipIndex - matchIndex) > LZ4_DISTANCE_MAX) break;       /* :573 */
        nbAttempts -= 1;
        if (matchIndex >= lowestMatchIndex) {
            const uint8_t* matchPtr = src + matchIndex;
            if (rd32(matchPtr) == pattern) {
                int32_t mlt = (int32_t)(MINMATCH + lz4Count(ip + MINMATCH, matchPtr + MINMATCH, iHighLimit));
                int32_t totalLength = mlt;                           /* back == 0 (:596 false) */
                if (totalLength > result.len) {
                    result.len = totalLength;
                    result.off = (int32_t)(ipIndex - matchIndex);
                    if (totalLength > maxNbAttempts) break;          /* :613 */
                }
            }
        }
        uint16_t delta = ctx->chainTable[matchIndex & LZ4HC_MAXD_MASK];                     /* :619 */
        if (delta == 0 || delta > matchIndex) break;
        matchIndex -= delta;
    }

    if (patternAnalysis && result.len > 0) {                         /* :626 */
        uint16_t delta = ctx->chainTable[matchIndex & LZ4HC_MAXD_MASK];
        if (delta == 1) {
            if (isRepetitivePattern(pattern)) {
                size_t srcPatternLength = countPattern(ip + 4, iHighLimit, pattern) + 4;
                if (matchIndex == 0) {
                    g_f8_hits++;                                     /* F8 guard, see header */
                    return result;
                }
                uint32_t matchCandidateIdx = matchIndex - 1;         /* :636 */
                if (matchCandidateIdx >= lowestMatchIndex && matchCandidateIdx >= dictIdx) {
                    const uint8_t* matchPtr = src + matchCandidateIdx;
                    if (rd32(matchPtr) == pattern) {
                        size_t forwardPatternLength = countPattern(matchPtr + 4, iHighLimit, pattern) + 4;
                        size_t backLength = reverseCountPattern(matchPtr, src, pattern);
                        uint32_t a = matchCandidateIdx - (uint32_t)backLength;
                        uint32_t lim = a > lowestMatchIndex ? a : lowestMatchIndex;
                        uint32_t limitedBackLength = matchCandidateIdx - lim;                /* :653 */
                        size_t currentSegmentLength = (size_t)limitedBackLength + forwardPatternLength;
                        uint32_t newMatchIndex;
                        size_t mn = currentSegmentLength < srcPatternLength ? currentSegmentLength : srcPatternLength;
                        int32_t maxML = (int32_t)mn;
                        if (currentSegmentLength >= srcPatternLength && forwardPatternLength <= srcPatternLength)
                            newMatchIndex = matchCandidateIdx + (uint32_t)forwardPatternLength - (uint32_t)srcPatternLength;
                        else
                            newMatchIndex = matchCandidateIdx - limitedBackLength;
                        if (maxML > result.len && (ipIndex - newMatchIndex) <= LZ4_DISTANCE_MAX) {
                            result.len = maxML;
                            result.off = (int32_t)(ipIndex - newMatchIndex);
                        }
                    }
                }
            }
        }
    }
    return result;
}

/* src/lz4hc.zig:1394-1425 */
static int encodeLiterals(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out) {
    if (cap < n + 1 + (n / 255)) return B2O_OutputTooSmall;
    uint8_t* op = dst;
    size_t litLen = n;
    if (litLen >= RUN_MASK) {
        size_t len = litLen - RUN_MASK;
        *op++ = (uint8_t)(RUN_MASK << ML_BITS);
        while (len >= 255) { *op++ = 255; len -= 255; }
        *op++ = (uint8_t)len;
    } else {
        *op++ = (uint8_t)(litLen << ML_BITS);
    }
    memcpy(op, src, litLen);
    op += litLen;
    *out = (size_t)(op - dst);
    return B2O_OK;
}

/* src/lz4hc.zig:976-1064 */
static int compressHashChain(hc_ctx* ctx, const uint8_t* src, size_t inputSize, uint8_t* dst, size_t cap,
                             int32_t maxNbAttempts, size_t* out) {
    const int patternAnalysis = (maxNbAttempts > 128);               /* :983 */
    if (inputSize < MFLIMIT + 1) return encodeLiterals(src, inputSize, dst, cap, out);      /* :995 */
    const uint8_t* ip = src;
    const uint8_t* anchor = ip;
    const uint8_t* iend = ip + inputSize;
    const uint8_t* mflimit = iend - MFLIMIT;
    const uint8_t* matchlimit = iend - LASTLITERALS;
    uint8_t* op = dst;
    uint8_t* oend = dst + cap;
    ctx->nextToUpdate = 0;                                           /* :1001 */

    while (ip <= mflimit) {                                          /* :1009 */
        hc_match match = insertAndFindBestMatch(ctx, src, (uint32_t)(ip - src), matchlimit, maxNbAttempts,
                                                patternAnalysis);
        if (match.len < MINMATCH || match.off == 0) { ip += 1; continue; }                  /* :1013 */

        /* encodeSequence, limitedOutput — src/lz4hc.zig:308-386 */
        size_t litLen = (size_t)(ip - anchor);
        size_t needed = (litLen / 255) + litLen + (2 + 1 + LASTLITERALS);
        if (op + needed > oend) return B2O_OutputTooSmall;           /* :320-325 */
        uint8_t* token = op++;
        if (litLen >= RUN_MASK) {
            size_t len = litLen - RUN_MASK;
            *token = (uint8_t)(RUN_MASK << ML_BITS);
            while (len >= 255) { *op++ = 255; len -= 255; }
            *op++ = (uint8_t)len;
        } else {
            *token = (uint8_t)(litLen << ML_BITS);
        }
        memcpy(op, anchor, litLen);
        op += litLen;
        op[0] = (uint8_t)(match.off & 0xFF);
        op[1] = (uint8_t)((match.off >> 8) & 0xFF);
        op += 2;
        size_t mlCode = (size_t)(match.len - MINMATCH);
        if (op + (mlCode / 255) + (1 + LASTLITERALS) > oend) return B2O_OutputTooSmall;     /* :355-359 */
        if (mlCode >= ML_MASK) {
            *token += ML_MASK;
            size_t remaining = mlCode - ML_MASK;
            while (remaining >= 510) { op[0] = 255; op[1] = 255; op += 2; remaining -= 510; }
            if (remaining >= 255) { *op++ = 255; remaining -= 255; }
            *op++ = (uint8_t)remaining;
        } else {
            *token += (uint8_t)mlCode;
        }
        ip += match.len;
        anchor = ip;
    }

    size_t finalLiterals = (size_t)(iend - anchor);                  /* :1035 */
    if (finalLiterals > 0) {
        if (op + finalLiterals + 1 > oend) return B2O_OutputTooSmall;                        /* :1037 */
        size_t ext = finalLiterals >= RUN_MASK ? (finalLiterals - RUN_MASK) / 255 + 1 : 0;
        if (op + 1 + ext + finalLiterals > oend) return B2O_OutputTooSmall;  /* deviation: no overflow */
        if (finalLiterals >= RUN_MASK) {
            size_t len = finalLiterals - RUN_MASK;
            *op++ = (uint8_t)(RUN_MASK << ML_BITS);
            while (len >= 255) { *op++ = 255; len -= 255; }
            *op++ = (uint8_t)len;
        } else {
            *op++ = (uint8_t)(finalLiterals << ML_BITS);
        }
        memcpy(op, anchor, finalLiterals);
        op += finalLiterals;
    }
    *out = (size_t)(op - dst);
    return B2O_OK;
}

/* src/lz4hc.zig:72-97 */
static int nb_searches_for_level(int level) {
    switch (level) {
        case 3: return 4; case 4: return 8; case 5: return 16; case 6: return 32;
        case 7: return 64; case 8: return 128; case 9: return 256;
        default: return -1;  /* 2 = lz4mid, 10..12 = lz4opt: not restated */
    }
}

/* src/lz4hc.zig:1440-1453 + :1457-1489 */
int b2o_compress_hc(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, int compressionLevel, size_t* out) {
    *out = 0;
    if (n > LZ4_MAX_INPUT_SIZE) return B2O_InputTooLarge;            /* :1442 */
    if (n == 0) return B2O_OK;                                       /* :1443 */
    int level = compressionLevel < 2 ? 9 : (compressionLevel > 12 ? 12 : compressionLevel);  /* :1445 */
    if (cap == 0) return B2O_OutputTooSmall;                         /* :1461 */
    int nb = nb_searches_for_level(level);
    if (nb < 0) return B2O_UnsupportedLevel;
    hc_ctx* ctx = (hc_ctx*)calloc(1, sizeof(hc_ctx));                /* :1450 Context.init(): zero tables */
    if (!ctx) return B2O_AllocationFailed;
    int rc = compressHashChain(ctx, src, n, dst, cap, nb, out);
    free(ctx);
    return rc;
}
