/*
 * oracle_xxh32.c — CPU ORACLE (test infrastructure only; see b2o.h).
 * XXH32 as used by the reference through Zig's std.hash.XxHash32 (Zig std 0.15.1 — NOT under
 * /root/reference; pinned by build.zig.zon:28 `minimum_zig_version`).  Call sites in the reference:
 * src/lz4f.zig:139 (header), :375/:385/:438 (content, streaming), :424 (block), :560/:595/:618/:630.
 * This is the published XXH32 algorithm (Yann Collet, xxHash spec §XXH32); it is pinned in tests
 * against libxxhash.so.0 and python-xxhash, and by stock liblz4 accepting our frame checksums.
 */
#include "b2o.h"
#include <string.h>

#define P1 2654435761u
#define P2 2246822519u
#define P3 3266489917u
#define P4 668265263u
#define P5 374761393u

static inline uint32_t rotl(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
static inline uint32_t rd32(const uint8_t* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static inline uint32_t xround(uint32_t acc, uint32_t x) { return rotl(acc + x * P2, 13) * P1; }

void b2o_xxh32_init(b2o_xxh32_state* s, uint32_t seed) {
    s->v[0] = seed + P1 + P2;
    s->v[1] = seed + P2;
    s->v[2] = seed;
    s->v[3] = seed - P1;
    s->buf_len = 0;
    s->total = 0;
    s->seed = seed;
}

void b2o_xxh32_update(b2o_xxh32_state* s, const void* data, size_t n) {
    const uint8_t* p = (const uint8_t*)data;
    s->total += n;
    if (s->buf_len) {
        size_t take = 16 - s->buf_len;
        if (take > n) take = n;
        memcpy(s->buf + s->buf_len, p, take);
        s->buf_len += (uint32_t)take;
        p += take;
        n -= take;
        if (s->buf_len < 16) return;
        for (int i = 0; i < 4; i++) s->v[i] = xround(s->v[i], rd32(s->buf + 4 * i));
        s->buf_len = 0;
    }
    while (n >= 16) {
        for (int i = 0; i < 4; i++) s->v[i] = xround(s->v[i], rd32(p + 4 * i));
        p += 16;
        n -= 16;
    }
    if (n) {
        memcpy(s->buf, p, n);
        s->buf_len = (uint32_t)n;
    }
}

uint32_t b2o_xxh32_final(const b2o_xxh32_state* s) {
    uint32_t h;
    if (s->total >= 16)
        h = rotl(s->v[0], 1) + rotl(s->v[1], 7) + rotl(s->v[2], 12) + rotl(s->v[3], 18);
    else
        h = s->seed + P5;
    h += (uint32_t)s->total;
    const uint8_t* p = s->buf;
    uint32_t n = s->buf_len;
    while (n >= 4) {
        h = rotl(h + rd32(p) * P3, 17) * P4;
        p += 4;
        n -= 4;
    }
    while (n) {
        h = rotl(h + (uint32_t)(*p) * P5, 11) * P1;
        p++;
        n--;
    }
    h ^= h >> 15;
    h *= P2;
    h ^= h >> 13;
    h *= P3;
    h ^= h >> 16;
    return h;
}

uint32_t b2o_xxh32(const void* p, size_t n, uint32_t seed) {
    b2o_xxh32_state s;
    b2o_xxh32_init(&s, seed);
    b2o_xxh32_update(&s, p, n);
    return b2o_xxh32_final(&s);
}
