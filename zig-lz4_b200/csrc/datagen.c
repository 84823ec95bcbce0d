/*
 * datagen.c — synthetic corpus generator for tests and bench (host C, pthreads).  Not part of the
 * codec: it only manufactures the inputs SURVEY.md §8(d) describes, deterministically and with
 * integer arithmetic only, so every run (here, on the GPU box, on any thread count) sees the same bytes.
 *
 * The buffer is produced in independent 64 KiB units; unit u draws from xoshiro256** seeded with
 * splitmix64(seed + u).  Its class is fixed (mode 0..3) or rotates every `span` bytes (mode 4):
 *   0 text-like   : words from a 4096-word synthetic vocabulary, log-uniform rank (Zipf-like), spaces /
 *                   newlines / full stops
 *   1 binary      : 32-byte records {u32 counter, u32 id<1024, 2 x f32-looking noise, 16 B mostly zero}
 *   2 redundant   : runs (1..512) over an 8-symbol alphabet mixed with repeats of a 64-byte pattern
 *   3 random      : incompressible PRNG bytes
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define UNIT 65536u

static inline uint64_t splitmix64(uint64_t* x) {
    uint64_t z = (*x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
typedef struct { uint64_t s[4]; } xo_t;
static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t xo_next(xo_t* g) {
    uint64_t* s = g->s;
    uint64_t r = rotl64(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl64(s[3], 45);
    return r;
}
static void xo_seed(xo_t* g, uint64_t seed) {
    uint64_t x = seed;
    for (int i = 0; i < 4; i++) g->s[i] = splitmix64(&x);
}

/* vocabulary: 4096 words of 2..10 lowercase letters, fixed for all seeds */
static uint8_t g_words[4096][12];
static uint8_t g_wlen[4096];
static pthread_once_t g_once = PTHREAD_ONCE_INIT;
static void init_vocab(void) {
    static const char freq[] = "eeeeetttaaaooiinnsshhrrdlcumwfgypbvk";  /* skewed letter mix */
    for (uint32_t w = 0; w < 4096; w++) {
        uint64_t x = 0xB200B200ull + w;
        uint64_t r = splitmix64(&x);
        uint32_t len = 2 + (uint32_t)(r % 9);
        g_wlen[w] = (uint8_t)len;
        for (uint32_t i = 0; i < len; i++) {
            r = splitmix64(&x);
            g_words[w][i] = (uint8_t)freq[r % (sizeof(freq) - 1)];
        }
    }
}

static void gen_text(uint8_t* d, uint32_t n, xo_t* g) {
    uint32_t p = 0;
    while (p < n) {
        uint64_t r = xo_next(g);
        uint32_t k = (uint32_t)(r % 13);
        uint32_t rank = ((1u << k) - 1) + ((uint32_t)(r >> 8) & ((1u << k) - 1));
        if (rank > 4095) rank = 4095;
        uint32_t len = g_wlen[rank];
        for (uint32_t i = 0; i < len && p < n; i++) d[p++] = g_words[rank][i];
        uint32_t sep = (uint32_t)(r >> 40) & 31;
        if (sep == 0) { if (p < n) d[p++] = '.'; if (p < n) d[p++] = ' '; }
        else if (sep < 3) { if (p < n) d[p++] = '\n'; }
        else if (p < n) d[p++] = ' ';
    }
}

static void gen_binary(uint8_t* d, uint32_t n, xo_t* g, uint64_t unit) {
    uint32_t counter = (uint32_t)(unit * (UNIT / 32));
    for (uint32_t p = 0; p + 32 <= n; p += 32, counter++) {
        uint64_t r = xo_next(g), r2 = xo_next(g);
        uint32_t id = (uint32_t)(r % 1024);
        uint32_t f0 = 0x3F800000u | ((uint32_t)(r >> 16) & 0x7FFFFFu);
        uint32_t f1 = 0x40000000u | ((uint32_t)(r2 >> 8) & 0x7FFFFFu);
        memcpy(d + p, &counter, 4); memcpy(d + p + 4, &id, 4); memcpy(d + p + 8, &f0, 4); memcpy(d + p + 12, &f1, 4);
        memset(d + p + 16, 0, 16);
        if (((r2 >> 40) & 3) == 0) d[p + 16 + ((r2 >> 44) & 15)] = (uint8_t)(r2 >> 52);
    }
    for (uint32_t p = n & ~31u; p < n; p++) d[p] = 0;
}

static void gen_redundant(uint8_t* d, uint32_t n, xo_t* g) {
    uint8_t pat[64], alpha[8];
    for (int i = 0; i < 64; i += 8) { uint64_t r = xo_next(g); memcpy(pat + i, &r, 8); }
    { uint64_t r = xo_next(g); memcpy(alpha, &r, 8); }
    uint32_t p = 0;
    while (p < n) {
        uint64_t r = xo_next(g);
        if ((r & 3) == 0) {
            uint32_t reps = 1 + (uint32_t)((r >> 8) % 16);
            for (uint32_t k = 0; k < reps * 64 && p < n; k++) d[p++] = pat[k & 63];
        } else {
            uint32_t len = 1 + (uint32_t)((r >> 8) % 512);
            uint8_t sym = alpha[(r >> 32) & 7];
            for (uint32_t k = 0; k < len && p < n; k++) d[p++] = sym;
        }
    }
}

static void gen_random(uint8_t* d, uint32_t n, xo_t* g) {
    uint32_t p = 0;
    for (; p + 8 <= n; p += 8) { uint64_t r = xo_next(g); memcpy(d + p, &r, 8); }
    if (p < n) { uint64_t r = xo_next(g); memcpy(d + p, &r, n - p); }
}

typedef struct {
    uint8_t* dst; uint64_t n, seed, span; uint32_t mode; uint64_t next, units;
} job_t;

static void* worker(void* arg) {
    job_t* j = (job_t*)arg;
    for (;;) {
        uint64_t u = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (u >= j->units) break;
        uint64_t off = u * UNIT;
        uint32_t len = (uint32_t)(j->n - off < UNIT ? j->n - off : UNIT);
        uint32_t cls = j->mode < 4 ? j->mode : (uint32_t)((off / j->span) % 4);
        xo_t g;
        xo_seed(&g, j->seed + u);
        uint8_t* d = j->dst + off;
        switch (cls) {
            case 0: gen_text(d, len, &g); break;
            case 1: gen_binary(d, len, &g, u); break;
            case 2: gen_redundant(d, len, &g); break;
            default: gen_random(d, len, &g); break;
        }
    }
    return NULL;
}

/* mode 0..3: single class; mode 4: classes rotate every `span` bytes (span a multiple of 64 KiB). */
__attribute__((visibility("default")))
int b2gen_fill(uint8_t* dst, uint64_t n, uint64_t seed, uint32_t mode, uint64_t span, int nthreads) {
    pthread_once(&g_once, init_vocab);
    if (span < UNIT) span = UNIT;
    job_t j = {dst, n, seed, span, mode, 0, (n + UNIT - 1) / UNIT};
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t t[256];
    int started = 0;
    for (int i = 0; i < nthreads - 1; i++) if (pthread_create(&t[started], NULL, worker, &j) == 0) started++;
    worker(&j);
    for (int i = 0; i < started; i++) pthread_join(t[i], NULL);
    return 0;
}
