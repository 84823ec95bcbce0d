// How long does the host-side header chain of a frame take?  16384 dependent 4-byte reads, ~36 KB apart, in pinned memory.
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <chrono>
#include <cuda_runtime.h>
int main() {
    const size_t n = 600u << 20;
    uint8_t* h = nullptr;
    cudaHostAlloc(&h, n, cudaHostAllocDefault);
    memset(h, 1, n);
    // records: header (size) + payload
    size_t p = 7; uint32_t nb = 0; uint64_t rng = 88172645463325252ull;
    while (p + 4 + 70000 < n && nb < 16384) {
        rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
        uint32_t sz = 30000 + (uint32_t)(rng % 12000);
        memcpy(h + p, &sz, 4); p += 4 + sz; nb++;
    }
    uint32_t z = 0; memcpy(h + p, &z, 4);
    // evict caches
    uint8_t* junk = (uint8_t*)malloc(256u << 20); memset(junk, 3, 256u << 20);
    for (int rep = 0; rep < 3; rep++) {
        volatile uint64_t sink = 0; for (size_t i = 0; i < (256u << 20); i += 64) sink += junk[i];
        auto t0 = std::chrono::steady_clock::now();
        size_t q = 7; uint32_t cnt = 0;
        for (;;) { uint32_t hd; memcpy(&hd, h + q, 4); q += 4; if (hd == 0) break; q += hd & 0x7FFFFFFFu; cnt++; }
        auto t1 = std::chrono::steady_clock::now();
        printf("walk of %u headers: %.3f ms\n", cnt, std::chrono::duration<double, std::milli>(t1 - t0).count());
    }
    return 0;
}
