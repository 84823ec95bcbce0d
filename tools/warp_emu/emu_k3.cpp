// emu_k3.cpp — runs K3's device code (k_compress_hc.cu, compiled for the host with -DB2_EMU) on the one-warp emulator and
// compares every block with the CPU oracle's compressHC: both kernel variants (hop-by-hop walk, jump tables), levels 3..9,
// several blocks in a row on ONE work area (the epoch base and the lazily built jump levels persist across blocks).
// Usage: emu_k3 [class 0..4] [blocks] [block_bytes] [seed]
#include "emu_cuda.h"
#include "../../zig-lz4_b200/csrc/k_compress_hc.cu"
#include "../../oracle/b2o.h"

extern "C" int b2gen_fill(uint8_t* dst, uint64_t n, uint64_t seed, uint32_t mode, uint64_t span, int nthreads);

static uint64_t g_checked = 0, g_failed = 0;
static int nb_searches(int level) {      // clevelTable of the hash-chain levels, src/lz4hc.zig:72-97
    static const int t[10] = {0, 0, 0, 4, 8, 16, 32, 64, 128, 256};
    return t[level];
}

static void check(b2::HcWork* work, bool jump, const uint8_t* src0, uint32_t n, uint32_t cap, int level, const char* what) {
    std::vector<uint8_t> want((size_t)cap + 64);
    size_t wlen = 0;
    const int wst = b2o_compress_hc(src0, n, want.data(), cap, level, &wlen);
    emu::Guarded gs(n ? n : 1, (g_checked & 1) == 0, (size_t)(g_checked * 5 % 16));
    memcpy(gs.p, src0, n);
    emu::Guarded gd(cap ? cap : 1, true);
    uint32_t r_olen = 0; int r_st = 0;
    const uint8_t* src = gs.p; uint8_t* dst = gd.p;
    emu::run_warp([&] {
        uint32_t ol = 0; int st = 0;
        if (jump) b2::compress_block_hc<true>(src, n, dst, cap, work, nb_searches(level), 24u, b2::lane_id(), ol, st);
        else b2::compress_block_hc<false>(src, n, dst, cap, work, nb_searches(level), 24u, b2::lane_id(), ol, st);
        if (b2::lane_id() == 0) { r_olen = ol; r_st = st; }
    });
    g_checked++;
    bool ok = r_st == wst;
    if (ok && wst == 0) ok = r_olen == wlen && memcmp(dst, want.data(), wlen) == 0;
    if (!ok) {
        g_failed++;
        fprintf(stderr, "MISMATCH %s n=%u cap=%u level=%d jump=%d: status %d/%d len %u/%zu\n", what, n, cap, level, (int)jump, r_st, wst,
                r_olen, wlen);
    }
}

static std::vector<uint8_t> read_file(const char* path) {
    std::vector<uint8_t> d;
    FILE* f = fopen(path, "rb");
    if (!f) { perror(path); exit(2); }
    uint8_t buf[65536];
    size_t k;
    while ((k = fread(buf, 1, sizeof buf, f)) > 0) d.insert(d.end(), buf, buf + k);
    fclose(f);
    return d;
}

int main(int argc, char** argv) {
    if (argc > 1 && argv[1][0] == '@') {
        // file mode: every @path as one block at levels 3, 6 and 9, both chain walks, on one work area
        b2::HcWork* work = (b2::HcWork*)calloc(1, sizeof(b2::HcWork));
        for (int a = 1; a < argc; a++) {
            const std::vector<uint8_t> d = read_file(argv[a] + 1);
            const uint32_t n = (uint32_t)d.size();
            for (int level : {3, 6, 9})
                for (bool jump : {false, true}) check(work, jump, d.data(), n, (uint32_t)b2o_compress_bound(n), level, argv[a]);
        }
        free(work);
        printf("files: %llu blocks checked, %llu failed\n", (unsigned long long)g_checked, (unsigned long long)g_failed);
        return g_failed ? 1 : 0;
    }
    const int cls = argc > 1 ? atoi(argv[1]) : 0;
    const uint32_t blocks = argc > 2 ? (uint32_t)atoi(argv[2]) : 2;
    const uint32_t bs = argc > 3 ? (uint32_t)atoi(argv[3]) : 16384;
    const uint64_t seed = argc > 4 ? strtoull(argv[4], nullptr, 0) : 0x4C5A3442ull;
    const uint64_t stride = bs < 65536 ? 65536 : bs;
    std::vector<uint8_t> data(stride * blocks + 64);
    b2gen_fill(data.data(), stride * blocks, seed, (uint32_t)cls, 65536, 4);
    b2::HcWork* work = (b2::HcWork*)calloc(1, sizeof(b2::HcWork));      // one area for everything, as a resident warp has
    const uint32_t bound = (uint32_t)b2o_compress_bound(bs);
    for (uint32_t b = 0; b < blocks; b++) {
        const uint8_t* src = data.data() + b * stride;
        for (int level : {9, 3, 6, 9})
            for (bool jump : {false, true}) check(work, jump, src, bs, bound, level, "full");
        if (b == 0) {
            for (uint32_t n : {0u, 1u, 12u, 13u, 14u, 40u, 300u, 4097u})
                check(work, true, src, n, (uint32_t)b2o_compress_bound(n), 9, "small");
            for (uint32_t cap : {0u, 1u, 17u, 200u, 1000u}) check(work, true, src, 2000, cap, 9, "cap");      // limitedOutput exits
        }
    }
    free(work);
    printf("K3 class %d: %llu blocks checked, %llu failed\n", cls, (unsigned long long)g_checked, (unsigned long long)g_failed);
    return g_failed ? 1 : 0;
}
