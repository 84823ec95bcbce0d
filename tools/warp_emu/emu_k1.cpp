// emu_k1.cpp — runs K1's device code (k_compress_fast.cu, compiled for the host with -DB2_EMU) on the one-warp
// emulator and compares every block with the CPU oracle.  Test tooling: lets a change of the kernel's logic be
// checked where there is no GPU.  Usage: emu_k1 [class 0..4] [blocks] [block_bytes] [accel] [seed]
#include "emu_cuda.h"
#include "../../zig-lz4_b200/csrc/k_compress_fast.cu"
#include <vector>
#include <string>
#include "../../oracle/b2o.h"

extern "C" int b2gen_fill(uint8_t* dst, uint64_t n, uint64_t seed, uint32_t mode, uint64_t span, int nthreads);

namespace b2 { EmuStats g_emu_stats; }

template <typename TableT>
static int run_block(const uint8_t* src, uint32_t n, uint32_t cap, uint32_t accel, std::vector<uint8_t>& out, uint32_t& olen) {
    static TableT table[b2::HASH_ENTRIES];
    out.assign((size_t)cap + 64, 0xEE);
    uint32_t r_olen = 0;
    int r_st = 0;
    emu::Guarded gd(cap ? cap : 1, true);          // a write past the capacity faults
    uint8_t* dst = gd.p;
    emu::run_warp([&] {
        b2::Ring ring{};
        uint32_t ol;
        int st;
        b2::compress_block<TableT, false>(src, n, dst, cap, table, accel, b2::lane_id(), ol, st, ring);
        if (b2::lane_id() == 0) { r_olen = ol; r_st = st; }
    });
    olen = r_olen;
    memcpy(out.data(), dst, cap);
    return r_st;
}

static uint64_t g_checked = 0, g_failed = 0;

static void check(const uint8_t* src0, uint32_t n, uint32_t cap, uint32_t accel, bool wide, const char* what) {
    std::vector<uint8_t> got, want((size_t)cap + 64);
    uint32_t olen = 0;
    // the input sits against an inaccessible page: behind its last byte on even cases, in front of the granule of its first
    // byte on odd ones (emu::Guarded) — an access outside the promised granules is a crash, not a silent read
    static uint64_t flip = 0;
    const bool at_end = (flip++ & 1) == 0;
    emu::Guarded g(n ? n : 1, at_end, (size_t)(flip * 5 % 16));
    memcpy(g.p, src0, n);
    const uint8_t* src = g.p;
    const int st = wide ? run_block<uint32_t>(src, n, cap, accel, got, olen) : run_block<uint16_t>(src, n, cap, accel, got, olen);
    size_t wlen = 0;
    const int wst = b2o_compress_fast(src, n, want.data(), cap, accel, &wlen);
    g_checked++;
    bool ok = st == wst;
    if (ok && st == 0) ok = olen == wlen && memcmp(got.data(), want.data(), wlen) == 0;
    if (!ok) {
        g_failed++;
        size_t d = 0;
        while (d < wlen && d < olen && got[d] == want[d]) d++;
        fprintf(stderr, "MISMATCH %s n=%u cap=%u accel=%u table=%s: status %d/%d len %u/%zu first diff at %zu\n", what, n, cap,
                accel, wide ? "u32" : "u16", st, wst, olen, wlen, d);
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
        // file mode: every @path is one block (u16 tables up to 64 KiB, u32 above) at accelerations 1 and 3, and cut into
        // 64 KiB blocks — the reference's own test inputs go through here (tests/test_warp_emu.py)
        for (int a = 1; a < argc; a++) {
            const std::vector<uint8_t> d = read_file(argv[a] + 1);
            const uint32_t n = (uint32_t)d.size();
            for (uint32_t accel : {1u, 3u}) check(d.data(), n, (uint32_t)b2o_compress_bound(n), accel, n > 65536, argv[a]);
            if (n <= 65536) check(d.data(), n, (uint32_t)b2o_compress_bound(n), 1, true, argv[a]);
            for (uint32_t o = 0; n > 65536 && o < n; o += 65536) {
                const uint32_t len = n - o < 65536 ? n - o : 65536;
                check(d.data() + o, len, (uint32_t)b2o_compress_bound(len), 1, false, argv[a]);
            }
        }
        printf("files: %llu cases, %llu failed\n", (unsigned long long)g_checked, (unsigned long long)g_failed);
        return g_failed ? 1 : 0;
    }
    const int cls = argc > 1 ? atoi(argv[1]) : 4;
    const uint32_t blocks = argc > 2 ? (uint32_t)atoi(argv[2]) : 8;
    const uint32_t bs = argc > 3 ? (uint32_t)atoi(argv[3]) : 65536;
    const uint32_t accel = argc > 4 ? (uint32_t)atoi(argv[4]) : 1;
    const uint64_t seed = argc > 5 ? strtoull(argv[5], nullptr, 0) : 0x4C5A3442ull;
    const uint64_t total = (uint64_t)blocks * (bs < 65536 ? 65536 : bs);
    std::vector<uint8_t> data(total + 64);
    b2gen_fill(data.data(), total, seed, (uint32_t)cls, 65536, 4);
    const uint32_t bound = (uint32_t)b2o_compress_bound(bs);
    for (uint32_t b = 0; b < blocks; b++) {
        const uint8_t* src = data.data() + (uint64_t)b * (bs < 65536 ? 65536 : bs);
        check(src, bs, bound, accel, bs > 65536, "full");
        if (b == 0) {
            // edge sizes and capacities on the first block
            for (uint32_t n : {0u, 1u, 4u, 11u, 12u, 13u, 14u, 20u, 31u, 32u, 33u, 44u, 45u, 63u, 64u, 65u, 100u, 127u, 128u, 129u, 255u, 300u, 1000u, 4095u, 4096u, 4097u})
                if (n <= bs) check(src, n, (uint32_t)b2o_compress_bound(n), accel, false, "small");
            for (uint32_t n : {300u, 4096u})
                if (n <= bs)
                    for (uint32_t cap : {0u, 1u, 5u, 17u, 100u, 200u, 1000u, 3000u}) check(src, n, cap, accel, false, "cap");
            if (bs >= 8192) check(src, 8192, (uint32_t)b2o_compress_bound(8192), accel, true, "u32-table");
            // blocks that do not start on a word boundary (batch API with arbitrary offsets)
            for (uint32_t sh : {1u, 2u, 3u, 5u})
                for (uint32_t n : {13u, 40u, 77u, 1000u, 20000u})
                    if (sh + n <= bs) check(src + sh, n, (uint32_t)b2o_compress_bound(n), accel, false, "unaligned");
        }
    }
    const b2::EmuStats& s = b2::g_emu_stats;
    printf("class %d: %llu cases, %llu failed | windows %llu (fast-chain %llu, bailed %llu), sequences %llu (batched %llu)\n", cls,
           (unsigned long long)g_checked, (unsigned long long)g_failed, (unsigned long long)s.windows,
           (unsigned long long)s.fast_windows, (unsigned long long)s.bailed_windows, (unsigned long long)s.sequences,
           (unsigned long long)s.batched_sequences);
    return g_failed ? 1 : 0;
}
