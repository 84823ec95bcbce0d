// emu_k2.cpp — runs K2's device code (k_decompress.cu, compiled for the host with -DB2_EMU) on the one-warp emulator:
// streams produced by the oracle's compressor (and damaged copies of them) are decoded by the chunked front end, the
// serial front end and the exact tier, and bytes, sizes and status codes are compared with the oracle's decoder.
// Usage: emu_k2 [class 0..4] [blocks] [block_bytes] [seed]
#include "emu_cuda.h"
#include "../../zig-lz4_b200/csrc/k_decompress.cu"
#include "../../oracle/b2o.h"

extern "C" int b2gen_fill(uint8_t* dst, uint64_t n, uint64_t seed, uint32_t mode, uint64_t span, int nthreads);

static uint64_t g_checked = 0, g_failed = 0;

// tier: 0 chunked front end, 1 serial front end, 2 exact tier only
static int run_decode(int tier, const uint8_t* src, uint32_t n, uint8_t* dst, uint32_t cap, const uint8_t* dict, uint32_t dict_len,
                      uint32_t& olen) {
    static b2::WarpStage stage;
    memset(&stage, 0, sizeof stage);
    uint32_t r_olen = 0;
    int r_st = 0;
    emu::run_warp([&] {
        const uint32_t lane = b2::lane_id();
        uint32_t ol = 0;
        int st = 0;
        if (tier == 0) b2::decode_block_fast(src, n, dst, cap, lane, &stage, ol, st);
        else if (tier == 1) b2::decode_block_fast_v1(src, n, dst, cap, dict, dict_len, dict != nullptr, lane, 0, 0, ol, st);
        else b2::decode_block<true>(src, n, dst, cap, dict, dict_len, dict != nullptr, lane, 0, 0, ol, st);
        if (lane == 0) { r_olen = ol; r_st = st; }
    });
    olen = r_olen;
    return r_st;
}

static void check(const uint8_t* comp, uint32_t clen, uint32_t cap, const char* what) {
    std::vector<uint8_t> want((size_t)cap + 64, 0xEE), got((size_t)cap + 64);
    size_t wlen = 0;
    const int wst = b2o_decompress_safe(comp, clen, want.data(), cap, &wlen);
    for (int tier = 0; tier < 3; tier++) {
        std::fill(got.begin(), got.end(), 0xEE);
        uint32_t olen = 0;
        // the stream sits at the very end of its allocation: a read past it would fault under a checker, and damaged
        // streams exercise the bounds logic
        static uint64_t flip = 0;
        const bool at_end = (flip++ & 1) == 0;
        emu::Guarded gs(clen ? clen : 1, at_end, (size_t)(flip * 7 % 16));
        memcpy(gs.p, comp, clen);
        emu::Guarded gd(cap ? cap : 1, (flip & 2) != 0, (size_t)(flip * 3 % 16));   // output: guard behind its end or in front of its start
        const int st = run_decode(tier, gs.p, clen, gd.p, cap, nullptr, 0, olen);
        memcpy(got.data(), gd.p, cap);
        g_checked++;
        bool ok = st == wst;
        if (ok && st == 0) ok = olen == wlen && memcmp(got.data(), want.data(), wlen) == 0;
        for (size_t i = cap; ok && i < (size_t)cap + 64; i++) ok = got[i] == 0xEE;
        if (!ok) {
            g_failed++;
            fprintf(stderr, "MISMATCH %s tier %d clen=%u cap=%u: status %d/%d len %u/%zu\n", what, tier, clen, cap, st, wst, olen, wlen);
        }
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
        // file mode: every @path is compressed by the oracle (fast mode and HC 9) and decoded by all three tiers
        for (int a = 1; a < argc; a++) {
            const std::vector<uint8_t> d = read_file(argv[a] + 1);
            const uint32_t n = (uint32_t)d.size();
            std::vector<uint8_t> comp(b2o_compress_bound(n) + 64);
            size_t clen = 0;
            b2o_compress_fast(d.data(), n, comp.data(), comp.size(), 1, &clen);
            check(comp.data(), (uint32_t)clen, n, argv[a]);
            if (n) check(comp.data(), (uint32_t)clen, n - 1, argv[a]);
            b2o_compress_hc(d.data(), n, comp.data(), comp.size(), 9, &clen);
            check(comp.data(), (uint32_t)clen, n + 3, argv[a]);
        }
        printf("files: %llu decodes checked, %llu failed\n", (unsigned long long)g_checked, (unsigned long long)g_failed);
        return g_failed ? 1 : 0;
    }
    const int cls = argc > 1 ? atoi(argv[1]) : 4;
    const uint32_t blocks = argc > 2 ? (uint32_t)atoi(argv[2]) : 4;
    const uint32_t bs = argc > 3 ? (uint32_t)atoi(argv[3]) : 65536;
    const uint64_t seed = argc > 4 ? strtoull(argv[4], nullptr, 0) : 0x4C5A3442ull;
    const uint64_t stride = bs < 65536 ? 65536 : bs;
    std::vector<uint8_t> data(stride * blocks + 64);
    b2gen_fill(data.data(), stride * blocks, seed, (uint32_t)cls, 65536, 4);
    std::vector<uint8_t> comp(b2o_compress_bound(bs) + 64);
    uint64_t rng = seed * 0x9E3779B97F4A7C15ull + 1;
    auto rnd = [&]() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; };
    for (uint32_t b = 0; b < blocks; b++) {
        const uint8_t* src = data.data() + b * stride;
        size_t clen = 0;
        b2o_compress_fast(src, bs, comp.data(), comp.size(), 1, &clen);
        check(comp.data(), (uint32_t)clen, bs, "exact-capacity");
        check(comp.data(), (uint32_t)clen, bs + 100, "roomy");
        if (b == 0) {
            check(comp.data(), (uint32_t)clen, bs - 1, "one-short");          // OutputTooSmall
            check(comp.data(), (uint32_t)clen, bs / 2, "half");
            check(comp.data(), (uint32_t)clen, 0, "cap0");
            check(comp.data(), 0, bs, "empty");
            for (uint32_t cut : {1u, 2u, 5u, 17u, 18u, 19u, 100u, 1000u})
                if (cut < clen) check(comp.data(), (uint32_t)clen - cut, bs, "truncated");
            for (int k = 0; k < 24; k++) {                                     // damaged streams: status codes must agree
                std::vector<uint8_t> bad(comp.begin(), comp.begin() + clen);
                const size_t at = rnd() % clen;
                bad[at] ^= (uint8_t)(1u << (rnd() & 7));
                if (k & 1) bad[rnd() % clen] = 0;
                check(bad.data(), (uint32_t)clen, bs, "damaged");
            }
            for (uint32_t n : {13u, 40u, 300u, 5000u}) {                       // short blocks
                size_t c2 = 0;
                b2o_compress_fast(src, n, comp.data(), comp.size(), 1, &c2);
                check(comp.data(), (uint32_t)c2, n, "short");
            }
        }
    }
    // hand-built streams: every small offset (each period class of the overlapping-match copies) with short and long lengths,
    // at varying output alignments
    if (cls == 2) {
        const uint32_t offs[] = {1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 24, 31, 32, 33, 48, 64, 65, 100, 255, 256, 1000};
        const uint32_t lens[] = {4, 5, 15, 16, 17, 19, 31, 32, 33, 63, 64, 65, 100, 127, 128, 129, 255, 256, 300, 511, 1024, 4000};
        for (uint32_t off : offs)
            for (uint32_t ml : lens) {
                std::vector<uint8_t> st;
                auto put_len = [&](uint32_t v) { while (v >= 255) { st.push_back(255); v -= 255; } st.push_back((uint8_t)v); };
                uint32_t total = 0;
                for (int rep = 0; rep < 3; rep++) {
                    const uint32_t LL = off + (uint32_t)(rnd() % 7) + (rep == 0 ? 0 : 1);     // enough history, shifting alignment
                    const uint32_t mlc = ml - 4;
                    st.push_back((uint8_t)(((LL < 15 ? LL : 15) << 4) | (mlc < 15 ? mlc : 15)));
                    if (LL >= 15) put_len(LL - 15);
                    for (uint32_t i = 0; i < LL; i++) st.push_back((uint8_t)rnd());
                    st.push_back((uint8_t)off); st.push_back((uint8_t)(off >> 8));
                    if (mlc >= 15) put_len(mlc - 15);
                    total += LL + ml;
                }
                st.push_back(0x50); for (int i = 0; i < 5; i++) st.push_back((uint8_t)rnd());   // last literals
                total += 5;
                check(st.data(), (uint32_t)st.size(), total, "hand-built");
                check(st.data(), (uint32_t)st.size(), total + 17, "hand-built roomy");
            }
    }
    printf("K2 class %d: %llu decodes checked, %llu failed\n", cls, (unsigned long long)g_checked, (unsigned long long)g_failed);
    return g_failed ? 1 : 0;
}
