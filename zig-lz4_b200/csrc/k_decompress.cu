// k_decompress.cu — K2: LZ4 block decompressor, many independent blocks, one warp per block.
//
// Semantics: lz4.decompressSafe / decompressSafeUsingDict == decompressGeneric with
// targetOutputSize == dst.len, /root/reference/src/lz4.zig:89-251, including its exact error kinds
// and order of checks, its leniency (no last-5-literals / MFLIMIT end rules) and its two quirks
// (src.len == 0 -> 0, dst.len == 0 -> 0 without error, :97-98).
//
// Two tiers share one block:
//   FAST tier (decode_block_fast): the token chain is the only serial part of LZ4 decoding.  Where a token
//   starts depends on the whole chain before it, but how long a token is does not: it is a function of the
//   bytes at its position alone.  So the warp first loads the next 256 bytes of the stream coalesced into
//   shared memory and computes, for EVERY byte position at once (8 positions per lane, four per 32-bit SIMD
//   word), the distance to the next token *if* a token started there: 3 + LL for a plain token, the same plus
//   the length-extension bytes for one LL byte and up to two ML bytes, 0 ("halt") for anything longer, too
//   close to the end of the stream, or running off it.  The serial part is then a chain of 32 dependent
//   shared-memory reads `p += delta[p]` (3 instructions per sequence instead of ~19: the round-1 walk issued a
//   global byte load and decoded the token on all 32 lanes for every sequence — half of the kernel's
//   instructions).  Sequence k goes to lane k, which decodes its own token from the staged bytes, and then
//   all 32 sequences are expanded at once: a warp scan of (LL + ML) gives every lane its output position, each lane copies
//   its own literal run, and the match copies are resolved in dependency rounds (a lane may copy once
//   no pending match of an earlier lane overlaps its source range; the first pending lane always may,
//   so the rounds terminate; a lane's own forward overlap, offset < length, is safe because one thread
//   copies in ascending order).  Sequences that are long (> LONG_SEQ bytes) are copied cooperatively by
//   the whole warp with 16-byte vector accesses.  Anything irregular — a stream end within 18 bytes,
//   offset 0, a match reaching before dst (dictionary or corruption), output overflow — is not
//   decided here: the fast tier stops at the start of that batch and the exact tier takes over.
//   EXACT tier (decode_block): lock-step parse with the reference's checks in the reference's order — behind the fast
//   tier for everything irregular, and the whole decoder of streams that shrink their block more than 6 x (runs and long
//   matches: nearly every sequence is long, and this loop is compact — the kernel is ~9700 instructions, K2 notes in
//   DESIGN.md);
//   the byte work is spread over the lanes:
//   * 255-run length extensions are scanned 32 bytes at a time with a ballot,
//   * literals are copied with 16-byte aligned stores (b2::warp_copy, read-only source path),
//   * a match copy with offset >= length is a plain vector copy; an overlapping match (offset <
//     length, :235-241) is produced by period doubling: after `offset` bytes are copied the region
//     [match, op+offset) is periodic with period `offset`, so each round copies twice as much from
//     the fixed start of the match — same bytes as the reference's forward byte loop.
//   * matches that start in the external dictionary (:181-228) copy the dictionary tail first, then
//     continue at dst[0..].
#include "b2_common.cuh"
#ifndef B2_EMU   // (B2_EMU: host build of the device code for the one-warp emulator, tools/warp_emu)
#include "b2_kernels.h"
#endif

namespace b2 {

constexpr int K2_WARPS = 4;
constexpr int K2_THREADS = K2_WARPS * 32;

// 255-run extension (src/lz4.zig:123-131 / :160-168).  Returns false => CorruptedData.
// `len` saturates at 2^31 so that the callers' range checks stay exact for hostile inputs.
__device__ __forceinline__ bool read_len_ext(const uint8_t* __restrict__ src, uint32_t& ip, uint32_t iend, uint32_t& len,
                                             uint32_t lane) {
    for (;;) {
        uint32_t idx = ip + lane;
        uint32_t b = idx < iend ? (uint32_t)__ldg(src + idx) : 256u;  // 256 = ran off the input
        uint32_t m = __ballot_sync(FULL, b != 255u);
        if (m) {
            uint32_t f = (uint32_t)__ffs(m) - 1;
            uint32_t bf = __shfl_sync(FULL, b, f);
            if (bf == 256u) return false;
            uint64_t t = (uint64_t)len + 255ull * f + bf;
            len = t > 0x80000000ull ? 0x80000000u : (uint32_t)t;
            ip += f + 1;
            return true;
        }
        uint64_t t = (uint64_t)len + 255ull * 32;
        len = t > 0x80000000ull ? 0x80000000u : (uint32_t)t;
        ip += 32;
    }
}

__device__ __forceinline__ uint32_t byte_of(const uint4& v, uint32_t i) {   // byte i (0..15) of a 16-byte vector
    const uint32_t w = i < 8 ? (i < 4 ? v.x : v.y) : (i < 12 ? v.z : v.w);
    return (w >> ((i & 3) * 8)) & 0xFFu;
}

// Overlapping match whose period divides 16 (byte / u16 / u32 / u64 runs, 16-byte patterns): the output is one 16-byte
// vector repeated, so it is read once and written with aligned 16-byte stores — no read-back of what was just written
// (the doubling rounds below are a store -> load round trip through L2 each: eight of them for a 255-byte run).
__device__ __forceinline__ void match_fill_pow2(uint8_t* d, const uint8_t* s, uint32_t offset, uint32_t ml, uint32_t lane) {
    const uint32_t bo = (uint32_t)(reinterpret_cast<uintptr_t>(s) & 15);
    const uint4* a16 = reinterpret_cast<const uint4*>(s - bo);
    const uint4 A = a16[0];
    const uint4 B = bo + offset > 16 ? a16[1] : A;            // only if the pattern reaches into it (then it holds pattern bytes)
    const uint4 X = extract16(A, B, bo);                       // s[0..15]; only the first `offset` bytes are pattern
    uint4 P;                                                   // P[b] = s[b mod offset]
    if (offset == 16) P = X;
    else if (offset == 8) P = make_uint4(X.x, X.y, X.x, X.y);
    else {
        const uint32_t w = offset == 4 ? X.x : (offset == 2 ? (X.x & 0xFFFFu) * 0x00010001u : (X.x & 0xFFu) * 0x01010101u);
        P = make_uint4(w, w, w, w);
    }
    const uint32_t head = (uint32_t)((16 - (reinterpret_cast<uintptr_t>(d) & 15)) & 15);
    if (ml <= head) { if (lane < ml) d[lane] = (uint8_t)byte_of(P, lane); return; }
    if (lane < head) d[lane] = (uint8_t)byte_of(P, lane);
    const uint4 V = extract16(P, P, head & (offset - 1));       // the vector at every 16-aligned output position
    const uint32_t rest = ml - head, nvec = rest >> 4, tail = rest & 15;
    uint4* d16 = reinterpret_cast<uint4*>(d + head);
    for (uint32_t i = lane; i < nvec; i += 32) d16[i] = V;
    if (lane < tail) d[head + (nvec << 4) + lane] = (uint8_t)byte_of(V, lane);
}

// Self-overlapping match (offset < ml), every method.  Used at ONE site, the serial front end's long matches (runs and
// repeated patterns are what that front end is chosen for).  The decoder is 9700 instructions: with this code inlined at
// all four match-copy sites the mixed workload waited 3.4 instead of 1.2 cycles per issue for instructions (blocks of
// different kinds are decoded side by side) and lost more than the runs gained; as an out-of-line function the call
// cost the chunked decoder's hot loop its registers (text 4.45 -> 4.73 ms per GiB).  The other sites keep the plain
// doubling copy.
__device__ __forceinline__ void match_copy_overlap(uint8_t* d, const uint8_t* s, uint32_t offset, uint32_t ml, uint32_t lane) {
    if (offset <= 16 && (offset & (offset - 1)) == 0) {
        match_fill_pow2(d, s, offset, ml, lane);
        return;
    }
    if (offset >= 16 && ml >= 2 * offset + 64) {
        // One period first; then [s, s + 2 * offset) is valid and periodic, and every 16-byte vector of the rest can be
        // read from inside it at its own phase: two rounds whatever the length (doubling: log2(ml / offset) + 1).
        warp_copy<false>(d, s, offset, lane);
        __syncwarp();
        uint8_t* d2 = d + offset;
        const uint32_t rest0 = ml - offset;
        const uint32_t head = (uint32_t)((16 - (reinterpret_cast<uintptr_t>(d2) & 15)) & 15);
        if (lane < head) d2[lane] = s[lane];                    // (head < 16 <= offset)
        const uint32_t rest = rest0 - head, nvec = rest >> 4, tail = rest & 15;
        uint4* d16 = reinterpret_cast<uint4*>(d2 + head);
        for (uint32_t i = lane; i < nvec; i += 32) {
            const uint8_t* ps = s + (head + 16 * i) % offset;
            const uint32_t bo = (uint32_t)(reinterpret_cast<uintptr_t>(ps) & 15);
            const uint4* a16 = reinterpret_cast<const uint4*>(ps - bo);
            const uint4 A = a16[0];
            const uint4 B = bo ? a16[1] : A;
            d16[i] = extract16(A, B, bo);
        }
        if (lane < tail) d2[head + (nvec << 4) + lane] = s[(head + (nvec << 4) + lane) % offset];
        __syncwarp();
        return;
    }
    uint32_t copied = 0, avail = offset;
    while (copied < ml) {
        uint32_t chunk = ml - copied < avail ? ml - copied : avail;
        warp_copy<false>(d + copied, s, chunk, lane);
        __syncwarp();
        copied += chunk;
        avail += chunk;
    }
}

// Forward-overlap-safe copy of `ml` bytes to d from s = d - offset (both inside the output buffer).
template <bool FULL_METHODS = false>
__device__ __forceinline__ void match_copy(uint8_t* d, const uint8_t* s, uint32_t offset, uint32_t ml, uint32_t lane) {
    if (offset >= ml) { warp_copy<false>(d, s, ml, lane); return; }
    if (FULL_METHODS) { match_copy_overlap(d, s, offset, ml, lane); return; }
    uint32_t copied = 0, avail = offset;     // period doubling
    while (copied < ml) {
        uint32_t chunk = ml - copied < avail ? ml - copied : avail;
        warp_copy<false>(d + copied, s, chunk, lane);
        __syncwarp();
        copied += chunk;
        avail += chunk;
    }
}

// One sequence of the reference's loop (src/lz4.zig:111-248) at (ip, op), checks in the reference's order.
// Returns 0 = go on, 1 = the stream ended normally (:113 / :146), 2 = error (st set).
template <bool WRITE>
__device__ __forceinline__ int exact_step(const uint8_t* __restrict__ src, uint32_t n, uint8_t* dst, uint32_t cap,
                                          const uint8_t* __restrict__ dict, uint32_t dict_len, bool has_dict, uint32_t lane,
                                          uint32_t& ip, uint32_t& op, int& st) {
    const uint32_t iend = n, oend = cap;
    if (ip >= iend) return 1;                                       // :113
    const uint32_t token = __ldg(src + ip);
    ip += 1;
    uint32_t LL = token >> 4;
    if (LL == RUN_MASK) {                                           // :123
        if (!read_len_ext(src, ip, iend, LL, lane)) { st = ST_CORRUPTED; return 2; }
    }
    if (LL > 0) {                                                   // :134
        if ((uint64_t)ip + LL > iend) { st = ST_CORRUPTED; return 2; }
        if ((uint64_t)op + LL > oend) { st = ST_OUTPUT_TOO_SMALL; return 2; }
        if (WRITE) warp_copy<true>(dst + op, src + ip, LL, lane);
        ip += LL;
        op += LL;
    }
    if (ip >= iend) return 1;                                       // :146
    if (ip + 2 > iend) { st = ST_CORRUPTED; return 2; }             // :149
    const uint32_t offset = (uint32_t)__ldg(src + ip) | ((uint32_t)__ldg(src + ip + 1) << 8);
    ip += 2;
    if (offset == 0) { st = ST_CORRUPTED; return 2; }               // :154
    uint32_t ML = token & ML_MASK;
    if (ML == ML_MASK) {                                            // :160
        if (!read_len_ext(src, ip, iend, ML, lane)) { st = ST_CORRUPTED; return 2; }
    }
    ML += MINMATCH;                                                 // :171
    if ((uint64_t)op + ML > oend) { st = ST_OUTPUT_TOO_SMALL; return 2; }  // :174
    if (offset > op) {                                              // :181 match starts before dst
        if (!has_dict) { st = ST_CORRUPTED; return 2; }             // :183-186
        if ((uint64_t)offset > (uint64_t)op + dict_len) { st = ST_CORRUPTED; return 2; }  // :190
        const uint32_t back = offset - op;                          // lowPrefixOffset, :195
        const uint8_t* dm = dict + dict_len - back;                 // :196
        if (WRITE) {
            __syncwarp();
            if (ML <= back) {                                       // :199
                warp_copy<true>(dst + op, dm, ML, lane);
            } else {
                warp_copy<true>(dst + op, dm, back, lane);
                __syncwarp();
                // rest continues at dst[0..] (:213-227): forward copy, may overlap itself
                match_copy(dst + op + back, dst, op + back, ML - back, lane);
            }
            __syncwarp();
        }
        op += ML;
    } else {
        if (WRITE) {
            __syncwarp();  // literals written by other lanes must be visible
            match_copy<true>(dst + op, dst + op - offset, offset, ML, lane);  // :232-246
            __syncwarp();
        }
        op += ML;
    }
    return 0;
}

template <bool WRITE>
__device__ void decode_block(const uint8_t* __restrict__ src, uint32_t n, uint8_t* dst, uint32_t cap,
                             const uint8_t* __restrict__ dict, uint32_t dict_len, bool has_dict, uint32_t lane,
                             uint32_t ip, uint32_t op, uint32_t& olen, int& st) {
    st = ST_OK;
    olen = 0;
    if (n == 0) return;    // :97
    if (cap == 0) return;  // :98
    for (;;) {
        const int r = exact_step<WRITE>(src, n, dst, cap, dict, dict_len, has_dict, lane, ip, op, st);
        if (r == 2) return;
        if (r == 1) break;
    }
    olen = op;
}


// ------------------------------------------------------------------------------------------------
// FAST tier.  Sequences whose copies exceed LONG_SEQ bytes are copied by the whole warp.
constexpr uint32_t LONG_SEQ = 24;
// A plain token (LL < 15, ML nibble < 15) at position t touches bytes up to t + 1 + 14 + 2.
constexpr uint32_t PLAIN_SPAN = 18;
// Length fields above this are left to the exact tier (keeps the 32-bit prefix sums exact).
constexpr uint32_t FAST_LEN_MAX = 1u << 24;

#ifdef B2_EMU
__device__ __forceinline__ void prefetch_l2(const void*) {}
__device__ __forceinline__ void prefetch_l1(const void*) {}
#else
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
#endif

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, uint32_t lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(FULL, v, d);
        if (lane >= (uint32_t)d) v += t;
    }
    return v;
}

// Number of lanes whose (non-decreasing across lanes) value is <= x, capped at 31.
__device__ __forceinline__ uint32_t count_le(uint32_t sorted_v, uint32_t x) {
    uint32_t c = 0;
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
        uint32_t v = __shfl_sync(FULL, sorted_v, c + step - 1);
        if (v <= x) c += step;
    }
    return c;
}

// Expands one batch: lanes < k hold a sequence each (literals at src + myLit, myLL of them, then a 2-byte offset; match
// length myML), in stream order; output continues at op.  Returns false — nothing this batch wrote matters — when a
// sequence needs the exact tier's judgement: offset 0 (:154), a match reaching before dst (:181 / :231), output
// overflow (:137 / :174).
template <bool TWO_ROUND>   // true (serial front end: runs, repeated patterns): overlapping long matches by match_copy_overlap
__device__ __forceinline__ bool expand_batch(const uint8_t* __restrict__ src, uint8_t* dst, uint32_t cap, uint32_t lane,
                                             uint32_t k, uint32_t myLit, uint32_t myLL, uint32_t myML, uint32_t& op) {
    const bool valid = lane < k;
    uint32_t off = 1;
    if (valid) off = (uint32_t)__ldg(src + myLit + myLL) | ((uint32_t)__ldg(src + myLit + myLL + 1) << 8);
    const uint32_t len = myLL + myML;  // 0 on idle lanes
    const uint32_t incl = warp_incl_scan(len, lane);
    const uint32_t o0 = op + incl - len;         // where this sequence's literals go
    const uint32_t ms = o0 + myLL;               // where its match goes
    const uint32_t batch_len = __shfl_sync(FULL, incl, 31);
    const bool odd = valid && (off == 0 || off > ms);
    if (__ballot_sync(FULL, odd) != 0 || batch_len > cap - op) return false;

    // literals: every lane copies its own short run, long runs go warp-wide
    {
        const bool lng = myLL > LONG_SEQ;
        const uint32_t nl = lng ? 0u : myLL;
        const uint32_t mx = __reduce_max_sync(FULL, nl);
        const uint8_t* ls = src + myLit;
        uint8_t* ld = dst + o0;
        uint32_t rem = nl;
        for (uint32_t i0 = 0; i0 < mx; i0 += 4) {
            uint32_t r[4] = {0u, 0u, 0u, 0u};   // (initialised: otherwise the predicated loads below make the compiler carry the old values through local memory)
#pragma unroll
            for (int u = 0; u < 4; u++)
                if ((uint32_t)u < rem) r[u] = __ldg(ls + u);
#pragma unroll
            for (int u = 0; u < 4; u++)
                if ((uint32_t)u < rem) ld[u] = (uint8_t)r[u];
            ls += 4; ld += 4;
            rem = rem > 4 ? rem - 4 : 0u;
        }
        uint32_t lm = __ballot_sync(FULL, lng);
        while (lm) {
            const int j = __ffs(lm) - 1;
            lm &= lm - 1;
            warp_copy<true>(dst + __shfl_sync(FULL, o0, j), src + __shfl_sync(FULL, myLit, j),
                            __shfl_sync(FULL, myLL, j), lane);
        }
    }
    __syncwarp();

    // matches: dependency rounds
    {
        const uint32_t s = ms - off;                           // source start
        const uint32_t e = (off < myML) ? ms : s + myML;       // source end outside its own output
        const uint32_t mend = valid ? ms + myML : 0xFFFFFFFFu; // non-decreasing across lanes
        const uint32_t mstart = valid ? ms : 0xFFFFFFFFu;
        uint32_t dep = 0;
        if (__ballot_sync(FULL, valid && e > op) != 0) {
            // lanes [lo, hi) hold the matches that overlap [s, e): mend_j > s and ms_j < e
            const uint32_t lo = count_le(mend, s);
            const uint32_t hi = count_le(mstart, e - 1);
            dep = hi > lo ? ((1u << hi) - 1u) & ~((1u << lo) - 1u) : 0u;
        }
        const bool lng = myML > LONG_SEQ;
        // a short self-overlapping match with offset < 8 repeats a pattern that fits one register pair
        const bool tiny = !lng && off < 8 && off < myML;
        bool pending = valid;
        uint32_t pm = __ballot_sync(FULL, pending);
        while (pm) {
            const bool go = pending && (pm & dep) == 0;
            uint8_t* md = dst + ms;
            const uint8_t* msrc = dst + s;
            {   // Non-overlapping short matches (offset >= length, the common case): the whole source, up to 24 bytes, is
                // read at once as up to four aligned 8-byte words (one memory round trip for the round instead of one
                // per 8 bytes — these reads go to L2 or DRAM, the output of 4700 warps does not fit any cache), realigned
                // in registers, and stored byte by byte.  Words may hold bytes around the source range (not yet written
                // or another sequence's): they are never stored.
                const bool wide = go && !lng && off >= myML;
                const uint32_t nw = wide ? myML : 0u;
                const uint32_t mxw = __reduce_max_sync(FULL, nw);
                if (mxw) {
                    const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(msrc) & 7);
                    const uint2* wp = reinterpret_cast<const uint2*>(msrc - sh);
                    const uint32_t nwords = wide ? (sh + myML + 7) >> 3 : 0u;     // 1..4
                    uint2 A = make_uint2(0u, 0u), B = A, C = A, D = A;
                    if (nwords > 0) A = wp[0];
                    if (nwords > 1) B = wp[1];
                    if (nwords > 2) C = wp[2];
                    if (nwords > 3) D = wp[3];
                    const bool hiw = sh >= 4;
                    const uint32_t fs = (sh & 3) * 8;
                    const uint32_t x0 = hiw ? A.y : A.x, x1 = hiw ? B.x : A.y, x2 = hiw ? B.y : B.x, x3 = hiw ? C.x : B.y,
                                   x4 = hiw ? C.y : C.x, x5 = hiw ? D.x : C.y, x6 = hiw ? D.y : D.x;
                    uint32_t o[6];
                    o[0] = __funnelshift_r(x0, x1, fs); o[1] = __funnelshift_r(x1, x2, fs); o[2] = __funnelshift_r(x2, x3, fs);
                    o[3] = __funnelshift_r(x3, x4, fs); o[4] = __funnelshift_r(x4, x5, fs); o[5] = __funnelshift_r(x5, x6, fs);
#pragma unroll
                    for (int u = 0; u < 8; u++)
                        if ((uint32_t)u < nw) md[u] = (uint8_t)(o[u >> 2] >> (8 * (u & 3)));
                    if (mxw > 8) {
#pragma unroll
                        for (int u = 8; u < 16; u++)
                            if ((uint32_t)u < nw) md[u] = (uint8_t)(o[u >> 2] >> (8 * (u & 3)));
                    }
                    if (mxw > 16) {
#pragma unroll
                        for (int u = 16; u < 24; u++)
                            if ((uint32_t)u < nw) md[u] = (uint8_t)(o[u >> 2] >> (8 * (u & 3)));
                    }
                }
                // Self-overlapping short matches with offset >= 8: 8 bytes at a time, all loads of a chunk before its
                // stores (a chunk only reads bytes of earlier chunks).
                const uint32_t nm = (go && !lng && !tiny && off < myML) ? myML : 0u;
                const uint32_t mx = __reduce_max_sync(FULL, nm);
                const uint8_t* ps = msrc;
                uint8_t* pd = md;
                uint32_t rem = nm;
                for (uint32_t i0 = 0; i0 < mx; i0 += 8) {
                    uint32_t r[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll
                    for (int u = 0; u < 8; u++)
                        if ((uint32_t)u < rem) r[u] = ps[u];
#pragma unroll
                    for (int u = 0; u < 8; u++)
                        if ((uint32_t)u < rem) pd[u] = (uint8_t)r[u];
                    ps += 8; pd += 8;
                    rem = rem > 8 ? rem - 8 : 0u;
                }
            }
            if (__ballot_sync(FULL, go && tiny)) {
                if (go && tiny) {
                    uint64_t pat = 0;
                    for (uint32_t i = 0; i < off; i++) pat |= (uint64_t)msrc[i] << (8 * i);
                    uint32_t j = 0;
                    for (uint32_t i = 0; i < myML; i++) {
                        md[i] = (uint8_t)(pat >> (8 * j));
                        j = (j + 1 == off) ? 0u : j + 1;
                    }
                }
                __syncwarp();
            }
            uint32_t gm = __ballot_sync(FULL, go && lng);
            while (gm) {
                const int j = __ffs(gm) - 1;
                gm &= gm - 1;
                const uint32_t jms = __shfl_sync(FULL, ms, j), joff = __shfl_sync(FULL, off, j);
                match_copy<TWO_ROUND>(dst + jms, dst + jms - joff, joff, __shfl_sync(FULL, myML, j), lane);
            }
            if (go) pending = false;
            __syncwarp();
            pm = __ballot_sync(FULL, pending);
        }
    }
    op += batch_len;
    return true;
}

// Serial front end: the warp walks the token chain itself, one token per step, every kind of length extension
// handled in place.  It is the round-1 decoder (b2lz4_debug_tune("k2_variant", 1) selects it for A/B runs) and the
// chunked decoder's fallback for streams whose tokens keep defeating its speculation (long runs: 255-extended lengths).
__device__ void decode_block_fast_v1(const uint8_t* __restrict__ src, uint32_t n, uint8_t* dst, uint32_t cap,
                                     const uint8_t* __restrict__ dict, uint32_t dict_len, bool has_dict, uint32_t lane,
                                     uint32_t ip, uint32_t op, uint32_t& olen, int& st) {
    if (n > PLAIN_SPAN && cap > 0) {
        const uint32_t iend = n;
        const uint32_t isafe = n - PLAIN_SPAN;  // plain tokens below this read in bounds without checks
        for (;;) {
            const uint32_t ip0 = ip;
            uint32_t myLit = 0, myLL = 0, myML = 0;
            uint32_t k = 0;
            bool stop = false;  // the exact tier must take over after this batch
            if (lane < 4 && ip0 + 1024 + lane * 128 < iend) prefetch_l2(src + ip0 + 1024 + lane * 128);
#pragma unroll 4
            for (; k < 32; k++) {
                if (ip >= isafe) { stop = true; break; }
                const uint32_t t = __ldg(src + ip);
                uint32_t LL = t >> 4, ML = t & ML_MASK;
                if (LL == RUN_MASK || ML == ML_MASK) {
                    // extended lengths: bounds-checked walk (src/lz4.zig:120-131, :157-168)
                    uint32_t p = ip + 1;
                    if (LL == RUN_MASK && !read_len_ext(src, p, iend, LL, lane)) { stop = true; break; }
                    const uint32_t lit = p;
                    if (LL > FAST_LEN_MAX || (uint64_t)p + LL + 2 > iend) { stop = true; break; }
                    p += LL + 2;
                    if (ML == ML_MASK && !read_len_ext(src, p, iend, ML, lane)) { stop = true; break; }
                    if (ML > FAST_LEN_MAX) { stop = true; break; }
                    if (lane == k) { myLit = lit; myLL = LL; myML = ML + MINMATCH; }
                    ip = p;
                } else {
                    if (lane == k) { myLit = ip + 1; myLL = LL; myML = ML + MINMATCH; }
                    ip += LL + 3;
                }
            }
            if (k == 0) break;
            if (!expand_batch<true>(src, dst, cap, lane, k, myLit, myLL, myML, op)) { ip = ip0; break; }
            if (stop) break;
        }
    }
    // the exact tier finishes the block (or decides the error) from a state where all earlier sequences are complete
    decode_block<true>(src, n, dst, cap, dict, dict_len, has_dict, lane, ip, op, olen, st);
}

// Out-of-line copies of the exact tier for the fast tier's rare exits (values in, values out: a reference parameter
// would pin the caller's ip / op in local memory for the whole hot loop).
__device__ __noinline__ uint4 exact_step_ool(const uint8_t* __restrict__ src, uint32_t n, uint8_t* dst, uint32_t cap,
                                             const uint8_t* __restrict__ dict, uint32_t dict_len, bool has_dict, uint32_t lane,
                                             uint32_t ip, uint32_t op) {
    int st = ST_OK;
    const int r = exact_step<true>(src, n, dst, cap, dict, dict_len, has_dict, lane, ip, op, st);
    return make_uint4(ip, op, (uint32_t)st, (uint32_t)r);
}
__device__ __noinline__ uint2 finish_exact_ool(const uint8_t* __restrict__ src, uint32_t n, uint8_t* dst, uint32_t cap,
                                               const uint8_t* __restrict__ dict, uint32_t dict_len, bool has_dict, uint32_t lane,
                                               uint32_t ip, uint32_t op) {
    uint32_t olen = 0;
    int st = ST_OK;
    decode_block<true>(src, n, dst, cap, dict, dict_len, has_dict, lane, ip, op, olen, st);
    return make_uint2(olen, (uint32_t)st);
}

__device__ __noinline__ uint2 fast_v1_ool(const uint8_t* __restrict__ src, uint32_t n, uint8_t* dst, uint32_t cap,
                                          const uint8_t* __restrict__ dict, uint32_t dict_len, bool has_dict, uint32_t lane,
                                          uint32_t ip, uint32_t op) {
    uint32_t olen = 0;
    int st = ST_OK;
    decode_block_fast_v1(src, n, dst, cap, dict, dict_len, has_dict, lane, ip, op, olen, st);
    return make_uint2(olen, (uint32_t)st);
}

// ---- chunked front end ----
constexpr uint32_t CHUNK = 256;        // stream bytes whose token lengths are computed at once (8 per lane)
constexpr uint32_t CHUNK_MARGIN = 32;  // staged bytes after the chunk: length-extension bytes of its last tokens
constexpr bool WALK_IN_REGS = false;   // the walk's 32 positions in shared memory (true: lane k keeps step k in a register)
constexpr uint32_t DELTA_SLOTS = 544;  // the walk may stand on any position < CHUNK + 2 + 269 + 2 + 2; slots >= CHUNK stay 0
struct __align__(16) WarpStage {
    uint8_t bytes[CHUNK + CHUNK_MARGIN];
    uint16_t delta[DELTA_SLOTS];
    uint16_t pos[40];                  // [32] = where the walk stands after its last step
};

// Token lengths of four positions at once, SPECULATIVELY: the bytes of `x` are candidate tokens, the bytes of `xs` the
// byte after each.  A plain token is 3 + LL bytes long; a token with an LL extension is assumed to have exactly one
// extension byte (19 + that byte), one with an ML extension exactly one more byte.  Whoever turns out to be a real
// token checks its own assumption after the walk (a 255 extension byte ends the batch there).  Returns the lengths
// doubled (= byte offsets into the u16 delta table: one add less per walk step) as two pairs of u16.
__device__ __forceinline__ void token_deltas(uint32_t x, uint32_t xs, uint32_t keep, uint32_t& lo16, uint32_t& hi16) {
    const uint32_t hi = (x >> 4) & 0x0F0F0F0Fu, lo = x & 0x0F0F0F0Fu;
    const uint32_t eh = ((hi + 0x01010101u) >> 4) & 0x01010101u;   // 1 where the LL nibble is 15
    const uint32_t el = ((lo + 0x01010101u) >> 4) & 0x01010101u;   // 1 where the ML nibble is 15
    const uint32_t d8 = (hi + 0x03030303u + el + eh) & keep;       // <= 20 per byte
    const uint32_t xm = xs & (eh * 0xFFu) & keep;                  // the LL extension byte where there is one
    lo16 = (__byte_perm(d8, 0u, 0x4140) + __byte_perm(xm, 0u, 0x4140)) << 1;   // <= 2 * 275: no carry between halves
    hi16 = (__byte_perm(d8, 0u, 0x4342) + __byte_perm(xm, 0u, 0x4342)) << 1;
}

// (no dictionary here: the kernel sends dictionary decodes to the serial front end)
__device__ void decode_block_fast(const uint8_t* __restrict__ src, uint32_t n, uint8_t* dst, uint32_t cap, uint32_t lane,
                                  WarpStage* __restrict__ ws, uint32_t& olen, int& st) {
    uint32_t ip = 0, op = 0;
    st = ST_OK;
    olen = 0;
    if (n == 0 || cap == 0) return;          // :97-98
    if (n > PLAIN_SPAN && n < 0x7F000000u) {   // (chunk positions are kept in int32)
        const uint32_t isafe = n - PLAIN_SPAN;   // a plain token below this reads its offset in bounds and is followed by more stream
        int32_t cb = 0;                          // block position of chunk byte 0 (>= -7)
        bool staged_ok = false;                  // the staged chunk still has >= 56 bytes after ip: walk it again
        while (ip < isafe) {
            if (!staged_ok) {
                // ---------------- stage the next 256 (+32) stream bytes, 8-byte aligned ----------------
                cb = (int32_t)ip - (int32_t)(reinterpret_cast<uintptr_t>(src + ip) & 7);
                const uint2* g = reinterpret_cast<const uint2*>(src + cb);
                uint2 w = make_uint2(0u, 0u), m = w;
                if (cb + 8 * (int32_t)lane < (int32_t)n) w = __ldg(g + lane);   // a word that holds a stream byte is mapped
                if (lane < CHUNK_MARGIN / 8 && cb + (int32_t)CHUNK + 8 * (int32_t)lane < (int32_t)n) m = __ldg(g + 32 + lane);
                reinterpret_cast<uint2*>(ws->bytes)[lane] = w;
                if (lane < CHUNK_MARGIN / 8) reinterpret_cast<uint2*>(ws->bytes)[32 + lane] = m;
                if (lane < 4 && ip + 1024 + lane * 128 < n) prefetch_l2(src + ip + 1024 + lane * 128);
                __syncwarp();
                const uint32_t nx = reinterpret_cast<const uint32_t*>(ws->bytes)[2 * lane + 2];   // the 4 bytes after mine
                // ---------------- distance to the next token from every position ----------------
                uint32_t keep0 = 0xFFFFFFFFu, keep1 = 0xFFFFFFFFu;
                const int32_t lim = (int32_t)isafe - cb;            // positions >= lim are left to the exact tier
                if (lim < (int32_t)CHUNK) {
                    const int32_t kv = lim - 8 * (int32_t)lane;      // valid positions of this lane
                    const uint64_t keep = kv <= 0 ? 0ull : (kv >= 8 ? ~0ull : ((1ull << (8 * kv)) - 1ull));
                    keep0 = (uint32_t)keep; keep1 = (uint32_t)(keep >> 32);
                }
                uint4 dd;
                token_deltas(w.x, __funnelshift_r(w.x, w.y, 8), keep0, dd.x, dd.y);
                token_deltas(w.y, __funnelshift_r(w.y, nx, 8), keep1, dd.z, dd.w);
                reinterpret_cast<uint4*>(ws->delta)[lane] = dd;
                __syncwarp();
            }
            // ---------------- the serial part: up to 32 dependent shared-memory reads, in groups of 8 ----------------
            uint32_t p2 = 2 * (uint32_t)((int32_t)ip - cb);          // twice the chunk position = byte offset into delta[]
            uint32_t steps = 32;
            uint32_t myp2 = 0;                                       // where the walk stood at step `lane` (kept in a register: the
                                                                     // kernel is bound by the load/store pipe, not by the ALUs)
#pragma unroll
            for (int g8 = 0; g8 < 4; g8++) {
                uint32_t d = 0;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    if (WALK_IN_REGS) { if (lane == (uint32_t)(g8 * 8 + k)) myp2 = p2; }
                    else ws->pos[g8 * 8 + k] = (uint16_t)p2;
                    d = *reinterpret_cast<const uint16_t*>(reinterpret_cast<const uint8_t*>(ws->delta) + p2);
                    p2 += d;
                }
                if (g8 < 3 && d == 0) { steps = 8 * (g8 + 1); break; }   // the walk stands still: nothing more in this chunk
            }
            if (!WALK_IN_REGS) { ws->pos[32] = (uint16_t)p2; myp2 = ws->pos[lane]; }   // (lanes >= steps: stale but in range)
            // ---------------- every lane decodes its own token from the staged bytes and checks the speculation ----------------
            const uint32_t mp = myp2 >> 1;                           // (lanes >= steps: position 0, in range, never live)
            const bool live = lane < steps && ws->delta[mp] != 0;    // halting (chunk end / exact-tier zone) is absorbing: a prefix of lanes
            uint32_t myLit = 0, myLL = 0, myML = 0;
            bool bad = false;
            if (live) {
                const uint32_t t = ws->bytes[mp];
                uint32_t q = mp + 1;
                myLL = t >> 4;
                if (myLL == RUN_MASK) { const uint32_t x = ws->bytes[q]; myLL += x; q += 1; bad = x == 255u; }
                myLit = (uint32_t)(cb + (int32_t)q);
                q += myLL + 2;                                       // past literals and offset
                myML = t & ML_MASK;
                if (myML == ML_MASK) {
                    const bool staged = q < CHUNK + CHUNK_MARGIN;
                    const uint32_t y = ws->bytes[staged ? q : 0u];
                    myML += y;
                    q += 1;
                    bad = bad || !staged || y == 255u;
                }
                myML += MINMATCH;
                bad = bad || cb + (int32_t)q > (int32_t)n;           // the whole token lies inside the stream
            }
            const uint32_t lm = __ballot_sync(FULL, live), bm = __ballot_sync(FULL, bad);
            uint32_t k = (uint32_t)__popc(lm);
            const bool cut = bm != 0;                                // a token needs the exact tier: the batch ends before it
            if (cut) k = (uint32_t)__ffs(bm) - 1;
            const uint32_t pk = WALK_IN_REGS ? __shfl_sync(FULL, myp2, k) : (uint32_t)ws->pos[k];   // (k == 32: where the walk stands after its last step)
            const uint32_t ipn = (uint32_t)(cb + (int32_t)((k == 32 ? p2 : pk) >> 1));
            if (lane >= k) { myLL = 0; myML = 0; }
            // the next chunk's lines into L1 while this batch is expanded (its staging then costs an L1 hit, not an L2 round trip)
            if (lane < 3 && ipn + lane * 128 < isafe) prefetch_l1(src + ipn + lane * 128);
            __syncwarp();                                            // staged bytes are dead from here (next chunk overwrites them)
            if (k != 0 && !expand_batch<false>(src, dst, cap, lane, k, myLit, myLL, myML, op)) break;   // ip still at the batch start
            ip = ipn;
            staged_ok = (int32_t)ip - cb <= (int32_t)CHUNK - 56;
            if (cut) {
                // a long, broken or stream-ending token: one exact sequence, warp-wide, with the reference's checks
                const uint4 r = exact_step_ool(src, n, dst, cap, nullptr, 0u, false, lane, ip, op);
                ip = r.x; op = r.y;
                if (r.w == 2) { st = (int)r.z; return; }
                if (r.w == 1) { olen = op; return; }
                staged_ok = staged_ok && (int32_t)ip - cb <= (int32_t)CHUNK - 56;
            }
        }
    }
    // the exact tier finishes the block (or decides the error) from a state where all earlier sequences are complete
    const uint2 r = finish_exact_ool(src, n, dst, cap, nullptr, 0u, false, lane, ip, op);
    olen = r.x;
    st = (int)r.y;
}

#ifndef B2_EMU
template <int MIN_CTAS, int VARIANT>
__global__ void __launch_bounds__(K2_THREADS, MIN_CTAS) k_decompress(BlockSet in, OutSet out, const uint32_t* __restrict__ hdr,
                                                           uint32_t* __restrict__ out_len, int32_t* __restrict__ status,
                                                           uint32_t nblocks, const uint8_t* __restrict__ dict,
                                                           uint32_t dict_len, int has_dict, uint32_t* ticket,
                                                           const uint32_t* __restrict__ order) {
    __shared__ WarpStage stage[VARIANT == 1 ? 1 : K2_WARPS];
    const uint32_t lane = lane_id();
    WarpStage* ws = &stage[VARIANT == 1 ? 0 : (threadIdx.x >> 5)];
    if (VARIANT != 1) {
        for (uint32_t i = CHUNK + lane; i < DELTA_SLOTS; i += 32) ws->delta[i] = 0;   // never written again: the walk halts there
        __syncwarp();
    }
    for (;;) {
        uint32_t blk = 0;
        if (lane == 0) blk = atomicAdd(ticket, 1u);
        blk = __shfl_sync(FULL, blk, 0);
        if (blk >= nblocks) break;
        if (order) blk = order[blk];             // expensive blocks first (k_order_heavy_first): the kernel's tail is cheap blocks
        const uint8_t* src; uint32_t n;
        uint8_t* dst; uint32_t cap;
        in.get(blk, src, n);
        out.get(blk, dst, cap);
        // keep the two block pointers in registers (otherwise every access re-adds base + offset from the constant bank)
        asm volatile("" : "+l"(src));
        asm volatile("" : "+l"(dst));
        __builtin_assume(__isGlobal(src));
        __builtin_assume(__isGlobal(dst));
        uint32_t olen = 0; int st = ST_OK;
        if (hdr && (hdr[blk] & 0x80000000u)) {
            // stored block: reference src/lz4f.zig:603-608
            if (n > cap) st = ST_RAW_NO_ROOM;
            else { warp_copy<true>(dst, src, n, lane); olen = n; }
        } else if (VARIANT == 1) {
            decode_block_fast_v1(src, n, dst, cap, dict, dict_len, (has_dict & 1) != 0, lane, 0, 0, olen, st);
        } else if ((has_dict & 1) || (uint64_t)n * 6 < cap) {
            // Streams that shrink their block more than 6 x are runs and long matches: every few tokens one with a
            // 255-extended length, which the chunked front end's speculation has to cut the batch at.  They go to the
            // EXACT tier, one sequence at a time, every copy warp-wide (runs and repeated patterns without read-back,
            // match_copy_overlap): nearly every sequence of such a stream is long, so batching 32 of them gains nothing,
            // and the exact tier's loop is compact — the serial front end (the round-1 decoder) is as fast on such
            // streams alone (1.48 ms per GiB both) but walks so much code per sequence that it waits 2.6 cycles per
            // issue for instructions and, decoded side by side with other kinds of blocks, takes their instruction cache
            // with it: mixed workload 3.10 -> 2.85 ms per GiB.  Dictionary decodes (small records) keep the serial front
            // end: the chunked tier leaves every match that reaches before the block to the exact tier anyway.
            // (has_dict: bit 0 = dictionary, bit 1 = exact tier for compressible streams; b2lz4_debug_tune("spare2", 1)
            // restores the serial front end for A/B runs.)
            const uint2 r = ((has_dict & 3) == 2) ? finish_exact_ool(src, n, dst, cap, dict, dict_len, (has_dict & 1) != 0, lane, 0, 0)
                                           : fast_v1_ool(src, n, dst, cap, dict, dict_len, (has_dict & 1) != 0, lane, 0, 0);
            olen = r.x;
            st = (int)r.y;
        } else {
            decode_block_fast(src, n, dst, cap, lane, ws, olen, st);
        }
        if (lane == 0) {
            out_len[blk] = st == ST_OK ? olen : 0u;
            status[blk] = st;
        }
        __syncwarp();
    }
}

// Stable partition of the block indices: blocks whose stream is at least 1/6 of the block size first (text, binary:
// ~1.4 ms of a warp each), stored and highly compressible blocks after them.  One block takes a third of the whole
// kernel's duration on the bench workload, so with index order the kernel ends on a tail of expensive blocks.
__global__ void __launch_bounds__(1024) k_order_heavy_first(const uint32_t* __restrict__ hdr, uint32_t nblocks, uint32_t block_size,
                                                            uint32_t* __restrict__ order) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t total_h;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    auto heavy = [&](uint32_t i) { const uint32_t h = hdr[i]; return !(h & 0x80000000u) && (uint64_t)(h & 0x7FFFFFFFu) * 6 >= block_size; };
    uint32_t cnt = 0;
    for (uint32_t i = tid; i < nblocks; i += 1024) cnt += heavy(i) ? 1u : 0u;
    cnt = __reduce_add_sync(FULL, cnt);
    if (lane == 0) wsum[warp] = cnt;
    __syncthreads();
    if (warp == 0) {
        const uint32_t t = __reduce_add_sync(FULL, wsum[lane]);
        if (lane == 0) total_h = t;
    }
    __syncthreads();
    uint32_t run_h = 0, run_l = total_h;
    for (uint32_t t0 = 0; t0 < nblocks; t0 += 1024) {
        const uint32_t i = t0 + tid;
        const bool in = i < nblocks, h = in && heavy(i);
        const uint32_t bm = __ballot_sync(FULL, h);
        __syncthreads();
        if (lane == 0) wsum[warp] = (uint32_t)__popc(bm);
        __syncthreads();
        uint32_t before = 0, tile_h = 0;
        for (uint32_t w = 0; w < 32; w++) { const uint32_t v = wsum[w]; if (w < warp) before += v; tile_h += v; }
        const uint32_t rank_h = before + (uint32_t)__popc(bm & lanemask_lt());
        const uint32_t valid = nblocks - t0 < 1024 ? nblocks - t0 : 1024;
        if (h) order[run_h + rank_h] = i;
        else if (in) order[run_l + (tid - rank_h)] = i;
        run_h += tile_h;
        run_l += valid - tile_h;
    }
}

// Decode order by stored size, largest first, in buckets of 1 KiB (stored blocks last): a counting sort in one CTA.
// Blocks of one kind compress to similar sizes, so the warps resident at any time mostly run the same paths of this
// 9700-instruction kernel (the instruction caches hold 2000), and the expensive blocks still go first.
__global__ void __launch_bounds__(1024) k_order_by_size(const uint32_t* __restrict__ hdr, uint32_t nblocks, uint32_t block_size,
                                                        uint32_t* __restrict__ order) {
    constexpr uint32_t NB = 258;                 // bucket 0: largest ... 256: smallest, 257: stored
    __shared__ uint32_t cnt[NB];
    __shared__ uint32_t base[NB];
    const uint32_t tid = threadIdx.x;
    auto bucket = [&](uint32_t i) -> uint32_t {
        const uint32_t h = hdr[i];
        if (h & 0x80000000u) return 257u;
        const uint32_t kib = (h & 0x7FFFFFFFu) * 256u / (block_size ? block_size : 1u);   // 0..256 relative to the block size
        return 256u - (kib > 256u ? 256u : kib);
    };
    for (uint32_t b = tid; b < NB; b += 1024) cnt[b] = 0;
    __syncthreads();
    for (uint32_t i = tid; i < nblocks; i += 1024) atomicAdd(&cnt[bucket(i)], 1u);
    __syncthreads();
    if (tid == 0) { uint32_t run = 0; for (uint32_t b = 0; b < NB; b++) { base[b] = run; run += cnt[b]; } }
    __syncthreads();
    for (uint32_t i = tid; i < nblocks; i += 1024) order[atomicAdd(&base[bucket(i)], 1u)] = i;
}

// Natural decoded size of every block (unbounded output, nothing written).
__global__ void __launch_bounds__(K2_THREADS) k_decoded_size(BlockSet in, const uint32_t* __restrict__ hdr,
                                                             uint32_t* __restrict__ out_len,
                                                             int32_t* __restrict__ status, uint32_t nblocks) {
    const uint32_t lane = lane_id();
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t blk = warp; blk < nblocks; blk += nwarps) {
        const uint8_t* src; uint32_t n;
        in.get(blk, src, n);
        uint32_t olen = 0; int st = ST_OK;
        if (hdr && (hdr[blk] & 0x80000000u)) olen = n;
        else decode_block<false>(src, n, nullptr, 0xFFFFFFFFu, nullptr, 0, false, lane, 0, 0, olen, st);
        if (lane == 0) { out_len[blk] = olen; status[blk] = st; }
    }
}

cudaError_t launch_decompress(const BlockSet& in, const OutSet& out, const uint32_t* hdr, uint32_t* out_len,
                              int32_t* status, uint32_t nblocks, const uint8_t* dict, uint32_t dict_len,
                              uint32_t* ticket, int num_sms, cudaStream_t stream, uint32_t* order_scratch, uint32_t block_size) {
    if (nblocks == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(ticket, 0, sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
    const uint32_t* order = nullptr;
    if (order_scratch && hdr && block_size && nblocks > (uint32_t)num_sms * 8 && tune().spare[0] == 0) {
        if (tune().spare[1] == 1) k_order_heavy_first<<<1, 1024, 0, stream>>>(hdr, nblocks, block_size, order_scratch);
        else k_order_by_size<<<1, 1024, 0, stream>>>(hdr, nblocks, block_size, order_scratch);
        count_launch();
        order = order_scratch;
    }
    uint32_t want = (nblocks + K2_WARPS - 1) / K2_WARPS;
    // CTAs of 4 warps per SM: 9 (56 registers, 36 warps) by default (8 / 9 / 10 measured 2.63 / 2.56 / 2.67 ms per GiB mixed); b2lz4_debug_tune("k2_occ") picks 10 or 12 for the
    // occupancy experiments of DESIGN.md, ("k2_variant", 1) the round-1 front end (10 per SM).
    const int occ_t = tune().k2_occ, variant = tune().k2_variant == 1 ? 1 : 2;
    const int occ = (occ_t == 6 || occ_t == 7 || occ_t == 8 || occ_t == 9 || occ_t == 10 || occ_t == 12) ? occ_t : (variant == 1 ? 10 : 9);   // the chunked decoder wants 64 registers
    uint32_t maxg = (uint32_t)(num_sms * occ);
    uint32_t grid = want < maxg ? want : maxg;
#define B2_K2_LAUNCH(N, V) k_decompress<N, V><<<grid, K2_THREADS, 0, stream>>>(in, out, hdr, out_len, status, nblocks, dict, \
                                                                              dict_len, (dict != nullptr ? 1 : 0) | (tune().spare[2] ? 0 : 2), ticket, order)
    if (variant == 1) {
        if (occ == 12) B2_K2_LAUNCH(12, 1); else if (occ == 8) B2_K2_LAUNCH(8, 1); else B2_K2_LAUNCH(10, 1);
    } else {
        if (occ == 12) B2_K2_LAUNCH(12, 2); else if (occ == 10) B2_K2_LAUNCH(10, 2); else if (occ == 9) B2_K2_LAUNCH(9, 2);
        else if (occ == 7) B2_K2_LAUNCH(7, 2); else if (occ == 6) B2_K2_LAUNCH(6, 2);
        else B2_K2_LAUNCH(8, 2);
    }
#undef B2_K2_LAUNCH
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_decoded_size(const BlockSet& in, const uint32_t* hdr, uint32_t* out_len, int32_t* status,
                                uint32_t nblocks, int num_sms, cudaStream_t stream) {
    if (nblocks == 0) return cudaSuccess;
    uint32_t want = (nblocks + K2_WARPS - 1) / K2_WARPS;
    uint32_t maxg = (uint32_t)(num_sms * 16);
    uint32_t grid = want < maxg ? want : maxg;
    k_decoded_size<<<grid, K2_THREADS, 0, stream>>>(in, hdr, out_len, status, nblocks);
    count_launch();
    return cudaGetLastError();
}
#endif  // !B2_EMU

}  // namespace b2
