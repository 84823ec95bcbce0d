// k_compress_hc.cu — K3: LZ4HC hash-chain compressor (levels 3..9), many independent blocks, one warp
// per block.
//
// Semantics: byte-identical to the reference's greedy hash-chain path
//   compressHashChain  /root/reference/src/lz4hc.zig:976-1064
//   insertHC :491-510, insertAndGetWiderMatch :538-681 (iLowLimit == ip, longest == 3, no chainSwap),
//   encodeSequence :308-386 (limitedOutput), pattern analysis :626-678 with the F8 guard of the oracle
//   (matchIndex == 0 => no pattern candidate; the reference underflows a u32 there).
//
// Warp-parallel restatement of a sequential algorithm:
//   * insertHC for a run of positions is done 32 positions per step: every lane hashes its position,
//     reads the bucket, and __match_any_sync restores the sequential order inside the step (a lane's
//     predecessor is the nearest earlier lane with the same hash, else the bucket value); the last lane
//     of each hash group writes the bucket.
//   * the chain walk is serial by nature (m -= chain[m]); the warp walks up to 32 hops, parks hop a's
//     candidate in lane a, then all lanes measure their candidate's match length at once.  The
//     reference's order-dependent rules are recovered with ballots: a candidate replaces the best only
//     if strictly longer (first lane holding the maximum wins), and the early exit `len > nbSearches`
//     (:613) cuts the chunk at the first such lane, which also defines the final matchIndex that the
//     pattern analysis reads.
//   * countPattern / reverseCountPattern / the final length of the chosen match are warp-wide
//     compares (128 bytes per round).
// State: hashTable u32[32768] + chainTable u16[65536] per resident warp, in global memory (L2
// resident).  The reference zero-fills 256 KiB per call (:1450); here bucket values carry a per-warp
// epoch base (value = base + index; anything below the base reads as 0 = empty), so nothing is
// re-zeroed between blocks, and the chain table never needs zeroing (only inserted slots are read).
#include "b2_common.cuh"
#ifndef B2_EMU   // (B2_EMU: host build of the device code for the one-warp emulator, tools/warp_emu)
#include "b2_kernels.h"
#endif

namespace b2 {

constexpr int HC_WARPS = 4;                 // warps per CTA
// K3 is bound by the latency of its dependent table reads, so blocks per second grow with the warps in flight
// (5.7k / 7.9k / 9.7k blocks/s at 32 / 48 / 64 warps per SM on 4 GiB of text).  64 warps per SM leave 32 registers per
// thread (a few spills, each block ~5 % slower), which only pays when there are more blocks than the 32-warp variant
// holds at once: the launcher picks.
constexpr int HC_CTAS_PER_SM = 16;          // workspace is sized for the 64-warp variant
constexpr int HC_CTAS_PER_SM_FEW = 8;       // 32 warps per SM, 56 registers, no spills
constexpr uint32_t HC_HASH = 32768, HC_CHAIN = 65536;


// Jump variant (round 2): next to the reference's chainTable (jump[0]: distance to the previous position of the same
// bucket) the distances to the 2nd, 4th, 8th and 16th predecessor, filled in when a position is inserted (each level is
// one read of the previous level at the predecessor).  A search then reaches its 32 next candidates with at most five
// dependent reads per lane (lane j follows the bits of j) instead of 32 dependent reads of one chain: the walk is what
// K3 spends its time on (every hop a DRAM round trip, issue slots 6 % busy).  Same candidates, same order, same bytes.
constexpr int HC_LEVELS = 5;
// ONE layout for every kernel variant: the epoch base and the bucket values of a work area must mean the same thing to
// whichever variant runs next on it (a variant that kept `base` elsewhere would read table bytes as its epoch).
struct HcWork {
    uint32_t hash[HC_HASH];
    uint16_t jump[HC_LEVELS][HC_CHAIN];   // jump[0] is the reference's chainTable
    uint32_t base;                        // epoch base of the bucket values (persists across launches)
    uint32_t pad[15];
};
#ifndef B2_EMU
static bool hc_jump() { return tune().k3_variant != 16; }     // 16 = the round-1 single-chain walk, for A/B runs

// one work area per warp that a launch over `nblocks` blocks can have in flight — not per resident warp of the device:
// a single small block needs a few MiB, not gigabytes
size_t hc_work_bytes(int num_sms, uint32_t nblocks) {
    const size_t ctas = (nblocks + HC_WARPS - 1) / HC_WARPS, max_ctas = (size_t)num_sms * HC_CTAS_PER_SM;
    return (ctas < max_ctas ? ctas : max_ctas) * HC_WARPS * sizeof(HcWork);
}
#endif

__device__ __forceinline__ uint32_t hashHC(uint32_t v) { return (v * HASH_MULTIPLIER) >> 17; }  // :129-131

__device__ __forceinline__ void write_len_ext_hc(uint8_t* p, uint32_t L, uint32_t cnt, uint32_t lane) {
    for (uint32_t i = lane; i < cnt; i += 32) p[i] = (i + 1 == cnt) ? (uint8_t)((L - 15u) % 255u) : (uint8_t)255;
}

// number of bytes equal to `b` in src[from, limit) counted forward from `from` (countPattern :170-199
// for a one-byte-repeat pattern)
__device__ uint32_t warp_count_byte_fwd(const uint8_t* __restrict__ src, uint32_t from, uint32_t limit, uint32_t pat32,
                                        uint32_t lane) {
    uint32_t total = 0;
    for (;;) {
        uint32_t a = from + 4 * lane;
        uint32_t nb = a >= limit ? 0u : (limit - a >= 4 ? 4u : limit - a);
        uint32_t cnt = 0;
        if (nb) {
            uint32_t x = ldg_u32(src + a) ^ pat32;
            uint32_t mm = x ? (uint32_t)(__ffs(x) - 1) >> 3 : 4u;
            cnt = mm < nb ? mm : nb;
        }
        uint32_t stopm = __ballot_sync(FULL, cnt < 4);
        if (stopm) {
            uint32_t f = (uint32_t)__ffs(stopm) - 1;
            return total + 4 * f + __shfl_sync(FULL, cnt, f);
        }
        total += 128; from += 128;
    }
}

// number of bytes equal to `b` immediately before `from`, not going below position 0
// (reverseCountPattern :202-222 with iLow = block start)
__device__ uint32_t warp_count_byte_bwd(const uint8_t* __restrict__ src, uint32_t from, uint32_t b, uint32_t lane) {
    uint32_t total = 0;
    for (;;) {
        // lane l looks at byte from-1-l
        bool inr = from > lane;
        bool eq = inr && (uint32_t)__ldg(src + (from - 1 - lane)) == b;
        uint32_t stopm = __ballot_sync(FULL, !eq);
        if (stopm) return total + ((uint32_t)__ffs(stopm) - 1);
        total += 32; from -= 32;
    }
}

// common prefix of src[a..] and src[m..], a limited to < limit (lz4Count :234-264), one lane
__device__ __forceinline__ uint32_t lane_count(const uint8_t* __restrict__ src, uint32_t a, uint32_t m, uint32_t limit) {
    uint32_t c = 0;
    while (a + 4 <= limit) {
        uint32_t x = ldg_u32(src + a) ^ ldg_u32(src + m);
        if (x) return c + ((uint32_t)(__ffs(x) - 1) >> 3);
        a += 4; m += 4; c += 4;
    }
    while (a < limit && __ldg(src + a) == __ldg(src + m)) { a++; m++; c++; }
    return c;
}

// One candidate of the chain, one lane: 4-byte check (:586) + lz4Count (:587-591), with ONE 16-byte read at the candidate
// serving the check and the first 5..12 bytes of the count (the candidate is a random position of the block: every read
// of it is a DRAM or L2 round trip, and the 4-bytes-at-a-time count made 3-5 of them per candidate).  f1..f3: the
// 12 bytes after ip + 4 (the same for every candidate of a search).  Returns 0 or the match length.
__device__ __forceinline__ uint32_t eval_candidate(const uint8_t* __restrict__ src, uint32_t ip, uint32_t cand, uint32_t pattern,
                                                   uint32_t f1, uint32_t f2, uint32_t f3, uint32_t mlimit) {
    const uintptr_t ca = reinterpret_cast<uintptr_t>(src + cand);
    const uint2* c8 = reinterpret_cast<const uint2*>(ca & ~uintptr_t(7));
    const uint2 A = __ldg(c8), B = __ldg(c8 + 1);     // the second word holds src[cand + 8 - c]: inside the block (cand <= n - 13)
    const uint32_t c = (uint32_t)(ca & 7), sh = (c & 3) * 8;
    const bool hiw = c >= 4;
    const uint32_t w0 = hiw ? A.y : A.x, w1 = hiw ? B.x : A.y, w2 = hiw ? B.y : B.x, w3 = hiw ? 0u : B.y;
    if (__funnelshift_r(w0, w1, sh) != pattern) return 0;
    const uint32_t m1 = __funnelshift_r(w1, w2, sh), m2 = __funnelshift_r(w2, w3, sh), m3 = __funnelshift_r(w3, 0u, sh);
    const uint32_t room = mlimit - (ip + MINMATCH);                   // bytes the count may still take on the ip side (>= 3)
    uint32_t avail = 12 - c;                                          // candidate bytes staged after the first four
    if (avail > room) avail = room;
    const uint32_t x1 = f1 ^ m1, x2 = f2 ^ m2, x3 = f3 ^ m3;
    const uint32_t c1 = (uint32_t)__clz(__brev(x1)) >> 3, c2 = (uint32_t)__clz(__brev(x2)) >> 3, c3 = (uint32_t)__clz(__brev(x3)) >> 3;
    const uint32_t d = c1 + (c1 == 4 ? c2 + (c2 == 4 ? c3 : 0u) : 0u);
    if (d < avail) return MINMATCH + d;
    if (avail == room) return MINMATCH + room;
    return MINMATCH + avail + lane_count(src, ip + MINMATCH + avail, cand + MINMATCH + avail, mlimit);
}

// JUMP: the block may switch to jump mode (jump tables maintained and used) once its chains prove long: the tables cost
// four more dependent reads per inserted batch, which short chains (binary records, low levels) never earn back.
template <bool JUMP>
__device__ void compress_block_hc(const uint8_t* __restrict__ src, uint32_t n, uint8_t* __restrict__ dst, uint32_t cap,
                                  void* wv, int nbs, uint32_t jump_after, uint32_t lane, uint32_t& olen, int& st) {
    HcWork* w = reinterpret_cast<HcWork*>(wv);
    HcWork* wj = w;
    st = ST_OK;
    olen = 0;
    if (n == 0) return;                                                  // :1443
    if (n > LZ4_MAX_INPUT_SIZE) { st = ST_INPUT_TOO_LARGE; return; }     // :1442
    if (cap == 0) { st = ST_OUTPUT_TOO_SMALL; return; }                  // :1461
    uint32_t op = 0, anchor = 0;
    if (n < MFLIMIT + 1) {                                               // :995 encodeLiterals :1394-1425
        if (cap < n + 1 + n / 255) { st = ST_OUTPUT_TOO_SMALL; return; }
        if (lane == 0) dst[0] = (uint8_t)(n << 4);
        warp_copy<true>(dst + 1, src, n, lane);
        olen = n + 1;
        return;
    }
    // epoch base for this block's bucket values
    uint32_t* basep = &w->base;
    uint32_t base = *basep;
    if (base > 0xFFFFFFFFu - n - 16u) {   // would wrap: re-zero once
        uint4 z = make_uint4(0, 0, 0, 0);
        uint4* t4 = reinterpret_cast<uint4*>(w->hash);
        for (uint32_t i = lane; i < HC_HASH * 4 / 16; i += 32) t4[i] = z;
        base = 0;
    }
    __syncwarp();
    if (lane == 0) *basep = base + n;
    uint32_t* H = w->hash;
    uint16_t* C = w->jump[0];
    const bool patternAnalysis = nbs > 128;                              // :983
    const uint32_t mflimit = n - MFLIMIT, mlimit = n - LASTLITERALS;
    const uint32_t lt = lanemask_lt(), gt = lanemask_gt();
    uint32_t ip = 0, ntu = 0;
    bool jm = false;                       // jump mode on
    uint32_t jv = 0;                       // jump mode: positions below jv have their jump levels (built lazily, in bulk)
    uint32_t nsearch = 0, nhops = 0, dsum = 0;   // chain statistics of the current window of searches: count, hops, hop distance / 256

    while (ip <= mflimit) {                                              // :1009
        // ---------------- insertHC(ctx, ip): positions [ntu, ip), :491-510 ----------------
        while (ntu < ip) {
            uint32_t idx = ntu + lane;
            bool act = idx < ip;
            uint32_t h = 0x80000000u | lane, prev = 0;
            if (act) {
                h = hashHC(ldg_u32(src + idx));
                uint32_t v = H[h];
                prev = v >= base ? v - base : 0;
            }
            uint32_t peers = __match_any_sync(FULL, h);
            uint32_t pm = peers & lt;
            if (pm) prev = idx - (lane - (31 - __clz(pm)));
            uint32_t delta = idx - prev;
            if (delta > MAX_DISTANCE) delta = MAX_DISTANCE;
            __syncwarp();
            if (act) {
                C[idx & (HC_CHAIN - 1)] = (uint16_t)delta;
                if ((peers & gt) == 0) H[h] = base + idx;
            }
            __syncwarp();
            ntu += 32;
        }
        ntu = ip;

        // ---------------- insertAndGetWiderMatch, :538-681 ----------------
        const uint32_t pattern = ldg_u32(src + ip);
        // the 12 bytes after ip + 4 (ip + 16 <= n + 4: the last word may reach past the block by up to 4 bytes only when
        // ip is within 4 of n - 12; those bytes are never counted — eval_candidate clips at mlimit — but must be readable)
        const uint32_t f1 = ldg_u32(src + ip + 4), f2 = ip + 12 <= n ? ldg_u32(src + ip + 8) : 0u, f3 = ip + 16 <= n ? ldg_u32(src + ip + 12) : 0u;
        uint32_t best_len = MINMATCH - 1, best_off = 0;
        uint32_t m;
        { uint32_t v = H[hashHC(pattern)]; m = v >= base ? v - base : 0; }   // :563
        if (m != 0) {                                                        // :566
            uint32_t final_m = m;
            if (JUMP && jm) {
                if (m >= jv) {
                    // The chain starts in the tail that has no jump levels yet: build them for every position a search can
                    // still reach, oldest first — level k = level k-1 here + level k-1 at that predecessor (an earlier
                    // position, complete; or a lane of this batch, written one step ago); 65535 = "out of reach", as in
                    // the reference's clamp (:505).  Predecessors below `lo` are out of every later search's reach,
                    // whatever their stale entries say, and so is everything reached through them.  Done lazily and in
                    // bulk because the four dependent reads per batch are only worth paying for positions a chain
                    // actually enters (incompressible stretches insert one position per search and never do).
                    const uint32_t lo = ip > MAX_DISTANCE ? ip - MAX_DISTANCE : 0;
                    for (uint32_t b0 = jv > lo ? jv : lo; b0 < ip; b0 += 32) {
                        const uint32_t idx = b0 + lane;
                        const bool act = idx < ip;
                        uint32_t dk = act ? (uint32_t)C[idx & (HC_CHAIN - 1)] : MAX_DISTANCE;
#pragma unroll
                        for (int k = 1; k < HC_LEVELS; k++) {
                            uint32_t nd = MAX_DISTANCE;
                            if (act) {
                                if (dk < MAX_DISTANCE && dk <= idx) {
                                    const uint32_t sum = dk + wj->jump[k - 1][(idx - dk) & (HC_CHAIN - 1)];
                                    nd = sum > MAX_DISTANCE ? MAX_DISTANCE : sum;
                                }
                                wj->jump[k][idx & (HC_CHAIN - 1)] = (uint16_t)nd;
                            }
                            __syncwarp();
                            dk = nd;
                        }
                    }
                    jv = ip;
                }
                // chain positions P_0 = m, P_{i+1} = P_i - chain[P_i] (:619-621), 32 per round: lane j jumps from the
                // round's first position by the bits of j.  A position is visited while it is > 0 (:571), within
                // 65535 of ip (:573) and attempts are left; positions only decrease, so the visited lanes are a prefix.
                uint32_t basepos = m;
                for (uint32_t round = 0;; round++) {
                    uint32_t pos = basepos;
                    bool ok = true;
#pragma unroll
                    for (int k = HC_LEVELS - 1; k >= 0; k--) {
                        if ((lane >> k) & 1) {
                            const uint32_t d = ok ? (uint32_t)wj->jump[k][pos & (HC_CHAIN - 1)] : 0u;
                            if (d == 0 || d > pos) ok = false; else pos -= d;
                        }
                    }
                    const bool visit = ok && pos > 0 && pos <= ip && ip - pos <= MAX_DISTANCE && round * 32 + lane < (uint32_t)nbs;
                    const uint32_t vm = __ballot_sync(FULL, visit);
                    uint32_t cnt = vm == FULL ? 32u : (uint32_t)__ffs(~vm) - 1;
                    // the hop after each visited position: next round's start, or where the walk stops
                    uint32_t nxt = pos;
                    bool stuck = false;                                                          // :620 delta == 0 or delta > matchIndex
                    if (lane < cnt) {
                        const uint32_t d0 = C[pos & (HC_CHAIN - 1)];
                        stuck = d0 == 0 || d0 > pos;
                        if (!stuck) nxt = pos - d0;
                    }
                    const uint32_t sm = __ballot_sync(FULL, stuck);
                    if (sm) cnt = (uint32_t)__ffs(sm);                                           // that position is still a candidate
                    if (cnt) {                                                                   // window statistics
                        nhops += cnt;
                        dsum += (basepos - __shfl_sync(FULL, pos, cnt - 1)) >> 8;
                    }
                    const uint32_t my_cand = pos;
                    uint32_t len = 0;
                    if (lane < cnt) len = eval_candidate(src, ip, my_cand, pattern, f1, f2, f3, mlimit);   // :586-591
                    const uint32_t xm = __ballot_sync(FULL, len > (uint32_t)nbs);               // :613
                    const uint32_t X = xm ? (uint32_t)__ffs(xm) - 1 : 32;
                    const uint32_t l = (lane < cnt && lane <= X) ? len : 0;
                    const uint32_t mx = __reduce_max_sync(FULL, l);
                    if (mx > best_len) {                                                         // :607
                        const uint32_t who = (uint32_t)__ffs(__ballot_sync(FULL, l == mx)) - 1;
                        best_len = mx;
                        best_off = ip - __shfl_sync(FULL, my_cand, who);
                    }
                    if (xm) { final_m = __shfl_sync(FULL, my_cand, X); break; }
                    if (sm) { final_m = __shfl_sync(FULL, my_cand, cnt - 1); break; }
                    if (cnt < 32 || (round + 1) * 32 >= (uint32_t)nbs) {
                        final_m = cnt == 0 ? basepos : __shfl_sync(FULL, nxt, cnt - 1);
                        break;
                    }
                    basepos = __shfl_sync(FULL, nxt, 31);
                }
            } else {
            int attempts = nbs;
            bool done = false;
            while (!done) {
                uint32_t my_cand = 0, cnt = 0;
                for (uint32_t a = 0; a < 32; a++) {
                    if (!(m > 0 && attempts > 0)) { done = true; final_m = m; break; }           // :571
                    if (m > ip || ip - m > MAX_DISTANCE) { done = true; final_m = m; break; }    // :573
                    attempts--;
                    if (lane == a) my_cand = m;
                    cnt = a + 1;
                    uint32_t delta = C[m & (HC_CHAIN - 1)];                                      // :619
                    if (delta == 0 || delta > m) { done = true; final_m = m; break; }
                    m -= delta;
                    if (JUMP) dsum += delta >> 8;
                }
                uint32_t len = 0;
                if (lane < cnt) len = eval_candidate(src, ip, my_cand, pattern, f1, f2, f3, mlimit);           // :586-591
                uint32_t xm = __ballot_sync(FULL, len > (uint32_t)nbs);                          // :613
                uint32_t X = xm ? (uint32_t)__ffs(xm) - 1 : 32;
                uint32_t l = (lane < cnt && lane <= X) ? len : 0;
                uint32_t mx = __reduce_max_sync(FULL, l);
                if (mx > best_len) {                                                             // :607
                    uint32_t who = (uint32_t)__ffs(__ballot_sync(FULL, l == mx)) - 1;
                    best_len = mx;
                    best_off = ip - __shfl_sync(FULL, my_cand, who);
                }
                if (xm) { done = true; final_m = __shfl_sync(FULL, my_cand, X); }
            }
            if (JUMP) nhops += (uint32_t)(nbs - attempts);
            }
            if (JUMP && ++nsearch == 128) {
                // Every 128 searches: jump mode pays for long chains whose hops land far apart (each one a DRAM round
                // trip): text.  Chains of near neighbours (records with a repeated field: hops of a few dozen positions,
                // served by L1/L2), short chains and incompressible stretches are faster walked as they are.  Positions
                // inserted while the mode is off simply have no levels yet (jv stays where it is).
                jm = nhops > jump_after * nsearch && dsum > 2 * nhops;
                nsearch = nhops = dsum = 0;
            }
            if (patternAnalysis) {                                                               // :626
                uint32_t delta = C[final_m & (HC_CHAIN - 1)];
                if (delta == 1 && ((pattern & 0xFFFF) == (pattern >> 16)) && ((pattern & 0xFF) == (pattern >> 24))) {
                    uint32_t srcPat = warp_count_byte_fwd(src, ip + 4, mlimit, pattern, lane) + 4;  // :633
                    if (final_m != 0) {                                                          // F8 guard
                        uint32_t cand = final_m - 1;                                             // :636
                        uint32_t lowest = ip < 65536 ? 0 : ip - MAX_DISTANCE;                    // :553-554
                        if (cand >= lowest && ldg_u32(src + cand) == pattern) {                  // :637,:644
                            uint32_t fwd = warp_count_byte_fwd(src, cand + 4, mlimit, pattern, lane) + 4;
                            uint32_t back = warp_count_byte_bwd(src, cand, pattern & 0xFF, lane);
                            uint32_t a = cand - back;
                            uint32_t limq = a > lowest ? a : lowest;
                            uint32_t limitedBack = cand - limq;                                  // :653
                            uint32_t seg = limitedBack + fwd;
                            uint32_t maxML = seg < srcPat ? seg : srcPat;
                            uint32_t newIdx = (seg >= srcPat && fwd <= srcPat) ? cand + fwd - srcPat : cand - limitedBack;
                            if (maxML > best_len && ip - newIdx <= MAX_DISTANCE) {               // :669
                                best_len = maxML;
                                best_off = ip - newIdx;
                            }
                        }
                    }
                }
            }
        }
        if (best_len < MINMATCH || best_off == 0) { ip += 1; continue; }                         // :1013

        // ---------------- encodeSequence (limitedOutput), :308-386 ----------------
        const uint32_t LL = ip - anchor;
        if ((uint64_t)op + LL / 255 + LL + 8 > cap) { st = ST_OUTPUT_TOO_SMALL; return; }        // :320-325
        const uint32_t nll = LL >= RUN_MASK ? (LL - RUN_MASK) / 255 + 1 : 0;
        const uint32_t mlc = best_len - MINMATCH;
        const uint32_t nml = mlc >= ML_MASK ? (mlc - ML_MASK) / 255 + 1 : 0;
        uint8_t* o = dst + op;
        uint8_t* o2 = o + 1 + nll + LL;
        if ((uint64_t)(op + 1 + nll + LL + 2) + mlc / 255 + 6 > cap) { st = ST_OUTPUT_TOO_SMALL; return; }  // :355-359
        if (lane == 0) {
            o[0] = (uint8_t)(((LL < 15 ? LL : 15u) << 4) | (mlc < 15 ? mlc : 15u));
            o2[0] = (uint8_t)(best_off & 0xFF);
            o2[1] = (uint8_t)(best_off >> 8);
        }
        write_len_ext_hc(o + 1, LL, nll, lane);
        warp_copy<true>(o + 1 + nll, src + anchor, LL, lane);
        write_len_ext_hc(o2 + 2, mlc, nml, lane);
        op += 1 + nll + LL + 2 + nml;
        ip += best_len;                                                                          // :382
        anchor = ip;
    }

    // ---------------- last literals, :1035-1061 ----------------
    const uint32_t LL = n - anchor;
    if (LL > 0) {
        const uint32_t nll = LL >= RUN_MASK ? (LL - RUN_MASK) / 255 + 1 : 0;
        if ((uint64_t)op + 1 + nll + LL > cap) { st = ST_OUTPUT_TOO_SMALL; return; }  // :1037 (+ no-overflow, as the oracle)
        uint8_t* o = dst + op;
        if (lane == 0) o[0] = (uint8_t)((LL < 15 ? LL : 15u) << 4);
        write_len_ext_hc(o + 1, LL, nll, lane);
        warp_copy<true>(o + 1 + nll, src + anchor, LL, lane);
        op += 1 + nll + LL;
    }
    olen = op;
}

#ifndef B2_EMU
template <int CTAS, bool JUMP>
__global__ void __launch_bounds__(HC_WARPS * 32, CTAS) k_compress_hc(BlockSet in, OutSet out, uint32_t* __restrict__ out_len,
                                                               int32_t* __restrict__ status, uint32_t nblocks, int nbs,
                                                               uint8_t* work, uint32_t* ticket, uint32_t jump_after) {
    const uint32_t lane = lane_id();
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    void* w = work + (size_t)gw * sizeof(HcWork);
    for (;;) {
        uint32_t blk = 0;
        if (lane == 0) blk = atomicAdd(ticket, 1u);
        blk = __shfl_sync(FULL, blk, 0);
        if (blk >= nblocks) break;
        const uint8_t* src; uint32_t n;
        uint8_t* dst; uint32_t cap;
        in.get(blk, src, n);
        out.get(blk, dst, cap);
        uint32_t olen; int st;
        compress_block_hc<JUMP>(src, n, dst, cap, w, nbs, jump_after, lane, olen, st);
        if (lane == 0) { out_len[blk] = st == ST_OK ? olen : 0u; status[blk] = st; }
        __syncwarp();
    }
}

cudaError_t launch_compress_hc(const BlockSet& in, const OutSet& out, uint32_t* out_len, int32_t* status, uint32_t nblocks,
                               int nb_searches, uint8_t* work, uint32_t* ticket, int num_sms, cudaStream_t stream) {
    if (nblocks == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(ticket, 0, sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
    const uint32_t want = (nblocks + HC_WARPS - 1) / HC_WARPS;
    const uint32_t few = (uint32_t)(num_sms * HC_CTAS_PER_SM_FEW);
    const int cap_ctas = tune().k3_variant;                 // experiment: 1..8 = at most this many CTAs (of 4 warps) per SM
    const bool jump = hc_jump() && nb_searches > 32;         // chains of at most 32 hops never earn the tables back
    // a block switches to jump mode when its searches average more than this many chain hops (spare5 overrides: experiments)
    const uint32_t jump_after = tune().spare[5] > 0 ? (uint32_t)tune().spare[5] : 24u;
#define B2_K3(C, J, G) k_compress_hc<C, J><<<(G), HC_WARPS * 32, 0, stream>>>(in, out, out_len, status, nblocks, nb_searches, work, ticket, jump_after)
    if (cap_ctas >= 1 && cap_ctas <= HC_CTAS_PER_SM_FEW) {
        const uint32_t g = (uint32_t)(num_sms * cap_ctas);
        B2_K3(HC_CTAS_PER_SM_FEW, true, want < g ? want : g);
    } else if (want <= few) {
        if (jump) B2_K3(HC_CTAS_PER_SM_FEW, true, want); else B2_K3(HC_CTAS_PER_SM_FEW, false, want);
    } else {
        const uint32_t maxg = (uint32_t)(num_sms * HC_CTAS_PER_SM);
        if (jump) B2_K3(HC_CTAS_PER_SM, true, want < maxg ? want : maxg); else B2_K3(HC_CTAS_PER_SM, false, want < maxg ? want : maxg);
    }
#undef B2_K3
    count_launch();
    return cudaGetLastError();
}
#endif  // !B2_EMU

}  // namespace b2
