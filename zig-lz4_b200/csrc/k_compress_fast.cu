// k_compress_fast.cu — K1: LZ4 fast block compressor, many independent blocks, one warp per block.
//
// Semantics: byte-identical to the reference's greedy single-probe matcher,
//   lz4.compressFast / compressDefault, /root/reference/src/lz4.zig:292-447
//   (+ compressAsLiterals :449-482, finishCompression :484-519).
//
// How a sequential hash-table algorithm is made warp-parallel without changing a single output byte:
// the search loop of the reference visits a *predictable* sequence of positions from a search start q
// (iteration k: ip_k, step s_k — src/lz4.zig:329-333) and stops at the first position whose table
// candidate is a valid match.  A warp evaluates 32 consecutive iterations at once:
//   * every lane computes its iteration's (ip_k, s_k) in closed form, applies the early-exit test
//     `forwardIp > mflimitPlusOne` (:335) *before* probing, hashes its 4 bytes and reads the table;
//   * intra-window ordering is restored with __match_any_sync on the hash: a lane's candidate is the
//     position of the nearest earlier lane with the same hash, else the table value (that is exactly
//     what the table would hold after the earlier lanes' put()s — :350);
//   * __ballot_sync finds the first lane with a valid match (:345-348); only lanes up to and
//     including it commit their put()s, later-lane-wins per bucket;
//   * the 63 repeated probes of ip0+1 that the reference's step schedule performs (step == 0 while
//     searchMatchNb < 64) are provable no-ops on table and output and are skipped; their
//     exit tests are implied by the next distinct iteration's test.
// Match extension (:401-413) compares 4 bytes per lane per round and finds the end with ballot/ffs.
// Literal runs are copied with 16-byte aligned stores (b2::warp_copy).
// That is the general path (any acceleration).  Acceleration 1 — compressDefault, every frame block — takes
// compress_block_a1 below: windows of 32 consecutive positions that yield every sequence starting inside them (static
// match chain + general step), one batched write per window, the input words of the window in a register ring.
//
// Memory: the 4096-entry hash table lives in shared memory: u16 entries when every block is <= 64 KiB (8 KiB/warp, 28
// warps per SM), u32 otherwise (16 KiB/warp, 14 per SM).  Input: the window's words come from the register ring (a1 path)
// or straight from global memory (general path); candidate bytes are read from global memory through L1/L2.  Blocks are
// handed to warps through a global ticket; large launches take them expensive-first (launch_compress_fast_ordered).
#include "b2_common.cuh"
#ifndef B2_EMU
#include "b2_kernels.h"
#include <mutex>
#endif

namespace b2 {

constexpr int HASH_ENTRIES = 4096;  // LZ4_HASH_SIZE_U32, src/lz4.zig:33

#ifdef B2_EMU   // counters of the host emulation (tools/warp_emu/emu_k1.cpp); nothing on the device
struct EmuStats { uint64_t windows, fast_windows, bailed_windows, sequences, batched_sequences; };
extern EmuStats g_emu_stats;
#define EMU_COUNT(field, n) do { if (lane == 0) g_emu_stats.field += (n); } while (0)
#else
#define EMU_COUNT(field, n) do {} while (0)
#endif

__device__ __forceinline__ uint32_t hash4(uint32_t v) { return (v * HASH_MULTIPLIER) >> 20; }  // :75-77

// Unaligned little-endian u32 from the (read-only) input: always two aligned words + funnel shift.  The
// second word holds a valid input byte whenever p + 4 <= n - 1, which every caller guarantees
// (search positions are <= n - 12, extension reads stop before n - 5).
__device__ __forceinline__ uint32_t ld_u32x(const uint8_t* __restrict__ p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
    return __funnelshift_r(__ldg(w), __ldg(w + 1), (uint32_t)(a & 3) * 8);
}
#ifdef B2_EMU
__device__ __forceinline__ void prefetch_l1(const void*) {}
__device__ __forceinline__ void prefetch_l2(const void*) {}
#else
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#endif
// F(x) = sum_{u<x} (u >> 6): cumulative step of the reference's `step = searchMatchNb >> 6` schedule.
__device__ __forceinline__ uint32_t step_prefix(uint32_t x) {
    uint32_t A = x >> 6, B = x & 63;
    return 32u * A * (A - 1u) + B * A;
}

// ------------------------------------------------------------------------------------------------
// Forward input ring (K1 "ring" variant): the bytes the search windows read next are staged in shared memory by
// TMA bulk copies (cp.async.bulk, one 128-byte line per copy, completion on an mbarrier) issued a few lines ahead,
// so a window's first dependent access is a shared-memory read instead of an L2 round trip (the u16/u32 hash tables
// leave the SM no L1 to speak of).  Four line slots per warp; a slot is re-armed only after its previous copy has
// been waited for, and everything in flight is drained before the warp leaves the kernel.
constexpr uint32_t RING_LINE = 128, RING_SLOTS = 4, RING_BYTES = RING_LINE * RING_SLOTS;

#ifdef B2_EMU   // the ring variant is not emulated (RING = false there): empty stand-ins for its PTX
__device__ __forceinline__ uint32_t smem_addr(const void*) { return 0; }
__device__ __forceinline__ void mbar_init(uint32_t, uint32_t) {}
__device__ __forceinline__ void mbar_expect_tx(uint32_t, uint32_t) {}
__device__ __forceinline__ void bulk_load(uint32_t, const void*, uint32_t, uint32_t) {}
__device__ __forceinline__ void mbar_wait(uint32_t, uint32_t) {}
__device__ __forceinline__ void fence_proxy_async() {}
__device__ __forceinline__ uint32_t lds_u32(uint32_t) { return 0; }
#else
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred P1;\n\tLAB_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
                 "@P1 bra DONE;\n\tbra LAB_WAIT;\n\tDONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
    uint32_t w;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(a));
    return w;
}
#endif

struct Ring {
    uint32_t data;      // shared-memory address of this warp's RING_BYTES
    uint32_t bars;      // shared-memory address of its RING_SLOTS mbarriers
    uint32_t next;      // next line to fetch, counted from `line0`; lines [next - 4, next) are resident or in flight
    uint32_t pending;   // bit s: slot s has a copy in flight that nobody has waited for yet
    uint32_t parity;    // bit s: phase parity the next wait on slot s must use
    uint64_t line0;     // global address of the 128-byte line that holds the block's first byte
    uint64_t lo16, hi16; // the block's bytes rounded out to 16-byte granules: nothing outside is ever read

    __device__ __forceinline__ void begin_block(uint64_t ga, uint32_t n) {
        line0 = ga & ~uint64_t(RING_LINE - 1);
        lo16 = ga & ~uint64_t(15);
        hi16 = (ga + n + 15) & ~uint64_t(15);
        next = 0;                                                    // nothing of this block is resident
    }
    // slot of line L: its global address picks it, so that a byte's place in the ring is (address & 511)
    __device__ __forceinline__ uint32_t slot_of(uint32_t L) const { return ((uint32_t)(line0 >> 7) + L) & (RING_SLOTS - 1); }
    __device__ __forceinline__ void wait_slot(uint32_t s) {
        if (pending & (1u << s)) {
            mbar_wait(bars + 8 * s, (parity >> s) & 1u);
            parity ^= 1u << s;
            pending &= ~(1u << s);
        }
    }
    // all lanes call; makes the lines holding global addresses [a, a + 56) readable and keeps two more coming
    __device__ __forceinline__ void ensure(uint64_t a, uint32_t lane) {
        const uint32_t Lb = (uint32_t)((a - line0) >> 7), Le = (uint32_t)((a + 55 - line0) >> 7);
        if (Lb >= next || next - Lb > RING_SLOTS) next = Lb;          // block start, or the search jumped past the ring
        const uint32_t want = Lb + RING_SLOTS;                       // never overwrite line Lb
        while (next < want && next <= Le + 2) {
            const uint32_t s = slot_of(next);
            wait_slot(s);                                            // a copy skipped by a jump: retire it first
            uint64_t g0 = line0 + ((uint64_t)next << 7), g1 = g0 + RING_LINE;
            if (g0 < lo16) g0 = lo16;
            if (g1 > hi16) g1 = hi16;
            if (g1 > g0) {
                __syncwarp();                                        // every lane is done reading the line this replaces
                if (lane == 0) {
                    fence_proxy_async();
                    mbar_expect_tx(bars + 8 * s, (uint32_t)(g1 - g0));
                    bulk_load(data + (uint32_t)(g0 & (RING_BYTES - 1)), reinterpret_cast<const void*>(g0), (uint32_t)(g1 - g0),
                              bars + 8 * s);
                }
                pending |= 1u << s;
            }
            next++;
        }
        wait_slot(slot_of(Lb));
        wait_slot(slot_of(Le));
    }
    __device__ __forceinline__ void drain() {
#pragma unroll
        for (uint32_t s = 0; s < RING_SLOTS; s++) wait_slot(s);
    }
    // unaligned little-endian u32 at global address a (its line(s) ensured)
    __device__ __forceinline__ uint32_t read_u32(uint64_t a) const {
        const uint32_t o = (uint32_t)a & (RING_BYTES - 1);
        const uint32_t w0 = lds_u32(data + (o & ~3u)), w1 = lds_u32(data + ((o + 4) & (RING_BYTES - 4)));
        return __funnelshift_r(w0, w1, (o & 3u) * 8);
    }
};

// 255-run length extension bytes (src/lz4.zig:368-382 / :416-429)
__device__ __forceinline__ void write_len_ext(uint8_t* p, uint32_t L, uint32_t cnt, uint32_t lane) {
    for (uint32_t i = lane; i < cnt; i += 32) p[i] = (i + 1 == cnt) ? (uint8_t)((L - 15u) % 255u) : (uint8_t)255;
}

template <typename TableT>
__device__ void compress_block_general(const uint8_t* __restrict__ src, uint32_t n, uint8_t* __restrict__ dst, uint32_t cap,
                               TableT* table, uint32_t accel, uint32_t lane, uint32_t& olen, int& st) {
    st = ST_OK;
    olen = 0;
    if (n == 0) return;                                          // :299
    if (n > LZ4_MAX_INPUT_SIZE) { st = ST_INPUT_TOO_LARGE; return; }  // :296
    uint32_t op = 0, anchor = 0;

    if (n >= MFLIMIT + 1) {                                      // :302
        {   // HashTable.init(), :307
            uint4 z = make_uint4(0, 0, 0, 0);
            uint4* t4 = reinterpret_cast<uint4*>(table);
            constexpr uint32_t NV = HASH_ENTRIES * sizeof(TableT) / 16;
#pragma unroll 4
            for (uint32_t i = lane; i < NV; i += 32) t4[i] = z;
            __syncwarp();
        }
        const uint32_t lim = n - MFLIMIT;         // mflimitPlusOne, :313
        const uint32_t mlimit = n - LASTLITERALS;  // matchLimit, :314
        const uint32_t a0 = accel < 1 ? 1u : (accel > ACCELERATION_MAX ? ACCELERATION_MAX : accel);  // :321
        const uint32_t skip = a0 < 64 ? 64 - a0 : 0;  // repeated (step == 0) iterations that are no-ops
        const uint32_t F0 = step_prefix(a0);
        const uint32_t lt = lanemask_lt(), gt = lanemask_gt();
        uint32_t q = 1;                                          // :317

        while (q < lim) {                                        // :320
            // the forward window is read through L1: keep the next lines coming
            if (lane == 0) {
                if (q + 384 < n) prefetch_l1(src + q + 384);
                if (q + 4096 < n) prefetch_l2(src + q + 4096);
            }
            // ---------------- search: windows of 32 iterations of the loop at :329-355 ----------------
            uint32_t j0 = 0, mpos = 0, mcand = 0;
            bool found = false;
            for (;;) {
                uint32_t j = j0 + lane;
                uint32_t k = j + 1 + (j >= 2 ? skip : 0);        // reference iteration number (1-based)
                uint32_t p, s;
                if (k == 1) { p = q; s = a0; }
                else { uint32_t x = a0 + k - 2; s = x >> 6; p = q + a0 + (step_prefix(x) - F0); }
                bool can = (p + s <= lim);                       // !(forwardIp > mflimitPlusOne), :335
                uint32_t em = __ballot_sync(FULL, !can);
                uint32_t E = em ? (uint32_t)__ffs(em) - 1 : 32;  // first iteration that exits
                bool active = lane < E;
                uint32_t v = 0, h = 0x80000000u | lane, cand = 0;
                if (active) { v = ld_u32x(src + p); h = hash4(v); cand = table[h]; }
                uint32_t peers = __match_any_sync(FULL, h);
                uint32_t prev = peers & lt;
                int sl = prev ? 31 - __clz(prev) : (int)lane;
                uint32_t pp = __shfl_sync(FULL, p, sl);
                if (prev) cand = pp;                             // what an earlier iteration just put()
                bool valid = active && cand > 0 && cand < p && cand + MAX_DISTANCE >= p;  // :345-347
                if (valid) valid = (ld_u32x(src + cand) == v);   // :348
                uint32_t vm = __ballot_sync(FULL, valid);
                uint32_t L = vm ? (uint32_t)__ffs(vm) - 1 : 31;
                uint32_t le = (L == 31) ? FULL : ((2u << L) - 1);
                bool commit = active && lane <= L && ((peers & gt & le) == 0);  // put(), :350
                __syncwarp();
                if (commit) table[h] = (TableT)p;
                __syncwarp();
                if (vm) { mpos = __shfl_sync(FULL, p, L); mcand = __shfl_sync(FULL, cand, L); found = true; break; }
                if (E < 32) break;                               // -> finishCompression, :335-338
                j0 += 32;
            }
            if (!found) break;

            // ---------------- match extension, :401-413 ----------------
            uint32_t ip = mpos;
            const uint32_t LL = ip - anchor;                     // :360
            const uint32_t offset = ip - mcand;                  // :395
            uint32_t a = ip + MINMATCH, b = mcand + MINMATCH, ml = 0;
            {   // first 16 bytes on four lanes (their sectors are in L1 by now); most matches end here
                uint32_t cnt = 4;
                if (lane < 4) {
                    const uint32_t al = a + 4 * lane;
                    const uint32_t nb = al >= mlimit ? 0u : (mlimit - al >= 4 ? 4u : mlimit - al);
                    cnt = 0;
                    if (nb) {
                        const uint32_t x = ld_u32x(src + al) ^ ld_u32x(src + b + 4 * lane);
                        const uint32_t mm = x ? (uint32_t)(__ffs(x) - 1) >> 3 : 4u;
                        cnt = mm < nb ? mm : nb;
                    }
                }
                const uint32_t stopm = __ballot_sync(FULL, cnt < 4);
                if (stopm) {
                    const uint32_t f = (uint32_t)__ffs(stopm) - 1;
                    ml = 4 * f + __shfl_sync(FULL, cnt, f);
                } else {
                    ml = 16; a += 16; b += 16;
                    for (;;) {
                        uint32_t al = a + 4 * lane;
                        uint32_t nb = al >= mlimit ? 0u : (mlimit - al >= 4 ? 4u : mlimit - al);
                        uint32_t c2 = 0;
                        if (nb) {
                            uint32_t x = ld_u32x(src + al) ^ ld_u32x(src + b + 4 * lane);
                            uint32_t mm = x ? (uint32_t)(__ffs(x) - 1) >> 3 : 4u;
                            c2 = mm < nb ? mm : nb;
                        }
                        uint32_t sm = __ballot_sync(FULL, c2 < 4);
                        if (sm) {
                            uint32_t f = (uint32_t)__ffs(sm) - 1;
                            ml += 4 * f + __shfl_sync(FULL, c2, f);
                            break;
                        }
                        ml += 128; a += 128; b += 128;
                    }
                }
            }
            ip += MINMATCH + ml;

            // ---------------- emit sequence, :362-432 ----------------
            if (LL < RUN_MASK && ml < ML_MASK) {
                // short form (no length extension bytes): token | literals | offset = LL + 3 <= 17 bytes, one byte per lane
                const uint32_t seq_end = op + LL + 3;
                if (seq_end > cap) { st = ST_OUTPUT_TOO_SMALL; return; }  // monotone in op: see DESIGN.md
                uint32_t bv = (LL << 4) | ml;
                if (lane >= 1 && lane <= LL) bv = __ldg(src + anchor + lane - 1);
                else if (lane == LL + 1) bv = offset;
                else if (lane == LL + 2) bv = offset >> 8;
                if (lane < LL + 3) dst[op + lane] = (uint8_t)bv;
                op = seq_end;
            } else {
                const uint32_t nll = LL >= RUN_MASK ? (LL - RUN_MASK) / 255 + 1 : 0;
                const uint32_t nml = ml >= ML_MASK ? (ml - ML_MASK) / 255 + 1 : 0;
                const uint32_t seq_end = op + 1 + nll + LL + 2 + nml;
                if (seq_end > cap) { st = ST_OUTPUT_TOO_SMALL; return; }
                uint8_t* o = dst + op;
                if (lane == 0) o[0] = (uint8_t)(((LL < 15 ? LL : 15u) << 4) | (ml < 15 ? ml : 15u));
                write_len_ext(o + 1, LL, nll, lane);
                warp_copy<true>(o + 1 + nll, src + anchor, LL, lane);
                uint8_t* o2 = o + 1 + nll + LL;
                if (lane == 0) { o2[0] = (uint8_t)(offset & 0xFF); o2[1] = (uint8_t)(offset >> 8); }
                write_len_ext(o2 + 2, ml, nml, lane);
                op = seq_end;
            }
            anchor = ip;                                         // :435
            if (ip < lim) {                                      // :438-442
                if (lane == 0) table[hash4(ld_u32x(src + ip))] = (TableT)ip;
                ip += 1;
            }
            __syncwarp();
            q = ip;
        }
    }

    // ---------------- last literals: compressAsLiterals :449-482 / finishCompression :484-519 ----------------
    const uint32_t LL = n - anchor;
    const uint32_t nll = LL >= RUN_MASK ? (LL - RUN_MASK) / 255 + 1 : 0;
    const uint32_t total = op + 1 + nll + LL;
    if (total > cap) { st = ST_OUTPUT_TOO_SMALL; return; }
    uint8_t* o = dst + op;
    if (lane == 0) o[0] = (uint8_t)((LL < 15 ? LL : 15u) << 4);
    write_len_ext(o + 1, LL, nll, lane);
    warp_copy<true>(o + 1 + nll, src + anchor, LL, lane);
    olen = total;
}

// ------------------------------------------------------------------------------------------------
// acceleration == 1 (compressDefault, every frame block): the lean path.
//
// With step == 1 the reference's iterations visit q, q+1, q+2, ... (p_j = q + j for the first
// 65 iterations), and the position e = q - 1 is the one put() after the previous match (:438-442).  A
// window is therefore the 32 consecutive positions p = e + lane.  One window costs two dependent memory
// round trips (the input words, then the bytes at the table candidates), so it is made to yield every
// sequence that starts inside it, not just the first:
//   * every lane reads its word, its bucket (`old`, the table before this window) and verifies `old`
//     against its word (range test :345-347 + 4-byte compare :348) — once;
//   * __match_any_sync groups the lanes by bucket.  While the window is walked, V is the set of lanes
//     the reference really visits (probed or put()); a lane's candidate is the nearest earlier lane of
//     its group that is in V (what the table would hold at that moment), else `old`;
//   * the walk: lane `lo` is the last put() position, lanes lo+1.. are iterations 1.. of the next search
//     (iterations 1 and 2 never take the exit at :335, later ones need p < lim); the first valid lane L
//     is the match, its literals are the low bytes of lanes lo..L-1, the match end becomes the new `lo`
//     and the lanes in between are never visited;
//   * at the end each bucket gets the position of its last visited lane (later put() wins), or `old`
//     if none of its lanes was visited.
// A search that reaches lane 31 without a match continues in the general schedule (iteration 32 - lo
// onwards, where the step grows).
// long form of a sequence (length-extension bytes and/or long literal run), :362-432.  Returns the new output
// position, or 0xFFFFFFFF if it does not fit (values, not references: a reference would pin `op` in local memory).
__device__ __noinline__ uint32_t emit_sequence_long(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint32_t cap,
                                                    uint32_t op, uint32_t anchor, uint32_t LL, uint32_t ml, uint32_t offset,
                                                    uint32_t lane) {
    const uint32_t nll = LL >= RUN_MASK ? (LL - RUN_MASK) / 255 + 1 : 0;
    const uint32_t nml = ml >= ML_MASK ? (ml - ML_MASK) / 255 + 1 : 0;
    const uint32_t seq_end = op + 1 + nll + LL + 2 + nml;
    if (seq_end > cap) return 0xFFFFFFFFu;
    uint8_t* o = dst + op;
    if (lane == 0) o[0] = (uint8_t)(((LL < 15 ? LL : 15u) << 4) | (ml < 15 ? ml : 15u));
    write_len_ext(o + 1, LL, nll, lane);
    warp_copy<true>(o + 1 + nll, src + anchor, LL, lane);
    uint8_t* o2 = o + 1 + nll + LL;
    if (lane == 0) { o2[0] = (uint8_t)(offset & 0xFF); o2[1] = (uint8_t)(offset >> 8); }
    write_len_ext(o2 + 2, ml, nml, lane);
    return seq_end;
}

template <typename TableT>
__device__ __forceinline__ bool emit_sequence(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint32_t cap,
                                              uint32_t& op, uint32_t anchor, uint32_t LL, uint32_t ml, uint32_t offset,
                                              bool lit_in_lanes, uint32_t litb, uint32_t lane) {
    if (LL < RUN_MASK && ml < ML_MASK) {
        // short form: token | literals | offset = LL + 3 <= 17 bytes, one byte per lane
        const uint32_t seq_end = op + LL + 3;
        if (seq_end > cap) return false;                         // monotone in op: see DESIGN.md
        const int32_t t = (int32_t)lane - 1 - (int32_t)LL;           // < 0: token / literal lanes, 0 and 1: offset bytes
        // Straight-line on purpose (predicated load and store in PTX): as `?:` / `if` this was two divergent regions
        // with their BSSY/BSYNC pairs, a third of the instructions issued per sequence.
        uint32_t bv = litb;
        {
            const uint32_t need = (!lit_in_lanes && t < 0) ? 1u : 0u;   // literals not in the window's registers (chained window)
            const uint8_t* lp = src + anchor + (lane ? lane - 1 : 0);
#ifdef B2_EMU
            if (need) bv = *lp;
#else
            asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p ld.global.nc.u8 %0, [%2];\n\t}"
                : "+r"(bv) : "r"(need), "l"(reinterpret_cast<uint64_t>(lp)));   // src is global memory (asserted by the kernel): its generic address IS its global address
#endif
        }
        const uint32_t ob = t == 0 ? offset : offset >> 8;
        bv = t >= 0 ? ob : bv;
        bv = lane == 0 ? ((LL << 4) | ml) : bv;
#ifdef B2_EMU
        if (t < 2) dst[op + lane] = (uint8_t)bv;
#else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.s32 p, %0, 2;\n\t@p st.global.u8 [%1], %2;\n\t}"
                     :: "r"(t), "l"(reinterpret_cast<uint64_t>(dst + op + lane)), "r"(bv) : "memory");
#endif
        op = seq_end;
    } else {
        // the caller computes the size itself, so that `op` never depends on a call result
        const uint32_t nll = LL >= RUN_MASK ? (LL - RUN_MASK) / 255 + 1 : 0;
        const uint32_t nml = ml >= ML_MASK ? (ml - ML_MASK) / 255 + 1 : 0;
        const uint32_t seq_end = op + 1 + nll + LL + 2 + nml;
        if (seq_end > cap) return false;
        emit_sequence_long(src, dst, cap, op, anchor, LL, ml, offset, lane);
        op = seq_end;
    }
    return true;
}

// number of equal bytes of src[a..] and src[b..] (b < a), stopping at mlimit = n - 5 — the loop at :401-413
__device__ __noinline__ uint32_t extend_bytes(const uint8_t* __restrict__ src, uint32_t a, uint32_t b, uint32_t mlimit,
                                                 uint32_t lane) {
    uint32_t ml = 0;
    uint32_t cnt = 4;   // first 16 bytes on four lanes; most matches end here
    if (lane < 4) {
        const uint32_t al = a + 4 * lane;
        const uint32_t nb = al >= mlimit ? 0u : (mlimit - al >= 4 ? 4u : mlimit - al);
        cnt = 0;
        if (nb) {
            const uint32_t x = ld_u32x(src + al) ^ ld_u32x(src + b + 4 * lane);
            const uint32_t mm = x ? (uint32_t)(__ffs(x) - 1) >> 3 : 4u;
            cnt = mm < nb ? mm : nb;
        }
    }
    const uint32_t stopm = __ballot_sync(FULL, cnt < 4);
    if (stopm) {
        const uint32_t f = (uint32_t)__ffs(stopm) - 1;
        return 4 * f + __shfl_sync(FULL, cnt, f);
    }
    ml = 16; a += 16; b += 16;
    for (;;) {
        uint32_t al = a + 4 * lane;
        uint32_t nb = al >= mlimit ? 0u : (mlimit - al >= 4 ? 4u : mlimit - al);
        uint32_t c2 = 0;
        if (nb) {
            uint32_t x = ld_u32x(src + al) ^ ld_u32x(src + b + 4 * lane);
            uint32_t mm = x ? (uint32_t)(__ffs(x) - 1) >> 3 : 4u;
            c2 = mm < nb ? mm : nb;
        }
        uint32_t sm = __ballot_sync(FULL, c2 < 4);
        if (sm) {
            uint32_t f = (uint32_t)__ffs(sm) - 1;
            return ml + 4 * f + __shfl_sync(FULL, c2, f);
        }
        ml += 128; a += 128; b += 128;
    }
}

// lanes a..b inclusive (a <= b <= 31)
__device__ __forceinline__ uint32_t extend_match(const uint8_t* __restrict__ src, uint32_t mpos, uint32_t mcand, uint32_t mlimit,
                                                 uint32_t lane) {
    return extend_bytes(src, mpos + MINMATCH, mcand + MINMATCH, mlimit, lane);
}

__device__ __forceinline__ uint32_t lane_range(uint32_t a, uint32_t b) { return ((2u << b) - 1u) & ~((1u << a) - 1u); }

// Writes the short sequences a window's walk has marked, all at once.  M: their match lanes; Lit: their literal lanes
// (the run of lanes before each match lane; lane = window position, its byte is the low byte of its word `v`); a match
// lane holds offset | ml << 16 in `mseq`.  Sequence k is token | literals | offset (src/lz4.zig:362-432 without length
// bytes: LL < 15, ml < 15), so a lane's output position is op + literals before it + 3 per match lane before it + 1.
// Returns false if the batch does not fit (nothing is written then).
__device__ __forceinline__ bool flush_batch(uint8_t* __restrict__ dst, uint32_t cap, uint32_t& op, uint32_t M, uint32_t Lit,
                                            uint32_t v, uint32_t mseq, uint32_t lane, uint32_t lt) {
    const uint32_t total = (uint32_t)__popc(Lit) + 3u * (uint32_t)__popc(M);
    if (op + total > cap) return false;
    uint8_t* o = dst + op + (uint32_t)__popc(Lit & lt) + 3u * (uint32_t)__popc(M & lt) + 1u;
    if ((Lit >> lane) & 1u) o[0] = (uint8_t)v;
    if ((M >> lane) & 1u) {
        const uint32_t LL = lane + (uint32_t)__clz(~Lit & lt) - 32u;     // literal lanes right before this one
        o[0] = (uint8_t)mseq;
        o[1] = (uint8_t)(mseq >> 8);
        *(o - LL - 1) = (uint8_t)((LL << 4) | (mseq >> 16));
    }
    op += total;
    return true;
}

// The search of compress_block_a1 past its first window: reference iterations j0, j0+1, ... of the search that started
// at q (acceleration 1: step = iteration >> 6, :329-333), 32 per round, until a match (returns its position, the
// candidate through *mcand) or the exit at :335 (returns 0xFFFFFFFF).  put()s are applied as the reference would.
template <typename TableT>
__device__ __noinline__ uint32_t search_later_windows(const uint8_t* __restrict__ src, TableT* table, uint32_t lim, uint32_t q,
                                                      uint32_t j0, uint32_t lane, uint32_t* mcand_out) {
    const uint32_t lt = lanemask_lt(), gt = lanemask_gt();
    for (;;) {
        uint32_t j = j0 + lane;
        uint32_t pj, s;
        if (j < 2) { pj = q + j; s = j == 0 ? 1u : 0u; }     // iterations 1 and 2 (only when lo == 31, 30)
        else { uint32_t x = j + 63; s = x >> 6; pj = q + 1 + step_prefix(x); }
        bool can = (pj + s <= lim);                      // :335
        uint32_t em = __ballot_sync(FULL, !can);
        uint32_t E = em ? (uint32_t)__ffs(em) - 1 : 32;
        bool active = lane < E;
        uint32_t v2 = 0, h2 = 0x80000000u | lane, cand = 0;
        if (active) { v2 = ld_u32x(src + pj); h2 = hash4(v2); cand = table[h2]; }
        uint32_t peers2 = __match_any_sync(FULL, h2);
        uint32_t prev = peers2 & lt;
        int sl = prev ? 31 - __clz(prev) : (int)lane;
        uint32_t pp = __shfl_sync(FULL, pj, sl);
        if (prev) cand = pp;
        bool valid = active && cand > 0 && cand < pj && cand + MAX_DISTANCE >= pj;
        if (valid) valid = (ld_u32x(src + cand) == v2);
        uint32_t vm = __ballot_sync(FULL, valid);
        uint32_t L = vm ? (uint32_t)__ffs(vm) - 1 : 31;
        uint32_t le = (L == 31) ? FULL : ((2u << L) - 1);
        bool commit = active && lane <= L && ((peers2 & gt & le) == 0);
        __syncwarp();
        if (commit) table[h2] = (TableT)pj;
        __syncwarp();
        if (vm) { *mcand_out = __shfl_sync(FULL, cand, L); return __shfl_sync(FULL, pj, L); }
        if (E < 32) return 0xFFFFFFFFu;
        j0 += 32;
    }
}

template <typename TableT, bool RING>
__device__ void compress_block_a1(const uint8_t* __restrict__ src, uint32_t n, uint8_t* __restrict__ dst, uint32_t cap,
                                  TableT* table, uint32_t lane, uint32_t& olen, int& st, Ring& ring) {
    st = ST_OK;
    olen = 0;
    if (n == 0) return;                                          // :299
    if (n > LZ4_MAX_INPUT_SIZE) { st = ST_INPUT_TOO_LARGE; return; }  // :296
    uint32_t op = 0, anchor = 0;

    if (n >= MFLIMIT + 1) {                                      // :302
        {   // HashTable.init(), :307
            uint4 z = make_uint4(0, 0, 0, 0);
            uint4* t4 = reinterpret_cast<uint4*>(table);
            constexpr uint32_t NV = HASH_ENTRIES * sizeof(TableT) / 16;
#pragma unroll 4
            for (uint32_t i = lane; i < NV; i += 32) t4[i] = z;
            __syncwarp();
        }
        const uint32_t lim = n - MFLIMIT;         // mflimitPlusOne, :313
        const uint32_t mlimit = n - LASTLITERALS;  // matchLimit, :314
        const uint32_t lt = lanemask_lt(), gt = lanemask_gt();
        uint32_t e = 0;  // last put() position; the search starts at e + 1 (0: nothing put, table value 0 == empty)
        const uint64_t ga = reinterpret_cast<uint64_t>(src);
        if (RING) ring.begin_block(ga, n);
        // Forward words in registers (not in the TMA ring variant): the 32 random candidate sectors every window of
        // every warp pulls through the SM's small L1 (the tables take the shared-memory side of it) evict the input
        // lines before the next window comes back to them, so the window's own words used to cost an L2 round trip
        // before the table could even be asked.  Lane j now holds the aligned input word whose index is congruent
        // to j mod 32, of the 32 words from `wlo` on; a window takes its words from there with shuffles; the slots of the
        // words the window has moved past are re-targeted 128 bytes further and loaded while the window is worked on.
        // Only a jump over more than 56 bytes waits for memory.
        const uint32_t* const w0p = reinterpret_cast<const uint32_t*>(ga & ~uint64_t(3));   // word that holds src[0]
        uint32_t W = 0, wlo = 0xFFFFFF00u;
        // Large blocks (u32 tables: few warps per SM, a block's own latency is what counts, registers to spare): the batch
        // of a window is written while the NEXT window waits for its candidate reads (the one long wait of a window);
        // pM / pLit / ppk (literal byte | mseq << 8 per lane) hold it until then.  4 MiB text blocks 136 -> 127 ms.  With
        // u16 tables (72 registers, 28 warps per SM hide the wait) the deferral only costs a spill: 15.9 -> 16.3 ms per GiB.
        constexpr bool DEFER = sizeof(TableT) == 4;
        uint32_t pM = 0, pLit = 0, ppk = 0;

        while (e + 1 < lim) {                                    // :320 with ip == e + 1
            if (RING) ring.ensure(ga + e, lane);
            else if (lane == 0 && e + 1024 < lim) prefetch_l2(src + e + 1024);
            // ---------------- window: p = base + lane ----------------
            const uint32_t base = e;
            const uint32_t p = base + lane;
            EMU_COUNT(windows, 1);
            const bool inwin = p <= lim;                         // positions the walk may touch (p + 3 <= n - 9)
            uint32_t v = 0, h = 0x80000000u | lane, old = 0;
            uint32_t F1, F2, F3;                                 // the words at p + 4, p + 8, p + 12 (for the extension)
            if (RING) {
                uint32_t v2 = 0;
                if (inwin) v = ring.read_u32(ga + p);
                if (lane < 16 && p + 32 <= lim) v2 = ring.read_u32(ga + p + 32);
                const uint32_t a1 = __shfl_sync(FULL, v, lane + 4), b1 = __shfl_sync(FULL, v2, lane + 4);
                const uint32_t a2 = __shfl_sync(FULL, v, lane + 8), b2 = __shfl_sync(FULL, v2, lane + 8);
                const uint32_t a3 = __shfl_sync(FULL, v, lane + 12), b3 = __shfl_sync(FULL, v2, lane + 12);
                F1 = lane + 4 < 32 ? a1 : b1; F2 = lane + 8 < 32 ? a2 : b2; F3 = lane + 12 < 32 ? a3 : b3;
            } else {
                // (recomputed from the pointer and lim every window: as block constants they were spilled, and a spill
                // reload goes through the same L1 the candidate reads keep flushing)
                uint32_t a0 = (uint32_t)ga, limv = lim;
#ifndef B2_EMU
                asm volatile("" : "+r"(a0), "+r"(limv));
#endif
                a0 &= 3u;
                const uint32_t nwords = (a0 + limv + MFLIMIT + 3) >> 2;      // words holding bytes of the block: nothing beyond is read
                const uint32_t i0 = (a0 + base) >> 2;
                if (i0 - wlo > 14u) {                                        // (also i0 < wlo): not covered, fetch and wait
                    wlo = i0;
                    const uint32_t idx = i0 + ((lane - i0) & 31u);
                    W = idx < nwords ? __ldg(w0p + idx) : 0u;
                }
                const uint32_t ab = ((a0 + base) & 3u) + lane;               // byte offset from word i0
                const uint32_t sl = i0 + (ab >> 2), sh = (ab & 3u) * 8u;
                const uint32_t x0 = __shfl_sync(FULL, W, sl), x1 = __shfl_sync(FULL, W, sl + 1), x2 = __shfl_sync(FULL, W, sl + 2),
                               x3 = __shfl_sync(FULL, W, sl + 3), x4 = __shfl_sync(FULL, W, sl + 4);
                if (inwin) v = __funnelshift_r(x0, x1, sh);
                F1 = __funnelshift_r(x1, x2, sh); F2 = __funnelshift_r(x2, x3, sh); F3 = __funnelshift_r(x3, x4, sh);
                if (i0 != wlo) {
                    // slide: the slots of words [wlo, i0) take [wlo + 32, i0 + 32).  Loaded straight into W: nothing reads W
                    // before the next window's shuffles, a whole window's work away.
                    const uint32_t r = (lane - wlo) & 31u;
                    const uint32_t idx = wlo + 32u + r;
                    if (r < i0 - wlo) W = idx < nwords ? __ldg(w0p + idx) : 0u;
                    wlo = i0;
                }
            }
            if (inwin) { h = hash4(v); old = table[h]; }
            const uint32_t peers = __match_any_sync(FULL, h);
            bool vold = inwin && old > 0 && old + MAX_DISTANCE >= p;     // :345-347 (old < e <= p always)
            // One 16-byte read at the candidate serves the 4-byte compare (:348) and stages the next
            // 5..12 bytes of the match (m1..m3, nmb of them valid) for the extension (:401-413).
            uint32_t m1 = 0, m2 = 0, m3 = 0, nmb = 0;
            uint2 A = make_uint2(0u, 0u), B = A;
            if (vold) {
                const uint2* c8 = reinterpret_cast<const uint2*>(reinterpret_cast<uintptr_t>(src + old) & ~uintptr_t(7));
                A = __ldg(c8);
                B = __ldg(c8 + 1);   // (holds input bytes: old + 8 - c < n always, old < p <= n - 12)
            }
            if (DEFER && pM) {                                   // the previous window's sequences, under the reads just issued
                if (!flush_batch(dst, cap, op, pM, pLit, ppk, ppk >> 8, lane, lt)) { st = ST_OUTPUT_TOO_SMALL; return; }
                pM = 0;
            }
            if (vold) {
                const uintptr_t ca = reinterpret_cast<uintptr_t>(src + old);
                const uint32_t c = (uint32_t)(ca & 7);
                const uint32_t sh = (c & 3) * 8;
                const bool hiw = c >= 4;
                const uint32_t w0 = hiw ? A.y : A.x, w1 = hiw ? B.x : A.y, w2 = hiw ? B.y : B.x, w3 = hiw ? 0u : B.y;
                vold = (__funnelshift_r(w0, w1, sh) == v);               // :348
                m1 = __funnelshift_r(w1, w2, sh);
                m2 = __funnelshift_r(w2, w3, sh);
                m3 = __funnelshift_r(w3, 0u, sh);
                nmb = 12 - c;
            }
            // Every lane extends its own (table) match as far as the staged bytes reach.  The forward words are
            // the reads of lanes +4, +8, +12; for the last lanes of the window they come from one more read of
            // the 16 positions after it (v2).  mlpk = length | "bytes exhausted, continue from memory" << 8.
            uint32_t mlpk;
            {
                uint32_t nfw = inwin ? (lim - p) >> 2 : 0u;              // forward words at positions <= lim
                nfw = nfw < 3 ? nfw : 3;
                const uint32_t navail = 4 * nfw < nmb ? 4 * nfw : nmb;
                const uint32_t x1 = F1 ^ m1, x2 = F2 ^ m2, x3 = F3 ^ m3;
                // equal leading bytes of each word, branch-free: clz(brev(0)) = 32 -> 4 (as nested ?: this was two
                // divergent regions per window)
                const uint32_t c1 = (uint32_t)__clz(__brev(x1)) >> 3, c2 = (uint32_t)__clz(__brev(x2)) >> 3,
                               c3 = (uint32_t)__clz(__brev(x3)) >> 3;
                const uint32_t d = c1 + (c1 == 4 ? c2 + (c2 == 4 ? c3 : 0u) : 0u);
                const bool more_bytes = d >= navail && p + MINMATCH + navail < mlimit;
                mlpk = (d < navail ? d : navail) | (more_bytes ? 0x100u : 0u);
            }
            const bool plt = p < lim;           // (implies inwin)
            uint32_t lo = 0, V = 1u;            // lane `lo`: last put(); V: lanes the reference has visited
            bool finish = false, more = false;  // finish: the block's search loop is over; more: search continues past lane 31
            // Short sequences whose literals sit in the window's lanes are not written one by one: the walk only marks
            // them (M: their match lanes, Lit: their literal lanes; a match lane keeps offset | ml << 16 in `mseq`) and
            // flush_batch writes all of them at once.
            uint32_t M = 0, Lit = 0;
            uint32_t mseq = ((p - old) & 0xFFFFu) | ((mlpk & 0xFFu) << 16);    // as a match against the table candidate
            // Static chain: a lane without an earlier lane of its bucket in this window sees the table as it was before
            // the window whatever the walk visits, so its candidate (old), its validity (vold) and its length (mlpk) are
            // known now.  J answers, for the search that starts after this lane (this lane = `lo`): where is the next
            // match and where does it end — if that is decided statically (next candidate is such a lane, fully measured,
            // fewer than 15 literals); everything else (J_SLOW) goes through the general step below.  Only in windows far
            // from the block's end (every lane eligible, no exit at :335, match ends < lim).
            constexpr uint32_t J_NONE = 0x10000u, J_SLOW = 0x20000u;
            const bool fast_ok = base + 64 <= lim;
            uint32_t J = J_SLOW;
            if (fast_ok) {
                const bool hasprev = (peers & lt) != 0;
                const uint32_t Cm = __ballot_sync(FULL, hasprev || vold);             // lanes that may be a match
                const uint32_t Sm = __ballot_sync(FULL, !hasprev && vold && (mlpk >> 8) == 0);   // ... decided statically
                const uint32_t nx = Cm & gt;
                const uint32_t nl = nx ? (uint32_t)__ffs(nx) - 1 : lane;
                const uint32_t mlen = __shfl_sync(FULL, mlpk & 0xFFu, nl);
                const bool slow = ((Sm >> nl) & 1u) == 0 || nl - lane >= RUN_MASK;
                J = nx == 0 ? J_NONE : (slow ? J_SLOW : ((nl + MINMATCH + mlen) | (nl << 8)));
            }
            for (;;) {
                if (fast_ok && anchor == base + lo) {
                    bool window_done = false;
                    for (;;) {
                        const uint32_t j = __shfl_sync(FULL, J, lo);
                        if (j & (J_NONE | J_SLOW)) {
                            if (j & J_NONE) { V |= ~((2u << lo) - 1u); more = true; window_done = true; }   // as `m == 0` below
                            break;
                        }
                        const uint32_t L = (j >> 8) & 31u, mendl = j & 0xFFu;
                        EMU_COUNT(sequences, 1);
                        EMU_COUNT(batched_sequences, 1);
                        M |= 1u << L;
                        const uint32_t lr = lane_range(lo, L - 1);
                        Lit |= lr;
                        V |= lr << 1;
                        if (mendl >= 31) {                                   // (mend < lim: fast_ok)
                            e = base + mendl;
                            anchor = e;
                            window_done = true;
                            break;
                        }
                        lo = mendl;
                        V |= 1u << lo;
                        anchor = base + lo;
                    }
                    if (window_done) break;
                }
                // lanes lo+1.. are iterations 1.. of the search that starts at e + lo + 1 (< lim)
                const bool elig = lane > lo && (plt || (inwin && lane == lo + 2));
                // candidate: nearest earlier lane of the same bucket that has been visited when this lane is probed
                const uint32_t seen = V | ~((2u << lo) - 1u);            // visited so far + every lane after lo (probed before me)
                const uint32_t pv = peers & lt & seen;
                const int pl = pv ? 31 - __clz(pv) : (int)lane;
                const uint32_t pvv = __shfl_sync(FULL, v, pl);
                const bool valid = elig && (pv ? (pvv == v && base + pl > 0) : vold);
                const uint32_t m = __ballot_sync(FULL, valid);
                if (m == 0) {
                    const uint32_t em = __ballot_sync(FULL, elig);       // a prefix range lo+1..hiE (may be empty)
                    V |= em;
                    if (em & 0x80000000u) more = true; else finish = true;   // ran off the window / hit the exit at :335
                    break;
                }
                const uint32_t L = (uint32_t)__ffs(m) - 1;
                V |= lane_range(lo + 1, L);
                const uint32_t mpos = base + L;
                EMU_COUNT(sequences, 1);
                // blocks <= 64 KiB (u16 table): candidate position and its pre-measured extension travel in one shuffle
                uint32_t mcand, pk;
                if (sizeof(TableT) == 2) {
                    const uint32_t both = __shfl_sync(FULL, (pv ? base + (uint32_t)pl : old) | (mlpk << 16), L);
                    mcand = both & 0xFFFFu;
                    pk = both >> 16;
                } else {
                    mcand = __shfl_sync(FULL, pv ? base + (uint32_t)pl : old, L);
                    pk = __shfl_sync(FULL, mlpk, L);
                }
                const uint32_t LL = mpos - anchor;                       // anchor == base + lo, except in a chained window's first search
                uint32_t ml;
                if (mcand < base) {                                      // table candidate: extension already measured
                    ml = pk & 0xFFu;
                    if (pk >> 8) ml += extend_bytes(src, mpos + MINMATCH + ml, mcand + MINMATCH + ml, mlimit, lane);
                } else {                                                 // candidate inside the window (runs, short periods)
                    ml = extend_bytes(src, mpos + MINMATCH, mcand + MINMATCH, mlimit, lane);
                }
                if (anchor == base + lo && LL < RUN_MASK && ml < ML_MASK) {
                    M |= 1u << L;                                        // joins the batch
                    Lit |= lane_range(lo, L - 1);
                    if (lane == L) mseq = (mpos - mcand) | (ml << 16);
                } else {
                    // long form, or literals that start before the window: written now, after what is pending
                    if (M) {
                        if (!flush_batch(dst, cap, op, M, Lit, v, mseq, lane, lt)) { st = ST_OUTPUT_TOO_SMALL; return; }
                        M = 0; Lit = 0;
                    }
                    if (!emit_sequence<TableT>(src, dst, cap, op, anchor, LL, ml, mpos - mcand, false, 0, lane)) {
                        st = ST_OUTPUT_TOO_SMALL;
                        return;
                    }
                }
                const uint32_t mend = mpos + MINMATCH + ml;
                anchor = mend;                                           // :435
                if (mend - base >= 31 || mend >= lim) {                  // next put() opens the next window, or there is none (:438)
                    e = mend;
                    break;
                }
                lo = mend - base;                                        // put(mend), :440
                V |= 1u << lo;
                if (mend + 1 >= lim) { finish = true; break; }           // :320
            }
            if (DEFER) {
                if (M) { pM = M; pLit = Lit; ppk = (v & 0xFFu) | (mseq << 8); }
            } else if (M && !flush_batch(dst, cap, op, M, Lit, v, mseq, lane, lt)) { st = ST_OUTPUT_TOO_SMALL; return; }
            // ---------------- table after the window: last visited lane of every bucket, else unchanged ----------------
            {
                const uint32_t vg = peers & V;
                const uint32_t wl = vg ? 31u - (uint32_t)__clz(vg) : 0u;
                const uint32_t val = vg ? base + wl : old;
                if (inwin && (peers & lt) == 0) table[h] = (TableT)val;
            }
            __syncwarp();
            if (finish) break;
            if (more && base + 31 - anchor <= 33 && base + 128 <= lim) {
                // The search goes on past lane 31 after fewer than 34 iterations: its next 31 iterations still visit
                // consecutive positions (step 1 up to iteration 64), so they are one more ordinary window whose lane 0
                // is the position just visited (re-putting it is a no-op).  anchor stays behind the window.
                e = base + 31;
                continue;
            }
            if (more) {
                // later windows of the same search (it started at anchor + 1 and has visited everything up to base + 31):
                // general step schedule, out of line, rare
                uint32_t mcand = 0;
                const uint32_t mpos = search_later_windows<TableT>(src, table, lim, anchor + 1, base + 31 - anchor, lane, &mcand);
                const bool found = mpos != 0xFFFFFFFFu;
                if (!found) break;
                if (DEFER && pM) {
                    if (!flush_batch(dst, cap, op, pM, pLit, ppk, ppk >> 8, lane, lt)) { st = ST_OUTPUT_TOO_SMALL; return; }
                    pM = 0;
                }
                const uint32_t ml = extend_match(src, mpos, mcand, mlimit, lane);
                if (!emit_sequence<TableT>(src, dst, cap, op, anchor, mpos - anchor, ml, mpos - mcand, false, 0, lane)) {
                    st = ST_OUTPUT_TOO_SMALL;
                    return;
                }
                anchor = mpos + MINMATCH + ml;
                e = anchor;
            }
        }
        if (DEFER && pM && !flush_batch(dst, cap, op, pM, pLit, ppk, ppk >> 8, lane, lt)) { st = ST_OUTPUT_TOO_SMALL; return; }
    }

    // ---------------- last literals: compressAsLiterals :449-482 / finishCompression :484-519 ----------------
    const uint32_t LL = n - anchor;
    const uint32_t nll = LL >= RUN_MASK ? (LL - RUN_MASK) / 255 + 1 : 0;
    const uint32_t total = op + 1 + nll + LL;
    if (total > cap) { st = ST_OUTPUT_TOO_SMALL; return; }
    uint8_t* o = dst + op;
    if (lane == 0) o[0] = (uint8_t)((LL < 15 ? LL : 15u) << 4);
    write_len_ext(o + 1, LL, nll, lane);
    warp_copy<true>(o + 1 + nll, src + anchor, LL, lane);
    olen = total;
}

template <typename TableT, bool RING>
__device__ __forceinline__ void compress_block(const uint8_t* __restrict__ src, uint32_t n, uint8_t* __restrict__ dst,
                                               uint32_t cap, TableT* table, uint32_t accel, uint32_t lane, uint32_t& olen,
                                               int& st, Ring& ring) {
    if (accel <= 1) compress_block_a1<TableT, RING>(src, n, dst, cap, table, lane, olen, st, ring);  // :321 clamps 0 to 1
    else compress_block_general<TableT>(src, n, dst, cap, table, accel, lane, olen, st);
}

#ifndef B2_EMU
// WARPS warps per CTA, CTAS CTAs per SM.  RING: one CTA per SM whose warps also own a forward input ring.
template <typename TableT, bool RING, int WARPS, int CTAS>
__global__ void __launch_bounds__(WARPS * 32, CTAS) k_compress_fast(BlockSet in, OutSet out, uint32_t* __restrict__ out_len,
                                                              int32_t* __restrict__ status, uint32_t nblocks,
                                                              uint32_t accel, uint32_t* ticket) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t warp = threadIdx.x >> 5;
    TableT* table = reinterpret_cast<TableT*>(smem_raw) + warp * HASH_ENTRIES;
    const uint32_t lane = lane_id();
    Ring ring;
    ring.data = ring.bars = ring.pending = ring.parity = ring.next = 0;
    ring.line0 = ring.lo16 = ring.hi16 = 0;
    if (RING) {
        uint8_t* rbase = smem_raw + (size_t)WARPS * HASH_ENTRIES * sizeof(TableT);       // multiple of 8 KiB: 512-aligned
        ring.data = smem_addr(rbase + warp * RING_BYTES);
        ring.bars = smem_addr(rbase + WARPS * RING_BYTES + warp * RING_SLOTS * 8);
        if (lane == 0) {
            for (uint32_t s = 0; s < RING_SLOTS; s++) mbar_init(ring.bars + 8 * s, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncwarp();
    }
    for (;;) {
        uint32_t blk = 0;
        if (lane == 0) blk = atomicAdd(ticket, 1u);
        blk = __shfl_sync(FULL, blk, 0);
        if (blk >= nblocks) break;
        const uint8_t* src; uint32_t n;
        uint8_t* dst; uint32_t cap;
        in.get(blk, src, n);
        out.get(blk, dst, cap);
        // keep the two block pointers in registers (otherwise every access re-adds base + offset from the constant bank)
        asm volatile("" : "+l"(src));
        asm volatile("" : "+l"(dst));
        __builtin_assume(__isGlobal(src));
        __builtin_assume(__isGlobal(dst));
        uint32_t olen; int st;
        compress_block<TableT, RING>(src, n, dst, cap, table, accel, lane, olen, st, ring);
        if (lane == 0) {
            out_len[blk] = st == ST_OK ? olen : 0u;
            status[blk] = st;
        }
        __syncwarp();
    }
    if (RING) ring.drain();   // no bulk copy may still be writing this CTA's shared memory when it exits
}

// Host launcher.  max_len decides the table entry width.
cudaError_t launch_compress_fast(const BlockSet& in, const OutSet& out, uint32_t* out_len, int32_t* status,
                                 uint32_t nblocks, uint32_t max_len, uint32_t accel, uint32_t* ticket, int num_sms,
                                 cudaStream_t stream) {
    if (nblocks == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(ticket, 0, sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
    const bool small = max_len <= 65536;
    // shared memory per SM is what bounds the blocks in flight (227 KiB per CTA, 228 per SM, 1 KiB reserved per CTA):
    //   u16 tables (8 KiB / warp):  CTAs of 7 warps, 4 per SM -> 28 blocks in flight (k1_variant 1: 3 warps, 9 per SM -> 27)
    //   u32 tables (16 KiB / warp): CTAs of 7 warps, 2 per SM -> 14 blocks in flight (k1_variant 1: 1 warp, 13 per SM -> 13)
    //   ring variant (b2lz4_debug_tune("k1_variant", 2)): ONE CTA per SM of 26 (u16) / 13 (u32) warps, each with its table,
    //   a 512-byte forward input ring fed by TMA bulk copies and four mbarriers
    const Tune& t = tune();
    const bool ring = t.k1_variant == 2;
    static std::once_flag attr_once[16];
    int dev = 0;
    cudaGetDevice(&dev);
    std::call_once(attr_once[dev & 15], [] {
        auto set = [](const void* f, int bytes) {
            cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
            cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        };
        set((const void*)k_compress_fast<uint16_t, false, 3, 9>, 3 * HASH_ENTRIES * 2);
        set((const void*)k_compress_fast<uint32_t, false, 1, 13>, HASH_ENTRIES * 4);
        set((const void*)k_compress_fast<uint16_t, false, 7, 4>, 7 * HASH_ENTRIES * 2);
        set((const void*)k_compress_fast<uint32_t, false, 7, 2>, 7 * HASH_ENTRIES * 4);
        // few blocks (at most one CTA per SM): leave the rest of the SM's memory to the L1, where the candidate reads of
        // the seven windows in flight then find their blocks' recent history (4 MiB text blocks 127 -> 120 ms), and let
        // the compiler have the registers it wants (no minimum of resident CTAs)
        cudaFuncSetAttribute((const void*)k_compress_fast<uint32_t, false, 7, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 7 * HASH_ENTRIES * 4);
        cudaFuncSetAttribute((const void*)k_compress_fast<uint32_t, false, 7, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, 50);
        cudaFuncSetAttribute((const void*)k_compress_fast<uint16_t, false, 7, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 7 * HASH_ENTRIES * 2);
        cudaFuncSetAttribute((const void*)k_compress_fast<uint16_t, false, 7, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, 25);
        set((const void*)k_compress_fast<uint16_t, true, 26, 1>, 26 * (HASH_ENTRIES * 2 + (int)RING_BYTES + (int)RING_SLOTS * 8));
        set((const void*)k_compress_fast<uint32_t, true, 13, 1>, 13 * (HASH_ENTRIES * 4 + (int)RING_BYTES + (int)RING_SLOTS * 8));
    });
    if (ring && accel <= 1) {
        const int warps = small ? 26 : 13;
        const size_t smem = (size_t)warps * (HASH_ENTRIES * (small ? 2 : 4) + RING_BYTES + RING_SLOTS * 8);
        uint32_t want = (nblocks + warps - 1) / warps;
        uint32_t grid = want < (uint32_t)num_sms ? want : (uint32_t)num_sms;
        if (small) k_compress_fast<uint16_t, true, 26, 1><<<grid, warps * 32, smem, stream>>>(in, out, out_len, status, nblocks, accel, ticket);
        else k_compress_fast<uint32_t, true, 13, 1><<<grid, warps * 32, smem, stream>>>(in, out, out_len, status, nblocks, accel, ticket);
    } else if (t.k1_variant == 1) {
        // the round-1/2 shape (27 / 13 blocks per SM), kept for A/B measurements
        const int warps = small ? 3 : 1;
        int ctas_per_sm = small ? 9 : 13;
        if (t.k1_ctas > 0 && t.k1_ctas < ctas_per_sm) ctas_per_sm = t.k1_ctas;   // occupancy experiments
        const size_t smem = (size_t)warps * HASH_ENTRIES * (small ? sizeof(uint16_t) : sizeof(uint32_t));
        uint32_t want = (nblocks + warps - 1) / warps;
        uint32_t maxg = (uint32_t)(num_sms * ctas_per_sm);
        uint32_t grid = want < maxg ? want : maxg;
        if (small) k_compress_fast<uint16_t, false, 3, 9><<<grid, warps * 32, smem, stream>>>(in, out, out_len, status, nblocks, accel, ticket);
        else k_compress_fast<uint32_t, false, 1, 13><<<grid, warps * 32, smem, stream>>>(in, out, out_len, status, nblocks, accel, ticket);
    } else {
        // 7-warp CTAs: 4 per SM with u16 tables (4 x (56 + 1) KiB = all 228 KiB of the SM: 28 blocks in flight, 72 registers),
        // 2 per SM with u32 tables (14 blocks in flight)
        const int warps = 7;
        int ctas_per_sm = small ? 4 : 2;
        if (t.k1_ctas > 0 && t.k1_ctas < ctas_per_sm) ctas_per_sm = t.k1_ctas;
        const size_t smem = (size_t)warps * HASH_ENTRIES * (small ? sizeof(uint16_t) : sizeof(uint32_t));
        uint32_t want = (nblocks + warps - 1) / warps;
        uint32_t maxg = (uint32_t)(num_sms * ctas_per_sm);
        uint32_t grid = want < maxg ? want : maxg;
        const bool few = want <= (uint32_t)num_sms && t.k1_variant != 4;
        if (small && few) k_compress_fast<uint16_t, false, 7, 1><<<grid, warps * 32, smem, stream>>>(in, out, out_len, status, nblocks, accel, ticket);
        else if (small) k_compress_fast<uint16_t, false, 7, 4><<<grid, warps * 32, smem, stream>>>(in, out, out_len, status, nblocks, accel, ticket);
        else if (few) k_compress_fast<uint32_t, false, 7, 1><<<grid, warps * 32, smem, stream>>>(in, out, out_len, status, nblocks, accel, ticket);
        else k_compress_fast<uint32_t, false, 7, 2><<<grid, warps * 32, smem, stream>>>(in, out, out_len, status, nblocks, accel, ticket);
    }
    count_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Expensive blocks first.  What a block costs K1 is its number of sequences: incompressible data (the search accelerates
// away) and runs (few long sequences) are cheap, everything in between is expensive — 4 ms against 0.25-0.7 ms of a warp
// at full occupancy — and with blocks taken in index order a launch ends on a tail of expensive blocks drawn late (28
// resident warps per SM, 24 active on average).  An estimate pass — K1 itself on the first KiB of every block, into the
// slots the real pass overwrites — gives est[i]; blocks whose prefix does not shrink at all go last, those that shrink to
// 15 % or less before them, everything else first: longest processing time first, as far as a prefix can tell (prefix
// ratios of the four data classes: text 0.78-0.89, binary records 0.60-0.72, runs 0.02-0.10, random 1.005).  The
// classification errs towards "expensive": ONE expensive block drawn last is a tail of its own (a 512-byte prefix with
// a 97 % threshold measured 10.6 ms on the mixed workload against 10.4 unordered and 9.95 with this rule).  The order is
// applied through the explicit (offset, length) form of the block sets, so the codec kernel itself is untouched (it is
// register-bound: two more parameters cost it 20 %); results are scattered back to block order afterwards.
struct OrderScratch {           // nb entries each, in one allocation of order_scratch_bytes(nb)
    uint64_t* in_off; uint64_t* out_off; uint32_t* in_len; uint32_t* out_cap; uint32_t* res_len; int32_t* res_st; uint32_t* est;
    uint32_t* ord;
};
size_t order_scratch_bytes(uint32_t nb) { return (size_t)nb * 40 + 64; }
static OrderScratch carve(void* p, uint32_t nb) {
    OrderScratch o;
    uint8_t* b = (uint8_t*)p;
    o.in_off = (uint64_t*)b; b += (size_t)nb * 8;
    o.out_off = (uint64_t*)b; b += (size_t)nb * 8;
    o.in_len = (uint32_t*)b; b += (size_t)nb * 4;
    o.out_cap = (uint32_t*)b; b += (size_t)nb * 4;
    o.res_len = (uint32_t*)b; b += (size_t)nb * 4;
    o.res_st = (int32_t*)b; b += (size_t)nb * 4;
    o.est = (uint32_t*)b; b += (size_t)nb * 4;
    o.ord = (uint32_t*)b;
    return o;
}

__global__ void k_est_prepare(uint64_t stride, uint64_t total, uint32_t clip, uint32_t nblocks, OrderScratch o) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nblocks) return;
    const uint64_t off = (uint64_t)i * stride, r = total > off ? total - off : 0;
    const uint32_t len = (uint32_t)(r < stride ? r : stride);
    o.in_off[i] = off;
    o.in_len[i] = len < clip ? len : clip;
}

__global__ void __launch_bounds__(1024) k_order_by_estimate(uint64_t stride, uint64_t total, uint64_t out_stride, uint32_t out_cap,
                                                            uint32_t clip, uint32_t nblocks, OrderScratch o) {
    __shared__ uint32_t cnt[3];
    __shared__ uint32_t base[3];
    const uint32_t tid = threadIdx.x;
    auto bucket = [&](uint32_t i) -> uint32_t {
        const uint64_t e = o.est[i], c = o.in_len[i];        // (in_len still holds the prefix lengths of the estimate pass)
        if (e >= c) return 2u;                                // not a byte gained: incompressible
        if (e * 100 <= c * 15) return 1u;                     // runs
        return 0u;
    };
    if (tid < 3) cnt[tid] = 0;
    __syncthreads();
    for (uint32_t i = tid; i < nblocks; i += 1024) atomicAdd(&cnt[bucket(i)], 1u);
    __syncthreads();
    if (tid == 0) { base[0] = 0; base[1] = cnt[0]; base[2] = cnt[0] + cnt[1]; }
    __syncthreads();
    for (uint32_t i = tid; i < nblocks; i += 1024) o.ord[atomicAdd(&base[bucket(i)], 1u)] = i;
    __syncthreads();
    for (uint32_t t = tid; t < nblocks; t += 1024) {          // the block sets of the real pass, in that order
        const uint32_t i = o.ord[t];
        o.in_off[t] = (uint64_t)i * stride;
        o.out_off[t] = (uint64_t)i * out_stride;
        o.out_cap[t] = out_cap;
    }
    __syncthreads();
    for (uint32_t t = tid; t < nblocks; t += 1024) {
        const uint32_t i = o.ord[t];
        const uint64_t off = (uint64_t)i * stride, r = total > off ? total - off : 0;
        o.in_len[t] = (uint32_t)(r < stride ? r : stride);
    }
}

__global__ void k_unpermute(uint32_t nblocks, OrderScratch o, uint32_t* __restrict__ out_len, int32_t* __restrict__ status) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nblocks) return;
    const uint32_t i = o.ord[t];
    out_len[i] = o.res_len[t];
    status[i] = o.res_st[t];
}

// Regular block sets only (frame bodies): `in` blocks at i * in.stride, slots at i * out.stride.
cudaError_t launch_compress_fast_ordered(const BlockSet& in, const OutSet& out, uint32_t* out_len, int32_t* status, uint32_t nblocks,
                                         uint32_t max_len, uint32_t* ticket, int num_sms, void* scratch, cudaStream_t stream) {
    constexpr uint32_t EST_CLIP = 1024;
    const OrderScratch o = carve(scratch, nblocks);
    const uint32_t g = (nblocks + 255) / 256;
    k_est_prepare<<<g, 256, 0, stream>>>(in.stride, in.total, EST_CLIP, nblocks, o);
    count_launch();
    BlockSet ein = in;
    ein.off = o.in_off; ein.len = o.in_len; ein.len_mask = 0xFFFFFFFFu;
    cudaError_t e = launch_compress_fast(ein, out, o.est, o.res_st, nblocks, max_len, 1, ticket, num_sms, stream);
    if (e != cudaSuccess) return e;
    k_order_by_estimate<<<1, 1024, 0, stream>>>(in.stride, in.total, out.stride, out.slot_cap, EST_CLIP, nblocks, o);
    count_launch();
    OutSet eout = out;
    eout.off = o.out_off; eout.cap = o.out_cap;
    e = launch_compress_fast(ein, eout, o.res_len, o.res_st, nblocks, max_len, 1, ticket, num_sms, stream);
    if (e != cudaSuccess) return e;
    k_unpermute<<<g, 256, 0, stream>>>(nblocks, o, out_len, status);
    count_launch();
    return cudaGetLastError();
}
#endif  // !B2_EMU

}  // namespace b2
