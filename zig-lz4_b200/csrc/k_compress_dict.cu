// k_compress_dict.cu — fast block compressor with a shared external dictionary.
//
// The reference's Stream.loadDict / compressFastContinue promise dictionary compression but never produce a
// match into the dictionary (the table is indexed against the current block only), so there is
// no reference output to be identical to.  This kernel defines the natural extension of the reference's own
// matcher (src/lz4.zig:292-447, same hash, same single-probe table, same step schedule, same emitter):
//   * the block is parsed as if it followed the last D = min(dict_len, 65536) dictionary bytes in one buffer
//     (virtual positions x < D are dictionary bytes, x >= D block bytes);
//   * the table starts primed with every dictionary position 1..D-4 (later position wins, exactly what
//     put() would have left), the search starts at the block's first byte;
//   * a match that starts in the dictionary ends at the dictionary's end (it is not continued into the
//     block; the stream stays valid, the match is just shorter);
//   * offsets are virtual distances, which is what decompressSafeUsingDict (src/lz4.zig:181-228, 960-964)
//     expects: an offset larger than the output position reaches into the dictionary tail.
// With an empty dictionary the output is byte-identical to compressFast.  Parity bar: round trip through
// the reference's dictionary decoder (oracle + K2), ratio reported beside the dictionary-blind one.
#include "b2_common.cuh"
#include "b2_kernels.h"

namespace b2 {

namespace {

constexpr int KD_HASH_ENTRIES = 4096;

__device__ __forceinline__ uint32_t hash4d(uint32_t v) { return (v * HASH_MULTIPLIER) >> 20; }
__device__ __forceinline__ uint32_t step_prefix_d(uint32_t x) {
    uint32_t A = x >> 6, B = x & 63;
    return 32u * A * (A - 1u) + B * A;
}
__device__ __forceinline__ uint32_t ldu32(const uint8_t* __restrict__ p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
    const uint32_t sh = (uint32_t)(a & 3) * 8;
    const uint32_t lo = __ldg(w);
    const uint32_t hi = sh ? __ldg(w + 1) : 0u;
    return __funnelshift_r(lo, hi, sh);
}
__device__ __forceinline__ void write_len_ext_d(uint8_t* p, uint32_t L, uint32_t cnt, uint32_t lane) {
    for (uint32_t i = lane; i < cnt; i += 32) p[i] = (i + 1 == cnt) ? (uint8_t)((L - 15u) % 255u) : (uint8_t)255;
}

// virtual buffer: dictionary tail [0, D) followed by the block [D, D + n)
struct VBuf {
    const uint8_t* dt;     // first of the D dictionary bytes
    uintptr_t blk_m;       // address of the block minus D
    uint32_t D;
    __device__ __forceinline__ const uint8_t* at(uint32_t x) const {
        return x < D ? dt + x : reinterpret_cast<const uint8_t*>(blk_m + x);
    }
};

}  // namespace

// table[h] = last dictionary position with hash h (positions 1..D-4; 0 stays "empty")
__global__ void k_prime_dict(const uint8_t* __restrict__ dt, uint32_t D, uint32_t* __restrict__ table) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (D >= 4 && x <= D - 4) atomicMax(&table[hash4d(ldu32(dt + x))], x);
}

__device__ void compress_block_dict(const uint8_t* __restrict__ src, uint32_t n, uint8_t* __restrict__ dst, uint32_t cap,
                                    uint32_t* table, const uint32_t* __restrict__ primed, const uint8_t* __restrict__ dt,
                                    uint32_t D, uint32_t accel, uint32_t lane, uint32_t& olen, int& st) {
    st = ST_OK;
    olen = 0;
    if (n == 0) return;                                          // :299
    if (n > LZ4_MAX_INPUT_SIZE) { st = ST_INPUT_TOO_LARGE; return; }  // :296
    VBuf V;
    V.dt = dt; V.D = D; V.blk_m = reinterpret_cast<uintptr_t>(src) - D;
    uint32_t op = 0, anchor = D;
    const uint32_t vend = D + n;

    if (n >= MFLIMIT + 1) {                                      // :302
        {   // the table as the dictionary left it (all zero without one), :307
            const uint4* p4 = reinterpret_cast<const uint4*>(primed);
            uint4* t4 = reinterpret_cast<uint4*>(table);
#pragma unroll 4
            for (uint32_t i = lane; i < KD_HASH_ENTRIES / 4; i += 32) t4[i] = __ldg(p4 + i);
            __syncwarp();
        }
        const uint32_t lim = vend - MFLIMIT;         // mflimitPlusOne, :313
        const uint32_t mlimit = vend - LASTLITERALS;  // matchLimit, :314
        const uint32_t a0 = accel < 1 ? 1u : (accel > ACCELERATION_MAX ? ACCELERATION_MAX : accel);  // :321
        const uint32_t skip = a0 < 64 ? 64 - a0 : 0;
        const uint32_t F0 = step_prefix_d(a0);
        const uint32_t lt = lanemask_lt(), gt = lanemask_gt();
        uint32_t q = D ? D : 1;                                  // :317 (position 0 is never matchable)

        while (q < lim) {                                        // :320
            uint32_t j0 = 0, mpos = 0, mcand = 0;
            bool found = false;
            for (;;) {
                uint32_t j = j0 + lane;
                uint32_t k = j + 1 + (j >= 2 ? skip : 0);        // reference iteration number (1-based)
                uint32_t p, s;
                if (k == 1) { p = q; s = a0; }
                else { uint32_t x = a0 + k - 2; s = x >> 6; p = q + a0 + (step_prefix_d(x) - F0); }
                bool can = (p + s <= lim);                       // :335
                uint32_t em = __ballot_sync(FULL, !can);
                uint32_t E = em ? (uint32_t)__ffs(em) - 1 : 32;
                bool active = lane < E;
                uint32_t v = 0, h = 0x80000000u | lane, cand = 0;
                if (active) { v = ldu32(V.at(p)); h = hash4d(v); cand = table[h]; }
                uint32_t peers = __match_any_sync(FULL, h);
                uint32_t prev = peers & lt;
                int sl = prev ? 31 - __clz(prev) : (int)lane;
                uint32_t pp = __shfl_sync(FULL, p, sl);
                if (prev) cand = pp;
                bool valid = active && cand > 0 && cand < p && cand + MAX_DISTANCE >= p;  // :345-347
                if (valid) valid = (ldu32(V.at(cand)) == v);     // :348 (dictionary candidates are <= D - 4: no straddle)
                uint32_t vm = __ballot_sync(FULL, valid);
                uint32_t L = vm ? (uint32_t)__ffs(vm) - 1 : 31;
                uint32_t le = (L == 31) ? FULL : ((2u << L) - 1);
                bool commit = active && lane <= L && ((peers & gt & le) == 0);  // put(), :350
                __syncwarp();
                if (commit) table[h] = p;
                __syncwarp();
                if (vm) { mpos = __shfl_sync(FULL, p, L); mcand = __shfl_sync(FULL, cand, L); found = true; break; }
                if (E < 32) break;
                j0 += 32;
            }
            if (!found) break;

            // ---------------- match extension, :401-413; a dictionary match stops at the dictionary's end ----------------
            uint32_t ip = mpos;
            const uint32_t LL = ip - anchor;
            const uint32_t offset = ip - mcand;
            uint32_t a = ip + MINMATCH, b = mcand + MINMATCH, ml = 0;
            uint32_t elimit = mlimit;
            if (mcand < D) { const uint32_t room = D - b; if (a + room < elimit) elimit = a + room; }
            for (;;) {
                uint32_t al = a + 4 * lane;
                uint32_t nb = al >= elimit ? 0u : (elimit - al >= 4 ? 4u : elimit - al);
                uint32_t cnt = 0;
                if (nb) {
                    uint32_t x = ldu32(V.at(al)) ^ ldu32(V.at(b + 4 * lane));
                    uint32_t mm = x ? (uint32_t)(__ffs(x) - 1) >> 3 : 4u;
                    cnt = mm < nb ? mm : nb;
                }
                uint32_t stopm = __ballot_sync(FULL, cnt < 4);
                if (stopm) {
                    uint32_t f = (uint32_t)__ffs(stopm) - 1;
                    ml += 4 * f + __shfl_sync(FULL, cnt, f);
                    break;
                }
                ml += 128; a += 128; b += 128;
            }
            ip += MINMATCH + ml;

            // ---------------- emit sequence, :362-432 ----------------
            const uint32_t nll = LL >= RUN_MASK ? (LL - RUN_MASK) / 255 + 1 : 0;
            const uint32_t nml = ml >= ML_MASK ? (ml - ML_MASK) / 255 + 1 : 0;
            const uint32_t seq_end = op + 1 + nll + LL + 2 + nml;
            if (seq_end > cap) { st = ST_OUTPUT_TOO_SMALL; return; }
            uint8_t* o = dst + op;
            if (lane == 0) o[0] = (uint8_t)(((LL < 15 ? LL : 15u) << 4) | (ml < 15 ? ml : 15u));
            write_len_ext_d(o + 1, LL, nll, lane);
            warp_copy<true>(o + 1 + nll, V.at(anchor), LL, lane);          // anchor >= D: block bytes
            uint8_t* o2 = o + 1 + nll + LL;
            if (lane == 0) { o2[0] = (uint8_t)(offset & 0xFF); o2[1] = (uint8_t)(offset >> 8); }
            write_len_ext_d(o2 + 2, ml, nml, lane);
            op = seq_end;
            anchor = ip;                                         // :435
            if (ip < lim) {                                      // :438-442
                if (lane == 0) table[hash4d(ldu32(V.at(ip)))] = ip;
                ip += 1;
            }
            __syncwarp();
            q = ip;
        }
    }

    // ---------------- last literals: compressAsLiterals :449-482 / finishCompression :484-519 ----------------
    const uint32_t LL = vend - anchor;
    const uint32_t nll = LL >= RUN_MASK ? (LL - RUN_MASK) / 255 + 1 : 0;
    const uint32_t total = op + 1 + nll + LL;
    if (total > cap) { st = ST_OUTPUT_TOO_SMALL; return; }
    uint8_t* o = dst + op;
    if (lane == 0) o[0] = (uint8_t)((LL < 15 ? LL : 15u) << 4);
    write_len_ext_d(o + 1, LL, nll, lane);
    warp_copy<true>(o + 1 + nll, V.at(anchor), LL, lane);
    olen = total;
}

constexpr int KD_WARPS = 1;   // 16 KiB table per warp: 13 single-warp CTAs per SM

__global__ void __launch_bounds__(KD_WARPS * 32) k_compress_fast_dict(BlockSet in, OutSet out, uint32_t* __restrict__ out_len,
                                                                      int32_t* __restrict__ status, uint32_t nblocks,
                                                                      const uint32_t* __restrict__ primed,
                                                                      const uint8_t* __restrict__ dt, uint32_t D, uint32_t accel,
                                                                      uint32_t* ticket) {
    __shared__ __align__(16) uint32_t table[KD_WARPS][KD_HASH_ENTRIES];
    const uint32_t lane = lane_id();
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
        compress_block_dict(src, n, dst, cap, table[threadIdx.x >> 5], primed, dt, D, accel, lane, olen, st);
        if (lane == 0) {
            out_len[blk] = st == ST_OK ? olen : 0u;
            status[blk] = st;
        }
        __syncwarp();
    }
}

// dict: device pointer to dict_len bytes (may be null when dict_len == 0); primed: 4096 u32 of scratch
cudaError_t launch_compress_fast_dict(const BlockSet& in, const OutSet& out, uint32_t* out_len, int32_t* status,
                                      uint32_t nblocks, const uint8_t* dict, uint64_t dict_len, uint32_t* primed,
                                      uint32_t accel, uint32_t* ticket, int num_sms, cudaStream_t stream) {
    if (nblocks == 0) return cudaSuccess;
    const uint32_t D = (uint32_t)(dict_len < 65536 ? dict_len : 65536);
    const uint8_t* dt = dict ? dict + (dict_len - D) : nullptr;
    cudaError_t e = cudaMemsetAsync(ticket, 0, sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(primed, 0, KD_HASH_ENTRIES * sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
    if (D >= 5) {
        k_prime_dict<<<(D + 255) / 256, 256, 0, stream>>>(dt, D, primed);
        count_launch();
    }
    uint32_t maxg = (uint32_t)(num_sms * 13);
    uint32_t grid = nblocks < maxg ? nblocks : maxg;
    k_compress_fast_dict<<<grid, KD_WARPS * 32, 0, stream>>>(in, out, out_len, status, nblocks, primed, dt, D, accel, ticket);
    count_launch();
    return cudaGetLastError();
}

}  // namespace b2
