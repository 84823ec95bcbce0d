// k_xxh32.cu — K4: XXH32 (seed-parameterised) on the device.
//
// The reference hashes through Zig's std.hash.XxHash32 (standard XXH32): block checksums over the
// *stored* bytes of each block (src/lz4f.zig:422-427, verified at :590-600) and one running content
// checksum over all raw bytes (:375,:384-386,:437-441 / :560,:617-619,:625-635).
//
// XXH32 is four dependent lane recurrences acc = rotl(acc + x*P2, 13) * P1 — a serial chain that
// cannot be split or prefix-scanned.  So:
//   * block checksums: ONE WARP PER BLOCK.  The 32 lanes stream the block through shared memory in 2 KiB
//     tiles (coalesced 16-byte loads, the next tile in flight while this one is hashed); lanes 0..3 each
//     run one accumulator chain.  (A single thread per block — the first version — exposed one DRAM round
//     trip per 64 bytes: 37 ms for 512 blocks of 4 MiB, now ~2 ms.)
//   * content checksum: one warp; 32 lanes stream 2 KiB tiles into shared memory (double buffered),
//     lanes 0..3 each run one accumulator chain.
//   The chain itself is the bound, so it is written for dependent-issue latency (chain_tile below): the
//   tile is stored one row per accumulator so a chain lane fetches four rounds of input per LDS.128, one
//   group ahead; x*P2 is computed off the chain; and the round is rewritten from
//   IMAD -> SHF(rotl) -> IMAD (three dependent ops) into b' = (b*C1 + y') + (b>>19)*P1 with b = acc + y,
//   C1 = P1<<13 (rotl(b,13)*P1 = b*C1 + (b>>19)*P1 mod 2^32): IMAD || SHF, then one IMAD — two deep.
//   Measured on one B200 warp (tools/exp/xxh_chain_bench.cu): 10.1 -> 7.3 ns per 16-byte stripe
//   (1.59 -> 2.20 GB/s); the mul.hi form of b>>19 (all on the fma pipe) is slower (9.1 ns).
#include "b2_common.cuh"
#include "b2_kernels.h"

namespace b2 {

constexpr uint32_t P1 = 2654435761u, P2 = 2246822519u, P3 = 3266489917u, P4 = 668265263u, P5 = 374761393u;

__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return __funnelshift_l(x, x, r); }
__device__ __forceinline__ uint32_t xround(uint32_t acc, uint32_t x) { return rotl32(acc + x * P2, 13) * P1; }

__device__ __forceinline__ uint32_t xxh_finish(uint32_t h, const uint8_t* p, uint32_t n) {
    // tail of < 16 bytes: 4-byte words then single bytes, then avalanche
    while (n >= 4) {
        h = rotl32(h + ldg_u32(p) * P3, 17) * P4;
        p += 4; n -= 4;
    }
    while (n) {
        h = rotl32(h + (uint32_t)__ldg(p) * P5, 11) * P1;
        p += 1; n -= 1;
    }
    h ^= h >> 15; h *= P2; h ^= h >> 13; h *= P3; h ^= h >> 16;
    return h;
}

constexpr uint32_t TILE = 2048;  // bytes per shared-memory tile (128 stripes)
constexpr int XXH_WARPS = 4;
constexpr uint32_t ROW = 132;                 // words per accumulator row: 128 + 4 pad (the four chain lanes hit different banks)
constexpr uint32_t TILE_WORDS = 4 * ROW;      // one staged tile: row j holds word j of every stripe
constexpr uint32_t C1 = P1 << 13;

// stripes held in registers (lane + 32u of the tile) -> rows of the staged tile
__device__ __forceinline__ void stage_tile(uint32_t* tile, const uint4 (&r)[4], uint32_t cnt, uint32_t lane) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const uint32_t s = lane + 32u * u;
        if (s < cnt) { uint32_t* t = tile + s; t[0] = r[u].x; t[ROW] = r[u].y; t[2 * ROW] = r[u].z; t[3 * ROW] = r[u].w; }
    }
}

__device__ __forceinline__ uint32_t chain_step(uint32_t b, uint32_t ynext) { return (b >> 19) * P1 + (b * C1 + ynext); }

// `cnt` rounds of one accumulator over its row of the staged tile
__device__ __forceinline__ uint32_t chain_tile(uint32_t acc, const uint32_t* row, uint32_t cnt) {
    if (cnt == 128) {
        uint4 cur = *reinterpret_cast<const uint4*>(row);
        uint32_t b = acc + cur.x * P2;
#pragma unroll 4
        for (uint32_t g = 0; g < 32; g++) {
            uint4 nxt = make_uint4(0, 0, 0, 0);
            if (g + 1 < 32) nxt = *reinterpret_cast<const uint4*>(row + 4 * (g + 1));
            b = chain_step(b, cur.y * P2);
            b = chain_step(b, cur.z * P2);
            b = chain_step(b, cur.w * P2);
            b = chain_step(b, nxt.x * P2);   // after the last group y = 0: b is the accumulator again
            cur = nxt;
        }
        return b;
    }
    for (uint32_t s = 0; s < cnt; s++) acc = xround(acc, row[s]);
    return acc;
}

// XXH32 of [p, p + len) by one warp; `tile` is this warp's 2 x 2 KiB staging area.  Result in every lane.
__device__ uint32_t xxh32_warp(const uint8_t* __restrict__ p, uint32_t len, uint32_t seed, uint32_t (*tile)[TILE_WORDS],
                               uint32_t lane) {
    uint32_t acc = seed;
    if (lane == 0) acc = seed + P1 + P2;
    else if (lane == 1) acc = seed + P2;
    else if (lane == 3) acc = seed - P1;
    const uint32_t nst = len >> 4;
    const uint32_t bo = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 15);
    const uint4* s16 = reinterpret_cast<const uint4*>(p - bo);
    uint4 r[4];
    auto load_tile = [&](uint32_t first) {          // stripes [first, first+128) -> registers
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t sidx = first + lane + 32u * u;
            if (sidx < nst) {
                if (bo == 0) r[u] = __ldg(s16 + sidx);
                else r[u] = extract16(__ldg(s16 + sidx), __ldg(s16 + sidx + 1), bo);
            }
        }
    };
    uint32_t done = 0;
    int buf = 0;
    if (nst) load_tile(0);
    while (done < nst) {
        const uint32_t cnt = nst - done < 128 ? nst - done : 128;
        stage_tile(tile[buf], r, cnt, lane);
        __syncwarp();
        if (done + 128 < nst) load_tile(done + 128);  // next tile in flight while the chains run
        if (lane < 4) acc = chain_tile(acc, tile[buf] + lane * ROW, cnt);
        done += cnt;
        buf ^= 1;
        __syncwarp();
    }
    const uint32_t v1 = __shfl_sync(FULL, acc, 0), v2 = __shfl_sync(FULL, acc, 1), v3 = __shfl_sync(FULL, acc, 2),
                   v4 = __shfl_sync(FULL, acc, 3);
    uint32_t h = len >= 16 ? rotl32(v1, 1) + rotl32(v2, 7) + rotl32(v3, 12) + rotl32(v4, 18) : seed + P5;
    h += len;
    return xxh_finish(h, p + ((size_t)nst << 4), len & 15);
}

__global__ void __launch_bounds__(XXH_WARPS * 32) k_xxh32_stored(BlockSet slots, BlockSet raw, const uint32_t* __restrict__ csize,
                                                                uint32_t* __restrict__ sums, uint32_t nblocks) {
    __shared__ __align__(16) uint32_t tiles[XXH_WARPS][2][TILE_WORDS];
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t i = blockIdx.x * XXH_WARPS + w;
    if (i >= nblocks) return;
    const uint8_t* rp; uint32_t rn;
    raw.get(i, rp, rn);
    const uint32_t c = csize[i];
    const uint8_t* p = rp; uint32_t n = rn;
    if (c < rn) { const uint8_t* sp; uint32_t sn; slots.get(i, sp, sn); p = sp; n = c; }  // src/lz4f.zig:407-408
    const uint32_t h = xxh32_warp(p, n, 0, tiles[w], lane);
    if (lane == 0) sums[i] = h;
}

__global__ void __launch_bounds__(XXH_WARPS * 32) k_xxh32_ranges(const uint8_t* __restrict__ base, const uint64_t* __restrict__ off,
                                                                const uint32_t* __restrict__ hdr, uint32_t* __restrict__ sums,
                                                                uint32_t nblocks) {
    __shared__ __align__(16) uint32_t tiles[XXH_WARPS][2][TILE_WORDS];
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t i = blockIdx.x * XXH_WARPS + w;
    if (i >= nblocks) return;
    const uint32_t h = xxh32_warp(base + off[i], hdr[i] & 0x7FFFFFFFu, 0, tiles[w], lane);
    if (lane == 0) sums[i] = h;
}

cudaError_t launch_xxh32_stored(const BlockSet& slots, const BlockSet& raw, const uint32_t* csize, uint32_t* sums,
                                uint32_t nblocks, cudaStream_t stream) {
    if (nblocks == 0) return cudaSuccess;
    k_xxh32_stored<<<(nblocks + XXH_WARPS - 1) / XXH_WARPS, XXH_WARPS * 32, 0, stream>>>(slots, raw, csize, sums, nblocks);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_xxh32_ranges(const uint8_t* base, const uint64_t* off, const uint32_t* hdr, uint32_t* sums,
                                uint32_t nblocks, cudaStream_t stream) {
    if (nblocks == 0) return cudaSuccess;
    k_xxh32_ranges<<<(nblocks + XXH_WARPS - 1) / XXH_WARPS, XXH_WARPS * 32, 0, stream>>>(base, off, hdr, sums, nblocks);
    count_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------ running (content) checksum ----
__global__ void k_xxh32_init(XxhState* st, uint32_t seed) {
    st->v[0] = seed + P1 + P2; st->v[1] = seed + P2; st->v[2] = seed; st->v[3] = seed - P1;
    st->tail[0] = st->tail[1] = st->tail[2] = st->tail[3] = 0;
    st->tail_len = 0; st->seed = seed; st->total_lo = 0; st->total_hi = 0;
}

// One warp.  Consumes n bytes at p, continuing from *st (src/lz4f.zig:385 XxHash32.update).
__global__ void __launch_bounds__(32) k_xxh32_update(XxhState* st, const uint8_t* __restrict__ p, uint64_t n) {
    __shared__ __align__(16) uint32_t tile[2][TILE_WORDS];
    __shared__ uint8_t tailb[32];
    const uint32_t lane = threadIdx.x;
    uint32_t acc = lane < 4 ? st->v[lane] : 0;
    uint32_t tail_len = st->tail_len;
    uint64_t total = ((uint64_t)st->total_hi << 32) | st->total_lo;
    if (lane < 16) tailb[lane] = (uint8_t)(st->tail[lane >> 2] >> ((lane & 3) * 8));
    __syncwarp();
    total += n;
    // 1) top up a pending partial stripe
    if (tail_len) {
        uint32_t take = 16 - tail_len;
        if ((uint64_t)take > n) take = (uint32_t)n;
        if (lane < take) tailb[tail_len + lane] = __ldg(p + lane);
        __syncwarp();
        tail_len += take; p += take; n -= take;
        if (tail_len == 16) {
            if (lane < 4) {
                uint32_t x = (uint32_t)tailb[4 * lane] | ((uint32_t)tailb[4 * lane + 1] << 8) |
                             ((uint32_t)tailb[4 * lane + 2] << 16) | ((uint32_t)tailb[4 * lane + 3] << 24);
                acc = xround(acc, x);
            }
            tail_len = 0;
        }
        __syncwarp();
    }
    // 2) full stripes, tile by tile
    uint64_t nst = n >> 4;                          // stripes remaining
    const uint32_t bo = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 15);
    const uint4* s16 = reinterpret_cast<const uint4*>(p - bo);
    uint64_t done = 0;                               // stripes consumed
    uint4 r[4];
    auto load_tile = [&](uint64_t first) {          // stripes [first, first+128) -> registers
#pragma unroll
        for (int u = 0; u < 4; u++) {
            uint64_t sidx = first + lane + 32u * u;
            if (sidx < nst) {
                if (bo == 0) r[u] = __ldg(s16 + sidx);
                else r[u] = extract16(__ldg(s16 + sidx), __ldg(s16 + sidx + 1), bo);
            }
        }
    };
    int buf = 0;
    if (nst) load_tile(0);
    while (done < nst) {
        uint32_t cnt = (uint32_t)(nst - done < 128 ? nst - done : 128);
        stage_tile(tile[buf], r, cnt, lane);
        __syncwarp();
        if (done + 128 < nst) load_tile(done + 128);  // prefetch the next tile while the chains run
        if (lane < 4) acc = chain_tile(acc, tile[buf] + lane * ROW, cnt);
        done += cnt;
        buf ^= 1;
    }
    __syncwarp();
    // 3) stash the new tail
    uint32_t rem = (uint32_t)(n & 15);
    const uint8_t* tp = p + (nst << 4);
    if (lane < rem) tailb[tail_len + lane] = __ldg(tp + lane);
    tail_len += rem;
    __syncwarp();
    if (lane < 4) {
        st->v[lane] = acc;
        uint32_t w = 0;
        for (int b = 0; b < 4; b++) {
            uint32_t idx = 4 * lane + b;
            if (idx < tail_len) w |= (uint32_t)tailb[idx] << (8 * b);
        }
        st->tail[lane] = w;
    }
    if (lane == 0) { st->tail_len = tail_len; st->total_lo = (uint32_t)total; st->total_hi = (uint32_t)(total >> 32); }
}

__global__ void k_xxh32_final(const XxhState* st, uint32_t* out) {
    uint64_t total = ((uint64_t)st->total_hi << 32) | st->total_lo;
    uint32_t h;
    if (total >= 16) h = rotl32(st->v[0], 1) + rotl32(st->v[1], 7) + rotl32(st->v[2], 12) + rotl32(st->v[3], 18);
    else h = st->seed + P5;
    h += (uint32_t)total;
    uint32_t n = st->tail_len;
    uint32_t i = 0;
    while (n >= 4) { h = rotl32(h + st->tail[i] * P3, 17) * P4; i++; n -= 4; }
    uint32_t w = i < 4 ? st->tail[i] : 0;
    while (n) { h = rotl32(h + (w & 0xFF) * P5, 11) * P1; w >>= 8; n--; }
    h ^= h >> 15; h *= P2; h ^= h >> 13; h *= P3; h ^= h >> 16;
    *out = h;
}

cudaError_t launch_xxh32_init(XxhState* st, uint32_t seed, cudaStream_t stream) {
    k_xxh32_init<<<1, 1, 0, stream>>>(st, seed);
    count_launch();
    return cudaGetLastError();
}
cudaError_t launch_xxh32_update(XxhState* st, const uint8_t* p, uint64_t n, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    k_xxh32_update<<<1, 32, 0, stream>>>(st, p, n);
    count_launch();
    return cudaGetLastError();
}
cudaError_t launch_xxh32_final(const XxhState* st, uint32_t* out, cudaStream_t stream) {
    k_xxh32_final<<<1, 1, 0, stream>>>(st, out);
    count_launch();
    return cudaGetLastError();
}

}  // namespace b2
