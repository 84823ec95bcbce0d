// k_index.cu — K7': the frame block index built in parallel.
//
// lz4f.decompressFrame finds its blocks by chasing the 4-byte block headers (src/lz4f.zig:563-587): a
// linked list in which every hop is one dependent DRAM access (~0.5 us per hop on the
// device, 8 ms for 1 GiB of 64 KiB blocks).  This file removes the dependency chain:
//
//   1. k_idx_count / k_idx_write — every byte position p >= start is tested as a *possible* block
//      header: h = u32le(p) != 0, size = h & 0x7FFFFFFF <= bound (the frame's block size), and the
//      record p + 4 + size (+4) fits in the frame.  Random payload passes with probability
//      2 * bound / 2^32 (3e-5 for 64 KiB blocks), but payload made of small little-endian integers passes
//      often, so a candidate must also *lead somewhere*: the IDX_HOPS records that follow it must be
//      possible headers too (or the end mark / the end of the frame).  Real records always do; a false
//      one survives with probability density^IDX_HOPS.  Survivors are written in position order
//      (16-bit mask per 16-byte chunk -> tile counts -> scan -> ordered write).
//   2. k_idx_link — each candidate finds the candidate that starts where its record ends (binary
//      search); a record whose successor is not a candidate points to the sentinel.
//   3. k_idx_jump — jump tables J_k[i] = node reached after 2^k hops (pointer doubling, log2 N levels).
//   4. k_idx_emit — the number of records on the chain that starts at `start` is read off the tables
//      (one thread, log2 N steps); record m is start advanced by the bits of m (log2 N steps, all m in
//      parallel).  The position after the last record is classified exactly like the serial walk
//      (end mark / ran off the end / truncated).  Only if that position holds a header that the
//      candidate filter rejected (a stored size above the bound) is the caller told to fall back to
//      the serial walk, which reproduces the reference's acceptance of such frames.
#include "b2_common.cuh"
#include "b2_kernels.h"

namespace b2 {

constexpr int IDX_THREADS = 256;
constexpr uint32_t IDX_TILE = 64u << 10;                    // bytes of frame per CTA
constexpr uint32_t IDX_ITERS = IDX_TILE / (IDX_THREADS * 16);

struct IdxParams {
    const uint8_t* frame;
    uint64_t n;          // frame bytes
    uint64_t start;      // first header position
    uint64_t a0;         // frame address rounded down to 16 (absolute), chunks are relative to it
    uint64_t lead;       // frame - a0 (0..15)
    uint32_t bound;      // largest stored size that the filter accepts
    uint32_t trailer;    // 4 if block checksums follow the data, else 0
};

constexpr int IDX_HOPS = 4;

__device__ __forceinline__ uint32_t rd_u32(const uint8_t* __restrict__ p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
    const uint32_t sh = (uint32_t)(a & 3) * 8;
    const uint32_t lo = __ldg(w);
    const uint32_t hi = sh ? __ldg(w + 1) : 0u;
    return __funnelshift_r(lo, hi, sh);
}

// Does the record chain that starts at the possible header p stay plausible for IDX_HOPS more records?
__device__ __forceinline__ bool leads_somewhere(const IdxParams& P, uint64_t p, uint32_t sz) {
    uint64_t x = p + 4 + sz + P.trailer;
#pragma unroll 1
    for (int hop = 0; hop < IDX_HOPS; hop++) {
        if (x + 4 > P.n) return true;                        // chain leaves the frame: the walk ends there
        const uint32_t h = rd_u32(P.frame + x);
        if (h == 0) return true;                             // end mark
        const uint32_t s2 = h & 0x7FFFFFFFu;
        if (s2 > P.bound || x + 4 + s2 + P.trailer > P.n) return false;
        x += 4 + (uint64_t)s2 + P.trailer;
    }
    return true;
}

// 16-bit mask of the candidate headers among the 16 byte positions of one aligned 16-byte chunk.
__device__ __forceinline__ uint32_t chunk_mask(const IdxParams& P, uint64_t chunk) {
    const uint64_t cbase = chunk * 16;                       // relative to a0
    const uint64_t span = P.lead + P.n;                      // bytes of [a0, frame + n)
    if (cbase >= span) return 0;
    const uint4 A = __ldg(reinterpret_cast<const uint4*>(P.a0 + cbase));
    uint32_t w4 = 0;
    if (cbase + 16 < span) w4 = __ldg(reinterpret_cast<const uint32_t*>(P.a0 + cbase + 16));
    const uint32_t w[5] = {A.x, A.y, A.z, A.w, w4};
    // cheap filter first: a non-zero word whose size field is within the bound (random bytes pass with
    // probability 2 * bound / 2^32); position range, record fit and the hop test only for those
    uint32_t mask = 0;
#pragma unroll
    for (int b = 0; b < 16; b++) {
        const uint32_t h = __funnelshift_r(w[b >> 2], w[(b >> 2) + 1], (b & 3) * 8);
        const bool ok = h != 0 && (h & 0x7FFFFFFFu) <= P.bound;
        mask |= ok ? (1u << b) : 0u;
    }
    uint32_t m = mask;
    while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        const uint32_t h = __funnelshift_r(b < 4 ? w[0] : b < 8 ? w[1] : b < 12 ? w[2] : w[3],
                                           b < 4 ? w[1] : b < 8 ? w[2] : b < 12 ? w[3] : w[4], (b & 3) * 8);
        const uint32_t sz = h & 0x7FFFFFFFu;
        const int64_t p = (int64_t)(cbase + b) - (int64_t)P.lead;   // position in the frame
        const bool ok = p >= (int64_t)P.start && (uint64_t)p + 4 + sz + P.trailer <= P.n &&
                        leads_somewhere(P, (uint64_t)p, sz);
        if (!ok) mask &= ~(1u << b);
    }
    return mask;
}

__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* s_warp, uint32_t& total) {
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(FULL, inc, d);
        if (lane >= (uint32_t)d) inc += t;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < IDX_THREADS / 32; i++) {
        uint32_t c = s_warp[i];
        if ((uint32_t)i < wid) base += c;
        tot += c;
    }
    __syncthreads();
    total = tot;
    return base + inc - v;
}

__global__ void __launch_bounds__(IDX_THREADS) k_idx_count(IdxParams P, uint32_t* __restrict__ tile_count,
                                                           uint16_t* __restrict__ masks) {
    __shared__ uint32_t s_warp[IDX_THREADS / 32];
    const uint64_t chunk0 = (uint64_t)blockIdx.x * (IDX_TILE / 16);
    uint32_t cnt = 0;
#pragma unroll 1
    for (uint32_t it = 0; it < IDX_ITERS; it++) {
        const uint64_t chunk = chunk0 + it * IDX_THREADS + threadIdx.x;
        const uint32_t m = chunk_mask(P, chunk);
        masks[chunk] = (uint16_t)m;
        cnt += __popc(m);
    }
    uint32_t total;
    block_excl_scan(cnt, s_warp, total);
    if (threadIdx.x == 0) tile_count[blockIdx.x] = total;
}

// exclusive scan of the tile counts (one CTA); total -> *n_nodes
__global__ void __launch_bounds__(1024) k_idx_scan(const uint32_t* __restrict__ tile_count, uint64_t* __restrict__ tile_base,
                                                   uint32_t ntiles, uint64_t* __restrict__ n_nodes) {
    __shared__ uint64_t s_part[1024];
    const uint32_t t = threadIdx.x;
    const uint32_t per = (ntiles + 1023) / 1024;
    const uint32_t lo = t * per, hi = lo + per < ntiles ? lo + per : ntiles;
    uint64_t sum = 0;
    for (uint32_t i = lo; i < hi; i++) sum += tile_count[i];
    s_part[t] = sum;
    __syncthreads();
    for (uint32_t d = 1; d < 1024; d <<= 1) {
        uint64_t v = t >= d ? s_part[t - d] : 0;
        __syncthreads();
        s_part[t] += v;
        __syncthreads();
    }
    uint64_t base = s_part[t] - sum;
    for (uint32_t i = lo; i < hi; i++) { tile_base[i] = base; base += tile_count[i]; }
    if (t == 1023) *n_nodes = s_part[1023];
}

__global__ void __launch_bounds__(IDX_THREADS) k_idx_write(IdxParams P, const uint16_t* __restrict__ masks,
                                                           const uint64_t* __restrict__ tile_base,
                                                           uint64_t* __restrict__ pos, uint64_t capacity) {
    __shared__ uint32_t s_warp[IDX_THREADS / 32];
    const uint64_t chunk0 = (uint64_t)blockIdx.x * (IDX_TILE / 16);
    uint64_t base = tile_base[blockIdx.x];
    for (uint32_t it = 0; it < IDX_ITERS; it++) {
        const uint64_t chunk = chunk0 + it * IDX_THREADS + threadIdx.x;
        uint32_t mask = masks[chunk];
        uint32_t total;
        uint64_t at = base + block_excl_scan(__popc(mask), s_warp, total);
        while (mask) {
            const int b = __ffs(mask) - 1;
            mask &= mask - 1;
            if (at < capacity) pos[at] = chunk * 16 + b - P.lead;
            at++;
        }
        base += total;
    }
}

constexpr uint32_t IDX_NONE = 0xFFFFFFFFu;

// index of the candidate at position x, or IDX_NONE
__device__ __forceinline__ uint32_t find_node(const uint64_t* __restrict__ pos, uint32_t n_nodes, uint64_t x) {
    uint32_t lo = 0, hi = n_nodes;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (pos[mid] < x) lo = mid + 1; else hi = mid;
    }
    return (lo < n_nodes && pos[lo] == x) ? lo : IDX_NONE;
}

// J_0: successor of every candidate.  Slot n_nodes is the sentinel (maps to itself).
__global__ void k_idx_link(IdxParams P, const uint64_t* __restrict__ pos, uint32_t n_nodes, uint32_t* __restrict__ J0) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n_nodes) return;
    if (i == n_nodes) { J0[i] = n_nodes; return; }
    const uint64_t p = pos[i];
    const uint32_t sz = rd_u32(P.frame + p) & 0x7FFFFFFFu;
    const uint32_t j = find_node(pos, n_nodes, p + 4 + sz + P.trailer);
    J0[i] = j == IDX_NONE ? n_nodes : j;
}

__global__ void k_idx_jump(const uint32_t* __restrict__ Jprev, uint32_t* __restrict__ Jnext, uint32_t n_slots) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_slots) Jnext[i] = Jprev[Jprev[i]];
}

// Records on the chain from `start`, in order.  One CTA: thread 0 measures the chain, then all threads emit.
__global__ void __launch_bounds__(1024) k_idx_emit(IdxParams P, const uint64_t* __restrict__ pos, uint32_t n_nodes,
                                                   const uint32_t* __restrict__ J, uint32_t levels, uint64_t* __restrict__ off,
                                                   uint32_t* __restrict__ hdr, uint32_t capacity, WalkResult* res) {
    __shared__ uint32_t s_first, s_count, s_last;
    __shared__ uint32_t s_max[32];
    const uint32_t slots = n_nodes + 1;
    if (threadIdx.x == 0) {
        uint32_t first = n_nodes ? find_node(pos, n_nodes, P.start) : IDX_NONE;
        uint32_t count = 0, cur = first;
        if (first != IDX_NONE) {
            count = 1;
            for (int k = (int)levels - 1; k >= 0; k--) {
                uint32_t nx = J[(size_t)k * slots + cur];
                if (nx != n_nodes) { cur = nx; count += 1u << k; }
            }
        }
        s_first = first; s_count = count; s_last = cur;
    }
    __syncthreads();
    const uint32_t first = s_first, count = s_count;
    uint32_t mx = 0;
    for (uint32_t m = threadIdx.x; m < count; m += blockDim.x) {
        uint32_t cur = first;
        for (uint32_t k = 0; k < levels; k++)
            if (m & (1u << k)) cur = J[(size_t)k * slots + cur];
        const uint64_t p = pos[cur];
        const uint32_t h = rd_u32(P.frame + p);
        if (m < capacity) { off[m] = p + 4; hdr[m] = h; }
        const uint32_t sz = h & 0x7FFFFFFFu;
        mx = sz > mx ? sz : mx;
    }
    mx = __reduce_max_sync(FULL, mx);
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < blockDim.x / 32; i++) mx = s_max[i] > mx ? s_max[i] : mx;
        // where the chain stops: same classification as the serial walk (k_walk, src/lz4f.zig:563-591)
        uint64_t p = P.start;
        if (count) {
            const uint64_t lp = pos[s_last];
            p = lp + 4 + (rd_u32(P.frame + lp) & 0x7FFFFFFFu) + P.trailer;
        }
        uint32_t terminal;
        uint64_t end_pos = p;
        if (p >= P.n) terminal = 1;                                   // ran off the end, no end mark
        else if (p + 4 > P.n) terminal = 2;                           // FrameSizeWrong
        else {
            const uint32_t h = rd_u32(P.frame + p);
            end_pos = p + 4;
            if (h == 0) terminal = 0;                                 // end mark
            else {
                const uint64_t sz = h & 0x7FFFFFFFu;
                if (p + 4 + sz > P.n || p + 4 + sz + P.trailer > P.n) terminal = 2;
                else terminal = 3;                                    // a record the filter rejected: serial walk decides
            }
        }
        res->nblocks = count;
        res->terminal = terminal;
        res->end_pos = end_pos;
        res->max_stored = mx;
    }
}

static IdxParams make_params(const uint8_t* frame, uint64_t n, uint64_t start, uint32_t bound, uint32_t block_checksum) {
    IdxParams P;
    P.frame = frame; P.n = n; P.start = start;
    P.a0 = reinterpret_cast<uint64_t>(frame) & ~uint64_t(15);
    P.lead = reinterpret_cast<uint64_t>(frame) - P.a0;
    P.bound = bound; P.trailer = block_checksum ? 4u : 0u;
    return P;
}

uint32_t index_tiles(const uint8_t* frame, uint64_t n) {
    const uint64_t lead = reinterpret_cast<uint64_t>(frame) & 15;
    return (uint32_t)((lead + n + IDX_TILE - 1) / IDX_TILE);
}

cudaError_t launch_index_candidates(const uint8_t* frame, uint64_t n, uint64_t start, uint32_t bound, uint32_t block_checksum,
                                    uint32_t* tile_count, uint64_t* tile_base, uint16_t* masks, uint64_t* pos,
                                    uint64_t capacity, uint64_t* n_nodes, cudaStream_t stream) {
    const IdxParams P = make_params(frame, n, start, bound, block_checksum);
    const uint32_t ntiles = index_tiles(frame, n);
    if (ntiles == 0) return cudaMemsetAsync(n_nodes, 0, sizeof(uint64_t), stream);
    k_idx_count<<<ntiles, IDX_THREADS, 0, stream>>>(P, tile_count, masks);
    count_launch();
    k_idx_scan<<<1, 1024, 0, stream>>>(tile_count, tile_base, ntiles, n_nodes);
    count_launch();
    k_idx_write<<<ntiles, IDX_THREADS, 0, stream>>>(P, masks, tile_base, pos, capacity);
    count_launch();
    return cudaGetLastError();
}

uint32_t index_levels(uint32_t n_nodes) {
    uint32_t levels = 1;
    while ((1ull << levels) <= (uint64_t)n_nodes) levels++;
    return levels;
}

cudaError_t launch_index_resolve(const uint8_t* frame, uint64_t n, uint64_t start, uint32_t bound, uint32_t block_checksum,
                                 const uint64_t* pos, uint32_t n_nodes, uint32_t* jump, uint64_t* off, uint32_t* hdr,
                                 uint32_t capacity, WalkResult* res, cudaStream_t stream) {
    const IdxParams P = make_params(frame, n, start, bound, block_checksum);
    const uint32_t slots = n_nodes + 1;
    const uint32_t levels = index_levels(n_nodes);
    const uint32_t g = (slots + 255) / 256;
    k_idx_link<<<g, 256, 0, stream>>>(P, pos, n_nodes, jump);
    count_launch();
    for (uint32_t k = 1; k < levels; k++) {
        k_idx_jump<<<g, 256, 0, stream>>>(jump + (size_t)(k - 1) * slots, jump + (size_t)k * slots, slots);
        count_launch();
    }
    k_idx_emit<<<1, 1024, 0, stream>>>(P, pos, n_nodes, jump, levels, off, hdr, capacity, res);
    count_launch();
    return cudaGetLastError();
}

}  // namespace b2
