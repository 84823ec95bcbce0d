// k_decompress.cu — K2: LZ4 block decompressor, many independent blocks, one warp per block.
//
// Semantics: lz4.decompressSafe / decompressSafeUsingDict == decompressGeneric with
// targetOutputSize == dst.len, /root/reference/src/lz4.zig:89-251, including its exact error kinds
// and order of checks, its leniency (no last-5-literals / MFLIMIT end rules) and its two quirks
// (src.len == 0 -> 0, dst.len == 0 -> 0 without error, :97-98).
//
// The token chain is parsed by the whole warp in lock-step (every lane holds ip/op); the byte work is
// spread over the lanes:
//   * 255-run length extensions are scanned 32 bytes at a time with a ballot,
//   * literals are copied with 16-byte aligned stores (b2::warp_copy, read-only source path),
//   * a match copy with offset >= length is a plain vector copy; an overlapping match (offset <
//     length, :235-241) is produced by period doubling: after `offset` bytes are copied the region
//     [match, op+offset) is periodic with period `offset`, so each round copies twice as much from
//     the fixed start of the match — same bytes as the reference's forward byte loop.
//   * matches that start in the external dictionary (:181-228) copy the dictionary tail first, then
//     continue at dst[0..].
#include "b2_common.cuh"
#include "b2_kernels.h"

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

// Forward-overlap-safe copy of `ml` bytes to d from s = d - offset (both inside the output buffer).
__device__ __forceinline__ void match_copy(uint8_t* d, const uint8_t* s, uint32_t offset, uint32_t ml, uint32_t lane) {
    if (offset >= ml) {
        warp_copy<false>(d, s, ml, lane);
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

template <bool WRITE>
__device__ void decode_block(const uint8_t* __restrict__ src, uint32_t n, uint8_t* dst, uint32_t cap,
                             const uint8_t* __restrict__ dict, uint32_t dict_len, bool has_dict, uint32_t lane,
                             uint32_t& olen, int& st) {
    st = ST_OK;
    olen = 0;
    if (n == 0) return;    // :97
    if (cap == 0) return;  // :98
    uint32_t ip = 0, op = 0;
    const uint32_t iend = n, oend = cap;
    for (;;) {
        if (ip >= iend) break;                                          // :113
        const uint32_t token = __ldg(src + ip);
        ip += 1;
        uint32_t LL = token >> 4;
        if (LL == RUN_MASK) {                                           // :123
            if (!read_len_ext(src, ip, iend, LL, lane)) { st = ST_CORRUPTED; return; }
        }
        if (LL > 0) {                                                   // :134
            if ((uint64_t)ip + LL > iend) { st = ST_CORRUPTED; return; }
            if ((uint64_t)op + LL > oend) { st = ST_OUTPUT_TOO_SMALL; return; }
            if (WRITE) warp_copy<true>(dst + op, src + ip, LL, lane);
            ip += LL;
            op += LL;
        }
        if (ip >= iend) break;                                          // :146
        if (ip + 2 > iend) { st = ST_CORRUPTED; return; }               // :149
        const uint32_t offset = (uint32_t)__ldg(src + ip) | ((uint32_t)__ldg(src + ip + 1) << 8);
        ip += 2;
        if (offset == 0) { st = ST_CORRUPTED; return; }                 // :154
        uint32_t ML = token & ML_MASK;
        if (ML == ML_MASK) {                                            // :160
            if (!read_len_ext(src, ip, iend, ML, lane)) { st = ST_CORRUPTED; return; }
        }
        ML += MINMATCH;                                                 // :171
        if ((uint64_t)op + ML > oend) { st = ST_OUTPUT_TOO_SMALL; return; }  // :174
        if (offset > op) {                                              // :181 match starts before dst
            if (!has_dict) { st = ST_CORRUPTED; return; }               // :183-186
            if ((uint64_t)offset > (uint64_t)op + dict_len) { st = ST_CORRUPTED; return; }  // :190
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
                match_copy(dst + op, dst + op - offset, offset, ML, lane);  // :232-246
                __syncwarp();
            }
            op += ML;
        }
    }
    olen = op;
}

__global__ void __launch_bounds__(K2_THREADS) k_decompress(BlockSet in, OutSet out, const uint32_t* __restrict__ hdr,
                                                           uint32_t* __restrict__ out_len, int32_t* __restrict__ status,
                                                           uint32_t nblocks, const uint8_t* __restrict__ dict,
                                                           uint32_t dict_len, int has_dict, uint32_t* ticket) {
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
        uint32_t olen = 0; int st = ST_OK;
        if (hdr && (hdr[blk] & 0x80000000u)) {
            // stored block: reference src/lz4f.zig:603-608
            if (n > cap) st = ST_RAW_NO_ROOM;
            else { warp_copy<true>(dst, src, n, lane); olen = n; }
        } else {
            decode_block<true>(src, n, dst, cap, dict, dict_len, has_dict != 0, lane, olen, st);
        }
        if (lane == 0) {
            out_len[blk] = st == ST_OK ? olen : 0u;
            status[blk] = st;
        }
        __syncwarp();
    }
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
        else decode_block<false>(src, n, nullptr, 0xFFFFFFFFu, nullptr, 0, false, lane, olen, st);
        if (lane == 0) { out_len[blk] = olen; status[blk] = st; }
    }
}

cudaError_t launch_decompress(const BlockSet& in, const OutSet& out, const uint32_t* hdr, uint32_t* out_len,
                              int32_t* status, uint32_t nblocks, const uint8_t* dict, uint32_t dict_len,
                              uint32_t* ticket, int num_sms, cudaStream_t stream) {
    if (nblocks == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(ticket, 0, sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
    uint32_t want = (nblocks + K2_WARPS - 1) / K2_WARPS;
    uint32_t maxg = (uint32_t)(num_sms * 16);  // 16 CTAs x 4 warps = 64 warps / SM
    uint32_t grid = want < maxg ? want : maxg;
    k_decompress<<<grid, K2_THREADS, 0, stream>>>(in, out, hdr, out_len, status, nblocks, dict, dict_len,
                                                  dict != nullptr ? 1 : 0, ticket);
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

}  // namespace b2
