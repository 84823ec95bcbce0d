// k_dest_size.cu — per-block search state of compressDestSize (reference src/lz4.zig:551-616).
//
// The reference finds "the largest prefix of src whose compressed form fits dst" with a bisection over
// prefix lengths, each probe a full compressDefault of that prefix (:575, :594).  The device version keeps
// one search state per block and runs every probe of every block as one launch of the fast compressor
// (K1) with per-block prefix lengths: k_dest_size_init picks the first probe, k_dest_size_step folds the
// probe's outcome into the state and picks the next one.  The probe order — the estimate `dst.len` first
// when `dst.len <= srcSize` (:573-586), then `mid = low + (high - low) / 2` (:590) — is the reference's, so
// the prefix it settles on is the reference's even where compressed size is not monotone in prefix length.
#include "b2_kernels.h"

namespace b2 {

namespace {

constexpr uint32_t DS_DONE = 0, DS_ESTIMATE = 1, DS_BISECT = 2;

__device__ __forceinline__ uint32_t bound_of(uint32_t n) { return n + n / 255 + 16; }  // :80-83, n <= LZ4_MAX_INPUT_SIZE

// :589-591 — next probe of the bisection, or done
__device__ __forceinline__ void next_probe(DestSizeState& s, uint32_t max_len) {
    s.phase = DS_DONE;
    s.cur = 0;
    if (s.low > s.high) return;
    const uint32_t mid = s.low + (s.high - s.low) / 2;
    if (mid == 0 || mid > max_len) return;
    s.cur = mid;
    s.phase = DS_BISECT;
}

__global__ void k_dest_size_init(const uint32_t* __restrict__ src_len, const uint32_t* __restrict__ dst_cap,
                                 DestSizeState* __restrict__ state, uint32_t* __restrict__ probe_len, uint32_t nblocks) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nblocks) return;
    const uint32_t max_len = src_len[i], cap = dst_cap[i];
    DestSizeState s;
    s.low = 1; s.high = max_len; s.best = 0; s.cur = 0; s.phase = DS_DONE; s.err = ST_OK;
    if (max_len == 0) {
        // :553-556
    } else if (max_len > LZ4_MAX_INPUT_SIZE) {
        s.err = ST_INPUT_TOO_LARGE;  // compressBound == 0 -> first branch -> compressDefault fails, :559-561, :296
    } else if (cap >= bound_of(max_len)) {
        s.best = max_len;            // :560-564, the final pass compresses all of it
    } else if (cap <= max_len) {     // :573
        if (cap == 0) {              // an empty prefix "fits" (size 0): bestSize stays 0, low = 1
            next_probe(s, max_len);
        } else {
            s.cur = cap;
            s.phase = DS_ESTIMATE;
        }
    } else {
        next_probe(s, max_len);
    }
    state[i] = s;
    probe_len[i] = s.cur;
}

__global__ void k_dest_size_step(const uint32_t* __restrict__ src_len, const int32_t* __restrict__ probe_status,
                                 DestSizeState* __restrict__ state, uint32_t* __restrict__ probe_len, uint32_t nblocks) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nblocks) return;
    DestSizeState s = state[i];
    if (s.phase == DS_DONE) { probe_len[i] = 0; return; }
    const uint32_t max_len = src_len[i];
    const bool fits = probe_status[i] == ST_OK;  // K1 was given dst's capacity: OK <=> size <= dst.len
    bool done = false;
    if (s.phase == DS_ESTIMATE) {                // :575-585
        if (fits) { s.best = s.cur; s.low = s.cur + 1; }
        else s.high = s.cur - 1;
    } else {                                     // :594-611
        if (fits) {
            s.best = s.cur;
            if (s.cur == max_len) done = true;
            else s.low = s.cur + 1;
        } else {
            s.high = s.cur - 1;
        }
        if (s.low > max_len) done = true;
    }
    if (done) { s.phase = DS_DONE; s.cur = 0; }
    else next_probe(s, max_len);
    state[i] = s;
    probe_len[i] = s.cur;
}

// after the final compression of the chosen prefixes: consumed sizes, and the error of blocks that never ran
__global__ void k_dest_size_finish(const DestSizeState* __restrict__ state, uint32_t* __restrict__ consumed,
                                   uint32_t* __restrict__ out_len, int32_t* __restrict__ status, uint32_t nblocks) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nblocks) return;
    const DestSizeState s = state[i];
    if (s.err != ST_OK) { consumed[i] = 0; out_len[i] = 0; status[i] = s.err; }
    else consumed[i] = s.best;
}

__global__ void k_dest_size_best(const DestSizeState* __restrict__ state, uint32_t* __restrict__ probe_len, uint32_t nblocks) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nblocks) probe_len[i] = state[i].err == ST_OK ? state[i].best : 0u;
}

}  // namespace

cudaError_t launch_dest_size_init(const uint32_t* src_len, const uint32_t* dst_cap, DestSizeState* state,
                                  uint32_t* probe_len, uint32_t nblocks, cudaStream_t stream) {
    if (nblocks == 0) return cudaSuccess;
    k_dest_size_init<<<(nblocks + 255) / 256, 256, 0, stream>>>(src_len, dst_cap, state, probe_len, nblocks);
    count_launch();
    return cudaGetLastError();
}
cudaError_t launch_dest_size_step(const uint32_t* src_len, const int32_t* probe_status, DestSizeState* state,
                                  uint32_t* probe_len, uint32_t nblocks, cudaStream_t stream) {
    if (nblocks == 0) return cudaSuccess;
    k_dest_size_step<<<(nblocks + 255) / 256, 256, 0, stream>>>(src_len, probe_status, state, probe_len, nblocks);
    count_launch();
    return cudaGetLastError();
}
cudaError_t launch_dest_size_best(const DestSizeState* state, uint32_t* probe_len, uint32_t nblocks, cudaStream_t stream) {
    if (nblocks == 0) return cudaSuccess;
    k_dest_size_best<<<(nblocks + 255) / 256, 256, 0, stream>>>(state, probe_len, nblocks);
    count_launch();
    return cudaGetLastError();
}
cudaError_t launch_dest_size_finish(const DestSizeState* state, uint32_t* consumed, uint32_t* out_len, int32_t* status,
                                    uint32_t nblocks, cudaStream_t stream) {
    if (nblocks == 0) return cudaSuccess;
    k_dest_size_finish<<<(nblocks + 255) / 256, 256, 0, stream>>>(state, consumed, out_len, status, nblocks);
    count_launch();
    return cudaGetLastError();
}

}  // namespace b2
