// k_frame.cu — K5 (record-size scan), K6 (frame assembly / compaction), K7 (block-header walk) and the
// decode summary.  These are the device side of lz4f.compressFrame's `dstPos` arithmetic
// (/root/reference/src/lz4f.zig:389,406-427,432-441) and of decompressFrame's header loop (:563-600).
#include "b2_common.cuh"
#include "b2_kernels.h"

namespace b2 {

// ------------------------------------------------------------------ K5: exclusive scan of record sizes ----
// record_i = 4 (block header) + stored_i + (block checksum ? 4 : 0),
// stored_i = csize_i >= len_i ? len_i : csize_i            (src/lz4f.zig:407-408)
// One CTA; each thread owns a contiguous run of blocks.  nblocks is at most a few hundred thousand.
constexpr int SCAN_THREADS = 1024;

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_records(const uint32_t* __restrict__ csize,
                                                               const int32_t* __restrict__ status, uint32_t nblocks,
                                                               uint64_t stride, uint64_t total, uint32_t block_checksum,
                                                               uint64_t* __restrict__ rec_off, FrameTotals* totals) {
    __shared__ uint64_t part[SCAN_THREADS];
    __shared__ uint32_t bad[SCAN_THREADS];
    const uint32_t t = threadIdx.x;
    const uint32_t per = (nblocks + SCAN_THREADS - 1) / SCAN_THREADS;
    const uint32_t lo = t * per < nblocks ? t * per : nblocks;
    const uint32_t hi = lo + per < nblocks ? lo + per : nblocks;
    uint64_t sum = 0;
    uint32_t fb = 0xFFFFFFFFu;
    for (uint32_t i = lo; i < hi; i++) {
        uint64_t o = (uint64_t)i * stride;
        uint64_t len = total - o < stride ? total - o : stride;
        uint32_t c = csize[i];
        uint64_t stored = c >= len ? len : c;
        sum += 4 + stored + (block_checksum ? 4 : 0);
        if (status[i] != 0 && fb == 0xFFFFFFFFu) fb = i;
    }
    part[t] = sum;
    bad[t] = fb;
    __syncthreads();
    // Hillis-Steele inclusive scan over 1024 partials
    for (uint32_t d = 1; d < SCAN_THREADS; d <<= 1) {
        uint64_t v = t >= d ? part[t - d] : 0;
        uint32_t b = t >= d ? bad[t - d] : 0xFFFFFFFFu;
        __syncthreads();
        part[t] += v;
        bad[t] = bad[t] < b ? bad[t] : b;
        __syncthreads();
    }
    uint64_t run = part[t] - sum;  // exclusive prefix of this thread's run
    for (uint32_t i = lo; i < hi; i++) {
        rec_off[i] = run;
        uint64_t o = (uint64_t)i * stride;
        uint64_t len = total - o < stride ? total - o : stride;
        uint32_t c = csize[i];
        uint64_t stored = c >= len ? len : c;
        run += 4 + stored + (block_checksum ? 4 : 0);
    }
    if (t == SCAN_THREADS - 1) {
        rec_off[nblocks] = part[t];
        totals->body_bytes = part[t];
        uint32_t b = bad[t];
        totals->first_bad = b;
        totals->bad_status = b != 0xFFFFFFFFu ? status[b] : 0;
    }
}

cudaError_t launch_scan_records(const uint32_t* csize, const int32_t* status, uint32_t nblocks, uint64_t stride,
                                uint64_t total, uint32_t block_checksum, uint64_t* rec_off, FrameTotals* totals,
                                cudaStream_t stream) {
    k_scan_records<<<1, SCAN_THREADS, 0, stream>>>(csize, status, nblocks, stride, total, block_checksum, rec_off, totals);
    count_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------ K6: assembly ----
// One warp per block record: header word, payload (compressed slot or raw input), optional checksum.
constexpr int ASM_WARPS = 8;

// One CTA per record (was one warp: 58 registers, a quarter of the warps active, every block the latency of one warp's
// loop): 256 threads stream the payload with 16-byte accesses, the source realigned in registers.
__global__ void __launch_bounds__(ASM_WARPS * 32, 8) k_assemble(BlockSet slots, BlockSet raw, const uint32_t* __restrict__ csize,
                                                                const uint32_t* __restrict__ sums,
                                                                const uint64_t* __restrict__ rec_off, uint8_t* __restrict__ body,
                                                                uint32_t nblocks, uint32_t block_checksum) {
    const uint32_t tid = threadIdx.x;
    constexpr uint32_t NT = ASM_WARPS * 32;
    for (uint32_t i = blockIdx.x; i < nblocks; i += gridDim.x) {
        const uint8_t* rp; uint32_t rn;
        raw.get(i, rp, rn);
        const uint32_t c = csize[i];
        const bool store_raw = c >= rn;                                  // src/lz4f.zig:407
        const uint8_t* p = rp; uint32_t n = rn;
        if (!store_raw) { const uint8_t* sp; uint32_t sn; slots.get(i, sp, sn); p = sp; n = c; }
        uint8_t* d = body + rec_off[i];
        const uint32_t hw = n | (store_raw ? 0x80000000u : 0u);          // :411-414
        if (tid < 4) d[tid] = (uint8_t)(hw >> (8 * tid));                // :418
        if (block_checksum && tid >= 32 && tid < 36) d[4 + n + (tid - 32)] = (uint8_t)(sums[i] >> (8 * (tid - 32)));  // :422-427
        d += 4;
        // head bytes up to the first 16-byte boundary of the destination, then vectors, then the tail
        uint32_t head = (uint32_t)((16 - (reinterpret_cast<uintptr_t>(d) & 15)) & 15);
        if (head > n) head = n;
        if (tid < head) d[tid] = __ldg(p + tid);
        const uint8_t* ps = p + head;
        uint8_t* pd = d + head;
        const uint32_t rest = n - head, nvec = rest >> 4;
        const uint32_t bo = (uint32_t)(reinterpret_cast<uintptr_t>(ps) & 15);
        const uint4* s16 = reinterpret_cast<const uint4*>(ps - bo);
        uint4* d16 = reinterpret_cast<uint4*>(pd);
        if (bo == 0) {
#pragma unroll 4
            for (uint32_t v = tid; v < nvec; v += NT) d16[v] = __ldg(s16 + v);
        } else {
#pragma unroll 4
            for (uint32_t v = tid; v < nvec; v += NT) d16[v] = extract16(__ldg(s16 + v), __ldg(s16 + v + 1), bo);
        }
        const uint32_t done = nvec << 4, tail = rest - done;
        if (tid < tail) pd[done + tid] = __ldg(ps + done + tid);
    }
}

cudaError_t launch_assemble(const BlockSet& slots, const BlockSet& raw, const uint32_t* csize, const uint32_t* sums,
                            const uint64_t* rec_off, uint8_t* body, uint32_t nblocks, uint32_t block_checksum,
                            int num_sms, cudaStream_t stream) {
    if (nblocks == 0) return cudaSuccess;
    uint32_t maxg = (uint32_t)num_sms * 8;
    k_assemble<<<nblocks < maxg ? nblocks : maxg, ASM_WARPS * 32, 0, stream>>>(slots, raw, csize, sums, rec_off, body, nblocks,
                                                                            block_checksum);
    count_launch();
    return cudaGetLastError();
}

// end mark (src/lz4f.zig:433) and content checksum (:437-441)
__global__ void k_finalize(uint8_t* frame, uint64_t header_size, const FrameTotals* totals, const uint32_t* content_sum) {
    uint8_t* p = frame + header_size + totals->body_bytes;
    uint32_t t = threadIdx.x;
    if (t < 4) p[t] = 0;
    if (content_sum && t >= 4 && t < 8) p[t] = (uint8_t)(*content_sum >> (8 * (t - 4)));
}

cudaError_t launch_finalize(uint8_t* frame, uint64_t header_size, const FrameTotals* totals, const uint32_t* content_sum,
                            cudaStream_t stream) {
    k_finalize<<<1, 32, 0, stream>>>(frame, header_size, totals, content_sum);
    count_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------ K7: block-header walk ----
// The headers form a linked list (each header tells where the next one is): a single
// thread chases it.  Mirrors the loop control of src/lz4f.zig:563-600 without touching payloads.
__global__ void k_walk(const uint8_t* __restrict__ frame, uint64_t n, uint64_t start, uint32_t block_checksum,
                       uint64_t* __restrict__ off, uint32_t* __restrict__ hdr, uint32_t capacity, WalkResult* res) {
    uint64_t pos = start;
    uint32_t count = 0, terminal = 1, max_stored = 0;
    while (pos < n) {                                                    // :563
        if (pos + 4 > n) { terminal = 2; break; }                        // :565 FrameSizeWrong
        uint32_t h = ldg_u32(frame + pos);
        pos += 4;
        if (h == 0) { terminal = 0; break; }                             // :573 end mark
        uint64_t sz = h & 0x7FFFFFFFu;
        if (pos + sz > n) { terminal = 2; break; }                       // :582
        if (block_checksum && pos + sz + 4 > n) {                        // :591 (checked after the data)
            // the reference fails here with FrameSizeWrong as well; the block is not decoded
            terminal = 2; break;
        }
        if (count < capacity) { off[count] = pos; hdr[count] = h; }
        if ((uint32_t)sz > max_stored) max_stored = (uint32_t)sz;
        count++;
        pos += sz + (block_checksum ? 4 : 0);
    }
    res->nblocks = count;
    res->terminal = terminal;
    res->end_pos = pos;
    res->max_stored = max_stored;
}

cudaError_t launch_walk(const uint8_t* frame, uint64_t n, uint64_t start, uint32_t block_checksum, uint64_t* off,
                        uint32_t* hdr, uint32_t capacity, WalkResult* res, cudaStream_t stream) {
    k_walk<<<1, 1, 0, stream>>>(frame, n, start, block_checksum, off, hdr, capacity, res);
    count_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------ decode summary ----
// First failing block in frame order, checksum failures taking precedence inside a block (the reference
// verifies the block checksum before decoding, src/lz4f.zig:590-614), plus the total decoded size and
// whether the "every non-final block is exactly blockSize" layout assumption held.
__global__ void __launch_bounds__(1024) k_decode_summary(const uint32_t* __restrict__ out_len, const int32_t* __restrict__ status,
                                                         const uint32_t* __restrict__ sums_calc, const uint8_t* __restrict__ frame,
                                                         const uint64_t* __restrict__ off, const uint32_t* __restrict__ hdr,
                                                         uint32_t nblocks, uint32_t block_size, uint32_t block_checksum,
                                                         DecodeSummary* out) {
    __shared__ uint32_t s_bad[1024];
    __shared__ uint32_t s_layout[1024];
    const uint32_t t = threadIdx.x;
    uint32_t fb = 0xFFFFFFFFu, layout_ok = 1;
    for (uint32_t i = t; i < nblocks; i += 1024) {
        bool bad = status[i] != 0;
        if (block_checksum) {
            uint32_t stored = ldg_u32(frame + off[i] + (hdr[i] & 0x7FFFFFFFu));
            if (stored != sums_calc[i]) bad = true;
        }
        if (bad && i < fb) fb = i;
        if (i + 1 < nblocks && out_len[i] != block_size) layout_ok = 0;
    }
    s_bad[t] = fb;
    s_layout[t] = layout_ok;
    __syncthreads();
    for (uint32_t d = 512; d > 0; d >>= 1) {
        if (t < d) {
            s_bad[t] = s_bad[t] < s_bad[t + d] ? s_bad[t] : s_bad[t + d];
            s_layout[t] = s_layout[t] & s_layout[t + d];
        }
        __syncthreads();
    }
    if (t == 0) {
        uint32_t b = s_bad[0];
        out->first_bad = b;
        out->bad_kind = 0;
        out->bad_status = 0;
        if (b != 0xFFFFFFFFu) {
            bool ck_bad = false;
            if (block_checksum) ck_bad = ldg_u32(frame + off[b] + (hdr[b] & 0x7FFFFFFFu)) != sums_calc[b];
            if (ck_bad) out->bad_kind = 1;
            else if (status[b] == ST_RAW_NO_ROOM) out->bad_kind = 3;
            else out->bad_kind = 2;
            out->bad_status = status[b];
        }
        out->layout_ok = s_layout[0];
        out->total = nblocks ? (uint64_t)(nblocks - 1) * block_size + out_len[nblocks - 1] : 0;
    }
}

cudaError_t launch_decode_summary(const uint32_t* out_len, const int32_t* status, const uint32_t* sums_calc,
                                  const uint8_t* frame, const uint64_t* off, const uint32_t* hdr, uint32_t nblocks,
                                  uint32_t block_size, uint32_t block_checksum, DecodeSummary* out, cudaStream_t stream) {
    k_decode_summary<<<1, 1024, 0, stream>>>(out_len, status, sums_calc, frame, off, hdr, nblocks, block_size,
                                             block_checksum, out);
    count_launch();
    return cudaGetLastError();
}

}  // namespace b2
