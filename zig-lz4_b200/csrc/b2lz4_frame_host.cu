// b2lz4_frame_host.cu — host-pointer frame entry points (the reference's own calling convention:
// caller-owned host slices, synchronous) and the README streaming trio, layered on the device path.
//   lz4f.compressFrame    /root/reference/src/lz4f.zig:354-446
//   lz4f.decompressFrame  /root/reference/src/lz4f.zig:541-638
//   compressBegin/Update/End + create/freeCompressionContext — README.md:98-122
// The host side only moves bytes (H2D / D2H) and keeps the < blockSize tail of a stream; all codec
// and checksum work runs in the kernels.
#include <algorithm>
#include <cstdlib>
#include <new>
#include <thread>
#include <vector>
#include "b2_host.h"

using namespace b2;

namespace {

static bool block_size_of(uint32_t id, size_t& bs) {
    switch (id) {
        case 0: case 4: bs = 64u << 10; return true;
        case 5: bs = 256u << 10; return true;
        case 6: bs = 1u << 20; return true;
        case 7: bs = 4u << 20; return true;
        default: return false;
    }
}
static size_t compress_bound(size_t n) { return n > LZ4_MAX_INPUT_SIZE ? 0 : n + n / 255 + 16; }

// Input is uploaded and output downloaded in chunks on two copy streams so that PCIe traffic overlaps
// itself in both directions; the codec runs once over the resident buffer.
constexpr size_t COPY_CHUNK = 64u << 20;

// Ordinary (pageable) caller memory goes through the context's pinned bounce rings (b2_mover.h); for a download the
// bytes are in `h` only after c->mover.flush().
static int upload(b2lz4_ctx* c, void* d, const void* h, size_t n, cudaStream_t s) {
    if (n == 0) return B2LZ4_OK;
    if (b2::HostMover::pageable(h)) { B2_CUDA(c->mover.h2d(d, h, n, s, true)); return B2LZ4_OK; }
    const uint8_t* hp = (const uint8_t*)h;
    uint8_t* dp = (uint8_t*)d;
    for (size_t o = 0; o < n; o += COPY_CHUNK) {
        size_t k = std::min(COPY_CHUNK, n - o);
        B2_CUDA(cudaMemcpyAsync(dp + o, hp + o, k, cudaMemcpyHostToDevice, s));
    }
    return B2LZ4_OK;
}
static int download(b2lz4_ctx* c, void* h, const void* d, size_t n, cudaStream_t s) {
    if (n == 0) return B2LZ4_OK;
    if (b2::HostMover::pageable(h)) { B2_CUDA(c->mover.d2h(h, d, n, s, true)); return B2LZ4_OK; }
    uint8_t* hp = (uint8_t*)h;
    const uint8_t* dp = (const uint8_t*)d;
    for (size_t o = 0; o < n; o += COPY_CHUNK) {
        size_t k = std::min(COPY_CHUNK, n - o);
        B2_CUDA(cudaMemcpyAsync(hp + o, dp + o, k, cudaMemcpyDeviceToHost, s));
    }
    return B2LZ4_OK;
}

// ---------------------------------------------------------------- one-shot (unpipelined) host paths
// Whole input up, one device call, whole output down.  Exact reference semantics in every corner; the
// pipelined paths below fall back to these whenever something is unusual.
static int compress_frame_simple(b2lz4_ctx* c, const void* src, size_t n, void* dst, size_t bound, const b2lz4f_prefs* prefs,
                                 size_t* out) {
    cudaStream_t s = c->stream;
    B2_CUDA(c->stage_in[0].ensure(n + 16));
    B2_CUDA(c->stage_out[0].ensure(bound + 16));
    int rc = upload(c, c->stage_in[0].p, src, n, s);
    if (rc) return rc;
    size_t produced = 0;
    rc = b2_compress_dev_impl(c, c->stage_in[0].p, n, c->stage_out[0].p, bound, prefs, &produced, s, false);
    if (rc) return rc;
    rc = download(c, dst, c->stage_out[0].p, produced, s);
    if (rc) { c->mover.abandon(); return rc; }
    B2_CUDA(cudaStreamSynchronize(s));
    B2_CUDA(c->mover.flush());
    *out = produced;
    return B2LZ4_OK;
}

static int decompress_frame_simple(b2lz4_ctx* c, const void* src, size_t n, void* dst, size_t cap, size_t* out) {
    cudaStream_t s = c->stream;
    B2_CUDA(c->stage_in[0].ensure(n + 16));
    B2_CUDA(c->stage_out[0].ensure(cap + 16));
    int rc = upload(c, c->stage_in[0].p, src, n, s);
    if (rc) return rc;
    size_t produced = 0;
    rc = b2_decompress_dev_impl(c, c->stage_in[0].p, n, c->stage_out[0].p, cap, &produced, s);
    if (rc) return rc;
    rc = download(c, dst, c->stage_out[0].p, produced, s);
    if (rc) { c->mover.abandon(); return rc; }
    B2_CUDA(cudaStreamSynchronize(s));
    B2_CUDA(c->mover.flush());
    *out = produced;
    return B2LZ4_OK;
}

// ---------------------------------------------------------------- pipelined host paths
// The frame is cut into chunks of whole blocks.  Three streams work on different chunks at the same
// time: copy_in uploads chunk k+1, the compute stream runs the kernels of chunk k, copy_out downloads
// chunk k-1.  Input and output are double-buffered on the device (2 x chunk), so the workspace no longer
// grows with the frame.  PCIe moves N + C bytes each way per round trip; the kernels hide behind that.
// A chunk must hold enough blocks to fill the GPU (one warp per block, ~4000 warps resident), so chunks
// are sized in blocks: 2048 blocks (three chunks in flight keep the GPU full), at most 256 MiB of raw data; frames whose block size leaves fewer
// than 1024 blocks per chunk (1 MiB / 4 MiB blocks) and frames shorter than three chunks take the
// one-shot path.  b2lz4_debug_tune("pipe_blocks") overrides the block count (tests use it to pipeline small frames).
constexpr size_t PIPE_MAX_CHUNK = 256u << 20;
static size_t pipe_blocks_for(size_t bs, bool* forced) {
    const int forced_blocks = b2::tune().pipe_blocks;
    if (forced_blocks > 0) { *forced = true; return (size_t)forced_blocks; }
    *forced = false;
    size_t blocks = 2048;
    if (blocks * bs > PIPE_MAX_CHUNK) blocks = PIPE_MAX_CHUNK / bs;
    return blocks;
}
static bool pipeline_applies(size_t raw_bytes, size_t bs, size_t* chunk_blocks) {
    if (b2::tune().no_pipeline) return false;
    bool forced;
    const size_t blocks = pipe_blocks_for(bs, &forced);
    *chunk_blocks = blocks;
    if (!forced && blocks < 1024) return false;
    return raw_bytes > 2 * blocks * bs;
}

static inline uint32_t rd32(const uint8_t* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

// First block of every chunk (+ sentinel).  The pipeline's head (nothing to compute until the first upload is
// in) and tail (nothing to overlap the last download with) shrink with the chunk, so decode chunks ramp up
// (1/4, 1/2 of the nominal size) and down again at the end.
static std::vector<size_t> chunk_plan(size_t nb, size_t chunk_blocks, bool ramp_head, bool ramp_tail) {
    std::vector<size_t> first;
    const size_t q = std::max<size_t>(1, chunk_blocks / 4), h = std::max<size_t>(1, chunk_blocks / 2);
    size_t i = 0;
    const bool big = nb >= 4 * chunk_blocks;
    const bool head = ramp_head && big, tl = ramp_tail && big;
    if (head) { first.push_back(0); first.push_back(q); i = q + h; }
    const size_t tail = tl ? q + h : 0;
    while (i + tail < nb) { first.push_back(i); i += std::min(chunk_blocks, nb - tail - i); }
    if (tl) { first.push_back(nb - q - h); first.push_back(nb - q); }
    first.push_back(nb);
    return first;
}

constexpr int PIPE_DEPTH = 3;   // chunks in flight: staging buffers, workspaces and compute streams

static void pipe_drain(b2lz4_ctx* c) {
    c->mover.abandon();
    cudaStreamSynchronize(c->stream);
    for (auto xs : c->x_stream) cudaStreamSynchronize(xs);
    cudaStreamSynchronize(c->copy_in); cudaStreamSynchronize(c->copy_out); cudaStreamSynchronize(c->side);
}

static int compress_frame_pipelined(b2lz4_ctx* c, const uint8_t* src, size_t n, uint8_t* dst, size_t cap,
                                    const b2lz4f_prefs* prefs, size_t bs, size_t chunk_blocks, size_t* out) {
    const bool bc = prefs->block_checksum == 1, cc = prefs->content_checksum == 1;
    const int level = prefs->compression_level;
    const bool page_src = b2::HostMover::pageable(src), page_dst = b2::HostMover::pageable(dst);
    size_t hsize = 0;
    { int rc = b2lz4f_write_frame_header(dst, cap, prefs, &hsize); if (rc) return rc; }      // src/lz4f.zig:369
    const size_t chunk = chunk_blocks * bs;
    // uniform chunks: a compress chunk is only done when its slowest block is (~5 ms), smaller chunks just add such waits
    // (measured: quarter/half-size chunks at the end 27.3 ms against 25.9 ms per GiB; 1024-block chunks 30.4 ms)
    const std::vector<size_t> cfirst = chunk_plan((n + bs - 1) / bs, chunk_blocks, false, false);
    const size_t nchunks = cfirst.size() - 1;
    const size_t rec_bound = 4 + compress_bound(bs) + (bc ? 4 : 0);
    const size_t out_bound = (chunk / bs) * rec_bound;
    // HC keeps its chain tables per resident warp in one shared buffer: its chunks run on one stream
    const int nws = level > 0 ? 1 : PIPE_DEPTH;
    for (int b = 0; b < PIPE_DEPTH; b++) {
        B2_CUDA(c->stage_in[b].ensure(chunk + 16));
        B2_CUDA(c->stage_out[b].ensure(out_bound + 16));
    }
    cudaEvent_t* ev_up = &c->ev_pipe[0];     // [3] chunk uploaded
    cudaEvent_t* ev_done = &c->ev_pipe[3];   // [3] chunk computed (input buffer free, totals on the host)
    cudaEvent_t* ev_down = &c->ev_pipe[6];   // [3] chunk downloaded (output buffer free)
    cudaEvent_t* ev_cc = &c->ev_pipe[9];     // [3] content checksum consumed the chunk
    if (cc) B2_CUDA(launch_xxh32_init(c->d_xxh(), 0, c->side));
    size_t pos = hsize;
    int err = B2LZ4_OK;
    const size_t lead = PIPE_DEPTH - 1;      // chunks enqueued ahead of the one being retired
    for (size_t k = 0; k < nchunks + lead; k++) {
        if (k < nchunks) {                                   // ---- enqueue chunk k
            const int b = (int)(k % PIPE_DEPTH);
            const b2_ws_ref w = c->ws(nws == 1 ? 0 : b);
            cudaStream_t s = w.stream;
            const size_t o = cfirst[k] * bs, len = std::min(cfirst[k + 1] * bs, n) - o;
            if (k >= (size_t)PIPE_DEPTH) {                   // buffers of chunk k - DEPTH must be free again
                B2_CUDA(cudaStreamWaitEvent(c->copy_in, ev_done[b], 0));
                if (cc) B2_CUDA(cudaStreamWaitEvent(c->copy_in, ev_cc[b], 0));
                B2_CUDA(cudaStreamWaitEvent(s, ev_down[b], 0));
            }
            B2_CUDA(c->mover.h2d(c->stage_in[b].p, src + o, len, c->copy_in, page_src));
            B2_CUDA(cudaEventRecord(ev_up[b], c->copy_in));
            B2_CUDA(cudaStreamWaitEvent(s, ev_up[b], 0));
            if (cc) {                                        // serial content chain, chunk after chunk
                B2_CUDA(cudaStreamWaitEvent(c->side, ev_up[b], 0));
                B2_CUDA(launch_xxh32_update(c->d_xxh(), c->stage_in[b].as<uint8_t>(), len, c->side));
                B2_CUDA(cudaEventRecord(ev_cc[b], c->side));
            }
            int rc = b2_enqueue_body(c, w, c->stage_in[b].p, len, bs, level, bc, c->stage_out[b].as<uint8_t>(), s, false);
            if (rc) { err = rc; break; }
            B2_CUDA(cudaMemcpyAsync(&c->h()->pipe_totals[k & 3], w.totals, sizeof(FrameTotals), cudaMemcpyDeviceToHost, s));
            B2_CUDA(cudaEventRecord(ev_done[b], s));
        }
        if (k >= lead) {                                     // ---- retire chunk k - lead: its size is known now
            const size_t kk = k - lead;
            const int b = (int)(kk % PIPE_DEPTH);
            B2_CUDA(cudaEventSynchronize(ev_done[b]));
            const FrameTotals t = c->h()->pipe_totals[kk & 3];
            if (t.first_bad != 0xFFFFFFFFu) {                // mapCompressionError, src/lz4f.zig:144-149
                err = t.bad_status == B2LZ4_ERR_OUTPUT_TOO_SMALL ? B2LZ4F_ERR_DST_MAX_SIZE_TOO_SMALL : B2LZ4F_ERR_GENERIC;
                break;
            }
            B2_CUDA(c->mover.d2h(dst + pos, c->stage_out[b].p, t.body_bytes, c->copy_out, page_dst));
            B2_CUDA(cudaEventRecord(ev_down[b], c->copy_out));
            pos += t.body_bytes;
        }
    }
    if (err) { pipe_drain(c); return err; }
    if (cc) {
        B2_CUDA(launch_xxh32_final(c->d_xxh(), c->d_content_sum(), c->side));
        B2_CUDA(cudaMemcpyAsync(&c->h()->content_sum, c->d_content_sum(), 4, cudaMemcpyDeviceToHost, c->side));
        B2_CUDA(cudaStreamSynchronize(c->side));
    }
    B2_CUDA(cudaStreamSynchronize(c->copy_out));
    B2_CUDA(c->mover.flush());
    dst[pos] = dst[pos + 1] = dst[pos + 2] = dst[pos + 3] = 0;                                  // end mark, :433
    pos += 4;
    if (cc) {                                                                                    // :437-441
        const uint32_t h = c->h()->content_sum;
        dst[pos] = (uint8_t)h; dst[pos + 1] = (uint8_t)(h >> 8); dst[pos + 2] = (uint8_t)(h >> 16); dst[pos + 3] = (uint8_t)(h >> 24);
        pos += 4;
    }
    *out = pos;
    return B2LZ4_OK;
}

// Returns 1 if the frame was decoded (or a CUDA error is reported through *rc_out), 0 if the caller must take
// the one-shot path (anything unusual).
//
// The block index is the header chain read straight from the caller's (host) frame, src/lz4f.zig:563-591 — 16 384
// dependent reads 36 KB apart for a 1 GiB frame of 64 KiB blocks: 1.85 ms of cache and TLB misses on the bench box
// (tools/exp/host_walk_bench.cu), 7 % of the whole call, during which neither the link nor the GPU did anything.  Long
// frames therefore start early: the chain is walked for the first two (ramp-sized) chunks only, those are enqueued, and
// the rest of the chain is walked while they upload and decode.
static int decompress_frame_pipelined(b2lz4_ctx* c, const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out,
                                      int* rc_out) {
    *rc_out = B2LZ4_OK;
    b2lz4f_prefs info; size_t hsize = 0;
    if (b2lz4f_parse_frame_header(src, n, &info, &hsize) != B2LZ4_OK) return 0;
    size_t bs; if (!block_size_of(info.block_size_id, bs)) return 0;
    const bool page_src = b2::HostMover::pageable(src), page_dst = b2::HostMover::pageable(dst);
    const bool bc = info.block_checksum == 1, cc = info.content_checksum == 1;
    const size_t tr = bc ? 4 : 0;
    if (b2::tune().no_pipeline) return 0;
    bool forced;
    const size_t chunk_blocks = pipe_blocks_for(bs, &forced);
    if (!forced && chunk_blocks < 1024) return 0;

    // ---- the header chain, walked on demand
    std::vector<uint64_t> off;
    std::vector<uint32_t> hdr;
    off.reserve(n / 1024 + 16); hdr.reserve(n / 1024 + 16);
    size_t p = hsize;
    bool end_mark = false, malformed = false;
    auto walk_until = [&](size_t want_blocks) {          // stops at the end mark, at anything unusual, or with want_blocks known
        while (!end_mark && !malformed && off.size() < want_blocks) {
            if (p >= n || p + 4 > n) { malformed = true; break; }       // ran off the frame without an end mark
            const uint32_t h = rd32(src + p);
            p += 4;
            if (h == 0) { end_mark = true; break; }
            const size_t sz = h & 0x7FFFFFFFu;
            if (p + sz + tr > n || sz > bs) { malformed = true; break; }
            off.push_back(p); hdr.push_back(h);
            p += sz + tr;
        }
    };
    const size_t q = std::max<size_t>(1, chunk_blocks / 4), hh = std::max<size_t>(1, chunk_blocks / 2);
    walk_until(q + hh);
    if (malformed) return 0;
    // early start only when the frame surely has the >= 4 chunks the ramped plan needs (judged by the records seen so far)
    bool early = false;
    if (!end_mark && off.size() == q + hh) {
        const double avg = (double)(p - hsize) / (double)(q + hh);
        early = (double)(n - hsize) / avg >= 4.4 * (double)chunk_blocks && cap / bs + 2 < 0x7FFFFFFFull;
    }
    if (!early) {
        walk_until(SIZE_MAX);
        if (malformed || !end_mark) return 0;
    }
    auto frame_ok = [&]() {                              // the checks that need the whole chain
        if (malformed || !end_mark || off.empty() || off.size() > 0x7FFFFFFFull) return false;
        if (cc && p + 4 > n) return false;
        if (cap < (off.size() - 1) * bs + 1) return false;            // optimistic layout needs room for every block start
        return off.size() * bs > 2 * chunk_blocks * bs;                // (pipeline_applies)
    };
    if (!early && !frame_ok()) return 0;

    // ---- chunk plan: first block of each chunk (+ sentinel).  With an early start only its head is known yet.
    std::vector<size_t> cfirst;
    if (early) { cfirst = {0, q, q + hh}; }
    else cfirst = chunk_plan(off.size(), chunk_blocks, true, true);
    size_t nchunks = early ? SIZE_MAX : cfirst.size() - 1;          // SIZE_MAX: not known yet
    // ---- buffers: exact sizes when the whole chain is known, bounds otherwise
    const size_t nb_bound = early ? cap / bs + 2 : off.size();
    size_t max_in = 0;
    if (early) max_in = chunk_blocks * (4 + bs + tr);
    else
        for (size_t k = 0; k + 1 < cfirst.size(); k++) {
            const size_t i0 = cfirst[k], i1 = cfirst[k + 1];
            max_in = std::max(max_in, (size_t)(off[i1 - 1] + (hdr[i1 - 1] & 0x7FFFFFFFu) + tr - (off[i0] - 4)));
        }
    auto fail = [&](cudaError_t e, const char* what) { set_cuda_error(e, what); *rc_out = B2LZ4_ERR_CUDA; pipe_drain(c); return 1; };
#define PIPE_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return fail(_e, #expr); } while (0)
    for (int b = 0; b < PIPE_DEPTH; b++) {
        PIPE_CUDA(c->stage_in[b].ensure(max_in + 32));
        PIPE_CUDA(c->stage_out[b].ensure(chunk_blocks * bs + 16));
    }
    PIPE_CUDA(c->walk_off.ensure(nb_bound * 8));
    PIPE_CUDA(c->walk_hdr.ensure(nb_bound * 4));
    PIPE_CUDA(c->out_len.ensure(nb_bound * 4 + 4));
    PIPE_CUDA(c->status.ensure(nb_bound * 4 + 4));
    PIPE_CUDA(c->sums.ensure(nb_bound * 4 + 4));
    // per-chunk relative offsets (each chunk's bytes land at the start of its staging buffer), uploaded as far as known
    std::vector<uint64_t> rel;
    size_t rel_done = 0, chunks_rel = 0;                             // blocks / chunks whose index entries are on the device
    auto upload_index = [&]() -> cudaError_t {
        const size_t known_chunks = cfirst.size() - 1;
        rel.resize(cfirst[known_chunks]);
        for (size_t k = chunks_rel; k < known_chunks; k++) {
            const uint64_t lo = off[cfirst[k]] - 4;
            for (size_t i = cfirst[k]; i < cfirst[k + 1]; i++) rel[i] = off[i] - lo;
        }
        const size_t upto = cfirst[known_chunks];
        cudaError_t e = cudaSuccess;
        if (upto > rel_done) {
            e = cudaMemcpyAsync(c->walk_off.as<uint64_t>() + rel_done, rel.data() + rel_done, (upto - rel_done) * 8,
                                cudaMemcpyHostToDevice, c->copy_in);
            if (e == cudaSuccess)
                e = cudaMemcpyAsync(c->walk_hdr.as<uint32_t>() + rel_done, hdr.data() + rel_done, (upto - rel_done) * 4,
                                    cudaMemcpyHostToDevice, c->copy_in);
        }
        rel_done = upto; chunks_rel = known_chunks;
        return e;
    };
    PIPE_CUDA(upload_index());
    cudaEvent_t* ev_up = &c->ev_pipe[0];
    cudaEvent_t* ev_done = &c->ev_pipe[3];
    cudaEvent_t* ev_down = &c->ev_pipe[6];
    cudaEvent_t* ev_cc = &c->ev_pipe[9];
    if (cc) PIPE_CUDA(launch_xxh32_init(c->d_xxh(), 0, c->side));
    bool unusual = false;
    size_t total = 0;
    const size_t lead = PIPE_DEPTH - 1;
    for (size_t k = 0; (nchunks == SIZE_MAX || k < nchunks + lead) && !unusual; k++) {
        if (nchunks == SIZE_MAX && k + 1 >= cfirst.size()) {
            // the chunks known so far are enqueued: walk the rest of the chain under their uploads, finish the plan
            walk_until(SIZE_MAX);
            if (!frame_ok() || off.size() > nb_bound) { unusual = true; break; }
            const size_t nb = off.size();
            size_t i = q + hh;
            const bool tl = nb >= 4 * chunk_blocks && nb > i + q + hh;        // ramp down at the end, as chunk_plan does
            const size_t tail = tl ? q + hh : 0;
            while (i + tail < nb) { cfirst.push_back(std::min(i + chunk_blocks, nb - tail)); i = cfirst.back(); }
            if (tl) { cfirst.push_back(nb - q); cfirst.push_back(nb); }
            if (cfirst.back() != nb) cfirst.push_back(nb);
            nchunks = cfirst.size() - 1;
            PIPE_CUDA(upload_index());
        }
        if (k < nchunks) {
            const int b = (int)(k % PIPE_DEPTH);
            const b2_ws_ref w = c->ws(b);
            cudaStream_t s = w.stream;
            const size_t i0 = cfirst[k], i1 = cfirst[k + 1], cnt = i1 - i0;
            const uint64_t lo = off[i0] - 4;
            const size_t in_len = (size_t)(off[i1 - 1] + (hdr[i1 - 1] & 0x7FFFFFFFu) + tr - lo);
            if (k >= (size_t)PIPE_DEPTH) {
                PIPE_CUDA(cudaStreamWaitEvent(c->copy_in, ev_done[b], 0));
                PIPE_CUDA(cudaStreamWaitEvent(s, ev_down[b], 0));
                if (cc) PIPE_CUDA(cudaStreamWaitEvent(s, ev_cc[b], 0));
            }
            PIPE_CUDA(c->mover.h2d(c->stage_in[b].p, src + lo, in_len, c->copy_in, page_src));
            PIPE_CUDA(cudaEventRecord(ev_up[b], c->copy_in));                 // also covers the index upload (same stream)
            PIPE_CUDA(cudaStreamWaitEvent(s, ev_up[b], 0));
            const uint64_t* d_off = c->walk_off.as<uint64_t>() + i0;
            const uint32_t* d_hdr = c->walk_hdr.as<uint32_t>() + i0;
            uint32_t* d_len = c->out_len.as<uint32_t>() + i0;
            int32_t* d_st = c->status.as<int32_t>() + i0;
            uint32_t* d_sum = c->sums.as<uint32_t>() + i0;
            const uint8_t* d_in = c->stage_in[b].as<uint8_t>();
            if (bc) PIPE_CUDA(launch_xxh32_ranges(d_in, d_off, d_hdr, d_sum, (uint32_t)cnt, s));
            BlockSet in; in.base = d_in; in.off = d_off; in.len = d_hdr; in.stride = 0; in.total = 0; in.len_mask = 0x7FFFFFFFu;
            const uint64_t room = cap > (uint64_t)i0 * bs ? cap - (uint64_t)i0 * bs : 0;
            const uint64_t out_room = std::min<uint64_t>((uint64_t)cnt * bs, room);
            OutSet o; o.base = c->stage_out[b].as<uint8_t>(); o.off = nullptr; o.cap = nullptr; o.stride = bs; o.total = out_room;
            o.slot_cap = (uint32_t)bs;
            PIPE_CUDA(launch_decompress(in, o, d_hdr, d_len, d_st, (uint32_t)cnt, nullptr, 0, w.ticket, c->num_sms, s));
            PIPE_CUDA(launch_decode_summary(d_len, d_st, d_sum, d_in, d_off, d_hdr, (uint32_t)cnt, (uint32_t)bs, bc ? 1 : 0,
                                            w.summary, s));
            PIPE_CUDA(cudaMemcpyAsync(&c->h()->pipe_summary[k & 3], w.summary, sizeof(DecodeSummary), cudaMemcpyDeviceToHost, s));
            PIPE_CUDA(cudaEventRecord(ev_done[b], s));
        }
        if (k >= lead && nchunks != SIZE_MAX) {
            const size_t kk = k - lead;
            const int b = (int)(kk % PIPE_DEPTH);
            PIPE_CUDA(cudaEventSynchronize(ev_done[b]));
            const DecodeSummary sm = c->h()->pipe_summary[kk & 3];
            const bool last = kk + 1 == nchunks;
            const size_t cnt = cfirst[kk + 1] - cfirst[kk];
            // every block of the chunk must be regular: no error, and exactly bs bytes unless it is the frame's last block
            if (sm.first_bad != 0xFFFFFFFFu || !sm.layout_ok || (!last && sm.total != (uint64_t)cnt * bs)) { unusual = true; break; }
            if (cfirst[kk] * bs + sm.total > cap) { unusual = true; break; }
            if (cc) {
                PIPE_CUDA(cudaStreamWaitEvent(c->side, ev_done[b], 0));
                PIPE_CUDA(launch_xxh32_update(c->d_xxh(), c->stage_out[b].as<uint8_t>(), sm.total, c->side));
                PIPE_CUDA(cudaEventRecord(ev_cc[b], c->side));
            }
            PIPE_CUDA(c->mover.d2h(dst + cfirst[kk] * bs, c->stage_out[b].p, sm.total, c->copy_out, page_dst));
            PIPE_CUDA(cudaEventRecord(ev_down[b], c->copy_out));
            total = cfirst[kk] * bs + sm.total;
        }
    }
    if (unusual) { pipe_drain(c); return 0; }
    if (cc) {
        PIPE_CUDA(launch_xxh32_final(c->d_xxh(), c->d_content_sum(), c->side));
        PIPE_CUDA(cudaMemcpyAsync(&c->h()->content_sum, c->d_content_sum(), 4, cudaMemcpyDeviceToHost, c->side));
        PIPE_CUDA(cudaStreamSynchronize(c->side));
    }
    PIPE_CUDA(cudaStreamSynchronize(c->copy_out));
    PIPE_CUDA(c->mover.flush());
    pipe_drain(c);
#undef PIPE_CUDA
    if (cc && rd32(src + p) != c->h()->content_sum) { *rc_out = B2LZ4F_ERR_CONTENT_CHECKSUM_INVALID; return 1; }   // :625-635
    *out = total;
    return 1;
}

}  // namespace

extern "C" {

int b2lz4f_compress_frame_ctx(b2lz4_ctx* c, const void* src, size_t n, void* dst, size_t cap, const b2lz4f_prefs* prefs,
                              size_t* out) {
    if (!c || !out) return B2LZ4F_ERR_PARAMETER_NULL;
    *out = 0;
    b2lz4f_prefs d; if (!prefs) { b2lz4f_prefs_init(&d); prefs = &d; }
    const size_t bound = b2lz4f_compress_frame_bound(n, prefs);
    if (cap < bound) return B2LZ4F_ERR_DST_MAX_SIZE_TOO_SMALL;   // src/lz4f.zig:363-366
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    b2::wait_previous(c);
    B2_CUDA(cudaSetDevice(c->device));
    size_t bs;
    const int level = prefs->compression_level;
    size_t chunk_blocks;
    if (block_size_of(prefs->block_size_id, bs) && (level <= 0 || b2_hc_supported(level)) && pipeline_applies(n, bs, &chunk_blocks))
        return compress_frame_pipelined(c, (const uint8_t*)src, n, (uint8_t*)dst, cap, prefs, bs, chunk_blocks, out);
    return compress_frame_simple(c, src, n, dst, bound, prefs, out);
}

int b2lz4f_decompress_frame_ctx(b2lz4_ctx* c, const void* src, size_t n, void* dst, size_t cap, size_t* out) {
    if (!c || !out) return B2LZ4F_ERR_PARAMETER_NULL;
    *out = 0;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    b2::wait_previous(c);
    B2_CUDA(cudaSetDevice(c->device));
    if (n > (1u << 20) && src && dst) {
        int rc = B2LZ4_OK;
        if (decompress_frame_pipelined(c, (const uint8_t*)src, n, (uint8_t*)dst, cap, out, &rc)) return rc;
        *out = 0;
    }
    return decompress_frame_simple(c, src, n, dst, cap, out);
}

int b2lz4f_compress_frame(const void* src, size_t n, void* dst, size_t cap, const b2lz4f_prefs* prefs, size_t* out) {
    b2lz4_ctx* c; int rc = b2_default_ctx(&c); if (rc) return rc;
    return b2lz4f_compress_frame_ctx(c, src, n, dst, cap, prefs, out);
}
int b2lz4f_decompress_frame(const void* src, size_t n, void* dst, size_t cap, size_t* out) {
    b2lz4_ctx* c; int rc = b2_default_ctx(&c); if (rc) return rc;
    return b2lz4f_decompress_frame_ctx(c, src, n, dst, cap, out);
}

}  // extern "C"

// ================================================================ one process, several GPUs
// SURVEY §8b asks for "multi-GPU variants taking ngpus": the same host-slice calls, the frame sharded by contiguous
// block range over devices 0..ngpus-1 (§8e), one host thread and one context per device, each range moving over its
// own GPU's PCIe link.  Sizes of the compressed ranges are only known after the codec has run, so compression is two
// phases (upload + codec on every GPU, then download of every body to its final offset); decoding knows every offset
// in advance.  Anything unusual (short frame, foreign layout, any error) takes the single-GPU path, which is exact
// in every corner.
namespace {

std::mutex g_mgpu_mu;
std::mutex g_mgpu_call_mu;            // one multi-GPU call at a time: its phases leave results in the per-device contexts
std::vector<b2lz4_ctx*> g_mgpu_ctx;   // one context per device, created on first use, kept for the process

static int mgpu_ctx(int dev, b2lz4_ctx** out) {
    std::lock_guard<std::mutex> lk(g_mgpu_mu);
    if ((int)g_mgpu_ctx.size() <= dev) g_mgpu_ctx.resize(dev + 1, nullptr);
    if (!g_mgpu_ctx[dev]) { int rc = b2lz4_ctx_create(dev, &g_mgpu_ctx[dev]); if (rc) return rc; }
    *out = g_mgpu_ctx[dev];
    return B2LZ4_OK;
}

static int mgpu_count(int ngpus) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count < 1) return 0;
    return std::max(1, std::min(ngpus, count));
}

struct MgpuJob {
    int dev = 0;
    b2lz4_ctx* c = nullptr;
    size_t in_lo = 0, in_len = 0;     // bytes of the caller's source this device reads
    size_t out_off = 0, out_cap = 0;  // where its result goes / how much room it has
    size_t produced = 0;
    size_t blocks = 0;
    int rc = B2LZ4_OK;
};

template <typename F>
static void mgpu_run(std::vector<MgpuJob>& jobs, F fn) {
    std::vector<std::thread> th;
    th.reserve(jobs.size());
    for (auto& j : jobs) th.emplace_back([&j, fn]() { j.rc = fn(j); });
    for (auto& t : th) t.join();
}

static int first_error(const std::vector<MgpuJob>& jobs) {
    for (const auto& j : jobs) if (j.rc) return j.rc;
    return B2LZ4_OK;
}

// content checksum over the device-resident ranges, in order (one serial chain)
static int mgpu_content_sum(std::vector<MgpuJob>& jobs, bool over_output, uint32_t* sum) {
    b2lz4_xxh32_state st;
    b2lz4_xxh32_state_init(&st, 0);
    for (auto& j : jobs) {
        const size_t len = over_output ? j.produced : j.in_len;
        if (!len) continue;
        const void* p = over_output ? j.c->stage_out[0].p : j.c->stage_in[0].p;
        int rc = b2lz4_xxh32_state_update_dev(j.c, &st, p, len, nullptr);
        if (rc) return rc;
    }
    *sum = b2lz4_xxh32_state_final(&st);
    return B2LZ4_OK;
}

}  // namespace

extern "C" {

int b2lz4f_compress_frame_mgpu(const void* src, size_t n, void* dst, size_t cap, const b2lz4f_prefs* prefs, int ngpus,
                               size_t* out) {
    if (!out) return B2LZ4F_ERR_PARAMETER_NULL;
    *out = 0;
    b2lz4f_prefs d; if (!prefs) { b2lz4f_prefs_init(&d); prefs = &d; }
    const int G = mgpu_count(ngpus);
    size_t bs;
    if (G <= 1 || !block_size_of(prefs->block_size_id, bs) || (prefs->compression_level > 0 && !b2_hc_supported(prefs->compression_level)))
        return b2lz4f_compress_frame(src, n, dst, cap, prefs, out);
    const size_t nb = (n + bs - 1) / bs;
    if (nb < (size_t)(4 * G)) return b2lz4f_compress_frame(src, n, dst, cap, prefs, out);
    const size_t bound = b2lz4f_compress_frame_bound(n, prefs);
    if (cap < bound) return B2LZ4F_ERR_DST_MAX_SIZE_TOO_SMALL;   // src/lz4f.zig:363-366
    const bool bc = prefs->block_checksum == 1, cc = prefs->content_checksum == 1;
    size_t hsize = 0;
    { int rc = b2lz4f_write_frame_header(dst, cap, prefs, &hsize); if (rc) return rc; }      // :369
    std::lock_guard<std::mutex> call_lock(g_mgpu_call_mu);
    std::vector<MgpuJob> jobs(G);
    for (int g = 0; g < G; g++) {
        const size_t b0 = (size_t)g * nb / G, b1 = (size_t)(g + 1) * nb / G;
        jobs[g].dev = g;
        jobs[g].in_lo = b0 * bs;
        jobs[g].in_len = std::min(b1 * bs, n) - b0 * bs;
        jobs[g].blocks = b1 - b0;
        jobs[g].out_cap = (b1 - b0) * (4 + compress_bound(bs) + (bc ? 4 : 0));
    }
    // phase 1: upload + codec; every body stays on its device
    mgpu_run(jobs, [&](MgpuJob& j) -> int {
        int rc = mgpu_ctx(j.dev, &j.c); if (rc) return rc;
        b2lz4_ctx* c = j.c;
        std::lock_guard<std::recursive_mutex> lk(c->mu);
        b2::wait_previous(c);
        B2_CUDA(cudaSetDevice(c->device));
        B2_CUDA(c->stage_in[0].ensure(j.in_len + 16));
        B2_CUDA(c->stage_out[0].ensure(j.out_cap + 16));
        rc = upload(c, c->stage_in[0].p, (const uint8_t*)src + j.in_lo, j.in_len, c->stream); if (rc) return rc;
        rc = b2_compress_dev_impl(c, c->stage_in[0].p, j.in_len, c->stage_out[0].p, j.out_cap, prefs, &j.produced, c->stream, true);
        if (rc) return rc;
        B2_CUDA(cudaStreamSynchronize(c->stream));
        return B2LZ4_OK;
    });
    { int rc = first_error(jobs); if (rc) return rc; }
    size_t pos = hsize;
    for (auto& j : jobs) { j.out_off = pos; pos += j.produced; }
    uint32_t csum = 0;
    if (cc) { int rc = mgpu_content_sum(jobs, false, &csum); if (rc) return rc; }
    // phase 2: every body to its place in the frame
    mgpu_run(jobs, [&](MgpuJob& j) -> int {
        b2lz4_ctx* c = j.c;
        std::lock_guard<std::recursive_mutex> lk(c->mu);
        b2::wait_previous(c);
        B2_CUDA(cudaSetDevice(c->device));
        int rc = download(c, (uint8_t*)dst + j.out_off, c->stage_out[0].p, j.produced, c->stream); if (rc) return rc;
        B2_CUDA(cudaStreamSynchronize(c->stream));
        B2_CUDA(c->mover.flush());
        return B2LZ4_OK;
    });
    { int rc = first_error(jobs); if (rc) return rc; }
    uint8_t* o = (uint8_t*)dst + pos;
    o[0] = o[1] = o[2] = o[3] = 0;                                                            // end mark, :433
    pos += 4;
    if (cc) { o[4] = (uint8_t)csum; o[5] = (uint8_t)(csum >> 8); o[6] = (uint8_t)(csum >> 16); o[7] = (uint8_t)(csum >> 24); pos += 4; }
    *out = pos;
    return B2LZ4_OK;
}

int b2lz4f_decompress_frame_mgpu(const void* srcv, size_t n, void* dst, size_t cap, int ngpus, size_t* out) {
    if (!out) return B2LZ4F_ERR_PARAMETER_NULL;
    *out = 0;
    const uint8_t* src = (const uint8_t*)srcv;
    const int G = mgpu_count(ngpus);
    auto single = [&]() { *out = 0; return b2lz4f_decompress_frame(srcv, n, dst, cap, out); };
    if (G <= 1 || !src || !dst) return single();
    b2lz4f_prefs info; size_t hsize = 0;
    if (b2lz4f_parse_frame_header(src, n, &info, &hsize) != B2LZ4_OK) return single();
    size_t bs; if (!block_size_of(info.block_size_id, bs)) return single();
    const bool bc = info.block_checksum == 1, cc = info.content_checksum == 1;
    const size_t tr = bc ? 4 : 0;
    // the header chain, read from the caller's frame (src/lz4f.zig:563-591)
    std::vector<uint64_t> off;
    size_t p = hsize, end_mark_pos = 0;
    bool end_mark = false;
    while (p < n) {
        if (p + 4 > n) return single();
        const uint32_t h = rd32(src + p);
        if (h == 0) { end_mark = true; end_mark_pos = p; p += 4; break; }
        p += 4;
        const size_t sz = h & 0x7FFFFFFFu;
        if (p + sz + tr > n || sz > bs) return single();
        off.push_back(p);
        p += sz + tr;
    }
    const size_t nb = off.size();
    if (!end_mark || nb < (size_t)(4 * G) || (cc && p + 4 > n)) return single();
    if (cap < (nb - 1) * bs + 1) return single();
    std::unique_lock<std::mutex> call_lock(g_mgpu_call_mu);
    std::vector<MgpuJob> jobs(G);
    for (int g = 0; g < G; g++) {
        const size_t b0 = (size_t)g * nb / G, b1 = (size_t)(g + 1) * nb / G;
        jobs[g].dev = g;
        jobs[g].in_lo = off[b0] - 4;
        jobs[g].in_len = (b1 < nb ? off[b1] - 4 : end_mark_pos) - jobs[g].in_lo;
        jobs[g].blocks = b1 - b0;
        jobs[g].out_off = b0 * bs;
        jobs[g].out_cap = std::min((b1 - b0) * bs, cap - b0 * bs);
    }
    mgpu_run(jobs, [&](MgpuJob& j) -> int {
        int rc = mgpu_ctx(j.dev, &j.c); if (rc) return rc;
        b2lz4_ctx* c = j.c;
        std::lock_guard<std::recursive_mutex> lk(c->mu);
        b2::wait_previous(c);
        B2_CUDA(cudaSetDevice(c->device));
        B2_CUDA(c->stage_in[0].ensure(j.in_len + 32));
        B2_CUDA(c->stage_out[0].ensure(j.out_cap + 16));
        rc = upload(c, c->stage_in[0].p, src + j.in_lo, j.in_len, c->stream); if (rc) return rc;
        rc = b2lz4f_decompress_blocks_dev(c, c->stage_in[0].p, j.in_len, c->stage_out[0].p, j.out_cap, bs, bc ? 1 : 0, &j.produced,
                                          c->stream);
        if (rc) return rc;
        // every range but the last must fill its blocks exactly for the offsets to hold; the caller checks
        rc = download(c, (uint8_t*)dst + j.out_off, c->stage_out[0].p, j.produced, c->stream); if (rc) return rc;
        B2_CUDA(cudaStreamSynchronize(c->stream));
        B2_CUDA(c->mover.flush());
        return B2LZ4_OK;
    });
    bool usual = first_error(jobs) == B2LZ4_OK;
    for (int g = 0; usual && g + 1 < G; g++) usual = jobs[g].produced == jobs[g].blocks * bs;
    if (!usual) { call_lock.unlock(); return single(); }   // exact error kind / foreign layout: the serial-order semantics of the one-GPU path
    if (cc) {                      // src/lz4f.zig:625-635
        uint32_t csum = 0;
        int rc = mgpu_content_sum(jobs, true, &csum); if (rc) return rc;
        if (rd32(src + p) != csum) return B2LZ4F_ERR_CONTENT_CHECKSUM_INVALID;
    }
    *out = jobs[G - 1].out_off + jobs[G - 1].produced;
    return B2LZ4_OK;
}

}  // extern "C"

// ================================================================ streaming trio
struct b2lz4f_cctx {
    b2lz4_ctx* ctx = nullptr;        // default context (not owned)
    b2lz4f_prefs prefs{};
    bool begun = false;
    size_t bs = 65536;
    std::vector<uint8_t> pending;    // < bs bytes not yet emitted as a block
    b2lz4_xxh32_state content{};
};

namespace {

// Emits the blocks of `data` (device-resident, n bytes) as frame records into dst (host).
static int emit_blocks(b2lz4f_cctx* cc, const uint8_t* h_a, size_t na, const uint8_t* h_b, size_t nb_, uint8_t* dst, size_t cap,
                       size_t* out) {
    b2lz4_ctx* c = cc->ctx;
    const size_t n = na + nb_;
    *out = 0;
    if (n == 0) return B2LZ4_OK;
    const size_t nblocks = (n + cc->bs - 1) / cc->bs;
    const size_t need = nblocks * (4 + compress_bound(cc->bs) + (cc->prefs.block_checksum == 1 ? 4 : 0));
    if (cap < need) return B2LZ4F_ERR_DST_MAX_SIZE_TOO_SMALL;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    b2::wait_previous(c);
    B2_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    B2_CUDA(c->stage_in[0].ensure(n + 16));
    B2_CUDA(c->stage_out[0].ensure(need + 16));
    uint8_t* d_in = c->stage_in[0].as<uint8_t>();
    if (na) B2_CUDA(cudaMemcpyAsync(d_in, h_a, na, cudaMemcpyHostToDevice, s));
    if (nb_) { int rc = upload(c, d_in + na, h_b, nb_, s); if (rc) return rc; }
    if (na) B2_CUDA(cudaStreamSynchronize(s));  // h_a is the cctx's pending buffer, about to be reused
    if (cc->prefs.content_checksum == 1) {
        int rc = b2lz4_xxh32_state_update_dev(c, &cc->content, d_in, n, s);
        if (rc) return rc;
    }
    size_t produced = 0;
    int rc = b2_compress_dev_impl(c, d_in, n, c->stage_out[0].p, need, &cc->prefs, &produced, s, true);
    if (rc) return rc;
    rc = download(c, dst, c->stage_out[0].p, produced, s);
    if (rc) return rc;
    B2_CUDA(cudaStreamSynchronize(s));
    B2_CUDA(c->mover.flush());
    *out = produced;
    return B2LZ4_OK;
}

}  // namespace

extern "C" {

int b2lz4f_create_compression_context(b2lz4f_cctx** out) {
    if (!out) return B2LZ4F_ERR_PARAMETER_NULL;
    *out = nullptr;
    b2lz4_ctx* c; int rc = b2_default_ctx(&c); if (rc) return rc;
    b2lz4f_cctx* cc = new (std::nothrow) b2lz4f_cctx();
    if (!cc) return B2LZ4F_ERR_ALLOCATION_FAILED;
    cc->ctx = c;
    *out = cc;
    return B2LZ4_OK;
}
void b2lz4f_free_compression_context(b2lz4f_cctx* cc) { delete cc; }

size_t b2lz4f_compress_bound(size_t src_size, const b2lz4f_prefs* prefs) {
    b2lz4f_prefs d; if (!prefs) { b2lz4f_prefs_init(&d); prefs = &d; }
    size_t bs; if (!block_size_of(prefs->block_size_id, bs)) bs = 65536;
    // update(): up to one block completed from the buffered tail plus the blocks of src; end(): the
    // last short block, the end mark and the content checksum
    size_t nb = src_size / bs + 2;
    return nb * (4 + compress_bound(bs) + 4) + 8;
}

int b2lz4f_compress_begin(b2lz4f_cctx* cc, void* dst, size_t cap, const b2lz4f_prefs* prefs, size_t* out) {
    if (!cc || !out) return B2LZ4F_ERR_PARAMETER_NULL;
    b2lz4f_prefs d; if (!prefs) { b2lz4f_prefs_init(&d); prefs = &d; }
    cc->prefs = *prefs;
    if (!block_size_of(prefs->block_size_id, cc->bs)) return B2LZ4F_ERR_MAX_BLOCK_SIZE_INVALID;
    int rc = b2lz4f_write_frame_header(dst, cap, prefs, out);
    if (rc) return rc;
    cc->pending.clear();
    b2lz4_xxh32_state_init(&cc->content, 0);
    cc->begun = true;
    return B2LZ4_OK;
}

int b2lz4f_compress_update(b2lz4f_cctx* cc, void* dst, size_t cap, const void* srcv, size_t n, size_t* out) {
    if (!cc || !out) return B2LZ4F_ERR_PARAMETER_NULL;
    *out = 0;
    if (!cc->begun) return B2LZ4F_ERR_COMPRESSION_STATE_UNINITIALIZED;
    const uint8_t* src = (const uint8_t*)srcv;
    const size_t have = cc->pending.size() + n;
    const size_t ready = have / cc->bs * cc->bs;            // bytes that form complete blocks
    if (ready == 0) { cc->pending.insert(cc->pending.end(), src, src + n); return B2LZ4_OK; }
    const size_t from_src = ready - cc->pending.size();     // >= 1 block boundary lies inside src
    int rc = emit_blocks(cc, cc->pending.data(), cc->pending.size(), src, from_src, (uint8_t*)dst, cap, out);
    if (rc) return rc;
    cc->pending.assign(src + from_src, src + n);
    return B2LZ4_OK;
}

int b2lz4f_compress_end(b2lz4f_cctx* cc, void* dstv, size_t cap, size_t* out) {
    if (!cc || !out) return B2LZ4F_ERR_PARAMETER_NULL;
    *out = 0;
    if (!cc->begun) return B2LZ4F_ERR_COMPRESSION_STATE_UNINITIALIZED;
    uint8_t* dst = (uint8_t*)dstv;
    size_t pos = 0;
    const size_t trailer = 4 + (cc->prefs.content_checksum == 1 ? 4 : 0);
    if (cap < trailer) return B2LZ4F_ERR_DST_MAX_SIZE_TOO_SMALL;
    int rc = emit_blocks(cc, cc->pending.data(), cc->pending.size(), nullptr, 0, dst, cap - trailer, &pos);
    if (rc) return rc;
    cc->pending.clear();
    dst[pos] = dst[pos + 1] = dst[pos + 2] = dst[pos + 3] = 0;   // end mark, src/lz4f.zig:433
    pos += 4;
    if (cc->prefs.content_checksum == 1) {                         // :437-441
        uint32_t h = b2lz4_xxh32_state_final(&cc->content);
        dst[pos] = (uint8_t)h; dst[pos + 1] = (uint8_t)(h >> 8); dst[pos + 2] = (uint8_t)(h >> 16); dst[pos + 3] = (uint8_t)(h >> 24);
        pos += 4;
    }
    cc->begun = false;
    *out = pos;
    return B2LZ4_OK;
}

}  // extern "C"
