// b2_host.h — internal host-side structures of libb2lz4 (context, workspace).
#pragma once
#include <atomic>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <string>
#include <cuda_runtime.h>
#include "../../include/b2lz4.h"
#include "b2_kernels.h"
#include "b2_mover.h"

namespace b2 {

void set_cuda_error(cudaError_t e, const char* what);

#define B2_CUDA(expr)                                  \
    do {                                               \
        cudaError_t _e = (expr);                       \
        if (_e != cudaSuccess) {                       \
            ::b2::set_cuda_error(_e, #expr);           \
            return B2LZ4_ERR_CUDA;                     \
        }                                              \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = (n + 255) & ~size_t(255);
        cudaError_t e = cudaMalloc(&p, want + 256);  // +256: 16-byte over-read slack for vector loads
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        cudaError_t e = cudaMallocHost(&p, n);
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

// Small results the device hands back once per call (lives in pinned memory).
struct HostResults {
    FrameTotals totals;
    WalkResult walk;
    DecodeSummary summary;
    uint32_t content_sum;
    uint32_t pad;
    XxhState xxh;
    uint64_t idx_nodes;
    FrameTotals pipe_totals[4];      // host-pointer pipeline: one slot per chunk in flight
    DecodeSummary pipe_summary[4];
};

}  // namespace b2

// Workspace of one in-flight chunk of the host-pointer pipeline (index 0 aliases the context's own buffers).
struct b2_ws_ref {
    b2::DevBuf *slots, *csize, *status, *sums, *rec_off;
    uint32_t* ticket;
    b2::FrameTotals* totals;
    b2::DecodeSummary* summary;
    cudaStream_t stream;
};

struct b2lz4_ctx {
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;   // the context's own stream
    cudaStream_t side = nullptr;     // content-checksum chain runs here, concurrently
    cudaStream_t copy_in = nullptr, copy_out = nullptr;  // host-pointer pipeline
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // Every call that touches the context's scratch (work ticket, HC tables, dictionary table, checksum state) is
    // ordered after the previous one on the DEVICE, whatever stream the caller passed: see b2::OrderGuard.
    cudaEvent_t ev_last = nullptr;
    cudaStream_t last_stream = nullptr;
    bool last_valid = false;
    cudaEvent_t ev_t[8] = {};
    cudaEvent_t ev_pipe[12] = {};
    std::recursive_mutex mu;
    bool timing = false;
    float phase_ms[5] = {0, 0, 0, 0, 0};
    // workspace
    b2::DevBuf slots, csize, status, sums, rec_off, small, walk_off, walk_hdr, out_len, order, hc_work;
    b2::DevBuf idx_tiles, idx_pos, idx_jump;   // parallel frame index scratch
    b2::DevBuf dict_table;                     // primed hash table of the dictionary compressor
    b2::DevBuf ds_work;                        // compressDestSize: per-block search state + probe lengths/results
    b2::DevBuf stage_in[3], stage_out[3], stage_aux;
    // two more chunk workspaces + compute streams for the host-pointer pipeline (three chunks in flight)
    b2::DevBuf x_slots[2], x_csize[2], x_status[2], x_sums[2], x_rec_off[2], x_small[2];
    cudaStream_t x_stream[2] = {nullptr, nullptr};
    b2::PinBuf results, pin_aux;
    b2::HostMover mover;                       // pinned bounce rings + copy threads for pageable caller buffers
    // layout of `small` (device): ticket u32 @0, FrameTotals @64, WalkResult @128, DecodeSummary @192,
    // content_sum u32 @256, XxhState @320, index node count u64 @512
    uint32_t* d_ticket() const { return small.as<uint32_t>(); }
    b2::FrameTotals* d_totals() const { return reinterpret_cast<b2::FrameTotals*>(small.as<uint8_t>() + 64); }
    b2::WalkResult* d_walk() const { return reinterpret_cast<b2::WalkResult*>(small.as<uint8_t>() + 128); }
    b2::DecodeSummary* d_summary() const { return reinterpret_cast<b2::DecodeSummary*>(small.as<uint8_t>() + 192); }
    uint32_t* d_content_sum() const { return reinterpret_cast<uint32_t*>(small.as<uint8_t>() + 256); }
    b2::XxhState* d_xxh() const { return reinterpret_cast<b2::XxhState*>(small.as<uint8_t>() + 320); }
    uint64_t* d_idx_nodes() const { return reinterpret_cast<uint64_t*>(small.as<uint8_t>() + 512); }
    b2::HostResults* h() const { return results.as<b2::HostResults>(); }
    size_t workspace_bytes() const;
    b2_ws_ref ws(int i) {
        if (i == 0) return b2_ws_ref{&slots, &csize, &status, &sums, &rec_off, d_ticket(), d_totals(), d_summary(), stream};
        uint8_t* sm = x_small[i - 1].as<uint8_t>();
        return b2_ws_ref{&x_slots[i - 1], &x_csize[i - 1], &x_status[i - 1], &x_sums[i - 1], &x_rec_off[i - 1],
                         reinterpret_cast<uint32_t*>(sm), reinterpret_cast<b2::FrameTotals*>(sm + 64),
                         reinterpret_cast<b2::DecodeSummary*>(sm + 192), x_stream[i - 1]};
    }
};

namespace b2 {
// Held (under c->mu) by every entry point that enqueues work using the context's scratch on stream s.  Two calls on one
// context with different streams would otherwise run concurrently on the device and trample the shared work ticket /
// HC tables / checksum state; the guard makes s wait for the previous call's event and records a new one on exit.
struct OrderGuard {
    b2lz4_ctx* c;
    cudaStream_t s;
    OrderGuard(b2lz4_ctx* ctx, cudaStream_t stream) : c(ctx), s(stream) {
        if (c->last_valid && c->last_stream != s) cudaStreamWaitEvent(s, c->ev_last, 0);
    }
    ~OrderGuard() {
        if (cudaEventRecord(c->ev_last, s) == cudaSuccess) { c->last_stream = s; c->last_valid = true; }
    }
};
// host-pointer entry points (synchronous, on the context's own streams): finish whatever an earlier asynchronous call left
inline void wait_previous(b2lz4_ctx* c) {
    if (c->last_valid) cudaEventSynchronize(c->ev_last);
}
}  // namespace b2

// internal entry points shared between translation units
int b2_compress_dev_impl(b2lz4_ctx* c, const void* src, size_t n, void* dst, size_t cap, const b2lz4f_prefs* prefs,
                         size_t* out, cudaStream_t s, bool body_only);
int b2_decompress_dev_impl(b2lz4_ctx* c, const void* src, size_t n, void* dst, size_t cap, size_t* out, cudaStream_t s);
int b2_default_ctx(b2lz4_ctx** out);
int b2_enqueue_body(b2lz4_ctx* c, const b2_ws_ref& w, const void* src, size_t n, size_t bs, int level, bool bc, uint8_t* body,
                    cudaStream_t s, bool timing);
int b2_hc_supported(int level);
