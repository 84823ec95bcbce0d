// b2lz4_api.cu — the C-ABI of include/b2lz4.h: context / workspace management, the frame writer and
// reader built on the kernels (host side of /root/reference/src/lz4f.zig:354-446 and :541-638), the
// batch and single-block entry points, the streaming trio and the multi-GPU shard helpers.
// No CPU codec lives here: every compress / decompress / checksum byte is produced by a kernel.  The
// only host arithmetic is the 7-19 byte frame header (its XXH32 header-checksum byte included).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <new>
#include <vector>
#include "b2_host.h"

namespace b2 {

static std::atomic<uint64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static Tune g_tune = {};
Tune& tune() { return g_tune; }

static thread_local std::string g_cuda_err;
void set_cuda_error(cudaError_t e, const char* what) {
    g_cuda_err = std::string(cudaGetErrorName(e)) + ": " + cudaGetErrorString(e) + " @ " + what;
    cudaGetLastError();  // clear sticky-less errors
}

// ---------------------------------------------------------------- tiny host XXH32 (frame header byte only)
static inline uint32_t rotl_h(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
static uint32_t xxh32_small_host(const uint8_t* p, size_t n) {  // n < 16 always (header descriptor is 2..14 bytes)
    const uint32_t P1 = 2654435761u, P2 = 2246822519u, P3 = 3266489917u, P4 = 668265263u, P5 = 374761393u;
    uint32_t h = P5 + (uint32_t)n;  // seed 0, len < 16
    while (n >= 4) {
        uint32_t w = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
        h = rotl_h(h + w * P3, 17) * P4;
        p += 4; n -= 4;
    }
    while (n) { h = rotl_h(h + (uint32_t)(*p) * P5, 11) * P1; p++; n--; }
    h ^= h >> 15; h *= P2; h ^= h >> 13; h *= P3; h ^= h >> 16;
    return h;
}

static inline uint32_t rd32h(const uint8_t* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static inline void wr32h(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }

static bool block_size_of(uint32_t id, size_t& bs) {  // BlockSizeID.toBlockSize, src/lz4f.zig:71-78
    switch (id) {
        case 0: case 4: bs = 64u << 10; return true;
        case 5: bs = 256u << 10; return true;
        case 6: bs = 1u << 20; return true;
        case 7: bs = 4u << 20; return true;
        default: return false;
    }
}

static size_t compress_bound(size_t n) { return n > LZ4_MAX_INPUT_SIZE ? 0 : n + n / 255 + 16; }  // src/lz4.zig:80-83
static size_t slot_stride_for(size_t bs) { return (compress_bound(bs) + 15) & ~size_t(15); }

// HC level routing, src/lz4hc.zig:1445 + :72-97.  Returns nbSearches, or -1 for strategies outside the path.
static int hc_nb_searches(int level) {
    int l = level < 2 ? 9 : (level > 12 ? 12 : level);
    if (l < 3 || l > 9) return -1;
    return 4 << (l - 3);
}

}  // namespace b2

using namespace b2;

int b2_hc_supported(int level) { return hc_nb_searches(level) >= 0 ? 1 : 0; }

size_t b2lz4_ctx::workspace_bytes() const {
    size_t t = slots.cap + csize.cap + status.cap + sums.cap + rec_off.cap + small.cap + walk_off.cap + walk_hdr.cap +
               out_len.cap + order.cap + hc_work.cap + stage_aux.cap + idx_tiles.cap + idx_pos.cap + idx_jump.cap + dict_table.cap + ds_work.cap;
    for (int i = 0; i < 3; i++) t += stage_in[i].cap + stage_out[i].cap;
    for (int i = 0; i < 2; i++) t += x_slots[i].cap + x_csize[i].cap + x_status[i].cap + x_sums[i].cap + x_rec_off[i].cap + x_small[i].cap;
    return t;
}

// ================================================================ misc exports
extern "C" {

const char* b2lz4_status_name(int s) {
    static const char* lz4e[] = {"ok", "lz4.OutputTooSmall", "lz4.InputTooLarge", "lz4.CorruptedData",
                                 "lz4.DecompressionFailed", "lz4.InvalidState", "lz4.AllocationFailed"};
    static const char* fe[] = {"lz4f.Generic", "lz4f.MaxBlockSizeInvalid", "lz4f.BlockModeInvalid", "lz4f.ParameterInvalid",
                               "lz4f.CompressionLevelInvalid", "lz4f.HeaderVersionWrong", "lz4f.BlockChecksumInvalid",
                               "lz4f.ReservedFlagSet", "lz4f.AllocationFailed", "lz4f.SrcSizeTooLarge",
                               "lz4f.DstMaxSizeTooSmall", "lz4f.FrameHeaderIncomplete", "lz4f.FrameTypeUnknown",
                               "lz4f.FrameSizeWrong", "lz4f.SrcPtrWrong", "lz4f.DecompressionFailed",
                               "lz4f.HeaderChecksumInvalid", "lz4f.ContentChecksumInvalid",
                               "lz4f.FrameDecodingAlreadyStarted", "lz4f.CompressionStateUninitialized",
                               "lz4f.ParameterNull", "lz4f.MaxCode", "lz4f.OutOfMemory"};
    if (s >= 0 && s <= 6) return lz4e[s];
    if (s >= 100 && s < 123) return fe[s - 100];
    if (s == B2LZ4_ERR_CUDA) return "b2lz4.CudaError";
    if (s == B2LZ4_ERR_UNSUPPORTED_LEVEL) return "b2lz4.UnsupportedLevel";
    return "b2lz4.UnknownStatus";
}
const char* b2lz4_last_cuda_error(void) { return g_cuda_err.c_str(); }
uint64_t b2lz4_kernel_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
#ifndef B2_SRC_HASH
#define B2_SRC_HASH "unknown"
#endif
const char* b2lz4_version(void) { return "b2lz4 0.2 (sm_100a) src:" B2_SRC_HASH; }
int b2lz4_debug_tune(const char* key, int value) {
    if (!key) return -1;
    b2::Tune& t = b2::tune();
    const std::string k(key);
    int* slot = k == "k1_ctas" ? &t.k1_ctas : k == "k1_variant" ? &t.k1_variant : k == "k2_occ" ? &t.k2_occ : k == "k2_variant" ? &t.k2_variant
              : k == "k3_variant" ? &t.k3_variant : k == "pipe_blocks" ? &t.pipe_blocks : k == "no_pipeline" ? &t.no_pipeline
              : k == "serial_walk" ? &t.serial_walk : k == "xxh_variant" ? &t.xxh_variant : nullptr;
    if (!slot && k.rfind("spare", 0) == 0 && k.size() == 6 && k[5] >= '0' && k[5] <= '7') slot = &t.spare[k[5] - '0'];
    if (!slot) return -1;
    const int old = *slot;
    *slot = value;
    return old;
}

// ================================================================ context
int b2lz4_ctx_create(int device, b2lz4_ctx** out) {
    if (!out) return B2LZ4F_ERR_PARAMETER_NULL;
    *out = nullptr;
    int count = 0;
    B2_CUDA(cudaGetDeviceCount(&count));
    if (count <= 0) { g_cuda_err = "no CUDA device"; return B2LZ4_ERR_CUDA; }
    if (device < 0) B2_CUDA(cudaGetDevice(&device));
    B2_CUDA(cudaSetDevice(device));
    b2lz4_ctx* c = new (std::nothrow) b2lz4_ctx();
    if (!c) return B2LZ4_ERR_ALLOCATION_FAILED;
    c->device = device;
    cudaDeviceProp prop;
    B2_CUDA(cudaGetDeviceProperties(&prop, device));
    c->num_sms = prop.multiProcessorCount;
    B2_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    B2_CUDA(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
    B2_CUDA(cudaStreamCreateWithFlags(&c->copy_in, cudaStreamNonBlocking));
    B2_CUDA(cudaStreamCreateWithFlags(&c->copy_out, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
        B2_CUDA(cudaStreamCreateWithFlags(&c->x_stream[i], cudaStreamNonBlocking));
        B2_CUDA(c->x_small[i].ensure(1024));
        B2_CUDA(cudaMemset(c->x_small[i].p, 0, 1024));
    }
    B2_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    B2_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    B2_CUDA(cudaEventCreateWithFlags(&c->ev_last, cudaEventDisableTiming));
    for (auto& e : c->ev_t) B2_CUDA(cudaEventCreate(&e));
    for (auto& e : c->ev_pipe) B2_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    B2_CUDA(c->small.ensure(1024));
    B2_CUDA(cudaMemset(c->small.p, 0, 1024));
    B2_CUDA(c->results.ensure(sizeof(HostResults)));
    memset(c->results.p, 0, sizeof(HostResults));
    *out = c;
    return B2LZ4_OK;
}

void b2lz4_ctx_destroy(b2lz4_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    DevBuf* bufs[] = {&c->slots, &c->csize, &c->status, &c->sums, &c->rec_off, &c->small, &c->walk_off, &c->walk_hdr,
                      &c->out_len, &c->order, &c->hc_work, &c->idx_tiles, &c->idx_pos, &c->idx_jump, &c->dict_table, &c->ds_work, &c->stage_in[0], &c->stage_in[1], &c->stage_in[2],
                      &c->stage_out[0], &c->stage_out[1], &c->stage_out[2], &c->stage_aux,
                      &c->x_slots[0], &c->x_slots[1], &c->x_csize[0], &c->x_csize[1], &c->x_status[0], &c->x_status[1],
                      &c->x_sums[0], &c->x_sums[1], &c->x_rec_off[0], &c->x_rec_off[1], &c->x_small[0], &c->x_small[1]};
    for (auto* b : bufs) b->release();
    c->mover.release();
    c->results.release();
    c->pin_aux.release();
    for (auto& e : c->ev_t) if (e) cudaEventDestroy(e);
    for (auto& e : c->ev_pipe) if (e) cudaEventDestroy(e);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->ev_last) cudaEventDestroy(c->ev_last);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->side) cudaStreamDestroy(c->side);
    if (c->copy_in) cudaStreamDestroy(c->copy_in);
    if (c->copy_out) cudaStreamDestroy(c->copy_out);
    for (auto& xs : c->x_stream) if (xs) cudaStreamDestroy(xs);
    delete c;
}
int b2lz4_ctx_device(const b2lz4_ctx* c) { return c ? c->device : -1; }
size_t b2lz4_ctx_workspace_bytes(const b2lz4_ctx* c) { return c ? c->workspace_bytes() : 0; }
void b2lz4_ctx_set_timing(b2lz4_ctx* c, int enabled) { if (c) c->timing = enabled != 0; }
int b2lz4_ctx_last_phase_ms(const b2lz4_ctx* c, float out_ms[5]) {
    if (!c || !out_ms) return B2LZ4F_ERR_PARAMETER_NULL;
    for (int i = 0; i < 5; i++) out_ms[i] = c->phase_ms[i];
    return B2LZ4_OK;
}

}  // extern "C"

// default context (functions without a ctx argument)
static std::mutex g_default_mu;
static b2lz4_ctx* g_default_ctx = nullptr;
int b2_default_ctx(b2lz4_ctx** out) {
    std::lock_guard<std::mutex> lk(g_default_mu);
    if (!g_default_ctx) {
        int rc = b2lz4_ctx_create(-1, &g_default_ctx);
        if (rc) return rc;
    }
    *out = g_default_ctx;
    return B2LZ4_OK;
}

// ================================================================ frame header codec (host)
extern "C" {

void b2lz4f_prefs_init(b2lz4f_prefs* p) { if (p) memset(p, 0, sizeof *p); }

size_t b2lz4_compress_bound(size_t n) { return compress_bound(n); }

size_t b2lz4f_compress_frame_bound(size_t srcSize, const b2lz4f_prefs* prefs) {  // src/lz4f.zig:274-301
    b2lz4f_prefs d; if (!prefs) { b2lz4f_prefs_init(&d); prefs = &d; }
    size_t bs; if (!block_size_of(prefs->block_size_id, bs)) bs = 65536;
    size_t nb = (srcSize + bs - 1) / bs;
    size_t r = 19 + nb * (4 + compress_bound(bs) + (prefs->block_checksum == 1 ? 4 : 0)) + 4;
    if (prefs->content_checksum == 1) r += 4;
    return r;
}

int b2lz4f_write_frame_header(void* dstv, size_t cap, const b2lz4f_prefs* p, size_t* out) {  // src/lz4f.zig:304-351
    if (!dstv || !p || !out) return B2LZ4F_ERR_PARAMETER_NULL;
    uint8_t* dst = (uint8_t*)dstv;
    if (cap < 7) return B2LZ4F_ERR_DST_MAX_SIZE_TOO_SMALL;
    size_t pos = 0;
    wr32h(dst, 0x184D2204u); pos = 4;
    uint8_t flg = 0x40;                                   // encodeFLG :152-184
    if (p->block_mode == 1) flg |= 0x20;
    if (p->block_checksum == 1) flg |= 0x10;
    if (p->content_size != 0) flg |= 0x08;
    if (p->content_checksum == 1) flg |= 0x04;
    if (p->dict_id != 0) flg |= 0x01;
    dst[pos++] = flg;
    uint32_t id = p->block_size_id;                       // encodeBD :224-232
    dst[pos++] = (uint8_t)(((id == 0 || id == 4) ? 4u : id) << 4);
    if (p->content_size != 0) {
        if (cap < pos + 8) return B2LZ4F_ERR_DST_MAX_SIZE_TOO_SMALL;
        wr32h(dst + pos, (uint32_t)p->content_size); wr32h(dst + pos + 4, (uint32_t)(p->content_size >> 32)); pos += 8;
    }
    if (p->dict_id != 0) {
        if (cap < pos + 4) return B2LZ4F_ERR_DST_MAX_SIZE_TOO_SMALL;
        wr32h(dst + pos, p->dict_id); pos += 4;
    }
    if (cap < pos + 1) return B2LZ4F_ERR_DST_MAX_SIZE_TOO_SMALL;
    dst[pos] = (uint8_t)((xxh32_small_host(dst + 4, pos - 4) >> 8) & 0xFF);  // headerChecksum :138-141
    pos += 1;
    *out = pos;
    return B2LZ4_OK;
}

int b2lz4f_header_size(const void* srcv, size_t n, size_t* out) {  // src/lz4f.zig:451-480
    if (!out) return B2LZ4F_ERR_PARAMETER_NULL;
    const uint8_t* src = (const uint8_t*)srcv;
    if (n < 5) return B2LZ4F_ERR_FRAME_HEADER_INCOMPLETE;
    uint32_t magic = rd32h(src);
    if (magic != 0x184D2204u) {
        if ((magic & 0xFFFFFFF0u) == 0x184D2A50u) { *out = 8; return B2LZ4_OK; }
        return B2LZ4F_ERR_FRAME_TYPE_UNKNOWN;
    }
    size_t size = 7;
    if (src[4] & 0x08) size += 8;
    if (src[4] & 0x01) size += 4;
    *out = size;
    return B2LZ4_OK;
}

int b2lz4f_parse_frame_header(const void* srcv, size_t n, b2lz4f_prefs* info, size_t* size) {  // src/lz4f.zig:483-538
    if (!info || !size) return B2LZ4F_ERR_PARAMETER_NULL;
    const uint8_t* src = (const uint8_t*)srcv;
    if (n < 7) return B2LZ4F_ERR_FRAME_HEADER_INCOMPLETE;
    if (rd32h(src) != 0x184D2204u) return B2LZ4F_ERR_FRAME_TYPE_UNKNOWN;
    size_t pos = 4;
    uint8_t flg = src[pos];
    b2lz4f_prefs_init(info);
    if (((flg >> 6) & 3) != 1) return B2LZ4F_ERR_HEADER_VERSION_WRONG;    // decodeFLG :187-221
    if (flg & 0x02) return B2LZ4F_ERR_RESERVED_FLAG_SET;
    info->block_mode = (flg & 0x20) ? 1 : 0;
    info->block_checksum = (flg & 0x10) ? 1 : 0;
    info->content_checksum = (flg & 0x04) ? 1 : 0;
    pos++;
    uint8_t bd = src[pos];                                                // decodeBD :235-249
    if (bd & 0x8F) return B2LZ4F_ERR_RESERVED_FLAG_SET;
    switch ((bd >> 4) & 7) {
        case 0: case 4: info->block_size_id = 4; break;
        case 5: info->block_size_id = 5; break;
        case 6: info->block_size_id = 6; break;
        case 7: info->block_size_id = 7; break;
        default: return B2LZ4F_ERR_MAX_BLOCK_SIZE_INVALID;
    }
    pos++;
    if (flg & 0x08) {
        if (n < pos + 8) return B2LZ4F_ERR_FRAME_HEADER_INCOMPLETE;
        info->content_size = (uint64_t)rd32h(src + pos) | ((uint64_t)rd32h(src + pos + 4) << 32);
        pos += 8;
    }
    if (flg & 0x01) {
        if (n < pos + 4) return B2LZ4F_ERR_FRAME_HEADER_INCOMPLETE;
        info->dict_id = rd32h(src + pos);
        pos += 4;
    }
    if (n < pos + 1) return B2LZ4F_ERR_FRAME_HEADER_INCOMPLETE;
    if (src[pos] != (uint8_t)((xxh32_small_host(src + 4, pos - 4) >> 8) & 0xFF)) return B2LZ4F_ERR_HEADER_CHECKSUM_INVALID;
    pos++;
    *size = pos;
    return B2LZ4_OK;
}

}  // extern "C"

// ================================================================ device-side frame writer
namespace {

struct Timer {
    b2lz4_ctx* c; cudaStream_t s; bool on;
    void mark(int i) { if (on) cudaEventRecord(c->ev_t[i], s); }
};

static BlockSet regular_in(const void* base, uint64_t stride, uint64_t total) {
    BlockSet b; b.base = (const uint8_t*)base; b.off = nullptr; b.len = nullptr; b.stride = stride; b.total = total;
    b.len_mask = 0xFFFFFFFFu; return b;
}
static BlockSet explicit_in(const void* base, const uint64_t* off, const uint32_t* len, uint32_t mask = 0xFFFFFFFFu) {
    BlockSet b; b.base = (const uint8_t*)base; b.off = off; b.len = len; b.stride = 0; b.total = 0; b.len_mask = mask; return b;
}
static OutSet regular_out(void* base, uint64_t stride, uint64_t total, uint32_t slot_cap) {
    OutSet o; o.base = (uint8_t*)base; o.off = nullptr; o.cap = nullptr; o.stride = stride; o.total = total; o.slot_cap = slot_cap;
    return o;
}
static OutSet explicit_out(void* base, const uint64_t* off, const uint32_t* cap) {
    OutSet o; o.base = (uint8_t*)base; o.off = off; o.cap = cap; o.stride = 0; o.total = 0; o.slot_cap = 0; return o;
}

// HC tables must start zeroed once (bucket values carry an epoch base afterwards).
// The workspace grows with the largest launch seen (a fresh, zeroed buffer is a valid state: epoch base 0, empty tables).
static int ensure_hc_work(b2lz4_ctx* c, cudaStream_t s, uint32_t nblocks) {
    const size_t need = hc_work_bytes(c->num_sms, nblocks);
    if (c->hc_work.cap >= need) return B2LZ4_OK;
    B2_CUDA(cudaStreamSynchronize(s));            // nothing in flight may still use the buffer that is about to be freed
    B2_CUDA(c->hc_work.ensure(need));
    B2_CUDA(cudaMemsetAsync(c->hc_work.p, 0, need, s));
    return B2LZ4_OK;
}

// Runs the block codec for the blocks of [src, src+n) into the context's slots.  level 0 = fast.
static int encode_blocks_to_slots(b2lz4_ctx* c, const b2_ws_ref& w, const void* src, size_t n, size_t bs, int level,
                                  uint32_t nb, cudaStream_t s) {
    const size_t stride = slot_stride_for(bs);
    B2_CUDA(w.slots->ensure((size_t)nb * stride));
    B2_CUDA(w.csize->ensure((size_t)nb * 4));
    B2_CUDA(w.status->ensure((size_t)nb * 4));
    BlockSet in = regular_in(src, bs, n);
    OutSet out = regular_out(w.slots->p, stride, (uint64_t)nb * stride, (uint32_t)compress_bound(bs));
    if (level > 0) {
        int nbs = hc_nb_searches(level);
        if (nbs < 0) return B2LZ4_ERR_UNSUPPORTED_LEVEL;
        { int rc = ensure_hc_work(c, s, nb); if (rc) return rc; }
        B2_CUDA(launch_compress_hc(in, out, w.csize->as<uint32_t>(), w.status->as<int32_t>(), nb, nbs,
                                   c->hc_work.as<uint8_t>(), w.ticket, c->num_sms, s));
    } else {
        // Many more blocks than the GPU holds at once (28 / 14 per SM): expensive blocks first, so that the launch does
        // not end on a tail of expensive blocks drawn late.  (b2lz4_debug_tune("spare3", 1) switches it off; the
        // host-pointer pipeline's chunks are below the threshold.)
        const uint32_t slots_in_flight = (uint32_t)c->num_sms * (bs <= 65536 ? 28u : 14u);
        if (nb >= 2 * slots_in_flight && bs >= 16384 && w.slots == &c->slots && b2::tune().spare[3] == 0) {   // (the context-wide scratch: workspace 0 only)
            B2_CUDA(c->order.ensure(order_scratch_bytes(nb)));
            B2_CUDA(launch_compress_fast_ordered(in, out, w.csize->as<uint32_t>(), w.status->as<int32_t>(), nb, (uint32_t)bs,
                                                 w.ticket, c->num_sms, c->order.p, s));
        } else {
            B2_CUDA(launch_compress_fast(in, out, w.csize->as<uint32_t>(), w.status->as<int32_t>(), nb, (uint32_t)bs, 1,
                                         w.ticket, c->num_sms, s));
        }
    }
    return B2LZ4_OK;
}

}  // namespace
// Enqueues block codec -> block checksums -> record scan -> assembly for the blocks of [src, src+n) (device)
// into `body` (device) on stream s.  Nothing is synchronised; the totals end up in c->d_totals().
int b2_enqueue_body(b2lz4_ctx* c, const b2_ws_ref& w, const void* src, size_t n, size_t bs, int level, bool bc, uint8_t* body,
                    cudaStream_t s, bool timing) {
    const uint32_t nb = (uint32_t)((n + bs - 1) / bs);
    Timer T{c, s, timing};
    if (nb) { int rc = encode_blocks_to_slots(c, w, src, n, bs, level, nb, s); if (rc) return rc; }
    else { B2_CUDA(w.csize->ensure(4)); B2_CUDA(w.status->ensure(4)); }
    T.mark(1);
    const size_t stride = slot_stride_for(bs);
    BlockSet slots = regular_in(w.slots->p, stride, (uint64_t)nb * stride);
    BlockSet raw = regular_in(src, bs, n);
    B2_CUDA(w.sums->ensure((size_t)nb * 4 + 4));
    if (bc && nb) B2_CUDA(launch_xxh32_stored(slots, raw, w.csize->as<uint32_t>(), w.sums->as<uint32_t>(), nb, s));
    T.mark(2);
    B2_CUDA(w.rec_off->ensure(((size_t)nb + 1) * 8));
    B2_CUDA(launch_scan_records(w.csize->as<uint32_t>(), w.status->as<int32_t>(), nb, bs, n, bc ? 1 : 0,
                                w.rec_off->as<uint64_t>(), w.totals, s));
    if (nb) B2_CUDA(launch_assemble(slots, raw, w.csize->as<uint32_t>(), w.sums->as<uint32_t>(), w.rec_off->as<uint64_t>(),
                                    body, nb, bc ? 1 : 0, c->num_sms, s));
    return B2LZ4_OK;
}

// Frame (or body-only) compression of a device buffer.  Enqueues everything on `s`, syncs once.
int b2_compress_dev_impl(b2lz4_ctx* c, const void* src, size_t n, void* dst, size_t cap, const b2lz4f_prefs* prefs,
                             size_t* out, cudaStream_t s, bool body_only) {
    b2lz4f_prefs d; if (!prefs) { b2lz4f_prefs_init(&d); prefs = &d; }
    if (!out) return B2LZ4F_ERR_PARAMETER_NULL;
    *out = 0;
    if (!s) s = c->stream;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    OrderGuard og(c, s);
    B2_CUDA(cudaSetDevice(c->device));
    size_t bs;
    const bool bs_ok = block_size_of(prefs->block_size_id, bs);
    const bool bc = prefs->block_checksum == 1, cc = prefs->content_checksum == 1 && !body_only;
    const size_t nb64 = bs_ok ? (n + bs - 1) / bs : 0;
    if (!body_only) {
        if (cap < b2lz4f_compress_frame_bound(n, prefs)) return B2LZ4F_ERR_DST_MAX_SIZE_TOO_SMALL;  // src/lz4f.zig:363-366
    } else if (bs_ok) {
        if (cap < nb64 * (4 + compress_bound(bs) + (bc ? 4 : 0))) return B2LZ4F_ERR_DST_MAX_SIZE_TOO_SMALL;
    }
    uint8_t hdr[32]; size_t hsize = 0;
    if (!body_only) { int rc = b2lz4f_write_frame_header(hdr, sizeof hdr, prefs, &hsize); if (rc) return rc; }  // :369
    if (!bs_ok) return B2LZ4F_ERR_MAX_BLOCK_SIZE_INVALID;                                                      // :372
    if (nb64 > 0x7FFFFFFFull) return B2LZ4F_ERR_SRC_SIZE_TOO_LARGE;
    const uint32_t nb = (uint32_t)nb64;
    const int level = prefs->compression_level;
    if (level > 0 && hc_nb_searches(level) < 0) return B2LZ4_ERR_UNSUPPORTED_LEVEL;

    Timer T{c, s, c->timing};
    T.mark(0);
    uint8_t* d8 = (uint8_t*)dst;
    if (hsize) {
        B2_CUDA(c->pin_aux.ensure(64));
        memcpy(c->pin_aux.p, hdr, hsize);
        B2_CUDA(cudaMemcpyAsync(d8, c->pin_aux.p, hsize, cudaMemcpyHostToDevice, s));
    }
    // content checksum: one serial chain, on the side stream, concurrent with the codec
    if (cc) {
        B2_CUDA(cudaEventRecord(c->ev_fork, s));
        B2_CUDA(cudaStreamWaitEvent(c->side, c->ev_fork, 0));
        if (c->timing) cudaEventRecord(c->ev_t[5], c->side);
        B2_CUDA(launch_xxh32_init(c->d_xxh(), 0, c->side));
        B2_CUDA(launch_xxh32_update(c->d_xxh(), (const uint8_t*)src, n, c->side));
        B2_CUDA(launch_xxh32_final(c->d_xxh(), c->d_content_sum(), c->side));
        if (c->timing) cudaEventRecord(c->ev_t[6], c->side);
        B2_CUDA(cudaEventRecord(c->ev_join, c->side));
    }
    { int rc = b2_enqueue_body(c, c->ws(0), src, n, bs, level, bc, d8 + hsize, s, c->timing); if (rc) return rc; }
    if (!body_only) {
        if (cc) B2_CUDA(cudaStreamWaitEvent(s, c->ev_join, 0));
        B2_CUDA(launch_finalize(d8, hsize, c->d_totals(), cc ? c->d_content_sum() : nullptr, s));
    }
    T.mark(3);
    B2_CUDA(cudaMemcpyAsync(&c->h()->totals, c->d_totals(), sizeof(FrameTotals), cudaMemcpyDeviceToHost, s));
    T.mark(4);
    B2_CUDA(cudaStreamSynchronize(s));
    if (c->timing) {
        cudaEventElapsedTime(&c->phase_ms[0], c->ev_t[0], c->ev_t[1]);
        cudaEventElapsedTime(&c->phase_ms[1], c->ev_t[1], c->ev_t[2]);
        cudaEventElapsedTime(&c->phase_ms[2], c->ev_t[2], c->ev_t[3]);
        c->phase_ms[3] = 0;
        if (cc) cudaEventElapsedTime(&c->phase_ms[3], c->ev_t[5], c->ev_t[6]);
        cudaEventElapsedTime(&c->phase_ms[4], c->ev_t[0], c->ev_t[4]);
    }
    const FrameTotals& t = c->h()->totals;
    if (nb && t.first_bad != 0xFFFFFFFFu)  // mapCompressionError, src/lz4f.zig:144-149
        return t.bad_status == B2LZ4_ERR_OUTPUT_TOO_SMALL ? B2LZ4F_ERR_DST_MAX_SIZE_TOO_SMALL : B2LZ4F_ERR_GENERIC;
    *out = hsize + t.body_bytes + (body_only ? 0 : 4 + (cc ? 4 : 0));
    return B2LZ4_OK;
}

// ================================================================ device-side frame reader
namespace {

static int decode_blocks_dev(b2lz4_ctx* c, const uint8_t* src, uint64_t n, const uint64_t* d_off, const uint32_t* d_hdr,
                             uint32_t nb, uint32_t terminal, uint8_t* dst, uint64_t cap, uint32_t bs, bool bc,
                             uint64_t* total_out, cudaStream_t s, Timer& T) {
    *total_out = 0;
    B2_CUDA(c->out_len.ensure((size_t)nb * 4 + 4));
    B2_CUDA(c->status.ensure((size_t)nb * 4 + 4));
    B2_CUDA(c->sums.ensure((size_t)nb * 4 + 4));
    if (bc && nb) B2_CUDA(launch_xxh32_ranges(src, d_off, d_hdr, c->sums.as<uint32_t>(), nb, s));
    T.mark(2);
    BlockSet in = explicit_in(src, d_off, d_hdr, 0x7FFFFFFFu);
    OutSet out = regular_out(dst, bs, cap, bs);
    B2_CUDA(c->order.ensure((size_t)nb * 4 + 4));
    if (nb) B2_CUDA(launch_decompress(in, out, d_hdr, c->out_len.as<uint32_t>(), c->status.as<int32_t>(), nb, nullptr, 0,
                                      c->d_ticket(), c->num_sms, s, c->order.as<uint32_t>(), bs));
    T.mark(3);
    B2_CUDA(launch_decode_summary(c->out_len.as<uint32_t>(), c->status.as<int32_t>(), c->sums.as<uint32_t>(), src, d_off, d_hdr,
                                  nb, bs, bc ? 1 : 0, c->d_summary(), s));
    B2_CUDA(cudaMemcpyAsync(&c->h()->summary, c->d_summary(), sizeof(DecodeSummary), cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    const DecodeSummary sm = c->h()->summary;
    if (sm.first_bad == 0xFFFFFFFFu && sm.layout_ok) {
        if (terminal == 2) return B2LZ4F_ERR_FRAME_SIZE_WRONG;
        *total_out = sm.total;
        return B2LZ4_OK;
    }
    // ---- general path: something is unusual (an error, or a foreign frame with short blocks).  Resolve
    // the reference's sequential semantics exactly: natural sizes on the device, dstPos chain on the host.
    std::vector<uint32_t> hdr(nb), nat(nb), calc(nb);
    std::vector<int32_t> nst(nb);
    std::vector<uint64_t> off(nb);
    B2_CUDA(launch_decoded_size(in, d_hdr, c->out_len.as<uint32_t>(), c->status.as<int32_t>(), nb, c->num_sms, s));
    B2_CUDA(cudaMemcpyAsync(hdr.data(), d_hdr, (size_t)nb * 4, cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaMemcpyAsync(off.data(), d_off, (size_t)nb * 8, cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaMemcpyAsync(nat.data(), c->out_len.p, (size_t)nb * 4, cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaMemcpyAsync(nst.data(), c->status.p, (size_t)nb * 4, cudaMemcpyDeviceToHost, s));
    if (bc) B2_CUDA(cudaMemcpyAsync(calc.data(), c->sums.p, (size_t)nb * 4, cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    std::vector<uint64_t> doff(nb);
    std::vector<uint32_t> dcap(nb);
    uint64_t dstPos = 0;
    int err = B2LZ4_OK;
    uint32_t good = nb;
    for (uint32_t i = 0; i < nb; i++) {
        const uint32_t sz = hdr[i] & 0x7FFFFFFFu;
        if (bc) {  // src/lz4f.zig:590-600
            uint32_t stored;
            B2_CUDA(cudaMemcpy(&stored, src + off[i] + sz, 4, cudaMemcpyDeviceToHost));
            if (stored != calc[i]) { err = B2LZ4F_ERR_BLOCK_CHECKSUM_INVALID; good = i; break; }
        }
        uint32_t len_i;
        if (hdr[i] & 0x80000000u) {  // :603-608
            if (dstPos + sz > cap) { err = B2LZ4F_ERR_DST_MAX_SIZE_TOO_SMALL; good = i; break; }
            len_i = sz;
        } else {                      // :610-613
            const uint64_t room = cap - dstPos;
            if (room == 0) len_i = 0;  // decompressGeneric: dst.len == 0 -> 0 (src/lz4.zig:98)
            else if (nst[i] != 0 || nat[i] > room) { err = B2LZ4F_ERR_DECOMPRESSION_FAILED; good = i; break; }
            else len_i = nat[i];
        }
        doff[i] = dstPos; dcap[i] = len_i; dstPos += len_i;
    }
    if (good) {  // the reference has already produced the blocks before the failing one
        B2_CUDA(c->stage_aux.ensure((size_t)good * 12));
        uint64_t* d_doff = c->stage_aux.as<uint64_t>();
        uint32_t* d_dcap = reinterpret_cast<uint32_t*>(c->stage_aux.as<uint8_t>() + (size_t)good * 8);
        B2_CUDA(cudaMemcpyAsync(d_doff, doff.data(), (size_t)good * 8, cudaMemcpyHostToDevice, s));
        B2_CUDA(cudaMemcpyAsync(d_dcap, dcap.data(), (size_t)good * 4, cudaMemcpyHostToDevice, s));
        OutSet eo = explicit_out(dst, d_doff, d_dcap);
        B2_CUDA(launch_decompress(in, eo, d_hdr, c->out_len.as<uint32_t>(), c->status.as<int32_t>(), good, nullptr, 0,
                                  c->d_ticket(), c->num_sms, s));
        B2_CUDA(cudaStreamSynchronize(s));
    }
    if (err) return err;
    if (terminal == 2) return B2LZ4F_ERR_FRAME_SIZE_WRONG;
    *total_out = dstPos;
    return B2LZ4_OK;
}

// Block index of a frame body: offsets and header words of its records, in order (reference loop control
// src/lz4f.zig:563-591).  Built in parallel (k_index.cu); the serial header chase (k_walk) only runs when
// the candidate list is implausibly long or a header above `bound` sits on the chain.
static int build_block_index(b2lz4_ctx* c, const uint8_t* src, uint64_t n, uint64_t start, uint32_t bound, bool bc,
                             cudaStream_t s, WalkResult* wout) {
    const bool force_serial = tune().serial_walk != 0;
    if (!force_serial && n > start) {
        const uint32_t ntiles = index_tiles(src, n);
        uint64_t cap_nodes = (n - start) / 64 + 4096;
        if (cap_nodes > (1u << 20)) cap_nodes = 1u << 20;
        // per tile: count u32, base u64; per 16-byte chunk: candidate mask u16
        const size_t tiles_bytes = (((size_t)ntiles * 4 + 15) & ~size_t(15)) + (size_t)ntiles * 8;
        B2_CUDA(c->idx_tiles.ensure(tiles_bytes + (size_t)ntiles * (65536 / 16) * 2 + 64));
        B2_CUDA(c->idx_pos.ensure((size_t)cap_nodes * 8));
        uint32_t* tile_count = c->idx_tiles.as<uint32_t>();
        uint64_t* tile_base = reinterpret_cast<uint64_t*>(c->idx_tiles.as<uint8_t>() + (((size_t)ntiles * 4 + 15) & ~size_t(15)));
        uint16_t* masks = reinterpret_cast<uint16_t*>(c->idx_tiles.as<uint8_t>() + ((tiles_bytes + 15) & ~size_t(15)));
        B2_CUDA(launch_index_candidates(src, n, start, bound, bc ? 1 : 0, tile_count, tile_base, masks,
                                        c->idx_pos.as<uint64_t>(), cap_nodes, c->d_idx_nodes(), s));
        B2_CUDA(cudaMemcpyAsync(&c->h()->idx_nodes, c->d_idx_nodes(), 8, cudaMemcpyDeviceToHost, s));
        B2_CUDA(cudaStreamSynchronize(s));
        const uint64_t nn = c->h()->idx_nodes;
        if (nn <= cap_nodes) {
            const uint32_t levels = index_levels((uint32_t)nn);
            B2_CUDA(c->idx_jump.ensure((size_t)levels * (nn + 1) * 4));
            B2_CUDA(c->walk_off.ensure((size_t)(nn + 1) * 8));
            B2_CUDA(c->walk_hdr.ensure((size_t)(nn + 1) * 4));
            B2_CUDA(launch_index_resolve(src, n, start, bound, bc ? 1 : 0, c->idx_pos.as<uint64_t>(), (uint32_t)nn,
                                         c->idx_jump.as<uint32_t>(), c->walk_off.as<uint64_t>(), c->walk_hdr.as<uint32_t>(),
                                         (uint32_t)nn, c->d_walk(), s));
            B2_CUDA(cudaMemcpyAsync(&c->h()->walk, c->d_walk(), sizeof(WalkResult), cudaMemcpyDeviceToHost, s));
            B2_CUDA(cudaStreamSynchronize(s));
            if (c->h()->walk.terminal != 3) { *wout = c->h()->walk; return B2LZ4_OK; }
        }
    }
    uint64_t capacity = (n - (n > start ? start : n)) / 256 + 1024;
    if (capacity > 0x7FFFFFFFull) capacity = 0x7FFFFFFFull;
    for (int attempt = 0; attempt < 2; attempt++) {
        B2_CUDA(c->walk_off.ensure((size_t)capacity * 8));
        B2_CUDA(c->walk_hdr.ensure((size_t)capacity * 4));
        B2_CUDA(launch_walk(src, n, start, bc ? 1 : 0, c->walk_off.as<uint64_t>(), c->walk_hdr.as<uint32_t>(), (uint32_t)capacity,
                            c->d_walk(), s));
        B2_CUDA(cudaMemcpyAsync(&c->h()->walk, c->d_walk(), sizeof(WalkResult), cudaMemcpyDeviceToHost, s));
        B2_CUDA(cudaStreamSynchronize(s));
        if (c->h()->walk.nblocks <= capacity) break;
        capacity = c->h()->walk.nblocks;
    }
    *wout = c->h()->walk;
    return B2LZ4_OK;
}

}  // namespace
int b2_decompress_dev_impl(b2lz4_ctx* c, const void* srcv, size_t n, void* dstv, size_t cap, size_t* out, cudaStream_t s) {
    if (!out) return B2LZ4F_ERR_PARAMETER_NULL;
    *out = 0;
    if (!s) s = c->stream;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    OrderGuard og(c, s);
    B2_CUDA(cudaSetDevice(c->device));
    const uint8_t* src = (const uint8_t*)srcv;
    uint8_t* dst = (uint8_t*)dstv;
    Timer T{c, s, c->timing};
    T.mark(0);
    // frame header: <= 19 bytes, parsed on the host (src/lz4f.zig:547)
    uint8_t hb[32];
    const size_t hn = n < 19 ? n : 19;
    B2_CUDA(c->pin_aux.ensure(64));
    if (hn) {
        B2_CUDA(cudaMemcpyAsync(c->pin_aux.p, src, hn, cudaMemcpyDeviceToHost, s));
        B2_CUDA(cudaStreamSynchronize(s));
        memcpy(hb, c->pin_aux.p, hn);
    }
    b2lz4f_prefs info; size_t hsize = 0;
    int rc = b2lz4f_parse_frame_header(hb, hn, &info, &hsize);
    if (rc) return rc;
    size_t bs; block_size_of(info.block_size_id, bs);
    const bool bc = info.block_checksum == 1, cc = info.content_checksum == 1;
    // K7': the block index (parallel; SURVEY F12 describes the serial chain it replaces)
    WalkResult w;
    rc = build_block_index(c, src, n, hsize, (uint32_t)bs, bc, s, &w);
    if (rc) return rc;
    T.mark(1);
    uint64_t total = 0;
    rc = decode_blocks_dev(c, src, n, c->walk_off.as<uint64_t>(), c->walk_hdr.as<uint32_t>(), w.nblocks, w.terminal, dst, cap,
                           (uint32_t)bs, bc, &total, s, T);
    if (rc) return rc;
    T.mark(4);
    if (cc) {  // src/lz4f.zig:625-635
        if (w.end_pos + 4 > n) return B2LZ4F_ERR_FRAME_SIZE_WRONG;
        B2_CUDA(launch_xxh32_init(c->d_xxh(), 0, s));
        B2_CUDA(launch_xxh32_update(c->d_xxh(), dst, total, s));
        B2_CUDA(launch_xxh32_final(c->d_xxh(), c->d_content_sum(), s));
        B2_CUDA(cudaMemcpyAsync(&c->h()->content_sum, c->d_content_sum(), 4, cudaMemcpyDeviceToHost, s));
        B2_CUDA(cudaMemcpyAsync(c->pin_aux.p, src + w.end_pos, 4, cudaMemcpyDeviceToHost, s));
    }
    T.mark(5);
    if (cc || c->timing) B2_CUDA(cudaStreamSynchronize(s));   // (without a content checksum the summary read-back was the last device work)
    if (c->timing) {
        cudaEventElapsedTime(&c->phase_ms[2], c->ev_t[0], c->ev_t[1]);  // header + walk
        cudaEventElapsedTime(&c->phase_ms[1], c->ev_t[1], c->ev_t[2]);  // block checksums
        cudaEventElapsedTime(&c->phase_ms[0], c->ev_t[2], c->ev_t[3]);  // decode kernel
        cudaEventElapsedTime(&c->phase_ms[3], c->ev_t[4], c->ev_t[5]);  // content checksum
        cudaEventElapsedTime(&c->phase_ms[4], c->ev_t[0], c->ev_t[5]);
    }
    if (cc && rd32h((const uint8_t*)c->pin_aux.p) != c->h()->content_sum) return B2LZ4F_ERR_CONTENT_CHECKSUM_INVALID;
    *out = total;
    return B2LZ4_OK;
}

extern "C" {

int b2lz4f_compress_frame_dev(b2lz4_ctx* c, const void* src, size_t n, void* dst, size_t cap, const b2lz4f_prefs* prefs,
                              size_t* out, void* stream) {
    if (!c) return B2LZ4F_ERR_PARAMETER_NULL;
    return b2_compress_dev_impl(c, src, n, dst, cap, prefs, out, (cudaStream_t)stream, false);
}
int b2lz4f_compress_blocks_dev(b2lz4_ctx* c, const void* src, size_t n, void* dst, size_t cap, const b2lz4f_prefs* prefs,
                               size_t* out, void* stream) {
    if (!c) return B2LZ4F_ERR_PARAMETER_NULL;
    return b2_compress_dev_impl(c, src, n, dst, cap, prefs, out, (cudaStream_t)stream, true);
}
int b2lz4f_decompress_frame_dev(b2lz4_ctx* c, const void* src, size_t n, void* dst, size_t cap, size_t* out, void* stream) {
    if (!c) return B2LZ4F_ERR_PARAMETER_NULL;
    return b2_decompress_dev_impl(c, src, n, dst, cap, out, (cudaStream_t)stream);
}

int b2lz4f_decompress_blocks_dev(b2lz4_ctx* c, const void* srcv, size_t n, void* dst, size_t cap, size_t block_size,
                                 int block_checksum, size_t* out, void* stream) {
    if (!c || !out) return B2LZ4F_ERR_PARAMETER_NULL;
    *out = 0;
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    OrderGuard og(c, s);
    B2_CUDA(cudaSetDevice(c->device));
    const uint8_t* src = (const uint8_t*)srcv;
    Timer T{c, s, c->timing};
    T.mark(0);
    WalkResult w;
    { int rc0 = build_block_index(c, src, n, 0, (uint32_t)block_size, block_checksum != 0, s, &w); if (rc0) return rc0; }
    T.mark(1);
    uint64_t total = 0;
    // a body has no end mark: running off the end (terminal 1) is the normal exit
    int rc = decode_blocks_dev(c, src, n, c->walk_off.as<uint64_t>(), c->walk_hdr.as<uint32_t>(), w.nblocks, w.terminal, (uint8_t*)dst,
                               cap, (uint32_t)block_size, block_checksum != 0, &total, s, T);
    T.mark(4); T.mark(5);
    if (c->timing) {
        cudaStreamSynchronize(s);
        cudaEventElapsedTime(&c->phase_ms[2], c->ev_t[0], c->ev_t[1]);
        cudaEventElapsedTime(&c->phase_ms[1], c->ev_t[1], c->ev_t[2]);
        cudaEventElapsedTime(&c->phase_ms[0], c->ev_t[2], c->ev_t[3]);
        c->phase_ms[3] = 0;
        cudaEventElapsedTime(&c->phase_ms[4], c->ev_t[0], c->ev_t[5]);
    }
    if (rc) return rc;
    *out = total;
    return B2LZ4_OK;
}

int b2lz4f_index_frame_dev(b2lz4_ctx* c, const void* srcv, size_t n, uint64_t* off_dev, uint32_t* hdr_dev, size_t capacity,
                           b2lz4f_frame_index* info, void* stream) {
    if (!c || !info) return B2LZ4F_ERR_PARAMETER_NULL;
    memset(info, 0, sizeof *info);
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    OrderGuard og(c, s);
    B2_CUDA(cudaSetDevice(c->device));
    const uint8_t* src = (const uint8_t*)srcv;
    uint8_t hb[32];
    const size_t hn = n < 19 ? n : 19;
    B2_CUDA(c->pin_aux.ensure(64));
    if (hn) {
        B2_CUDA(cudaMemcpyAsync(c->pin_aux.p, src, hn, cudaMemcpyDeviceToHost, s));
        B2_CUDA(cudaStreamSynchronize(s));
        memcpy(hb, c->pin_aux.p, hn);
    }
    b2lz4f_prefs fi; size_t hsize = 0;
    int rc = b2lz4f_parse_frame_header(hb, hn, &fi, &hsize);
    if (rc) return rc;
    size_t bs; block_size_of(fi.block_size_id, bs);
    WalkResult w;
    rc = build_block_index(c, src, n, hsize, (uint32_t)bs, fi.block_checksum == 1, s, &w);
    if (rc) return rc;
    info->nblocks = w.nblocks; info->end_pos = w.end_pos; info->content_size = fi.content_size; info->terminal = w.terminal;
    info->header_size = (uint32_t)hsize; info->block_size = (uint32_t)bs; info->block_checksum = fi.block_checksum == 1;
    info->content_checksum = fi.content_checksum == 1; info->max_stored = w.max_stored;
    const size_t take = std::min<size_t>(capacity, w.nblocks);
    if (take && off_dev) B2_CUDA(cudaMemcpyAsync(off_dev, c->walk_off.p, take * 8, cudaMemcpyDeviceToDevice, s));
    if (take && hdr_dev) B2_CUDA(cudaMemcpyAsync(hdr_dev, c->walk_hdr.p, take * 4, cudaMemcpyDeviceToDevice, s));
    return B2LZ4_OK;
}

// ================================================================ batch API (device pointers)
int b2lz4_compress_fast_batch_dev(b2lz4_ctx* c, const void* src, const uint64_t* src_off, const uint32_t* src_len, void* dst,
                                  const uint64_t* dst_off, const uint32_t* dst_cap, uint32_t* out_len, int32_t* status,
                                  size_t nblocks, uint32_t accel, void* stream) {
    if (!c) return B2LZ4F_ERR_PARAMETER_NULL;
    if (nblocks > 0x7FFFFFFFull) return B2LZ4_ERR_INPUT_TOO_LARGE;
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    OrderGuard og(c, s);
    B2_CUDA(cudaSetDevice(c->device));
    // block lengths live on the device: use the wide (u32) hash table, valid for every block size
    B2_CUDA(launch_compress_fast(explicit_in(src, src_off, src_len), explicit_out(dst, dst_off, dst_cap), out_len, status,
                                 (uint32_t)nblocks, 0xFFFFFFFFu, accel, c->d_ticket(), c->num_sms, s));
    return B2LZ4_OK;
}

int b2lz4_compress_fast_dict_batch_dev(b2lz4_ctx* c, const void* src, const uint64_t* src_off, const uint32_t* src_len, void* dst,
                                       const uint64_t* dst_off, const uint32_t* dst_cap, uint32_t* out_len, int32_t* status,
                                       size_t nblocks, const void* dict, size_t dict_len, uint32_t accel, void* stream) {
    if (!c) return B2LZ4F_ERR_PARAMETER_NULL;
    if (nblocks > 0x7FFFFFFFull) return B2LZ4_ERR_INPUT_TOO_LARGE;
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    OrderGuard og(c, s);
    B2_CUDA(cudaSetDevice(c->device));
    B2_CUDA(c->dict_table.ensure(4096 * 4));
    B2_CUDA(launch_compress_fast_dict(explicit_in(src, src_off, src_len), explicit_out(dst, dst_off, dst_cap), out_len, status,
                                      (uint32_t)nblocks, (const uint8_t*)dict, dict ? dict_len : 0, c->dict_table.as<uint32_t>(),
                                      accel, c->d_ticket(), c->num_sms, s));
    return B2LZ4_OK;
}

int b2lz4_decompress_safe_batch_dev(b2lz4_ctx* c, const void* src, const uint64_t* src_off, const uint32_t* src_len, void* dst,
                                    const uint64_t* dst_off, const uint32_t* dst_cap, uint32_t* out_len, int32_t* status,
                                    size_t nblocks, const void* dict, size_t dict_len, void* stream) {
    if (!c) return B2LZ4F_ERR_PARAMETER_NULL;
    if (nblocks > 0x7FFFFFFFull || dict_len > 0xFFFFFFFFull) return B2LZ4_ERR_INPUT_TOO_LARGE;
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    OrderGuard og(c, s);
    B2_CUDA(cudaSetDevice(c->device));
    B2_CUDA(launch_decompress(explicit_in(src, src_off, src_len), explicit_out(dst, dst_off, dst_cap), nullptr, out_len, status,
                              (uint32_t)nblocks, (const uint8_t*)dict, (uint32_t)dict_len, c->d_ticket(), c->num_sms, s));
    return B2LZ4_OK;
}

int b2lz4_compress_hc_batch_dev(b2lz4_ctx* c, const void* src, const uint64_t* src_off, const uint32_t* src_len, void* dst,
                                const uint64_t* dst_off, const uint32_t* dst_cap, uint32_t* out_len, int32_t* status,
                                size_t nblocks, int level, void* stream) {
    if (!c) return B2LZ4F_ERR_PARAMETER_NULL;
    if (nblocks > 0x7FFFFFFFull) return B2LZ4_ERR_INPUT_TOO_LARGE;
    int nbs = hc_nb_searches(level);
    if (nbs < 0) return B2LZ4_ERR_UNSUPPORTED_LEVEL;
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    OrderGuard og(c, s);
    B2_CUDA(cudaSetDevice(c->device));
    { int rc = ensure_hc_work(c, s, (uint32_t)nblocks); if (rc) return rc; }
    B2_CUDA(launch_compress_hc(explicit_in(src, src_off, src_len), explicit_out(dst, dst_off, dst_cap), out_len, status,
                               (uint32_t)nblocks, nbs, c->hc_work.as<uint8_t>(), c->d_ticket(), c->num_sms, s));
    return B2LZ4_OK;
}

// compressDestSize for many blocks (reference src/lz4.zig:551-616): every probe of the reference's bisection is one
// K1 launch over all blocks with per-block prefix lengths and dst's real capacity (K1 never writes past it and
// reports OutputTooSmall exactly when the prefix does not fit); the chosen prefixes are compressed last, so
// dst[0..out_len) is always compressDefault(src[0..consumed)).  max_len bounds src_len[] (0 = unknown).
static int dest_size_dev(b2lz4_ctx* c, const void* src, const uint64_t* src_off, const uint32_t* src_len, void* dst,
                         const uint64_t* dst_off, const uint32_t* dst_cap, uint32_t* consumed, uint32_t* out_len,
                         int32_t* status, size_t nb, uint32_t max_len, cudaStream_t s) {
    if (nb == 0) return B2LZ4_OK;
    B2_CUDA(c->ds_work.ensure(nb * (sizeof(DestSizeState) + 12) + 64));
    DestSizeState* st = c->ds_work.as<DestSizeState>();
    uint32_t* probe_len = reinterpret_cast<uint32_t*>(st + nb);
    uint32_t* probe_out = probe_len + nb;
    int32_t* probe_status = reinterpret_cast<int32_t*>(probe_out + nb);
    const uint32_t n32 = (uint32_t)nb;
    const uint32_t table_len = max_len ? max_len : 0xFFFFFFFFu;
    // the estimate probe + one bisection probe per bit of the longest block (+1: the interval [1, max] has max values)
    uint32_t bits = 0;
    for (uint32_t v = max_len ? max_len : 0x7E000000u; v; v >>= 1) bits++;
    const uint32_t probes = bits + 2;
    B2_CUDA(launch_dest_size_init(src_len, dst_cap, st, probe_len, n32, s));
    for (uint32_t it = 0; it < probes; it++) {
        B2_CUDA(launch_compress_fast(explicit_in(src, src_off, probe_len), explicit_out(dst, dst_off, dst_cap), probe_out,
                                     probe_status, n32, table_len, 1, c->d_ticket(), c->num_sms, s));
        B2_CUDA(launch_dest_size_step(src_len, probe_status, st, probe_len, n32, s));
    }
    B2_CUDA(launch_dest_size_best(st, probe_len, n32, s));
    B2_CUDA(launch_compress_fast(explicit_in(src, src_off, probe_len), explicit_out(dst, dst_off, dst_cap), out_len, status,
                                 n32, table_len, 1, c->d_ticket(), c->num_sms, s));
    B2_CUDA(launch_dest_size_finish(st, consumed, out_len, status, n32, s));
    return B2LZ4_OK;
}

int b2lz4_compress_dest_size_batch_dev(b2lz4_ctx* c, const void* src, const uint64_t* src_off, const uint32_t* src_len,
                                       void* dst, const uint64_t* dst_off, const uint32_t* dst_cap, uint32_t* consumed,
                                       uint32_t* out_len, int32_t* status, size_t nblocks, uint32_t max_src_len,
                                       void* stream) {
    if (!c) return B2LZ4F_ERR_PARAMETER_NULL;
    if (nblocks > 0x7FFFFFFFull) return B2LZ4_ERR_INPUT_TOO_LARGE;
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    OrderGuard og(c, s);
    B2_CUDA(cudaSetDevice(c->device));
    return dest_size_dev(c, src, src_off, src_len, dst, dst_off, dst_cap, consumed, out_len, status, nblocks, max_src_len, s);
}

int b2lz4_xxh32_dev(b2lz4_ctx* c, const void* src, size_t n, uint32_t seed, uint32_t* out_dev, void* stream) {
    if (!c || !out_dev) return B2LZ4F_ERR_PARAMETER_NULL;
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    OrderGuard og(c, s);
    B2_CUDA(cudaSetDevice(c->device));
    B2_CUDA(launch_xxh32_init(c->d_xxh(), seed, s));
    B2_CUDA(launch_xxh32_update(c->d_xxh(), (const uint8_t*)src, n, s));
    B2_CUDA(launch_xxh32_final(c->d_xxh(), out_dev, s));
    return B2LZ4_OK;
}

}  // extern "C"

// ================================================================ host-pointer batch / block API
namespace {

enum class Op { Fast, Decode, HC, FastDict };

// Stages a host batch through the device: one H2D of the byte span the blocks cover, the kernel, one
// D2H of the span the outputs cover.
static int host_batch(b2lz4_ctx* c, Op op, int param, const void* srcv, const uint64_t* src_off, const uint32_t* src_len,
                      void* dstv, const uint64_t* dst_off, const uint32_t* dst_cap, uint32_t* out_len, int32_t* status,
                      size_t nb, const void* dict, size_t dict_len) {
    if (!c) return B2LZ4F_ERR_PARAMETER_NULL;
    if (nb == 0) return B2LZ4_OK;
    if (nb > 0x7FFFFFFFull) return B2LZ4_ERR_INPUT_TOO_LARGE;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    b2::wait_previous(c);
    B2_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    uint64_t s_lo = ~0ull, s_hi = 0, d_lo = ~0ull, d_hi = 0;
    uint32_t max_len = 0;
    for (size_t i = 0; i < nb; i++) {
        s_lo = std::min(s_lo, src_off[i]); s_hi = std::max(s_hi, src_off[i] + src_len[i]);
        d_lo = std::min(d_lo, dst_off[i]); d_hi = std::max(d_hi, dst_off[i] + dst_cap[i]);
        max_len = std::max(max_len, src_len[i]);
    }
    const size_t s_span = (size_t)(s_hi - s_lo), d_span = (size_t)(d_hi - d_lo);
    B2_CUDA(c->stage_in[0].ensure(s_span + 16));
    B2_CUDA(c->stage_out[0].ensure(d_span + 16));
    // aux: src_off | dst_off (u64) | src_len | dst_cap | out_len | status (u32) | dict
    const size_t aux_bytes = nb * (8 + 8 + 4 + 4 + 4 + 4) + ((dict_len + 15) & ~size_t(15)) + 64;
    B2_CUDA(c->stage_aux.ensure(aux_bytes));
    uint8_t* a = c->stage_aux.as<uint8_t>();
    uint64_t* d_soff = (uint64_t*)a;
    uint64_t* d_doff = d_soff + nb;
    uint32_t* d_slen = (uint32_t*)(d_doff + nb);
    uint32_t* d_dcap = d_slen + nb;
    uint32_t* d_olen = d_dcap + nb;
    int32_t* d_stat = (int32_t*)(d_olen + nb);
    uint8_t* d_dict = (uint8_t*)(((uintptr_t)(d_stat + nb) + 15) & ~uintptr_t(15));
    std::vector<uint64_t> so(nb), dofs(nb);
    for (size_t i = 0; i < nb; i++) { so[i] = src_off[i] - s_lo; dofs[i] = dst_off[i] - d_lo; }
    B2_CUDA(cudaMemcpyAsync(d_soff, so.data(), nb * 8, cudaMemcpyHostToDevice, s));
    B2_CUDA(cudaMemcpyAsync(d_doff, dofs.data(), nb * 8, cudaMemcpyHostToDevice, s));
    B2_CUDA(cudaMemcpyAsync(d_slen, src_len, nb * 4, cudaMemcpyHostToDevice, s));
    B2_CUDA(cudaMemcpyAsync(d_dcap, dst_cap, nb * 4, cudaMemcpyHostToDevice, s));
    if (s_span) B2_CUDA(cudaMemcpyAsync(c->stage_in[0].p, (const uint8_t*)srcv + s_lo, s_span, cudaMemcpyHostToDevice, s));
    if (dict && dict_len) B2_CUDA(cudaMemcpyAsync(d_dict, dict, dict_len, cudaMemcpyHostToDevice, s));
    BlockSet in = explicit_in(c->stage_in[0].p, d_soff, d_slen);
    OutSet out = explicit_out(c->stage_out[0].p, d_doff, d_dcap);
    if (op == Op::Fast) {
        B2_CUDA(launch_compress_fast(in, out, d_olen, d_stat, (uint32_t)nb, max_len, (uint32_t)param, c->d_ticket(), c->num_sms, s));
    } else if (op == Op::FastDict) {
        B2_CUDA(c->dict_table.ensure(4096 * 4));
        B2_CUDA(launch_compress_fast_dict(in, out, d_olen, d_stat, (uint32_t)nb, (dict && dict_len) ? d_dict : nullptr,
                                          dict ? dict_len : 0, c->dict_table.as<uint32_t>(), (uint32_t)param, c->d_ticket(),
                                          c->num_sms, s));
    } else if (op == Op::Decode) {
        B2_CUDA(launch_decompress(in, out, nullptr, d_olen, d_stat, (uint32_t)nb, dict ? d_dict : nullptr, (uint32_t)dict_len,
                                  c->d_ticket(), c->num_sms, s));
    } else {
        { int rc = ensure_hc_work(c, s, (uint32_t)nb); if (rc) return rc; }
        B2_CUDA(launch_compress_hc(in, out, d_olen, d_stat, (uint32_t)nb, param, c->hc_work.as<uint8_t>(), c->d_ticket(),
                                   c->num_sms, s));
    }
    B2_CUDA(cudaMemcpyAsync(out_len, d_olen, nb * 4, cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaMemcpyAsync(status, d_stat, nb * 4, cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    // copy back only what each block produced (the caller's other dst bytes stay untouched): a few blocks one
    // by one; many blocks as one download of the whole span into pinned memory + host memcpys (one
    // cudaMemcpyAsync per block costs ~5 us, 0.3 s for 65536 records)
    if (nb <= 64) {
        for (size_t i = 0; i < nb; i++) {
            if (status[i] == 0 && out_len[i])
                B2_CUDA(cudaMemcpyAsync((uint8_t*)dstv + dst_off[i], c->stage_out[0].as<uint8_t>() + dofs[i], out_len[i],
                                        cudaMemcpyDeviceToHost, s));
        }
        B2_CUDA(cudaStreamSynchronize(s));
    } else {
        B2_CUDA(c->pin_aux.ensure(d_span + 64));
        B2_CUDA(cudaMemcpyAsync(c->pin_aux.p, c->stage_out[0].p, d_span, cudaMemcpyDeviceToHost, s));
        B2_CUDA(cudaStreamSynchronize(s));
        const uint8_t* t = c->pin_aux.as<uint8_t>();
        for (size_t i = 0; i < nb; i++)
            if (status[i] == 0 && out_len[i]) memcpy((uint8_t*)dstv + dst_off[i], t + dofs[i], out_len[i]);
    }
    return B2LZ4_OK;
}

// compressDestSize over host buffers: src_len[i] is the most block i may consume, dst_cap[i] the room it has.
static int host_dest_size_batch(b2lz4_ctx* c, const void* srcv, const uint64_t* src_off, const uint32_t* src_len, void* dstv,
                                const uint64_t* dst_off, const uint32_t* dst_cap, uint32_t* consumed, uint32_t* out_len,
                                int32_t* status, size_t nb) {
    if (!c) return B2LZ4F_ERR_PARAMETER_NULL;
    if (nb == 0) return B2LZ4_OK;
    if (nb > 0x7FFFFFFFull) return B2LZ4_ERR_INPUT_TOO_LARGE;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    b2::wait_previous(c);
    B2_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    uint64_t s_lo = ~0ull, s_hi = 0, d_lo = ~0ull, d_hi = 0;
    uint32_t max_len = 0;
    for (size_t i = 0; i < nb; i++) {
        s_lo = std::min(s_lo, src_off[i]); s_hi = std::max(s_hi, src_off[i] + src_len[i]);
        d_lo = std::min(d_lo, dst_off[i]); d_hi = std::max(d_hi, dst_off[i] + dst_cap[i]);
        max_len = std::max(max_len, src_len[i]);
    }
    const size_t s_span = (size_t)(s_hi - s_lo), d_span = (size_t)(d_hi - d_lo);
    B2_CUDA(c->stage_in[0].ensure(s_span + 16));
    B2_CUDA(c->stage_out[0].ensure(d_span + 16));
    B2_CUDA(c->stage_aux.ensure(nb * (8 + 8 + 4 * 5) + 64));
    uint64_t* d_soff = c->stage_aux.as<uint64_t>();
    uint64_t* d_doff = d_soff + nb;
    uint32_t* d_slen = (uint32_t*)(d_doff + nb);
    uint32_t* d_dcap = d_slen + nb;
    uint32_t* d_used = d_dcap + nb;
    uint32_t* d_olen = d_used + nb;
    int32_t* d_stat = (int32_t*)(d_olen + nb);
    std::vector<uint64_t> so(nb), dofs(nb);
    for (size_t i = 0; i < nb; i++) { so[i] = src_off[i] - s_lo; dofs[i] = dst_off[i] - d_lo; }
    B2_CUDA(cudaMemcpyAsync(d_soff, so.data(), nb * 8, cudaMemcpyHostToDevice, s));
    B2_CUDA(cudaMemcpyAsync(d_doff, dofs.data(), nb * 8, cudaMemcpyHostToDevice, s));
    B2_CUDA(cudaMemcpyAsync(d_slen, src_len, nb * 4, cudaMemcpyHostToDevice, s));
    B2_CUDA(cudaMemcpyAsync(d_dcap, dst_cap, nb * 4, cudaMemcpyHostToDevice, s));
    if (s_span) B2_CUDA(cudaMemcpyAsync(c->stage_in[0].p, (const uint8_t*)srcv + s_lo, s_span, cudaMemcpyHostToDevice, s));
    { int rc = dest_size_dev(c, c->stage_in[0].p, d_soff, d_slen, c->stage_out[0].p, d_doff, d_dcap, d_used, d_olen, d_stat, nb,
                             max_len, s); if (rc) return rc; }
    B2_CUDA(cudaMemcpyAsync(consumed, d_used, nb * 4, cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaMemcpyAsync(out_len, d_olen, nb * 4, cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaMemcpyAsync(status, d_stat, nb * 4, cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    if (nb <= 64) {
        for (size_t i = 0; i < nb; i++)
            if (status[i] == 0 && out_len[i])
                B2_CUDA(cudaMemcpyAsync((uint8_t*)dstv + dst_off[i], c->stage_out[0].as<uint8_t>() + dofs[i], out_len[i],
                                        cudaMemcpyDeviceToHost, s));
        B2_CUDA(cudaStreamSynchronize(s));
    } else {
        B2_CUDA(c->pin_aux.ensure(d_span + 64));
        B2_CUDA(cudaMemcpyAsync(c->pin_aux.p, c->stage_out[0].p, d_span, cudaMemcpyDeviceToHost, s));
        B2_CUDA(cudaStreamSynchronize(s));
        const uint8_t* t = c->pin_aux.as<uint8_t>();
        for (size_t i = 0; i < nb; i++)
            if (status[i] == 0 && out_len[i]) memcpy((uint8_t*)dstv + dst_off[i], t + dofs[i], out_len[i]);
    }
    return B2LZ4_OK;
}

static int host_single(Op op, int param, const void* src, size_t n, void* dst, size_t cap, const void* dict, size_t dict_len,
                       size_t* out) {
    if (!out) return B2LZ4F_ERR_PARAMETER_NULL;
    *out = 0;
    if (op != Op::Decode && n > LZ4_MAX_INPUT_SIZE) return B2LZ4_ERR_INPUT_TOO_LARGE;  // src/lz4.zig:296, lz4hc.zig:1442
    if (n > 0xFFFFFFFFull) return B2LZ4_ERR_INPUT_TOO_LARGE;
    b2lz4_ctx* c; int rc = b2_default_ctx(&c); if (rc) return rc;
    uint64_t so = 0, dofs = 0;
    uint32_t sl = (uint32_t)n, dc = (uint32_t)std::min<size_t>(cap, 0xFFFFFFFFull), ol = 0; int32_t st = 0;
    rc = host_batch(c, op, param, src, &so, &sl, dst, &dofs, &dc, &ol, &st, 1, dict, dict_len);
    if (rc) return rc;
    *out = ol;
    return st;
}

}  // namespace

extern "C" {

int b2lz4_compress_fast_batch(b2lz4_ctx* c, const void* src, const uint64_t* so, const uint32_t* sl, void* dst, const uint64_t* dofs,
                              const uint32_t* dc, uint32_t* ol, int32_t* st, size_t nb, uint32_t accel) {
    return host_batch(c, Op::Fast, (int)accel, src, so, sl, dst, dofs, dc, ol, st, nb, nullptr, 0);
}
int b2lz4_compress_fast_dict_batch(b2lz4_ctx* c, const void* src, const uint64_t* so, const uint32_t* sl, void* dst, const uint64_t* dofs,
                                   const uint32_t* dc, uint32_t* ol, int32_t* st, size_t nb, const void* dict, size_t dict_len,
                                   uint32_t accel) {
    return host_batch(c, Op::FastDict, (int)accel, src, so, sl, dst, dofs, dc, ol, st, nb, dict, dict ? dict_len : 0);
}
int b2lz4_decompress_safe_batch(b2lz4_ctx* c, const void* src, const uint64_t* so, const uint32_t* sl, void* dst, const uint64_t* dofs,
                                const uint32_t* dc, uint32_t* ol, int32_t* st, size_t nb, const void* dict, size_t dict_len) {
    return host_batch(c, Op::Decode, 0, src, so, sl, dst, dofs, dc, ol, st, nb, dict, dict_len);
}
int b2lz4_compress_hc_batch(b2lz4_ctx* c, const void* src, const uint64_t* so, const uint32_t* sl, void* dst, const uint64_t* dofs,
                            const uint32_t* dc, uint32_t* ol, int32_t* st, size_t nb, int level) {
    int nbs = hc_nb_searches(level);
    if (nbs < 0) return B2LZ4_ERR_UNSUPPORTED_LEVEL;
    return host_batch(c, Op::HC, nbs, src, so, sl, dst, dofs, dc, ol, st, nb, nullptr, 0);
}

int b2lz4_compress_dest_size_batch(b2lz4_ctx* c, const void* src, const uint64_t* so, const uint32_t* sl, void* dst,
                                   const uint64_t* dofs, const uint32_t* dc, uint32_t* consumed, uint32_t* ol, int32_t* st,
                                   size_t nb) {
    return host_dest_size_batch(c, src, so, sl, dst, dofs, dc, consumed, ol, st, nb);
}
int b2lz4_compress_dest_size(const void* src, void* dst, size_t cap, size_t* src_size, size_t* out) {  // src/lz4.zig:551-616
    if (!out || !src_size) return B2LZ4F_ERR_PARAMETER_NULL;
    *out = 0;
    const size_t max_len = *src_size;
    if (max_len == 0) return B2LZ4_OK;                                  // :553-556 (*src_size stays 0)
    if (max_len > LZ4_MAX_INPUT_SIZE) return B2LZ4_ERR_INPUT_TOO_LARGE;  // :559-561 -> :296; *src_size unchanged
    b2lz4_ctx* c; int rc = b2_default_ctx(&c); if (rc) return rc;
    uint64_t so = 0, dofs = 0;
    uint32_t sl = (uint32_t)max_len, dc = (uint32_t)std::min<size_t>(cap, 0xFFFFFFFFull), used = 0, ol = 0; int32_t st = 0;
    rc = host_dest_size_batch(c, src, &so, &sl, dst, &dofs, &dc, &used, &ol, &st, 1);
    if (rc) return rc;
    if (st) return st;
    *src_size = used;
    *out = ol;
    return B2LZ4_OK;
}

int b2lz4_compress_fast(const void* src, size_t n, void* dst, size_t cap, uint32_t accel, size_t* out) {
    return host_single(Op::Fast, (int)accel, src, n, dst, cap, nullptr, 0, out);
}
int b2lz4_compress_default(const void* src, size_t n, void* dst, size_t cap, size_t* out) {  // src/lz4.zig:283-285
    return b2lz4_compress_fast(src, n, dst, cap, 1, out);
}
int b2lz4_decompress_safe(const void* src, size_t n, void* dst, size_t cap, size_t* out) {
    return host_single(Op::Decode, 0, src, n, dst, cap, nullptr, 0, out);
}
int b2lz4_decompress_safe_using_dict(const void* src, size_t n, void* dst, size_t cap, const void* dict, size_t dict_len, size_t* out) {
    static const uint8_t empty = 0;
    return host_single(Op::Decode, 0, src, n, dst, cap, dict ? dict : &empty, dict_len, out);
}
int b2lz4_compress_fast_using_dict(const void* src, size_t n, void* dst, size_t cap, const void* dict, size_t dict_len,
                                   uint32_t accel, size_t* out) {
    return host_single(Op::FastDict, (int)accel, src, n, dst, cap, dict, dict ? dict_len : 0, out);
}
int b2lz4_compress_hc(const void* src, size_t n, void* dst, size_t cap, int level, size_t* out) {
    if (!out) return B2LZ4F_ERR_PARAMETER_NULL;
    *out = 0;
    if (n > LZ4_MAX_INPUT_SIZE) return B2LZ4_ERR_INPUT_TOO_LARGE;  // src/lz4hc.zig:1442
    if (n == 0) return B2LZ4_OK;                                   // :1443
    if (cap == 0) return B2LZ4_ERR_OUTPUT_TOO_SMALL;               // :1461
    int nbs = hc_nb_searches(level);
    if (nbs < 0) return B2LZ4_ERR_UNSUPPORTED_LEVEL;
    return host_single(Op::HC, nbs, src, n, dst, cap, nullptr, 0, out);
}

int b2lz4_xxh32(const void* src, size_t n, uint32_t seed, uint32_t* out) {
    if (!out) return B2LZ4F_ERR_PARAMETER_NULL;
    b2lz4_ctx* c; int rc = b2_default_ctx(&c); if (rc) return rc;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    b2::wait_previous(c);
    B2_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    B2_CUDA(c->stage_in[0].ensure(n + 16));
    if (n) B2_CUDA(cudaMemcpyAsync(c->stage_in[0].p, src, n, cudaMemcpyHostToDevice, s));
    B2_CUDA(launch_xxh32_init(c->d_xxh(), seed, s));
    B2_CUDA(launch_xxh32_update(c->d_xxh(), c->stage_in[0].as<uint8_t>(), n, s));
    B2_CUDA(launch_xxh32_final(c->d_xxh(), c->d_content_sum(), s));
    B2_CUDA(cudaMemcpyAsync(&c->h()->content_sum, c->d_content_sum(), 4, cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    *out = c->h()->content_sum;
    return B2LZ4_OK;
}

// ---- running XXH32 state hand-off (multi-GPU content checksum)
void b2lz4_xxh32_state_init(b2lz4_xxh32_state* st, uint32_t seed) {
    if (!st) return;
    memset(st, 0, sizeof *st);
    st->v[0] = seed + 2654435761u + 2246822519u; st->v[1] = seed + 2246822519u; st->v[2] = seed; st->v[3] = seed - 2654435761u;
    st->seed = seed;
}
int b2lz4_xxh32_state_update_dev(b2lz4_ctx* c, b2lz4_xxh32_state* st, const void* src, size_t n, void* stream) {
    if (!c || !st) return B2LZ4F_ERR_PARAMETER_NULL;
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    OrderGuard og(c, s);
    B2_CUDA(cudaSetDevice(c->device));
    XxhState& x = c->h()->xxh;
    for (int i = 0; i < 4; i++) { x.v[i] = st->v[i]; x.tail[i] = rd32h(st->tail + 4 * i); }
    x.tail_len = st->tail_len; x.seed = st->seed; x.total_lo = (uint32_t)st->total; x.total_hi = (uint32_t)(st->total >> 32);
    B2_CUDA(cudaMemcpyAsync(c->d_xxh(), &x, sizeof x, cudaMemcpyHostToDevice, s));
    B2_CUDA(launch_xxh32_update(c->d_xxh(), (const uint8_t*)src, n, s));
    B2_CUDA(cudaMemcpyAsync(&x, c->d_xxh(), sizeof x, cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    for (int i = 0; i < 4; i++) { st->v[i] = x.v[i]; wr32h(st->tail + 4 * i, x.tail[i]); }
    st->tail_len = x.tail_len; st->total = ((uint64_t)x.total_hi << 32) | x.total_lo;
    return B2LZ4_OK;
}
uint32_t b2lz4_xxh32_state_final(const b2lz4_xxh32_state* st) {
    // a few integer ops on 40 bytes of state: the merge + avalanche of XXH32 (no payload bytes touched)
    const uint32_t P1 = 2654435761u, P2 = 2246822519u, P3 = 3266489917u, P4 = 668265263u, P5 = 374761393u;
    uint32_t h = st->total >= 16 ? rotl_h(st->v[0], 1) + rotl_h(st->v[1], 7) + rotl_h(st->v[2], 12) + rotl_h(st->v[3], 18)
                                 : st->seed + P5;
    h += (uint32_t)st->total;
    const uint8_t* p = st->tail; uint32_t n = st->tail_len;
    while (n >= 4) { h = rotl_h(h + rd32h(p) * P3, 17) * P4; p += 4; n -= 4; }
    while (n) { h = rotl_h(h + (uint32_t)(*p) * P5, 11) * P1; p++; n--; }
    h ^= h >> 15; h *= P2; h ^= h >> 13; h *= P3; h ^= h >> 16;
    return h;
}

}  // extern "C"
