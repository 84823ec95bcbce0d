// b2lz4_frame_host.cu — host-pointer frame entry points (the reference's own calling convention:
// caller-owned host slices, synchronous) and the README streaming trio, layered on the device path.
//   lz4f.compressFrame    /root/reference/src/lz4f.zig:354-446
//   lz4f.decompressFrame  /root/reference/src/lz4f.zig:541-638
//   compressBegin/Update/End + create/freeCompressionContext — README.md:98-122
// The host side only moves bytes (H2D / D2H) and keeps the < blockSize tail of a stream; all codec
// and checksum work runs in the kernels.
#include <algorithm>
#include <new>
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

static int upload(b2lz4_ctx* c, void* d, const void* h, size_t n, cudaStream_t s) {
    const uint8_t* hp = (const uint8_t*)h;
    uint8_t* dp = (uint8_t*)d;
    for (size_t o = 0; o < n; o += COPY_CHUNK) {
        size_t k = std::min(COPY_CHUNK, n - o);
        B2_CUDA(cudaMemcpyAsync(dp + o, hp + o, k, cudaMemcpyHostToDevice, s));
    }
    return B2LZ4_OK;
}
static int download(b2lz4_ctx* c, void* h, const void* d, size_t n, cudaStream_t s) {
    uint8_t* hp = (uint8_t*)h;
    const uint8_t* dp = (const uint8_t*)d;
    for (size_t o = 0; o < n; o += COPY_CHUNK) {
        size_t k = std::min(COPY_CHUNK, n - o);
        B2_CUDA(cudaMemcpyAsync(hp + o, dp + o, k, cudaMemcpyDeviceToHost, s));
    }
    return B2LZ4_OK;
}

}  // namespace

extern "C" {

int b2lz4f_compress_frame_ctx(b2lz4_ctx* c, const void* src, size_t n, void* dst, size_t cap, const b2lz4f_prefs* prefs,
                              size_t* out) {
    if (!c || !out) return B2LZ4F_ERR_PARAMETER_NULL;
    *out = 0;
    const size_t bound = b2lz4f_compress_frame_bound(n, prefs);
    if (cap < bound) return B2LZ4F_ERR_DST_MAX_SIZE_TOO_SMALL;   // src/lz4f.zig:363-366
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    B2_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    B2_CUDA(c->stage_in[0].ensure(n + 16));
    B2_CUDA(c->stage_out[0].ensure(bound + 16));
    int rc = upload(c, c->stage_in[0].p, src, n, s);
    if (rc) return rc;
    size_t produced = 0;
    rc = b2_compress_dev_impl(c, c->stage_in[0].p, n, c->stage_out[0].p, bound, prefs, &produced, s, false);
    if (rc) return rc;
    rc = download(c, dst, c->stage_out[0].p, produced, s);
    if (rc) return rc;
    B2_CUDA(cudaStreamSynchronize(s));
    *out = produced;
    return B2LZ4_OK;
}

int b2lz4f_decompress_frame_ctx(b2lz4_ctx* c, const void* src, size_t n, void* dst, size_t cap, size_t* out) {
    if (!c || !out) return B2LZ4F_ERR_PARAMETER_NULL;
    *out = 0;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    B2_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    B2_CUDA(c->stage_in[0].ensure(n + 16));
    B2_CUDA(c->stage_out[0].ensure(cap + 16));
    int rc = upload(c, c->stage_in[0].p, src, n, s);
    if (rc) return rc;
    size_t produced = 0;
    rc = b2_decompress_dev_impl(c, c->stage_in[0].p, n, c->stage_out[0].p, cap, &produced, s);
    if (rc) return rc;
    rc = download(c, dst, c->stage_out[0].p, produced, s);
    if (rc) return rc;
    B2_CUDA(cudaStreamSynchronize(s));
    *out = produced;
    return B2LZ4_OK;
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
