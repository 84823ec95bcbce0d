// b2_kernels.h — host-visible launchers of the CUDA kernels (internal; the public ABI is include/b2lz4.h).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "b2_common.cuh"

namespace b2 {

void count_launch();  // bumps the process-wide kernel launch counter (b2lz4_kernel_launch_count)

// Diagnostic knobs (b2lz4_debug_tune, include/b2lz4.h): process-wide, 0 = the shipped default.  They exist for the
// occupancy / variant experiments recorded in DESIGN.md and for tests that force a rare path; nothing reads the
// environment on a launch path.
struct Tune {
    int k1_ctas;        // K1: CTAs per SM (u16-table kernel), 1..9
    int k1_variant;     // K1: 2 = ring variant (one CTA per SM, forward input ring fed by TMA bulk copies)
    int k2_occ;         // K2: CTAs of 4 warps per SM: 8, 10, 12
    int k2_variant;     // K2: 1 = round-1 decoder (serial token walk), else the chunked decoder
    int k3_variant;     // K3 experiments
    int pipe_blocks;    // host-pointer frame pipeline: blocks per chunk (tests pipeline small frames with it)
    int no_pipeline;    // host-pointer frame calls take the one-shot path
    int serial_walk;    // frame index: force the serial header walk (K7) instead of the parallel index (K7')
    int xxh_variant;    // content checksum experiments
    int spare[8];
};
Tune& tune();

// K1 — fast compressor (k_compress_fast.cu)
cudaError_t launch_compress_fast(const BlockSet& in, const OutSet& out, uint32_t* out_len, int32_t* status,
                                 uint32_t nblocks, uint32_t max_len, uint32_t accel, uint32_t* ticket, int num_sms,
                                 cudaStream_t stream);

// the same over regular block sets with the blocks taken expensive-first (estimate pass + order: k_compress_fast.cu);
// scratch: order_scratch_bytes(nblocks) of device memory
size_t order_scratch_bytes(uint32_t nblocks);
cudaError_t launch_compress_fast_ordered(const BlockSet& in, const OutSet& out, uint32_t* out_len, int32_t* status, uint32_t nblocks,
                                         uint32_t max_len, uint32_t* ticket, int num_sms, void* scratch, cudaStream_t stream);

// fast compressor with a shared dictionary (k_compress_dict.cu); primed: 4096 u32 of device scratch
cudaError_t launch_compress_fast_dict(const BlockSet& in, const OutSet& out, uint32_t* out_len, int32_t* status,
                                      uint32_t nblocks, const uint8_t* dict, uint64_t dict_len, uint32_t* primed,
                                      uint32_t accel, uint32_t* ticket, int num_sms, cudaStream_t stream);

// compressDestSize search state, one per block (k_dest_size.cu; reference src/lz4.zig:551-616)
struct DestSizeState {
    uint32_t low, high;    // bisection interval over prefix lengths
    uint32_t best;         // bestSize: longest prefix that fitted so far
    uint32_t cur;          // prefix length of the probe in flight
    uint32_t phase;        // 0 done, 1 estimate probe, 2 bisection probe
    int32_t err;           // status of a block that cannot run (InputTooLarge)
};
cudaError_t launch_dest_size_init(const uint32_t* src_len, const uint32_t* dst_cap, DestSizeState* state,
                                  uint32_t* probe_len, uint32_t nblocks, cudaStream_t stream);
cudaError_t launch_dest_size_step(const uint32_t* src_len, const int32_t* probe_status, DestSizeState* state,
                                  uint32_t* probe_len, uint32_t nblocks, cudaStream_t stream);
cudaError_t launch_dest_size_best(const DestSizeState* state, uint32_t* probe_len, uint32_t nblocks, cudaStream_t stream);
cudaError_t launch_dest_size_finish(const DestSizeState* state, uint32_t* consumed, uint32_t* out_len, int32_t* status,
                                    uint32_t nblocks, cudaStream_t stream);

// K2 — decompressor (k_decompress.cu).  hdr: optional frame block headers (bit31 = stored raw).
// order_scratch (nblocks u32, optional) + block_size: frame blocks are decoded expensive-first (see k_order_heavy_first)
cudaError_t launch_decompress(const BlockSet& in, const OutSet& out, const uint32_t* hdr, uint32_t* out_len,
                              int32_t* status, uint32_t nblocks, const uint8_t* dict, uint32_t dict_len,
                              uint32_t* ticket, int num_sms, cudaStream_t stream, uint32_t* order_scratch = nullptr,
                              uint32_t block_size = 0);
// size-only parse (no output written): natural decoded size of each block, for foreign frames whose
// non-final blocks are not exactly blockSize
cudaError_t launch_decoded_size(const BlockSet& in, const uint32_t* hdr, uint32_t* out_len, int32_t* status,
                                uint32_t nblocks, int num_sms, cudaStream_t stream);

// K3 — HC hash-chain compressor (k_compress_hc.cu).  work: per-resident-warp tables (hc_work_bytes()).
size_t hc_work_bytes(int num_sms, uint32_t nblocks);
cudaError_t launch_compress_hc(const BlockSet& in, const OutSet& out, uint32_t* out_len, int32_t* status,
                               uint32_t nblocks, int nb_searches, uint8_t* work, uint32_t* ticket, int num_sms,
                               cudaStream_t stream);

// K4 — XXH32 (k_xxh32.cu)
// per-block checksum of the *stored* bytes: slot data if csize < raw len, else the raw block
cudaError_t launch_xxh32_stored(const BlockSet& slots, const BlockSet& raw, const uint32_t* csize, uint32_t* sums,
                                uint32_t nblocks, cudaStream_t stream);
// per-block checksum of explicit (off,len) ranges — frame decode verification
cudaError_t launch_xxh32_ranges(const uint8_t* base, const uint64_t* off, const uint32_t* hdr, uint32_t* sums,
                                uint32_t nblocks, cudaStream_t stream);
// serial running XXH32 over one buffer; state is 12 u32: v[4], tail[4 words], tail_len, seed, total lo/hi
struct XxhState {
    uint32_t v[4];
    uint32_t tail[4];
    uint32_t tail_len;
    uint32_t seed;
    uint32_t total_lo, total_hi;
};
cudaError_t launch_xxh32_init(XxhState* st, uint32_t seed, cudaStream_t stream);
cudaError_t launch_xxh32_update(XxhState* st, const uint8_t* p, uint64_t n, cudaStream_t stream);
cudaError_t launch_xxh32_final(const XxhState* st, uint32_t* out, cudaStream_t stream);

// K5/K6/K7 — frame scan, assembly, index walk (k_frame.cu)
struct FrameTotals {       // written by the device, read back once per call
    uint64_t body_bytes;   // sum of block records
    uint32_t first_bad;    // first block with status != 0 (0xFFFFFFFF if none)
    int32_t bad_status;
};
cudaError_t launch_scan_records(const uint32_t* csize, const int32_t* status, uint32_t nblocks, uint64_t stride,
                                uint64_t total, uint32_t block_checksum, uint64_t* rec_off, FrameTotals* totals,
                                cudaStream_t stream);
cudaError_t launch_assemble(const BlockSet& slots, const BlockSet& raw, const uint32_t* csize, const uint32_t* sums,
                            const uint64_t* rec_off, uint8_t* body, uint32_t nblocks, uint32_t block_checksum,
                            int num_sms, cudaStream_t stream);
// writes end mark (+ content checksum from *content_sum if non-null) after the body
cudaError_t launch_finalize(uint8_t* frame, uint64_t header_size, const FrameTotals* totals,
                            const uint32_t* content_sum, cudaStream_t stream);

struct WalkResult {
    uint32_t nblocks;      // blocks found (may exceed capacity: then only the first `capacity` were stored)
    uint32_t terminal;     // 0 = end mark seen, 1 = ran off the end without end mark, 2 = truncated (FrameSizeWrong)
    uint64_t end_pos;      // srcPos after the walk (position of the content checksum if any)
    uint32_t max_stored;   // largest stored block size
    uint32_t pad;
};
cudaError_t launch_walk(const uint8_t* frame, uint64_t n, uint64_t start, uint32_t block_checksum, uint64_t* off,
                        uint32_t* hdr, uint32_t capacity, WalkResult* res, cudaStream_t stream);

// K7' — parallel frame index (k_index.cu): candidate headers -> links -> jump tables -> ordered records.
// res->terminal == 3 means "a header above `bound` sits on the chain": fall back to launch_walk.
uint32_t index_tiles(const uint8_t* frame, uint64_t n);
uint32_t index_levels(uint32_t n_nodes);
cudaError_t launch_index_candidates(const uint8_t* frame, uint64_t n, uint64_t start, uint32_t bound, uint32_t block_checksum,
                                    uint32_t* tile_count, uint64_t* tile_base, uint16_t* masks, uint64_t* pos,
                                    uint64_t capacity, uint64_t* n_nodes, cudaStream_t stream);
cudaError_t launch_index_resolve(const uint8_t* frame, uint64_t n, uint64_t start, uint32_t bound, uint32_t block_checksum,
                                 const uint64_t* pos, uint32_t n_nodes, uint32_t* jump, uint64_t* off, uint32_t* hdr,
                                 uint32_t capacity, WalkResult* res, cudaStream_t stream);

struct DecodeSummary {
    uint32_t first_bad;        // first block index with a checksum / decode problem (0xFFFFFFFF none)
    int32_t bad_kind;          // 1 = block checksum mismatch, 2 = decode error, 3 = raw block no room
    int32_t bad_status;        // the lz4 status of the failing decode
    uint32_t layout_ok;        // 1 if every non-final block decoded to exactly block_size
    uint64_t total;            // decoded bytes (valid when no error and layout_ok)
};
cudaError_t launch_decode_summary(const uint32_t* out_len, const int32_t* status, const uint32_t* sums_calc,
                                  const uint8_t* frame, const uint64_t* off, const uint32_t* hdr, uint32_t nblocks,
                                  uint32_t block_size, uint32_t block_checksum, DecodeSummary* out,
                                  cudaStream_t stream);

}  // namespace b2
