// b2_common.cuh — shared device helpers for the LZ4 kernels (sm_100a).
#pragma once
#include <cstdint>
#ifdef B2_EMU   // host build of the device code for the one-warp emulator (tools/warp_emu): test tooling only
#include "emu_cuda.h"
#else
#include <cuda_runtime.h>
#endif

namespace b2 {

// ---- format constants, reference src/lz4.zig:12-44 ----
constexpr uint32_t MINMATCH = 4;
constexpr uint32_t LASTLITERALS = 5;
constexpr uint32_t MFLIMIT = 12;
constexpr uint32_t ML_MASK = 15;
constexpr uint32_t RUN_MASK = 15;
constexpr uint32_t MAX_DISTANCE = 65535;
constexpr uint32_t LZ4_MAX_INPUT_SIZE = 0x7E000000u;
constexpr uint32_t HASH_MULTIPLIER = 2654435761u;
constexpr uint32_t ACCELERATION_MAX = 65537;

constexpr unsigned FULL = 0xffffffffu;

// status codes written by kernels (subset of include/b2lz4.h)
constexpr int ST_OK = 0;
constexpr int ST_OUTPUT_TOO_SMALL = 1;
constexpr int ST_INPUT_TOO_LARGE = 2;
constexpr int ST_CORRUPTED = 3;
// internal: raw (stored) frame block did not fit -> lz4f DstMaxSizeTooSmall
constexpr int ST_RAW_NO_ROOM = 110;

// A set of blocks: either explicit (off/len arrays) or regular (stride, total).
struct BlockSet {
    const uint8_t* base;
    const uint64_t* off;   // nullptr => regular
    const uint32_t* len;   // explicit lengths (explicit mode)
    uint64_t stride;       // regular mode: block i starts at i*stride
    uint64_t total;        // regular mode: total bytes (last block may be short)
    uint32_t len_mask;     // explicit mode: n = len[i] & len_mask (frame headers carry a flag in bit 31)
    __device__ __forceinline__ void get(uint32_t i, const uint8_t*& p, uint32_t& n) const {
        if (off) {
            p = base + off[i];
            n = len[i] & len_mask;
        } else {
            uint64_t o = (uint64_t)i * stride;
            uint64_t r = total > o ? total - o : 0;
            p = base + o;
            n = (uint32_t)(r < stride ? r : stride);
        }
    }
};

struct OutSet {
    uint8_t* base;
    const uint64_t* off;   // nullptr => regular
    const uint32_t* cap;   // explicit capacities (explicit mode)
    uint64_t stride;       // regular: slot i at i*stride
    uint64_t total;        // regular: bytes available in total (cap_i = min(slot_cap, total - i*stride))
    uint32_t slot_cap;     // regular: capacity of each slot
    __device__ __forceinline__ void get(uint32_t i, uint8_t*& p, uint32_t& c) const {
        if (off) {
            p = base + off[i];
            c = cap[i];
        } else {
            uint64_t o = (uint64_t)i * stride;
            uint64_t r = total > o ? total - o : 0;
            p = base + o;
            c = (uint32_t)(r < slot_cap ? r : slot_cap);
        }
    }
};

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }
#ifdef B2_EMU
__device__ __forceinline__ uint32_t lanemask_lt() { return (1u << lane_id()) - 1u; }
__device__ __forceinline__ uint32_t lanemask_gt() { return lane_id() == 31 ? 0u : ~((2u << lane_id()) - 1u); }
#else
__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
__device__ __forceinline__ uint32_t lanemask_gt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_gt;" : "=r"(m));
    return m;
}
#endif

// Unaligned little-endian u32 from read-only global memory (input streams): two aligned words
// through the non-coherent path + funnel shift.  Never touches a word that holds no valid byte.
__device__ __forceinline__ uint32_t ldg_u32(const uint8_t* p) {
    uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
    uint32_t sh = (uint32_t)(a & 3) * 8;
    uint32_t lo = __ldg(w);
    uint32_t hi = sh ? __ldg(w + 1) : 0u;
    return __funnelshift_r(lo, hi, sh);
}

// Same through the coherent path (for buffers this kernel also writes, e.g. decode output).
__device__ __forceinline__ uint32_t ld_u32(const uint8_t* p) {
    uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
    uint32_t sh = (uint32_t)(a & 3) * 8;
    uint32_t lo = *w;
    uint32_t hi = sh ? *(w + 1) : 0u;
    return __funnelshift_r(lo, hi, sh);
}

template <bool NC>
__device__ __forceinline__ uint4 ld16(const uint4* p) {
    if (NC) return __ldg(p);
    return *p;
}
template <bool NC>
__device__ __forceinline__ uint8_t ld8(const uint8_t* p) {
    if (NC) return __ldg(p);
    return *p;
}

// 16 bytes starting at byte offset `bo` (0..15, warp-uniform) of the 32-byte pair (A,B).
__device__ __forceinline__ uint4 extract16(const uint4& A, const uint4& B, uint32_t bo) {
    uint32_t sh = (bo & 3) * 8;
    uint32_t w0, w1, w2, w3, w4;
    switch (bo >> 2) {
        case 0: w0 = A.x; w1 = A.y; w2 = A.z; w3 = A.w; w4 = B.x; break;
        case 1: w0 = A.y; w1 = A.z; w2 = A.w; w3 = B.x; w4 = B.y; break;
        case 2: w0 = A.z; w1 = A.w; w2 = B.x; w3 = B.y; w4 = B.z; break;
        default: w0 = A.w; w1 = B.x; w2 = B.y; w3 = B.z; w4 = B.w; break;
    }
    uint4 r;
    r.x = __funnelshift_r(w0, w1, sh);
    r.y = __funnelshift_r(w1, w2, sh);
    r.z = __funnelshift_r(w2, w3, sh);
    r.w = __funnelshift_r(w3, w4, sh);
    return r;
}

// Warp-cooperative copy of `len` bytes, arbitrary alignment on both sides, non-overlapping.
// Long runs use 16-byte aligned stores; the source is read as aligned 16-byte pairs and realigned
// in registers.  NC selects the non-coherent (read-only) load path for the source.
// All 32 lanes must call it with identical arguments.
template <bool NC>
__device__ __forceinline__ void warp_copy(uint8_t* dst, const uint8_t* src, uint32_t len, uint32_t lane) {
    if (len < 64) {
        for (uint32_t i = lane; i < len; i += 32) dst[i] = ld8<NC>(src + i);
        return;
    }
    uint32_t head = (uint32_t)((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15);
    if (lane < head) dst[lane] = ld8<NC>(src + lane);
    dst += head;
    src += head;
    len -= head;
    uint32_t nvec = len >> 4;
    uint32_t bo = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15);
    const uint4* s16 = reinterpret_cast<const uint4*>(src - bo);
    uint4* d16 = reinterpret_cast<uint4*>(dst);
    if (bo == 0) {
        for (uint32_t i = lane; i < nvec; i += 32) d16[i] = ld16<NC>(s16 + i);
    } else {
        for (uint32_t i = lane; i < nvec; i += 32) {
            uint4 A = ld16<NC>(s16 + i);
            uint4 B = ld16<NC>(s16 + i + 1);
            d16[i] = extract16(A, B, bo);
        }
    }
    uint32_t done = nvec << 4;
    uint32_t tail = len - done;
    if (lane < tail) dst[done + lane] = ld8<NC>(src + done + lane);
}

}  // namespace b2
