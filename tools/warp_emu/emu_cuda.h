// emu_cuda.h — a one-warp SIMT emulator for the host (test tooling, never part of the product).
//
// The device code of a kernel file is compiled by g++ with -DB2_EMU; every lane of the warp is a coroutine (ucontext)
// running the same per-thread function, and every warp collective (__shfl_sync, __ballot_sync, __match_any_sync,
// __syncwarp) is a rendezvous of the 32 coroutines.  That executes the kernel's logic exactly as written — lane by
// lane, with the collectives' semantics — so that a change of K1/K2 can be checked against the oracle here, where
// there is no GPU, before it spends box time.  Only full-mask collectives are supported (all the kernels use).
#pragma once
#include <ucontext.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>
#include <string>
#include <algorithm>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
struct ushort2 { uint16_t x, y; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
typedef int cudaError_t;
typedef void* cudaStream_t;

#include <sys/mman.h>
#include <unistd.h>

namespace emu {

// A buffer of n bytes between two inaccessible pages: `at_end` puts its last byte right before the guard page behind it
// (any read or write past the end faults), otherwise its first byte sits `lead` bytes after the guard page in front (any
// access below the 16-byte granule that holds the first byte faults, for lead < 16).  The kernels promise to touch only
// aligned granules that hold at least one valid byte; on the GPU nothing would report a violation, here it is a SIGSEGV.
struct Guarded {
    uint8_t* map = nullptr; size_t map_len = 0; uint8_t* p = nullptr;
    Guarded(size_t n, bool at_end, size_t lead = 0) {
        const size_t pg = (size_t)sysconf(_SC_PAGESIZE);
        const size_t body = (n + lead + pg - 1) / pg * pg + pg;      // one spare page so that both placements fit
        map_len = body + 2 * pg;
        map = (uint8_t*)mmap(nullptr, map_len, PROT_NONE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (map == MAP_FAILED) { perror("mmap"); abort(); }
        if (mprotect(map + pg, body, PROT_READ | PROT_WRITE) != 0) { perror("mprotect"); abort(); }
        if (at_end) {
            // the accessible part must END at the byte after the buffer: shrink it so the guard page follows directly
            const size_t used = (n + pg - 1) / pg * pg;
            if (mprotect(map + pg + used, body - used, PROT_NONE) != 0) { perror("mprotect"); abort(); }
            p = map + pg + used - n;
        } else {
            p = map + pg + lead;
        }
    }
    ~Guarded() { if (map) munmap(map, map_len); }
    Guarded(const Guarded&) = delete;
    Guarded& operator=(const Guarded&) = delete;
};

constexpr int W = 32;
// Context switch: on x86-64 a dozen instructions (callee-saved registers + stack pointer) instead of swapcontext(), which
// makes a sigprocmask system call per switch — a third of the emulator's run time.
#if defined(__x86_64__)
#define EMU_FAST_SWITCH 1
extern "C" void emu_switch(void** save_sp, void* load_sp);
#endif

struct Warp {
#ifdef EMU_FAST_SWITCH
    void* sched_sp;
    void* lane_sp[W];
#else
    ucontext_t sched;
    ucontext_t lanes[W];
#endif
    bool done[W];
    int cur;
    uint32_t buf[2][W];
    int arrived[2], left[2];
    uint32_t gen[W];
    uint64_t n_collectives;
    std::function<void()> body;
};
extern Warp* g_warp;
extern uint32_t g_warp_index;

inline int lane() { return g_warp->cur; }
#ifdef EMU_FAST_SWITCH
inline void yield() { Warp* w = g_warp; emu_switch(&w->lane_sp[w->cur], w->sched_sp); }
#else
inline void yield() { Warp* w = g_warp; swapcontext(&w->lanes[w->cur], &w->sched); }
#endif

// every lane deposits `v`; returns a pointer to the 32 deposited values (valid until the lane's next-but-one collective)
inline const uint32_t* exchange(uint32_t v) {
    Warp* w = g_warp;
    const int l = w->cur;
    const int s = (int)(w->gen[l]++ & 1u);
    w->buf[s][l] = v;
    w->arrived[s]++;
    while (w->arrived[s] < W) yield();
    if (++w->left[s] == W) { w->arrived[s] = 0; w->left[s] = 0; w->n_collectives++; }
    else {
        // the slot may only be re-armed after everybody has read it: a lane that runs ahead waits at its next use of `s`
    }
    return w->buf[s];
}

void run_warp(const std::function<void()>& body, size_t stack_bytes = 1 << 20);

struct Dim { unsigned x, y, z; };
struct Tid { unsigned x, y, z; };
inline Tid tid() { return Tid{(unsigned)(g_warp_index * 32 + (uint32_t)lane()), 0, 0}; }

}  // namespace emu

#define threadIdx (emu::tid())

static inline uint32_t __shfl_sync(unsigned, uint32_t v, int src, int = 32) {
    uint32_t copy[32];
    const uint32_t* all = emu::exchange(v);
    memcpy(copy, all, sizeof(copy));
    return copy[src & 31];
}
static inline int __shfl_sync(unsigned m, int v, int src, int w = 32) { return (int)__shfl_sync(m, (uint32_t)v, src, w); }
static inline uint32_t __shfl_up_sync(unsigned, uint32_t v, unsigned d, int = 32) {
    const int l = emu::lane();
    const uint32_t* all = emu::exchange(v);
    return l >= (int)d ? all[l - d] : v;
}
static inline uint32_t __shfl_down_sync(unsigned, uint32_t v, unsigned d, int = 32) {
    const int l = emu::lane();
    const uint32_t* all = emu::exchange(v);
    return l + (int)d < 32 ? all[l + d] : v;
}
static inline uint32_t __shfl_xor_sync(unsigned, uint32_t v, int x, int = 32) {
    const int l = emu::lane();
    const uint32_t* all = emu::exchange(v);
    return all[(l ^ x) & 31];
}
static inline uint32_t __ballot_sync(unsigned, int pred) {
    const uint32_t* all = emu::exchange(pred ? 1u : 0u);
    uint32_t m = 0;
    for (int i = 0; i < 32; i++) m |= (all[i] & 1u) << i;
    return m;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xFFFFFFFFu; }
static inline uint32_t __match_any_sync(unsigned, uint32_t v) {
    const uint32_t* all = emu::exchange(v);
    uint32_t m = 0;
    for (int i = 0; i < 32; i++) m |= (uint32_t)(all[i] == v) << i;
    return m;
}
static inline void __syncwarp(unsigned = 0xffffffffu) { (void)emu::exchange(0); }
static inline uint32_t __reduce_add_sync(unsigned, uint32_t v) {
    const uint32_t* all = emu::exchange(v);
    uint32_t s = 0;
    for (int i = 0; i < 32; i++) s += all[i];
    return s;
}
static inline uint32_t __reduce_max_sync(unsigned, uint32_t v) {
    const uint32_t* all = emu::exchange(v);
    uint32_t s = 0;
    for (int i = 0; i < 32; i++) s = all[i] > s ? all[i] : s;
    return s;
}
static inline uint32_t __reduce_min_sync(unsigned, uint32_t v) {
    const uint32_t* all = emu::exchange(v);
    uint32_t s = 0xFFFFFFFFu;
    for (int i = 0; i < 32; i++) s = all[i] < s ? all[i] : s;
    return s;
}
static inline uint32_t __reduce_or_sync(unsigned, uint32_t v) {
    const uint32_t* all = emu::exchange(v);
    uint32_t s = 0;
    for (int i = 0; i < 32; i++) s |= all[i];
    return s;
}

template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcg(const T* p) { return *p; }
template <typename T> static inline T __ldcs(const T* p) { return *p; }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) {
    return (uint32_t)(((((uint64_t)hi) << 32) | lo) >> (sh & 31));
}
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t sh) {
    return (uint32_t)((((((uint64_t)hi) << 32) | lo) << (sh & 31)) >> 32);
}
static inline int __clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
static inline int __ffs(uint32_t x) { return __builtin_ffs((int)x); }
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline uint32_t __brev(uint32_t x) {
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i);
    return r;
}
static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t s) {
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        const uint32_t sel = (s >> (4 * i)) & 0xF;
        uint32_t byte = (uint32_t)(v >> (8 * (sel & 7))) & 0xFF;
        if (sel & 8) byte = (byte & 0x80) ? 0xFF : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
static inline uint32_t __vcmpeq4(uint32_t a, uint32_t b) {
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) if (((a >> (8 * i)) & 0xFF) == ((b >> (8 * i)) & 0xFF)) r |= 0xFFu << (8 * i);
    return r;
}
static inline uint32_t min(uint32_t a, uint32_t b) { return a < b ? a : b; }
static inline uint32_t max(uint32_t a, uint32_t b) { return a > b ? a : b; }
static inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { uint32_t o = *p; *p += v; return o; }
static inline bool __isGlobal(const void*) { return true; }
#define __builtin_assume(x) ((void)0)
