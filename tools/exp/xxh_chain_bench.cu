// Micro-benchmark of the serial XXH32 accumulator chain (one warp): variants of the round formulation.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o xxh_chain_bench xxh_chain_bench.cu ; run: ./xxh_chain_bench [MiB]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr uint32_t P1 = 2654435761u, P2 = 2246822519u;
constexpr uint32_t C1 = P1 << 13;  // rotl(b,13)*P1 = b*C1 + (b>>19)*P1 (mod 2^32)
__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return __funnelshift_l(x, x, r); }

// variant 0: baseline (row-major tile, scalar LDS, rotl form)
__global__ void __launch_bounds__(32) k_v0(const uint4* __restrict__ p, uint64_t nst, uint32_t* out) {
    __shared__ __align__(16) uint32_t tile[2][512];
    const uint32_t lane = threadIdx.x;
    uint32_t acc = lane + 1;
    uint4 r[4];
    auto load_tile = [&](uint64_t first) {
#pragma unroll
        for (int u = 0; u < 4; u++) { uint64_t s = first + lane + 32u * u; if (s < nst) r[u] = __ldg(p + s); }
    };
    int buf = 0; uint64_t done = 0;
    if (nst) load_tile(0);
    while (done < nst) {
        uint32_t cnt = (uint32_t)(nst - done < 128 ? nst - done : 128);
#pragma unroll
        for (int u = 0; u < 4; u++) { uint32_t s = lane + 32u * u; if (s < cnt) reinterpret_cast<uint4*>(tile[buf])[s] = r[u]; }
        __syncwarp();
        if (done + 128 < nst) load_tile(done + 128);
        if (lane < 4) {
            const uint32_t* t = tile[buf] + lane;
#pragma unroll 8
            for (uint32_t s = 0; s < cnt; s++) acc = rotl32(acc + t[4 * s] * P2, 13) * P1;
        }
        done += cnt; buf ^= 1;
    }
    if (lane < 4) out[lane] = acc;
}

// variants 1/2: column-major tile (one row per accumulator), LDS.128 one group ahead, two-deep round:
//   b' = (b*C1 + y_next) + (b>>19)*P1, with b = acc + y.   MODE 1: shift on the alu pipe; MODE 2: mul.hi on the fma pipe
constexpr uint32_t ROW = 132;  // 128 words + 4 pad: the four chain lanes hit different banks
template <int MODE>
__device__ __forceinline__ uint32_t step(uint32_t b, uint32_t ynext) {
    const uint32_t u = b * C1 + ynext;
    const uint32_t t = MODE == 2 ? __umulhi(b, 8192u) : (b >> 19);
    return t * P1 + u;
}
template <int MODE>
__global__ void __launch_bounds__(32) k_v12(const uint4* __restrict__ p, uint64_t nst, uint32_t* out) {
    __shared__ __align__(16) uint32_t tile[2][4 * ROW];
    const uint32_t lane = threadIdx.x;
    uint32_t acc = lane + 1;
    uint4 r[4];
    auto load_tile = [&](uint64_t first) {
#pragma unroll
        for (int u = 0; u < 4; u++) { uint64_t s = first + lane + 32u * u; if (s < nst) r[u] = __ldg(p + s); }
    };
    int buf = 0; uint64_t done = 0;
    if (nst) load_tile(0);
    while (done < nst) {
        uint32_t cnt = (uint32_t)(nst - done < 128 ? nst - done : 128);
#pragma unroll
        for (int u = 0; u < 4; u++) {
            uint32_t s = lane + 32u * u;
            if (s < cnt) { uint32_t* t = tile[buf] + s; t[0] = r[u].x; t[ROW] = r[u].y; t[2 * ROW] = r[u].z; t[3 * ROW] = r[u].w; }
        }
        __syncwarp();
        if (done + 128 < nst) load_tile(done + 128);
        if (lane < 4) {
            const uint32_t* row = tile[buf] + lane * ROW;
            if (cnt == 128) {
                uint4 cur = *reinterpret_cast<const uint4*>(row);
                uint32_t b = acc + cur.x * P2;
#pragma unroll 4
                for (uint32_t g = 0; g < 32; g++) {
                    uint4 nxt = make_uint4(0, 0, 0, 0);
                    if (g + 1 < 32) nxt = *reinterpret_cast<const uint4*>(row + 4 * (g + 1));
                    b = step<MODE>(b, cur.y * P2);
                    b = step<MODE>(b, cur.z * P2);
                    b = step<MODE>(b, cur.w * P2);
                    b = step<MODE>(b, nxt.x * P2);   // last group: y = 0 turns b into the accumulator
                    cur = nxt;
                }
                acc = b;
            } else {
                for (uint32_t s = 0; s < cnt; s++) acc = rotl32(acc + row[s] * P2, 13) * P1;
            }
        }
        done += cnt; buf ^= 1;
    }
    if (lane < 4) out[lane] = acc;
}

int main(int argc, char** argv) {
    size_t mib = argc > 1 ? atoi(argv[1]) : 256;
    size_t n = mib << 20;
    std::vector<uint32_t> h(n / 4);
    uint64_t x = 88172645463325252ull;
    for (auto& w : h) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; w = (uint32_t)x; }
    uint32_t want[4];
    for (int l = 0; l < 4; l++) {
        uint32_t acc = l + 1;
        for (size_t s = 0; s < n / 16; s++) { uint32_t v = acc + h[4 * s + l] * P2; acc = ((v << 13) | (v >> 19)) * P1; }
        want[l] = acc;
    }
    void* d; uint32_t* o;
    cudaMalloc(&d, n); cudaMalloc(&o, 16);
    cudaMemcpy(d, h.data(), n, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int v = 0; v < 3; v++) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (v == 0) k_v0<<<1, 32>>>((const uint4*)d, n / 16, o);
            if (v == 1) k_v12<1><<<1, 32>>>((const uint4*)d, n / 16, o);
            if (v == 2) k_v12<2><<<1, 32>>>((const uint4*)d, n / 16, o);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            uint32_t got[4]; cudaMemcpy(got, o, 16, cudaMemcpyDeviceToHost);
            bool ok = got[0] == want[0] && got[1] == want[1] && got[2] == want[2] && got[3] == want[3];
            printf("variant %d rep %d: %.2f ms  %.3f GB/s  %.2f ns/stripe  %s\n", v, rep, ms, n / ms / 1e6, ms * 1e6 / (n / 16), ok ? "ok" : "MISMATCH");
        }
    }
    printf("cuda: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
