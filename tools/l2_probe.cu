// Micro-probe: how much randomly probed data stays L2 resident on this GPU, for the two-phase scan's slice size.
// A region of R bytes is (optionally) streamed once with sequential 256-bit loads, then probed at random with one 32-byte
// sector load per thread-iteration (the scan's bucket load), several times; prints G probes/s per region size and policy.
// Build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tools/bin/l2_probe tools/l2_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

__device__ __forceinline__ void ld256(const void *p, uint32_t (&x)[8], int policy)
{
    if (policy == 0)
        asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]), "=r"(x[4]), "=r"(x[5]), "=r"(x[6]), "=r"(x[7]) : "l"(p));
    else
        asm volatile("ld.global.nc.L1::no_allocate.L2::evict_last.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]), "=r"(x[4]), "=r"(x[5]), "=r"(x[6]), "=r"(x[7]) : "l"(p));
}

__global__ void stream_kernel(const uint8_t *base, uint64_t n_sectors, int policy, uint32_t *sink)
{
    uint32_t acc = 0;
    for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_sectors; s += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t x[8];
        ld256(base + s * 32, x, policy);
        acc ^= x[0] ^ x[7];
    }
    if (acc == 0x12345u) *sink = acc;
}

__global__ void probe_kernel(const uint8_t *base, uint64_t n_sectors, uint64_t probes_per_thread, int policy, uint32_t *sink)
{
    uint64_t s = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345;
    uint32_t acc = 0;
    for (uint64_t i = 0; i < probes_per_thread; i += 4) {
        uint32_t x[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            s = s * 6364136223846793005ull + 1442695040888963407ull;
            const uint64_t sec = __umul64hi(s, n_sectors);
            ld256(base + sec * 32, x[u], policy);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) acc ^= x[u][0] ^ x[u][7];
    }
    if (acc == 0x12345u) *sink = acc;
}

int main()
{
    const size_t cap = 2048ull << 20;
    uint8_t *buf; uint32_t *sink;
    cudaMalloc(&buf, cap); cudaMalloc(&sink, 4);
    cudaMemset(buf, 1, cap);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 4, block = 256;
    const uint64_t per_thread = 256;                                  // 38.8 M probes per launch
    for (int policy = 0; policy < 2; ++policy)
        for (int pre = 0; pre < 2; ++pre)
            for (size_t mb : { 5, 10, 20, 40, 60, 80, 100, 120, 160, 320, 1280 }) {
                const uint64_t n_sectors = (mb << 20) / 32;
                // flush: stream another 400 MB region
                stream_kernel<<<grid, block>>>(buf + (1500ull << 20), (400ull << 20) / 32, 0, sink);
                if (pre) stream_kernel<<<grid, block>>>(buf, n_sectors, policy, sink);
                float best = 1e9f;
                for (int rep = 0; rep < 3; ++rep) {
                    cudaEventRecord(e0);
                    probe_kernel<<<grid, block>>>(buf, n_sectors, per_thread, policy, sink);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                    float ms; cudaEventElapsedTime(&ms, e0, e1);
                    if (rep == 0) printf("policy=%s prestream=%d region=%4zu MB: first pass %.3f ms = %.1f G probes/s", policy ? "evict_last" : "plain", pre, mb, ms,
                                         grid * block * per_thread / ms / 1e6);
                    else if (ms < best) best = ms;
                }
                printf(", later passes %.3f ms = %.1f G probes/s\n", best, grid * block * per_thread / best / 1e6);
            }
    return 0;
}
