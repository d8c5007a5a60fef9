// What pinned host memory costs and gives on this box: time to get N MB of it (cudaHostAlloc; malloc + touch +
// cudaHostRegister; the same on transparent huge pages) and the host -> device rate from each, with 1, 2 and 4 copies in
// flight.  Basis for the reader arenas of the executables and for the e2e ceiling of bench.py.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/bin/pinned_probe tools/pinned_probe.cu
// Usage: pinned_probe [mb=512]
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static void rate(const char *what, const uint8_t *h, uint8_t *d, size_t bytes)
{
    cudaStream_t st[4];
    for (auto &s : st) cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    for (int n : { 1, 2, 4 }) {
        const size_t part = bytes / n / 4096 * 4096;
        double best = 0;
        for (int rep = 0; rep < 4; ++rep) {
            cudaDeviceSynchronize();
            const double t0 = now();
            for (int k = 0; k < n; ++k) cudaMemcpyAsync(d + k * part, h + k * part, part, cudaMemcpyHostToDevice, st[k]);
            cudaDeviceSynchronize();
            const double gbs = part * n / 1e9 / (now() - t0);
            if (gbs > best) best = gbs;
        }
        printf("  %-34s H2D %d in flight: %.1f GB/s\n", what, n, best);
    }
    for (auto &s : st) cudaStreamDestroy(s);
}

int main(int argc, char **argv)
{
    const size_t bytes = (size_t)(argc > 1 ? atoi(argv[1]) : 512) << 20;
    cudaFree(0);
    uint8_t *d = nullptr;
    if (cudaMalloc(&d, bytes) != cudaSuccess) { printf("cudaMalloc failed\n"); return 1; }
    {
        FILE *f = fopen("/sys/kernel/mm/transparent_hugepage/enabled", "r");
        char line[128] = "?";
        if (f) { if (!fgets(line, sizeof line, f)) line[0] = 0; fclose(f); }
        printf("transparent_hugepage/enabled: %s", line);
    }
    for (int round = 0; round < 2; ++round) {
        uint8_t *h = nullptr;
        double t0 = now();
        cudaHostAlloc((void **)&h, bytes, cudaHostAllocDefault);
        printf("cudaHostAlloc of %zu MB: %.1f ms\n", bytes >> 20, (now() - t0) * 1e3);
        t0 = now();
        memset(h, 1, bytes);
        printf("  first touch (memset): %.1f ms\n", (now() - t0) * 1e3);
        rate("cudaHostAlloc", h, d, bytes);
        t0 = now();
        cudaFreeHost(h);
        printf("  cudaFreeHost: %.1f ms\n", (now() - t0) * 1e3);
    }
    {
        uint8_t *h = nullptr;
        double t0 = now();
        cudaHostAlloc((void **)&h, bytes, cudaHostAllocWriteCombined);
        printf("cudaHostAlloc write-combined of %zu MB: %.1f ms\n", bytes >> 20, (now() - t0) * 1e3);
        t0 = now();
        memset(h, 1, bytes);
        printf("  first touch (memset): %.1f ms\n", (now() - t0) * 1e3);
        rate("write-combined", h, d, bytes);
        cudaFreeHost(h);
    }
    for (int huge = 0; huge < 2; ++huge) {
        double t0 = now();
        uint8_t *h = (uint8_t *)aligned_alloc(2u << 20, bytes);
        if (huge) madvise(h, bytes, MADV_HUGEPAGE);
        const int nt = 8;                                       // touched by 8 threads, as reader threads would
        std::vector<std::thread> th;
        for (int k = 0; k < nt; ++k) th.emplace_back([=]() { memset(h + bytes / nt * k, 1, bytes / nt); });
        for (auto &t : th) t.join();
        const double t1 = now();
        const cudaError_t e = cudaHostRegister(h, bytes, cudaHostRegisterDefault);
        printf("aligned_alloc%s + touch on %d threads: %.1f ms, cudaHostRegister: %.1f ms (%s)\n", huge ? " + MADV_HUGEPAGE" : "", nt, (t1 - t0) * 1e3,
               (now() - t1) * 1e3, cudaGetErrorString(e));
        if (e == cudaSuccess) { rate(huge ? "registered (huge pages asked)" : "registered (4 KB pages)", h, d, bytes); cudaHostUnregister(h); }
        free(h);
    }
    {
        uint8_t *h = (uint8_t *)malloc(bytes);
        memset(h, 1, bytes);
        rate("pageable", h, d, bytes);
        free(h);
    }
    // many small pinned allocations (what 16 reader threads would do at once)
    {
        const int nt = 16; const size_t each = 32u << 20;
        std::vector<uint8_t *> hs(nt, nullptr);
        const double t0 = now();
        std::vector<std::thread> th;
        for (int k = 0; k < nt; ++k) th.emplace_back([&hs, k, each]() { cudaHostAlloc((void **)&hs[k], each, cudaHostAllocDefault); });
        for (auto &t : th) t.join();
        printf("%d threads x cudaHostAlloc of %zu MB at once: %.1f ms\n", nt, each >> 20, (now() - t0) * 1e3);
        for (auto p : hs) cudaFreeHost(p);
    }
    cudaFree(d);
    return 0;
}
