// Probe for the software gunzip kernel (strainer2_b200/csrc/s2_inflate.cuh): N ordinary single-member .gz images of
// FASTA text, one decoder per thread, checked byte for byte against the text they were made from.  Groundwork for the
// ordinary-.gz half of SURVEY 8(f) rank 1 (DESIGN.md "What comes next"); not part of the product library.
// Build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/bin/gunzip_probe tools/gunzip_probe.cu -lz
// Usage: gunzip_probe [files=1024] [text_kb=1024] [threads_per_block=32] [variant=0]
//   variant 0: one decoder per THREAD, tables in the thread's local memory (threads_per_block 1 = one decoder per warp)
//   variant 1: one decoder per WARP (lane 0 decodes), tables in shared memory, threads_per_block / 32 warps per block.
//              Why: a thread's local memory is interleaved over the 32 lanes of its warp, so the 3.3 KB of tables of a
//              lone decoder are spread over 105 KB of address space and miss L1 on nearly every lookup - the 460 cycles
//              per literal of variant 0.  Measured (profiles/r1s_gunzip_kernel_v0_probe.txt): variant 0 8.5 GB/s of text, variant 1 14.3.
#include "../strainer2_b200/csrc/s2_inflate.cuh"
#include <cuda_runtime.h>
#include <zlib.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CKP(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

// one thread = one .gz file (all its members); the decoder's tables live in the thread's local memory
__global__ void s2_gunzip_kernel(const uint8_t *comp, const unsigned long long *comp_off, uint8_t *text, const unsigned long long *text_off,
                                 unsigned long long *out_len, int *status, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    S2InfTables t;
    uint64_t got = 0;
    status[i] = s2_gunzip(comp + comp_off[i], comp_off[i + 1] - comp_off[i], text + text_off[i], text_off[i + 1] - text_off[i], &got, t);
    out_len[i] = got;
}

// one warp = one .gz file, lane 0 decodes, tables in shared memory (3,264 bytes per warp)
__global__ void s2_gunzip_warp_kernel(const uint8_t *comp, const unsigned long long *comp_off, uint8_t *text, const unsigned long long *text_off,
                                      unsigned long long *out_len, int *status, int n)
{
    extern __shared__ __align__(16) unsigned char smem[];
    S2InfTables *tables = reinterpret_cast<S2InfTables *>(smem);
    const int warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const int i = blockIdx.x * warps + warp;
    if (i >= n || (threadIdx.x & 31)) return;
    uint64_t got = 0;
    status[i] = s2_gunzip(comp + comp_off[i], comp_off[i + 1] - comp_off[i], text + text_off[i], text_off[i + 1] - text_off[i], &got, tables[warp]);
    out_len[i] = got;
}

static std::vector<uint8_t> make_fasta(size_t bytes, unsigned long long seed)
{
    std::vector<uint8_t> t;
    t.reserve(bytes + 128);
    unsigned long long s = seed * 0x9E3779B97F4A7C15ull + 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    int contig = 0;
    while (t.size() < bytes) {
        char hdr[64];
        const int h = snprintf(hdr, sizeof hdr, ">contig_%d len=125000\n", contig++);
        t.insert(t.end(), hdr, hdr + h);
        for (int line = 0; line < 1563 && t.size() < bytes; ++line) {
            for (int k = 0; k < 80; ++k) t.push_back("ACGT"[rnd() & 3]);
            t.push_back('\n');
        }
    }
    return t;
}

static std::vector<uint8_t> gzip6(const std::vector<uint8_t> &in)
{
    z_stream z; memset(&z, 0, sizeof z);
    deflateInit2(&z, 6, Z_DEFLATED, 31, 8, Z_DEFAULT_STRATEGY);          // 31: gzip wrapper
    std::vector<uint8_t> out(deflateBound(&z, in.size()) + 64);
    z.next_in = (Bytef *)in.data(); z.avail_in = (uInt)in.size();
    z.next_out = out.data(); z.avail_out = (uInt)out.size();
    deflate(&z, Z_FINISH);
    out.resize(z.total_out);
    deflateEnd(&z);
    return out;
}

int main(int argc, char **argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 1024;
    const size_t text_bytes = (size_t)(argc > 2 ? atoi(argv[2]) : 1024) << 10;
    int tpb = argc > 3 ? atoi(argv[3]) : 32;
    const int variant = argc > 4 ? atoi(argv[4]) : 0;
    if (variant == 1) tpb = tpb < 32 ? 32 : tpb / 32 * 32;
    const int distinct = 8;                                              // distinct images, cycled
    std::vector<std::vector<uint8_t>> texts, comps;
    for (int k = 0; k < distinct; ++k) { texts.push_back(make_fasta(text_bytes, k + 1)); comps.push_back(gzip6(texts.back())); }
    std::vector<unsigned long long> coff(n + 1, 0), toff(n + 1, 0);
    for (int i = 0; i < n; ++i) { coff[i + 1] = coff[i] + comps[i % distinct].size(); toff[i + 1] = toff[i] + texts[i % distinct].size(); }
    printf("%d files, %.1f MB of .gz -> %.1f MB of FASTA, %d threads per block, variant %d, decoder tables %zu bytes\n", n, coff[n] / 1e6, toff[n] / 1e6, tpb,
           variant, sizeof(S2InfTables));
    uint8_t *d_comp, *d_text; unsigned long long *d_coff, *d_toff, *d_len; int *d_status;
    CKP(cudaMalloc(&d_comp, coff[n] + 64)); CKP(cudaMalloc(&d_text, toff[n] + 64));
    CKP(cudaMalloc(&d_coff, (n + 1) * 8)); CKP(cudaMalloc(&d_toff, (n + 1) * 8)); CKP(cudaMalloc(&d_len, n * 8)); CKP(cudaMalloc(&d_status, n * 4));
    for (int i = 0; i < n; ++i) CKP(cudaMemcpy(d_comp + coff[i], comps[i % distinct].data(), comps[i % distinct].size(), cudaMemcpyHostToDevice));
    CKP(cudaMemcpy(d_coff, coff.data(), (n + 1) * 8, cudaMemcpyHostToDevice));
    CKP(cudaMemcpy(d_toff, toff.data(), (n + 1) * 8, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        CKP(cudaMemset(d_text, 0, toff[n]));
        cudaEventRecord(e0);
        if (variant == 1) {
            const int warps = tpb / 32;
            const size_t smem = (size_t)warps * sizeof(S2InfTables);
            CKP(cudaFuncSetAttribute(s2_gunzip_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            s2_gunzip_warp_kernel<<<(n + warps - 1) / warps, tpb, smem>>>(d_comp, d_coff, d_text, d_toff, d_len, d_status, n);
        } else {
            s2_gunzip_kernel<<<(n + tpb - 1) / tpb, tpb>>>(d_comp, d_coff, d_text, d_toff, d_len, d_status, n);
        }
        cudaEventRecord(e1);
        CKP(cudaDeviceSynchronize());
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        std::vector<int> status(n); std::vector<unsigned long long> len(n);
        CKP(cudaMemcpy(status.data(), d_status, n * 4, cudaMemcpyDeviceToHost));
        CKP(cudaMemcpy(len.data(), d_len, n * 8, cudaMemcpyDeviceToHost));
        int bad = 0;
        std::vector<uint8_t> back;
        for (int i = 0; i < n; ++i) {
            if (status[i] != 0 || len[i] != texts[i % distinct].size()) { ++bad; continue; }
            if (i % 61 == 0 || i == n - 1) {                             // sample the bytes
                back.resize(len[i]);
                CKP(cudaMemcpy(back.data(), d_text + toff[i], len[i], cudaMemcpyDeviceToHost));
                if (memcmp(back.data(), texts[i % distinct].data(), len[i])) ++bad;
            }
        }
        printf("rep %d: %.3f ms = %.2f GB/s of text (%.2f GB/s of .gz), %d of %d files wrong\n", rep, ms, toff[n] / 1e6 / ms, coff[n] / 1e6 / ms, bad, n);
    }
    return 0;
}
