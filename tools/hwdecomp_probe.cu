// Probe of the Blackwell hardware decompression engine through the CUDA driver API
// (the batch-decompress entry point of cuda.h, CUDA 12.8+): is DEFLATE offered on this device, what is the longest single
// operation, and how fast is one long raw-deflate stream / a batch of streams of FASTQ-like text?
// `hwdecomp_probe errors` instead probes how the engine reports bad input: bytes after the end of the stream (a gzip
// trailer + another member), a corrupt stream, a destination that is too small, a stream longer than the maximum.
// Build: nvcc -O2 -o hwdecomp_probe tools/hwdecomp_probe.cu -lcuda -lz
#include <cuda.h>
#include <cuda_runtime.h>
#include <zlib.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CKD(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char *s; cuGetErrorString(r_, &s); printf("driver error %d (%s) at line %d\n", (int)r_, s ? s : "?", __LINE__); return 1; } } while (0)

static std::vector<unsigned char> make_fastq(size_t approx)
{
    std::vector<unsigned char> t;
    unsigned long long s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    size_t n = 0;
    while (t.size() < approx) {
        char hdr[64]; int h = snprintf(hdr, sizeof hdr, "@r%zu\n", n++);
        t.insert(t.end(), hdr, hdr + h);
        for (int i = 0; i < 150; ++i) t.push_back("ACGT"[rnd() & 3]);
        t.push_back('\n'); t.push_back('+'); t.push_back('\n');
        for (int i = 0; i < 150; ++i) t.push_back('I');
        t.push_back('\n');
    }
    return t;
}

static std::vector<unsigned char> raw_deflate(const std::vector<unsigned char> &in)
{
    z_stream z; memset(&z, 0, sizeof z);
    deflateInit2(&z, 6, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
    std::vector<unsigned char> out(deflateBound(&z, in.size()));
    z.next_in = (Bytef *)in.data(); z.avail_in = (uInt)in.size();
    z.next_out = out.data(); z.avail_out = (uInt)out.size();
    deflate(&z, Z_FINISH);
    out.resize(z.total_out);
    deflateEnd(&z);
    return out;
}

// one decompress operation, everything reported: what the submit returns, what the stream says afterwards, the
// byte count the engine wrote, whether the text is right, and whether the context still works afterwards
static int probe_case(const char *what, const std::vector<unsigned char> &comp, size_t src_bytes, const std::vector<unsigned char> &text, size_t dst_bytes)
{
    cudaStream_t st; cudaStreamCreate(&st);
    CUdeviceptr dsrc, ddst, dact;
    CKD(cuMemAlloc(&dsrc, comp.size() + 64)); CKD(cuMemAlloc(&ddst, text.size() + 4096)); CKD(cuMemAlloc(&dact, 64));
    CKD(cuMemcpyHtoD(dsrc, comp.data(), comp.size()));
    cudaMemset((void *)ddst, 0x55, text.size() + 4096);
    cudaMemset((void *)dact, 0xEE, 64);
    CUmemDecompressParams p; memset(&p, 0, sizeof p);
    p.srcNumBytes = src_bytes; p.dstNumBytes = dst_bytes; p.dstActBytes = (cuuint32_t *)dact;
    p.src = (const void *)dsrc; p.dst = (void *)ddst; p.algo = CU_MEM_DECOMPRESS_ALGORITHM_DEFLATE;
    size_t err = 12345;
    const CUresult r = cuMemBatchDecompressAsync(&p, 1, 0, &err, st);
    const cudaError_t se = cudaStreamSynchronize(st);
    unsigned act = 0; const CUresult rc = cuMemcpyDtoH(&act, dact, 4);
    std::vector<unsigned char> back(text.size() + 16, 0);
    const CUresult rb = cuMemcpyDtoH(back.data(), ddst, text.size() + 16);
    size_t same = 0; while (same < text.size() && back[same] == text[same]) ++same;
    bool past_untouched = true; for (int i = 0; i < 16; ++i) past_untouched = past_untouched && back[text.size() + i] == 0x55;
    const cudaError_t after = cudaDeviceSynchronize();
    printf("%-46s submit=%d err_index=%zu sync=%d (%s) act=%u (text %zu) first %zu bytes right, bytes past the text untouched=%d, D2H=%d/%d, device afterwards=%d (%s)\n",
           what, (int)r, err, (int)se, cudaGetErrorName(se), act, text.size(), same, (int)past_untouched, (int)rc, (int)rb, (int)after, cudaGetErrorName(after));
    cudaGetLastError();
    cuMemFree(dsrc); cuMemFree(ddst); cuMemFree(dact);
    cudaStreamDestroy(st);
    return 0;
}

static int probe_errors(int maxlen, int only)
{
    int case_no = 0;
#define CASE(...) do { if (only < 0 || only == case_no) probe_case(__VA_ARGS__); ++case_no; } while (0)
    std::vector<unsigned char> text = make_fastq(1 << 20), comp = raw_deflate(text);
    CASE("clean 1 MB stream", comp, comp.size(), text, text.size());
    // what a single-member .gz looks like behind its header: stream + CRC32 + ISIZE, then possibly another member
    std::vector<unsigned char> tailed = comp;
    for (int i = 0; i < 8; ++i) tailed.push_back((unsigned char)(0xA0 + i));
    CASE("stream + 8 trailer bytes inside srcNumBytes", tailed, tailed.size(), text, text.size());
    std::vector<unsigned char> two = tailed;
    const unsigned char head[10] = { 0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 3 };
    two.insert(two.end(), head, head + 10); two.insert(two.end(), comp.begin(), comp.end());
    CASE("stream + trailer + a second gzip member", two, two.size(), text, text.size());
    CASE("destination larger than the text", comp, comp.size(), text, text.size() + 2048);
    CASE("destination 1000 bytes too small", comp, comp.size(), text, text.size() - 1000);
    CASE("stream cut 100 bytes short", comp, comp.size() - 100, text, text.size());
    std::vector<unsigned char> bad = comp;
    for (size_t i = bad.size() / 2; i < bad.size() / 2 + 64; ++i) bad[i] ^= 0x5A;
    CASE("64 corrupt bytes in the middle", bad, bad.size(), text, text.size());
    std::vector<unsigned char> junk(200000);
    for (size_t i = 0; i < junk.size(); ++i) junk[i] = (unsigned char)(i * 2654435761u >> 13);
    CASE("not deflate at all", junk, junk.size(), text, text.size());
    CASE("clean stream again (is the engine still fine?)", comp, comp.size(), text, text.size());
    std::vector<unsigned char> big = make_fastq((size_t)maxlen + (1 << 20)), bigc = raw_deflate(big);
    CASE("text 1 MB longer than the maximum length", bigc, bigc.size(), big, big.size());
    std::vector<unsigned char> edge = make_fastq((size_t)maxlen - 4096); edge.resize((size_t)maxlen);
    std::vector<unsigned char> edgec = raw_deflate(edge);
    CASE("text exactly the maximum length", edgec, edgec.size(), edge, edge.size());
    CASE("clean stream at the end", comp, comp.size(), text, text.size());
    return 0;
}

int main(int argc, char **argv)
{
    cudaFree(0);
    CUdevice dev; CKD(cuInit(0)); CKD(cuDeviceGet(&dev, 0));
    int mask = 0, maxlen = 0;
    CKD(cuDeviceGetAttribute(&mask, CU_DEVICE_ATTRIBUTE_MEM_DECOMPRESS_ALGORITHM_MASK, dev));
    CKD(cuDeviceGetAttribute(&maxlen, CU_DEVICE_ATTRIBUTE_MEM_DECOMPRESS_MAXIMUM_LENGTH, dev));
    printf("decompress algorithm mask = 0x%x (deflate=%d snappy=%d lz4=%d), maximum length = %d bytes\n", mask, mask & 1, (mask >> 1) & 1, (mask >> 2) & 1, maxlen);
    if (!(mask & CU_MEM_DECOMPRESS_ALGORITHM_DEFLATE)) { printf("no hardware deflate\n"); return 0; }
    if (argc > 1 && !strcmp(argv[1], "errors")) return probe_errors(maxlen, argc > 2 ? atoi(argv[2]) : -1);      // a failed case poisons the context: one case per process
    size_t sizes_mb[] = { 1, 2, 3 };
    int nsz = argc > 1 ? atoi(argv[1]) : 3;
    cudaStream_t st; cudaStreamCreate(&st);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int k = 0; k < nsz; ++k) {
        std::vector<unsigned char> text = make_fastq(sizes_mb[k] << 20), comp = raw_deflate(text);
        CUdeviceptr dsrc, ddst, dact;
        CKD(cuMemAlloc(&dsrc, comp.size() + 64)); CKD(cuMemAlloc(&ddst, text.size() + 64)); CKD(cuMemAlloc(&dact, 64));
        CKD(cuMemcpyHtoD(dsrc, comp.data(), comp.size()));
        cudaMemset((void *)ddst, 0, text.size());
        CUmemDecompressParams p; memset(&p, 0, sizeof p);
        p.srcNumBytes = comp.size(); p.dstNumBytes = text.size(); p.dstActBytes = (cuuint32_t *)dact;
        p.src = (const void *)dsrc; p.dst = (void *)ddst; p.algo = CU_MEM_DECOMPRESS_ALGORITHM_DEFLATE;
        size_t err = 0;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0, st);
            CUresult r = cuMemBatchDecompressAsync(&p, 1, 0, &err, st);
            cudaEventRecord(e1, st);
            cudaError_t se = cudaStreamSynchronize(st);
            float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
            unsigned act = 0; cuMemcpyDtoH(&act, dact, 4);
            std::vector<unsigned char> back(text.size());
            cuMemcpyDtoH(back.data(), ddst, text.size());
            printf("one stream %4zu MB text (%6.1f MB deflate): submit=%d sync=%d act=%u ok=%d  %.3f ms  %.2f GB/s of text\n", sizes_mb[k], comp.size() / 1048576.0,
                   (int)r, (int)se, act, (int)(act == text.size() && memcmp(back.data(), text.data(), text.size()) == 0), ms, text.size() / 1e6 / ms);
            if (r != CUDA_SUCCESS || se != cudaSuccess) return 1;
        }
        cuMemFree(dsrc); cuMemFree(ddst); cuMemFree(dact);
    }
    // batches of independent 64 KB-of-text streams (the BGZF block shape) and of 1 MB streams
    const size_t shapes[][2] = { { 65280, 1024 }, { 65280, 4096 }, { 1 << 20, 256 }, { 3 << 20, 64 } };
    for (auto &sh : shapes) {
        const size_t tsz = sh[0]; const int B = (int)sh[1];
        std::vector<unsigned char> big = make_fastq(tsz + 4096);
        std::vector<unsigned char> text(big.begin(), big.begin() + tsz), comp = raw_deflate(text);
        const size_t cs = (comp.size() + 255) & ~255ull, ts = tsz;            // outputs packed back to back
        CUdeviceptr dsrc, ddst, dact;
        CKD(cuMemAlloc(&dsrc, cs * B)); CKD(cuMemAlloc(&ddst, ts * B + 256)); CKD(cuMemAlloc(&dact, 4 * B));
        std::vector<CUmemDecompressParams> ps(B);
        for (int i = 0; i < B; ++i) {
            CKD(cuMemcpyHtoD(dsrc + cs * i, comp.data(), comp.size()));
            memset(&ps[i], 0, sizeof ps[i]);
            ps[i].srcNumBytes = comp.size(); ps[i].dstNumBytes = text.size(); ps[i].dstActBytes = (cuuint32_t *)(dact + 4 * i);
            ps[i].src = (const void *)(dsrc + cs * i); ps[i].dst = (void *)(ddst + ts * i); ps[i].algo = CU_MEM_DECOMPRESS_ALGORITHM_DEFLATE;
        }
        size_t err = 0;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0, st);
            CUresult r = cuMemBatchDecompressAsync(ps.data(), B, 0, &err, st);
            cudaEventRecord(e1, st);
            cudaError_t se = cudaStreamSynchronize(st);
            float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
            std::vector<unsigned char> back(text.size());
            cuMemcpyDtoH(back.data(), ddst + ts * (B - 1), text.size());
            printf("batch of %5d x %7zu B text (dst unaligned, packed): submit=%d sync=%d last ok=%d  %.3f ms  %.2f GB/s of text\n", B, tsz, (int)r, (int)se,
                   (int)(memcmp(back.data(), text.data(), text.size()) == 0), ms, B * text.size() / 1e6 / ms);
            if (r != CUDA_SUCCESS || se != cudaSuccess) break;
        }
        cuMemFree(dsrc); cuMemFree(ddst); cuMemFree(dact);
    }
    return 0;
}
