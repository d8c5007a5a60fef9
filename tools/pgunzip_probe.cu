// Probe for the chunk-parallel gunzip kernels (strainer2_b200/csrc/s2_gunzip.cu): N ordinary .gz images of FASTA or FASTQ
// text through decode -> chain -> translate -> CRC, each kernel timed with CUDA events, the text checked byte for byte.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/bin/pgunzip_probe tools/pgunzip_probe.cu strainer2_b200/csrc/s2_gunzip.cu -lz
// Usage: pgunzip_probe [files=256] [text_kb=5000] [sub_kb=32] [fastq=0] [level=6]
#include "../strainer2_b200/csrc/s2_gunzip.h"
#include "../strainer2_b200/csrc/s2_inflate.cuh"
#include <zlib.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CKP(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

static std::vector<uint8_t> make_text(size_t bytes, unsigned long long seed, bool fastq)
{
    std::vector<uint8_t> t;
    t.reserve(bytes + 512);
    unsigned long long s = seed * 0x9E3779B97F4A7C15ull + 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    int rec = 0;
    while (t.size() < bytes) {
        char hdr[96];
        if (fastq) {
            const int h = snprintf(hdr, sizeof hdr, "@SRR%llu.%d %d/1\n", seed, rec, rec); ++rec;
            t.insert(t.end(), hdr, hdr + h);
            for (int k = 0; k < 150; ++k) t.push_back("ACGT"[rnd() & 3]);
            t.push_back('\n'); t.push_back('+'); t.push_back('\n');
            for (int k = 0; k < 150; ++k) t.push_back("FFFFFFFFFFFF:F,#"[rnd() & 15]);
            t.push_back('\n');
        } else {
            const int h = snprintf(hdr, sizeof hdr, ">contig_%d len=125000\n", rec++);
            t.insert(t.end(), hdr, hdr + h);
            for (int line = 0; line < 1563 && t.size() < bytes; ++line) {
                for (int k = 0; k < 80; ++k) t.push_back("ACGT"[rnd() & 3]);
                t.push_back('\n');
            }
        }
    }
    return t;
}

static std::vector<uint8_t> gzip_level(const std::vector<uint8_t> &in, int level)
{
    z_stream z; memset(&z, 0, sizeof z);
    deflateInit2(&z, level, Z_DEFLATED, 31, 8, Z_DEFAULT_STRATEGY);
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
    const int n = argc > 1 ? atoi(argv[1]) : 256;
    const size_t text_bytes = (size_t)(argc > 2 ? atoi(argv[2]) : 5000) << 10;
    const uint32_t sub_bytes = (uint32_t)(argc > 3 ? atoi(argv[3]) : 32) << 10;
    const bool fastq = argc > 4 && atoi(argv[4]);
    const int level = argc > 5 ? atoi(argv[5]) : 6;
    const int distinct = 4;
    std::vector<std::vector<uint8_t>> texts, comps;
    for (int k = 0; k < distinct; ++k) { texts.push_back(make_text(text_bytes, k + 1, fastq)); comps.push_back(gzip_level(texts.back(), level)); }
    std::vector<GzFileDesc> files(n);
    std::vector<uint32_t> sub_file, slice0(n + 1);
    size_t coff = 0, toff = 0; uint32_t sub0 = 0, slices = 0;
    for (int i = 0; i < n; ++i) {
        const auto &c = comps[i % distinct]; const auto &t = texts[i % distinct];
        GzFileDesc &d = files[i];
        memset(&d, 0, sizeof d);
        d.comp_off = coff; d.comp_len = c.size(); d.first_bit = d.chain_bit = s2_gzip_header_len(c.data(), c.size()) * 8;
        d.text_off = toff; d.text_len = t.size(); d.sub0 = sub0; d.n_sub = (uint32_t)((c.size() + sub_bytes - 1) / sub_bytes);
        for (uint32_t k = 0; k < d.n_sub; ++k) sub_file.push_back(i);
        slice0[i] = slices; slices += (uint32_t)((t.size() + 4095) / 4096);
        coff += (c.size() + 15) / 16 * 16 + 16; toff += t.size(); sub0 += d.n_sub;
    }
    slice0[n] = slices;
    const uint32_t sub_cap = sub_bytes * 8 + (512u << 10) + 32768u;
    printf("%d %s files (level %d), %.1f MB of .gz -> %.1f MB of text, %u sub-chunks of %u KB, symbol area %.1f MB, tables %zu bytes per warp\n", n,
           fastq ? "FASTQ" : "FASTA", level, coff / 1e6, toff / 1e6, sub0, sub_bytes >> 10, (double)sub0 * sub_cap * 2 / 1e6, gz_tables_bytes());
    uint8_t *d_comp, *d_text, *d_win; uint16_t *d_sym; GzSubResult *d_res; uint64_t *d_sub_off; GzFileDesc *d_files; uint32_t *d_sub_file, *d_slice0, *d_crc; GzFileResult *d_fres;
    unsigned *d_act;
    CKP(cudaMalloc(&d_comp, coff + 64)); CKP(cudaMemset(d_comp, 0, coff + 64));
    CKP(cudaMalloc(&d_text, toff + 64)); CKP(cudaMalloc(&d_sym, gz_sym_slots(sub0, sub_cap) * 2)); d_sym = gz_launch_sym_init(d_sym, sub0, sub_cap, 0); CKP(cudaMalloc(&d_res, (size_t)sub0 * gz_sub_result_bytes()));
    CKP(cudaMalloc(&d_win, ((size_t)sub0 + n + 1) * 32768)); CKP(cudaMemset(d_win, 0, ((size_t)sub0 + n + 1) * 32768));
    CKP(cudaMalloc(&d_sub_off, (size_t)sub0 * 8)); CKP(cudaMalloc(&d_files, n * sizeof(GzFileDesc))); CKP(cudaMalloc(&d_sub_file, sub0 * 4));
    CKP(cudaMalloc(&d_slice0, (n + 1) * 4)); CKP(cudaMalloc(&d_crc, n * 4)); CKP(cudaMemset(d_crc, 0, n * 4)); CKP(cudaMalloc(&d_fres, n * sizeof(GzFileResult)));
    CKP(cudaMalloc(&d_act, n * 4));
    for (int i = 0; i < n; ++i) CKP(cudaMemcpy(d_comp + files[i].comp_off, comps[i % distinct].data(), comps[i % distinct].size(), cudaMemcpyHostToDevice));
    CKP(cudaMemcpy(d_files, files.data(), n * sizeof(GzFileDesc), cudaMemcpyHostToDevice));
    CKP(cudaMemcpy(d_sub_file, sub_file.data(), sub0 * 4, cudaMemcpyHostToDevice));
    CKP(cudaMemcpy(d_slice0, slice0.data(), (n + 1) * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e[5]; for (auto &x : e) cudaEventCreate(&x);
    for (int rep = 0; rep < 3; ++rep) {
        CKP(cudaMemset(d_text, 0, toff));
        cudaEventRecord(e[0]);
        gz_launch_decode(d_comp, d_files, d_sub_file, sub0, sub_bytes, d_sym, sub_cap, d_res, 0);
        cudaEventRecord(e[1]);
        gz_launch_chain(d_comp, d_files, n, d_sym, sub_cap, d_res, d_win, d_sub_off, d_fres, 0);
        cudaEventRecord(e[2]);
        gz_launch_translate(d_files, d_sub_file, 0, sub0, d_sym, sub_cap, d_win, d_sub_off, d_fres, d_text, 0);
        cudaEventRecord(e[3]);
        gz_launch_crc(d_files, 0, n, d_slice0, slices, d_text, d_fres, d_crc, d_act, 0);
        cudaEventRecord(e[4]);
        CKP(cudaDeviceSynchronize());
        float ms[4]; for (int k = 0; k < 4; ++k) cudaEventElapsedTime(&ms[k], e[k], e[k + 1]);
        std::vector<GzFileResult> fres(n);
        CKP(cudaMemcpy(fres.data(), d_fres, n * sizeof(GzFileResult), cudaMemcpyDeviceToHost));
        int bad = 0, first_bad = 0;
        std::vector<uint8_t> back;
        for (int i = 0; i < n; ++i) {
            if (fres[i].status != 0 || fres[i].text_len != texts[i % distinct].size()) { if (!bad) first_bad = fres[i].status; ++bad; continue; }
            if (i % 37 == 0 || i == n - 1) {
                back.resize(fres[i].text_len);
                CKP(cudaMemcpy(back.data(), d_text + files[i].text_off, back.size(), cudaMemcpyDeviceToHost));
                if (memcmp(back.data(), texts[i % distinct].data(), back.size())) { ++bad; first_bad = -999; }
            }
        }
        const float tot = ms[0] + ms[1] + ms[2] + ms[3];
        printf("rep %d: decode %.3f  chain %.3f  translate %.3f  crc %.3f ms = %.2f GB/s of text (decode alone %.2f), %d of %d files wrong (first status %d)\n", rep, ms[0], ms[1],
               ms[2], ms[3], toff / 1e6 / tot, toff / 1e6 / ms[0], bad, n, first_bad);
    }
    return 0;
}
