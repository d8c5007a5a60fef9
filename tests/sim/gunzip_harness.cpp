// TEST-ONLY harness: the host build of strainer2_b200/csrc/s2_gunzip.cuh (chunk-parallel gunzip: block finder,
// speculative decode with window markers, chain validation, window resolution, translate), run serially - one "warp" of
// one lane per sub-chunk - so that tests/test_host.py can check the whole scheme against zlib on the CPU.  Never loaded by
// the package or the product library.
#include "../../strainer2_b200/csrc/s2_gunzip.cuh"
#include "../../strainer2_b200/csrc/s2_inflate.cuh"      // s2_gzip_header_len

#include <cstring>
#include <vector>

extern "C" {
// src[0..n): a .gz file.  sub_bytes: size of a sub-chunk (multiple of 4).  Returns 0 and the text in dst (*out_len bytes),
// or a negative code: -100 not gzip, -101 chain broken (a sub-chunk did not start where its predecessor ended), -102 the
// stream did not end, -103 ISIZE mismatch, -104 trailing bytes, else the decoder's own error.  stats[0] = sub-chunks,
// [1] = sub-chunks that decoded something, [2] = marker symbols written, [3] = symbols written.
int sim_pgunzip(const unsigned char *src, unsigned long long n, unsigned sub_bytes, unsigned cap_ratio, unsigned char *dst, unsigned long long cap,
                unsigned long long *out_len, unsigned long long *stats)
{
    *out_len = 0;
    const uint64_t hl = s2_gzip_header_len(src, n);
    if (!hl) return -100;
    const uint64_t n_words = (n + 3) / 4;
    std::vector<uint32_t> words(n_words + 4, 0u);
    memcpy(words.data(), src, n);
    const uint64_t n_sub = (n + sub_bytes - 1) / sub_bytes;
    const uint32_t sub_cap = sub_bytes * cap_ratio + (1u << 19) + GZ_WINDOW;     // a sub-chunk runs on to the first block boundary behind the next cut
    std::vector<uint16_t> sym_alloc((size_t)n_sub * sub_cap + GZ_WINDOW);             // (the layout of gz_launch_sym_init)
    for (uint64_t i = 0; i < n_sub; ++i) gz_marker_prefix(sym_alloc.data() + i * sub_cap, 0, 1);
    uint16_t *const sym = sym_alloc.data() + GZ_WINDOW;
    std::vector<GzSubResult> res(n_sub);
    static GzTables t;
    static uint8_t kraft9[512];
    gz_kraft9_fill(kraft9, 0, 1);
    for (uint64_t i = 0; i < n_sub; ++i)
        gz_subchunk(words.data(), n_words, i == 0 ? hl * 8 : ~0ull, i * sub_bytes * 8ull, (i + 1) * sub_bytes * 8ull, 8ull << 20, sym + i * sub_cap,
                    sub_cap - GZ_WINDOW - 1u, t, kraft9, &res[i], 0, 1);
    // chain
    uint64_t cur = hl * 8, total = 0, end_bit = 0;
    bool done = false;
    std::vector<uint8_t> win((size_t)(n_sub + 1) * GZ_WINDOW, 0);
    std::vector<uint64_t> off(n_sub, 0);
    std::vector<uint32_t> len(n_sub, 0);
    stats[0] = n_sub; stats[1] = stats[2] = stats[3] = 0;
    for (uint64_t i = 0; i < n_sub; ++i) {
        uint8_t *prev = win.data() + i * GZ_WINDOW, *next = win.data() + (i + 1) * GZ_WINDOW;
        off[i] = total;
        if (done) { memcpy(next, prev, GZ_WINDOW); continue; }
        if (res[i].start_bit != cur) return -101;
        if (res[i].status < 0) return res[i].status;
        len[i] = res[i].n_out;
        gz_next_window(prev, sym + i * sub_cap, len[i], next, 0, 1);
        total += len[i];
        cur = res[i].end_bit;
        if (len[i]) ++stats[1];
        stats[3] += len[i];
        for (uint32_t k = 0; k < len[i]; ++k) stats[2] += sym[i * sub_cap + k] >= 256;
        if (res[i].status == GZ_FINAL) { done = true; end_bit = cur; }
    }
    if (!done) return -102;
    if (total > cap) return GZ_ERR_OUTPUT;
    for (uint64_t i = 0; i < n_sub; ++i) gz_translate(win.data() + i * GZ_WINDOW, sym + i * sub_cap, len[i], dst + off[i], 0, 1);
    *out_len = total;
    const uint64_t trailer = (end_bit + 7) / 8;
    if (trailer + 8 > n) return -102;
    const uint32_t isize = (uint32_t)src[trailer + 4] | (uint32_t)src[trailer + 5] << 8 | (uint32_t)src[trailer + 6] << 16 | (uint32_t)src[trailer + 7] << 24;
    if (isize != (uint32_t)total) return -103;
    for (uint64_t i = trailer + 8; i < n; ++i) if (src[i]) return -104;
    return 0;
}
}
