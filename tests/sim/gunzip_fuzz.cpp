// TEST-ONLY: fuzz of the host build of strainer2_b200/csrc/s2_gunzip.cuh (block finder + speculative decode of one sub-chunk)
// under AddressSanitizer / UBSan (tests/test_host.py::test_gunzip_fuzz_under_sanitizers).  Every sub-chunk gets a symbol region
// of its own, allocated to the byte - 32,768 marker slots, `cap` symbols, the guard slot - and the compressed words are
// allocated to the word, so any access the decoder makes outside what the kernels give it is a report here; on the GPU it
// would be silent corruption of a neighbouring region.  Damaged and truncated streams, regions that overflow, all levels.
// argv[1] = trials per (text, level).
#include "../../strainer2_b200/csrc/s2_gunzip.cuh"
#include "../../strainer2_b200/csrc/s2_inflate.cuh"      // s2_gzip_header_len
#include <zlib.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>
int main(int argc, char **argv)
{
    unsigned long long s = 0x9E3779B97F4A7C15ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    static GzTables tab;
    static uint8_t kraft9[512];
    gz_kraft9_fill(kraft9, 0, 1);
    long ok = 0, err = 0, overflow = 0, exact = 0;
    for (int round = 0; round < 6; ++round) {
        std::vector<uint8_t> t;
        const size_t n = 40000 + rnd() % 400000;
        if (round % 3 == 0) while (t.size() < n) { for (int k = 0; k < 80; ++k) t.push_back("ACGT"[rnd() & 3]); t.push_back('\n'); }
        else if (round % 3 == 1) while (t.size() < n) { t.push_back('@'); for (int k = 0; k < 100; ++k) t.push_back("ACGT"[rnd() & 3]); t.push_back('\n'); t.push_back('+'); t.push_back('\n'); for (int k = 0; k < 100; ++k) t.push_back(rnd() % 9 ? 'I' : (uint8_t)('#' + rnd() % 40)); t.push_back('\n'); }
        else { std::vector<uint8_t> unit; for (int k = 0; k < 700; ++k) unit.push_back("ACGT"[rnd() & 3]); while (t.size() < n) t.insert(t.end(), unit.begin(), unit.end()); }     // 200 : 1
        for (int lvl : { 1, 6, 9 }) {
            z_stream z; memset(&z, 0, sizeof z);
            deflateInit2(&z, lvl, Z_DEFLATED, 31, 8, Z_DEFAULT_STRATEGY);
            std::vector<uint8_t> c(deflateBound(&z, t.size()) + 64);
            z.next_in = t.data(); z.avail_in = t.size(); z.next_out = c.data(); z.avail_out = c.size();
            deflate(&z, Z_FINISH); c.resize(z.total_out); deflateEnd(&z);
            for (int trial = 0; trial < (argc > 1 ? atoi(argv[1]) : 60); ++trial) {
                std::vector<uint8_t> cc(c);
                const int flips = trial ? (int)(rnd() % 4) : 0;
                for (int f = 0; f < flips; ++f) cc[rnd() % cc.size()] ^= (uint8_t)(1u << (rnd() % 8));
                if (trial % 5 == 4) cc.resize(20 + rnd() % (cc.size() - 20));
                const uint64_t hl = s2_gzip_header_len(cc.data(), cc.size());
                if (!hl) { ++err; continue; }
                const uint64_t n_words = (cc.size() + 3) / 4;
                uint32_t *words = (uint32_t *)calloc(n_words, 4);                 // to the word: the reader may not look past n_words
                memcpy(words, cc.data(), cc.size());
                const uint32_t sub_bytes = 4096u << (rnd() % 5);
                // symbols a region holds: usually plenty, sometimes far too few (the literal / copy stores must stop at the guard slot)
                const uint32_t cap = trial % 4 == 3 ? 16u + (uint32_t)(rnd() % 40000) : (uint32_t)t.size() + 300000u;
                const uint64_t n_sub = (cc.size() + sub_bytes - 1) / sub_bytes;
                uint64_t total = 0, cur = hl * 8;
                bool chain = true, done = false;
                for (uint64_t i = 0; i < n_sub; ++i) {
                    uint16_t *region = (uint16_t *)malloc(((size_t)GZ_WINDOW + cap + 1) * sizeof(uint16_t));      // markers | cap symbols | guard
                    gz_marker_prefix(region, 0, 1);
                    GzSubResult r;
                    gz_subchunk(words, n_words, i == 0 ? hl * 8 : ~0ull, i * sub_bytes * 8ull, (i + 1) * sub_bytes * 8ull, 8ull << 20, region + GZ_WINDOW, cap, tab,
                                kraft9, &r, 0, 1);
                    if (r.status == GZ_ERR_OUTPUT) ++overflow;
                    if (r.status >= 0 && r.n_out > cap) { printf("n_out beyond the region\n"); return 1; }
                    if (chain && !done) {
                        if (r.start_bit != cur || r.status < 0) chain = false;
                        else { total += r.n_out; cur = r.end_bit; done = r.status == GZ_FINAL; }
                    }
                    free(region);
                }
                free(words);
                if (chain && done) {
                    ++ok;
                    if (!flips && trial % 5 != 4) { if (total != t.size()) { printf("MISMATCH: %llu symbols for %zu bytes\n", (unsigned long long)total, t.size()); return 1; } ++exact; }
                } else {
                    ++err;
                    if (!flips && trial % 5 != 4 && trial % 4 != 3) { printf("an intact stream did not decode (round %d level %d sub %u)\n", round, lvl, sub_bytes); return 1; }
                }
            }
        }
    }
    printf("fuzz done: %ld streams chained to their end (%ld intact ones with the exact size), %ld not, %ld sub-chunks overflowed their region\n", ok, exact, err, overflow);
    return 0;
}
