// TEST-ONLY: fuzz of the host build of strainer2_b200/csrc/s2_inflate.cuh under AddressSanitizer / UBSan with exact-size heap
// buffers (tests/test_host.py::test_inflate_fuzz_under_sanitizers): on the GPU an out-of-bounds byte would be silent corruption.
// argv[1] = trials per (text, level).
#include "../../strainer2_b200/csrc/s2_inflate.cuh"
#include <zlib.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>
int main(int argc, char **argv)
{
    unsigned long long s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    static S2InfTables tab;
    long ok = 0, err = 0;
    for (int round = 0; round < 6; ++round) {
        std::vector<uint8_t> t;
        const size_t n = 3000 + rnd() % 60000;
        if (round % 3 == 0) while (t.size() < n) { for (int k = 0; k < 80; ++k) t.push_back("ACGT"[rnd() & 3]); t.push_back('\n'); }
        else if (round % 3 == 1) while (t.size() < n) { t.push_back('@'); for (int k = 0; k < 100; ++k) t.push_back("ACGT"[rnd() & 3]); t.push_back('\n'); t.push_back('+'); t.push_back('\n'); for (int k = 0; k < 100; ++k) t.push_back('I'); t.push_back('\n'); }
        else while (t.size() < n) t.push_back((uint8_t)(rnd() % 7 ? 'x' : rnd()));
        for (int lvl : { 1, 6, 9 }) {
            z_stream z; memset(&z, 0, sizeof z);
            deflateInit2(&z, lvl, Z_DEFLATED, 31, 8, Z_DEFAULT_STRATEGY);
            std::vector<uint8_t> c(deflateBound(&z, t.size()) + 64);
            z.next_in = t.data(); z.avail_in = t.size(); z.next_out = c.data(); z.avail_out = c.size();
            deflate(&z, Z_FINISH); c.resize(z.total_out); deflateEnd(&z);
            // exact-size heap buffers so that ASan sees any byte out of place
            for (int trial = 0; trial < (argc > 1 ? atoi(argv[1]) : 1500); ++trial) {
                std::vector<uint8_t> cc(c);
                const int flips = trial ? 1 + rnd() % 3 : 0;
                for (int f = 0; f < flips; ++f) cc[rnd() % cc.size()] ^= (uint8_t)(1u << (rnd() % 8));
                if (trial % 5 == 4) cc.resize(rnd() % cc.size());
                uint8_t *in = (uint8_t *)malloc(cc.size() ? cc.size() : 1); memcpy(in, cc.data(), cc.size());
                const size_t cap = trial % 7 == 6 ? rnd() % (t.size() + 1) : t.size();
                uint8_t *out = (uint8_t *)malloc(cap ? cap : 1);
                uint64_t got = 0;
                const int rc = s2_gunzip(in, cc.size(), out, cap, &got, tab);
                if (rc == 0) { ++ok; if (!flips && cap == t.size() && (got != t.size() || memcmp(out, t.data(), got))) { printf("MISMATCH\n"); return 1; } }
                else ++err;
                if (got > cap) { printf("got > cap\n"); return 1; }
                free(in); free(out);
            }
        }
    }
    printf("fuzz done: %ld decoded, %ld rejected\n", ok, err);
    return 0;
}
