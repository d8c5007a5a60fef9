// TEST-ONLY harness: exposes the __host__ __device__ bit primitives of strainer2_b200/csrc/s2_kmer.cuh
// (compiled for the host by g++) so tests/test_host.py can check them against string arithmetic on
// the CPU before any GPU minute is spent.  Never loaded by the package or the product library.
#include "../../strainer2_b200/csrc/s2_kmer.cuh"
#include <string.h>

extern "C" {
void sim_pack16(const unsigned char *b, uint32_t *w, uint32_t *m)
{
    uint32_t v[4];
    memcpy(v, b, 16);
    s2_pack16(v[0], v[1], v[2], v[3], w, m);
}
uint32_t sim_rc16(uint32_t w) { return s2_rc16(w); }
uint64_t sim_extract31(uint32_t a, uint32_t b, uint32_t c, unsigned j) { return s2_extract31(a, b, c, j); }
int sim_window_valid(uint32_t a, uint32_t b, uint32_t c, unsigned j) { return s2_window_valid(a, b, c, j) ? 1 : 0; }
uint64_t sim_revcomp31(uint64_t k) { return s2_revcomp31(k); }
uint32_t sim_djb2(uint64_t k) { return s2_djb2_of_kmer(k); }
void sim_hash(uint64_t k, uint32_t *h, uint32_t *fp) { s2_hash_t r = s2_hash(k); *h = r.h; *fp = r.fp; }
uint32_t sim_bucket(uint32_t h, uint32_t n) { return s2_bucket_of(h, n); }
// canonical k-mer of window j of a 48-byte stretch, exactly the way the scan kernel computes it
uint64_t sim_window_canon(const unsigned char *b48, unsigned j, int *valid)
{
    uint32_t w[3], m[3];
    for (int i = 0; i < 3; ++i) sim_pack16(b48 + 16 * i, &w[i], &m[i]);
    const uint32_t r0 = s2_rc16(w[2]), r1 = s2_rc16(w[1]), r2 = s2_rc16(w[0]);
    *valid = s2_window_valid(m[0], m[1], m[2], j) ? 1 : 0;
    return s2_canonical(s2_extract31(w[0], w[1], w[2], j), s2_extract31(r0, r1, r2, 17u - j));
}
}
