// TEST-ONLY harness: the host build of strainer2_b200/csrc/s2_inflate.cuh (the DEFLATE / gzip decoder written for host
// and device), so that tests/test_host.py can check it against zlib on the CPU.  Never loaded by the package or the
// product library.
#include "../../strainer2_b200/csrc/s2_inflate.cuh"

extern "C" {
int sim_inflate_raw(const unsigned char *src, unsigned long long n, unsigned char *dst, unsigned long long cap, unsigned long long *out_len,
                    unsigned long long *consumed)
{
    static S2InfTables t;
    uint64_t o = 0, c = 0;
    const int rc = s2_inflate_raw(src, n, dst, cap, &o, &c, t);
    *out_len = o; *consumed = c;
    return rc;
}
int sim_gunzip(const unsigned char *src, unsigned long long n, unsigned char *dst, unsigned long long cap, unsigned long long *out_len)
{
    static S2InfTables t;
    uint64_t o = 0;
    const int rc = s2_gunzip(src, n, dst, cap, &o, t);
    *out_len = o;
    return rc;
}
unsigned long long sim_inflate_table_bytes(void) { return sizeof(S2InfTables); }
}
