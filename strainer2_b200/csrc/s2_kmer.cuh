// s2_kmer.cuh - bit-level k-mer primitives shared by every kernel (and unit-tested on the host).
//
// Replaces, for ACGT windows, the string machinery of the reference's scan loop:
//   BIO_stringToUpper        /root/reference/src/BIO_sequence.c:228-234
//   COMPLEMENT[] / orient_string / rc_strcmp   src/BIO_sequence.c:203-213, src/genome_compare.c:1100-1141
//   contains_N               src/genome_compare.c:443-451
//   hashU (djb2)             src/BIO_hash.c:208-216   (kept only to replay the reference's row order)
//
// Encoding: A=0 C=1 G=2 T=3, first base in the most significant 2-bit field.  It is order preserving
// (byte order A<C<G<T), so the reference's "lexicographically larger of window and reverse complement,
// forward wins ties" is simply max(fwd, rc) on 62-bit integers, and complement is x ^ 3.
// (The reference's dead up2bit codec uses A0 C1 T2 G3, src/up2bit.c:14; that one is only provided as
// s2_encode_2bit/s2_decode_2bit in the C ABI for bit compatibility and is not used for comparison.)
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define S2_HD __host__ __device__ __forceinline__
#else
#define S2_HD inline
#endif

#define S2_K 31
#define S2_KMER_MASK 0x3FFFFFFFFFFFFFFFull      /* 62 bits */
#define S2_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull      /* never a stored key: stored keys have bit 62 clear */
#define S2_INFORMATIVE_BIT 0x8000000000000000ull /* strain_detect's KMER_TYPE==INFORMATIVE payload   */
#define S2_BUCKET_SLOTS 16                       /* 16 x 16-bit fingerprints = one 32-byte sector     */

// ---- 4 ASCII bases (one little-endian 32-bit word, first base in the low byte) -----------------
// returns the 8-bit packed code (first base in bits 7:6) in *code and a 4-bit validity nibble (first
// base in bit 3) in *valid.  valid bit = byte is one of ACGTacgt.  Exact for all 256 byte values.
S2_HD void s2_pack4(uint32_t v, uint32_t *code, uint32_t *valid)
{
    const uint32_t u = v & 0xDFDFDFDFu;                       // toupper for letters; never aliases into ACGT
    const uint32_t x = (u >> 1) & 0x03030303u;                // A0 C1 G3 T2
    const uint32_t xs = (x >> 1) & 0x01010101u;
    const uint32_t y = x ^ xs;                                // A0 C1 G2 T3
    *code = (y * 0x40100401u) >> 24;                          // gather 4 x 2 bits, first base highest
    // expected upper-case letter for that code: 'A' + 2x + 15*[x==2]  ->  A C G T
    const uint32_t is_t = xs & ~x;                            // x == 2 (per byte, in bit 0)
    const uint32_t e = 0x41414141u + (x << 1) + is_t * 15u;
    const uint32_t d = u ^ e;                                 // zero byte <=> valid base
    const uint32_t nz = (((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d) & 0x80808080u;   // bit 7 of byte set <=> nonzero
    const uint32_t ok = (nz ^ 0x80808080u) >> 7;              // bit 0 of byte set <=> valid
    *valid = ((ok * 0x08040201u) >> 24) & 0xFu;
}

// 16 ASCII bases (four words in memory order) -> 32-bit packed word + 16-bit validity mask,
// base i of the 16 in bits (31-2i : 30-2i) of the word and bit (15-i) of the mask.
S2_HD void s2_pack16(uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t *word, uint32_t *mask)
{
    uint32_t c0, c1, c2, c3, v0, v1, v2, v3;
    s2_pack4(a, &c0, &v0); s2_pack4(b, &c1, &v1); s2_pack4(c, &c2, &v2); s2_pack4(d, &c3, &v3);
    *word = (c0 << 24) | (c1 << 16) | (c2 << 8) | c3;
    *mask = (v0 << 12) | (v1 << 8) | (v2 << 4) | v3;
}

// reverse complement of a packed 16-base word (still 16 bases, reversed order, complemented)
S2_HD uint32_t s2_rc16(uint32_t w)
{
#if defined(__CUDA_ARCH__)
    uint32_t r = __brev(w);
#else
    uint32_t r = w;
    r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
    r = ((r >> 2) & 0x33333333u) | ((r & 0x33333333u) << 2);
    r = ((r >> 4) & 0x0F0F0F0Fu) | ((r & 0x0F0F0F0Fu) << 4);
    r = ((r >> 8) & 0x00FF00FFu) | ((r & 0x00FF00FFu) << 8);
    r = (r >> 16) | (r << 16);
#endif
    // brev also swapped the two bits inside every field: swap them back, then complement (x ^ 3)
    r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
    return ~r;
}

S2_HD uint32_t s2_funnel_l(uint32_t hi, uint32_t lo, unsigned s)   // (hi:lo << s) >> 32, 0 <= s < 32
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, s);
#else
    return s ? (hi << s) | (lo >> (32 - s)) : hi;
#endif
}

// k-mer starting at base j (0..16) of the 48-base string w0:w1:w2 -> 62-bit value (first base highest)
S2_HD uint64_t s2_extract31(uint32_t w0, uint32_t w1, uint32_t w2, unsigned j)
{
    uint32_t a = w0, b = w1, c = w2;
    if (j >= 16) { a = w1; b = w2; c = 0; j -= 16; }
    const uint32_t hi = s2_funnel_l(a, b, 2 * j);
    const uint32_t lo = s2_funnel_l(b, c, 2 * j);
    return (((uint64_t)hi << 32) | lo) >> 2;
}

// validity of the window starting at base j (0..15) given the 48-bit mask string m0:m1:m2 (16 bits each)
S2_HD bool s2_window_valid(uint32_t m0, uint32_t m1, uint32_t m2, unsigned j)
{
    const uint32_t hi = (m0 << 16) | m1;                     // bases 0..31
    const uint32_t lo = m2 << 16;                            // bases 32..47 in the top half
    const uint32_t win = s2_funnel_l(hi, lo, j);             // 32 flags starting at base j
    return (win >> 1) == 0x7FFFFFFFu;                        // first 31 all valid
}

S2_HD uint64_t s2_canonical(uint64_t fwd, uint64_t rc) { return fwd > rc ? fwd : rc; }

// reverse complement of a full 62-bit k-mer (host-side helpers, table flagging, tests)
S2_HD uint64_t s2_revcomp31(uint64_t k)
{
    // k occupies bits 61..0 ; put it MSB-aligned in 64 bits: k << 2 (pad field = 00 at the bottom)
    const uint64_t a = k << 2;
    const uint32_t w0 = (uint32_t)(a >> 32), w1 = (uint32_t)a;
    // reverse 32 fields: rc16(w1):rc16(w0); the pad field (complemented to 11) comes out on top
    const uint64_t r = ((uint64_t)s2_rc16(w1) << 32) | s2_rc16(w0);
    return r & S2_KMER_MASK;
}

S2_HD uint32_t s2_mulhi32(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

// ---- hashing for the device table ---------------------------------------------------------------
// One 64-bit multiply.  Canonical k-mers of real sequence are close to uniform already (only skewed
// towards large values by the max(fwd, rc) rule), so a Fibonacci multiply is enough: every input bit
// reaches the high word, which picks the bucket (mulhi by n_buckets uses its top bits); the
// fingerprint comes from the top of the LOW product word, i.e. from different bits.
struct s2_hash_t { uint32_t h; uint32_t fp; };

S2_HD s2_hash_t s2_hash(uint64_t canon)
{
    const uint64_t x = canon * 0x9E3779B97F4A7C15ull;
    s2_hash_t r;
    r.h = (uint32_t)(x >> 32);
    // fingerprint: a *normal positive fp16 bit pattern* (0x0400..0x7BFF) so that the probe can test 16
    // fingerprints with 8 HSET2 (half2 equality) instructions; 0x0000 is the empty slot.
    r.fp = 0x0400u + s2_mulhi32((uint32_t)x, 30720u);
    return r;
}

S2_HD uint32_t s2_bucket_of(uint32_t h, uint32_t n_buckets) { return s2_mulhi32(h, n_buckets); }

// ---- djb2 of the 31-letter ASCII spelling of a packed k-mer (src/BIO_hash.c:208-216, before % M) --
S2_HD uint32_t s2_djb2_of_kmer(uint64_t k)
{
    uint32_t h = 5381u;
    for (int i = 0; i < S2_K; ++i) {
        const uint32_t code = (uint32_t)(k >> (2 * (S2_K - 1 - i))) & 3u;
        // A C G T = 65 67 71 84
        const uint32_t ch = 65u + 2u * code + (code >> 1) * (2u + 11u * (code & 1u));
        h = h * 33u + ch;
    }
    return h;
}

S2_HD char s2_letter(uint32_t code) { return (char)(65u + 2u * code + (code >> 1) * (2u + 11u * (code & 1u))); }
