// s2_gunzip.cuh - chunk-parallel gunzip: ONE ordinary .gz stream decoded by many warps at once.
//
// What it replaces: zlib's gzread under the reference's parser (/root/reference/src/genome_compare.c:194-203,
// src/strain_detect.c:417-433, src/kseq.h:68-101).  Every input the reference ships is an ordinary single-member .gz;
// one DEFLATE stream is sequential, zlib inflates it at 0.3 GB/s of text per core, and the Blackwell decompression
// engine cannot take it.  The scheme here is the published one of pugz / rapidgzip, laid out for warps:
//
//   1. the compressed bytes are cut into sub-chunks of fixed size.  The warp of sub-chunk i FINDS the first DEFLATE
//      block that starts at or after its cut (32 lanes test 32 bit offsets at a time: block type, code counts, a
//      complete code-length code; survivors get the whole header parsed and both Huffman codes checked for completeness)
//      and DECODES from there to the first block boundary at or after the next cut.  It does not know the 32 KB of text
//      before its start, so it writes 16-bit symbols: a byte, or 256 + the position inside that unknown window.
//   2. a chain pass per file checks that every sub-chunk ended exactly where the next one started (a false block start -
//      or a missed one - breaks the chain: the file is then NOT handled and goes to the host reader, nothing is guessed)
//      and resolves the windows in order: window i = last 32 KB of text up to the end of sub-chunk i.
//   3. a translate pass turns the 16-bit symbols into text with window i-1, all sub-chunks in parallel.
//
// The decoder itself is written for a warp: all 32 lanes run the same bit reader and the same table lookups (shared
// memory broadcasts, no divergence), lane 0 stores literals, and a match is copied by all lanes at once (coalesced),
// which is where a one-lane decoder spends most of its time.  Host code (lane 0 of 1) compiles from the same source:
// tests/sim/gunzip_harness.cpp checks the whole scheme against zlib on the CPU.
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define GZ_HD __host__ __device__
#else
#define GZ_HD
#endif
#define GZ_UNLIKELY(x) __builtin_expect(!!(x), 0)
#if defined(__CUDA_ARCH__)
#define GZ_SYNC() __syncwarp()
#define GZ_FENCE() asm volatile("" ::: "memory")
#define GZ_BALLOT(p) __ballot_sync(0xFFFFFFFFu, (p))                        /* the lanes' predicates as a mask */
#define GZ_BCAST0(v) __shfl_sync(0xFFFFFFFFu, (v), 0)                       /* lane 0's value */
#define GZ_POPC(m) ((uint32_t)__popc(m))
#define GZ_CTZ(m) (__ffs((int)(m)) - 1)
#define GZ_UNROLL _Pragma("unroll")
#define GZ_NOUNROLL _Pragma("unroll 1")
// The symbol loop addresses its tables and buffers through values the compiler must keep in registers: left to itself it
// re-derives them from the kernel parameters, block and thread ids inside the loop (a dozen instructions per symbol).
// (An empty asm statement is not enough: it leaves no trace in the PTX, and ptxas - which does the re-deriving under the 64
// register limit - sees through the copies.  A value that has been through a VOLATILE shared-memory slot has no other
// derivation.  Once per block: four instructions.)
#define GZ_KEEP64(p) do { unsigned long long v_ = (unsigned long long)(p); asm volatile("{ .reg .b64 t; mov.b64 t, %0; st.volatile.shared.b64 [%1], t; ld.volatile.shared.b64 %0, [%1]; }" : "+l"(v_) : "r"(gz_keep_slot) : "memory"); (p) = reinterpret_cast<decltype(p)>(v_); } while (0)
typedef uint32_t gz_tab_t;                                                  // shared-memory address of a table
__device__ __forceinline__ gz_tab_t gz_tab(const uint32_t *t, uint32_t)
{
    // the same for every lane, and said so: the address then lives in a uniform register and is the offset of the LDS
    return __shfl_sync(0xFFFFFFFFu, (uint32_t)__cvta_generic_to_shared(t), 0);
}
__device__ __forceinline__ uint32_t gz_tab_at(gz_tab_t t, uint32_t i) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(t + (i << 2))); return v; }
#else
// Host builds: one "warp" of one lane (tests/sim/gunzip_harness.cpp, gunzip_fuzz.cpp), or - GZ_EMULATE_WARP, tests/sim/
// gunzip_warp_emu.cpp - 32 threads that run the lanes' code with barriers where the lanes of a warp synchronise (the
// harness defines GZ_SYNC / GZ_FENCE / GZ_BALLOT / GZ_BCAST0 before it includes this file).
#if !defined(GZ_EMULATE_WARP)
#define GZ_SYNC() do { } while (0)
#define GZ_FENCE() do { } while (0)
#define GZ_BALLOT(p) ((p) ? 1u : 0u)
#define GZ_BCAST0(v) (v)
#endif
#define GZ_POPC(m) ((uint32_t)__builtin_popcount(m))
#define GZ_CTZ(m) __builtin_ctz(m)
#define GZ_UNROLL
#define GZ_NOUNROLL
#define GZ_KEEP64(p) do { } while (0)
typedef const uint32_t *gz_tab_t;
inline gz_tab_t gz_tab(const uint32_t *t, uint32_t) { return t; }
inline uint32_t gz_tab_at(gz_tab_t t, uint32_t i) { return t[i]; }
#endif

#define GZ_WINDOW 32768u
#define GZ_LIT_ROOT 9
#define GZ_DIST_ROOT 6
#define GZ_PRE_ROOT 7
#define GZ_LIT_ENTRIES 864u        /* zlib's bound for 286 symbols, 15 bits, root 9 is 852 */
#define GZ_DIST_ENTRIES 608u       /* ... for 30 symbols, 15 bits, root 6: 592 */

// table entry: value << 16 | flags | extra_bits << 4 | bits_to_consume
#define GZ_F_LIT 0x100u            /* value = the byte */
#define GZ_F_BASE 0x200u           /* value = base length / distance, extra_bits follow in the stream */
#define GZ_F_EOB 0x400u
#define GZ_F_SUB 0x800u            /* value = offset of a second-level table indexed by the next extra_bits bits */

enum {
    GZ_OK = 0,
    GZ_FINAL = 1,                  // decoding ended with the stream's last block
    GZ_ERR_INPUT = -1,             // ran past the end of the input
    GZ_ERR_HEADER = -2,            // block type 3, stored LEN/NLEN mismatch, impossible code lengths
    GZ_ERR_SYMBOL = -3,            // a bit pattern that is no code
    GZ_ERR_DISTANCE = -4,          // a match reaches back further than the window (or before the start of the member)
    GZ_ERR_OUTPUT = -5,            // the sub-chunk's symbol region is full
    GZ_ERR_NOT_FOUND = -6,         // no block start within the search limit
    GZ_ERR_TABLE = -7              // second-level tables do not fit (never for codes a DEFLATE stream can carry)
};

struct GzTables {                  // one per decoding warp (shared memory on the device)
    uint32_t lit[GZ_LIT_ENTRIES];
    uint32_t dist[GZ_DIST_ENTRIES];    // also: the code-length code's table, and scratch while `lit` is built
    uint8_t lens[320];
    uint8_t scratch[192];              // scratch while `dist` is built
};

// ------------------------------------------------------------------------------------------------
// bit reader over 32-bit words (the input buffer is 4-byte aligned and padded with >= 8 zero bytes)
// ------------------------------------------------------------------------------------------------
// Two words of the stream in registers and the bit position inside the first: 32 valid bits at any time through ONE
// funnel shift, no 64-bit arithmetic in the symbol loop (round 2's first version kept a 64-bit buffer: a variable
// 64-bit shift is half a dozen instructions, and the decoder is bound by instruction issue - profiles/r2d_gz_decode_ncu.txt).
struct GzBits {
    const uint32_t *w;
    uint32_t n_words;              // words that may be read (files below 16 GiB)
    uint32_t wi;                   // index of `lo`
    uint32_t lo, hi, nxt;          // w[wi], w[wi + 1], w[wi + 2] (zeros behind the end)
    uint32_t pos;                  // bits of `lo` already consumed, < 32
};

GZ_HD inline uint32_t gz_word(const GzBits &b, uint32_t i) { return i < b.n_words ? b.w[i] : 0u; }

GZ_HD inline void gz_bits_seek(GzBits &b, uint64_t bitpos)
{
    b.wi = (uint32_t)(bitpos >> 5);
    b.pos = (uint32_t)(bitpos & 31u);
    b.lo = gz_word(b, b.wi); b.hi = gz_word(b, b.wi + 1u); b.nxt = gz_word(b, b.wi + 2u);
}
GZ_HD inline uint64_t gz_bits_pos(const GzBits &b) { return (uint64_t)b.wi * 32u + b.pos; }
// the next 32 bits of the stream
GZ_HD inline uint32_t gz_peek(const GzBits &b)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(b.lo, b.hi, b.pos);
#else
    return b.pos ? (b.lo >> b.pos) | (b.hi << (32u - b.pos)) : b.lo;
#endif
}
GZ_HD inline void gz_skip(GzBits &b, uint32_t k)              // k <= 32
{
    b.pos += k;
    if (b.pos >= 32u) { b.pos -= 32u; b.lo = b.hi; b.hi = b.nxt; ++b.wi; b.nxt = gz_word(b, b.wi + 2u); }
}
GZ_HD inline void gz_refill(GzBits &) {}                      // (32 bits are always there)
GZ_HD inline uint32_t gz_take(GzBits &b, uint32_t k)          // k <= 32 bits
{
    const uint32_t v = gz_peek(b) & (k >= 32u ? 0xFFFFFFFFu : ((1u << k) - 1u));
    gz_skip(b, k);
    return v;
}

// back to pos < 32 after a symbol's bits were added to pos: one step for pos < 64, two for pos < 96
GZ_HD inline void gz_norm1(GzBits &b)
{
    if (b.pos >= 32u) { b.pos -= 32u; b.lo = b.hi; b.hi = b.nxt; ++b.wi; b.nxt = gz_word(b, b.wi + 2u); }
}
GZ_HD inline void gz_norm2(GzBits &b)
{
    if (b.pos >= 32u) {
        b.pos -= 32u; b.lo = b.hi; b.hi = b.nxt; ++b.wi; b.nxt = gz_word(b, b.wi + 2u);
        if (b.pos >= 32u) { b.pos -= 32u; b.lo = b.hi; b.hi = b.nxt; ++b.wi; b.nxt = gz_word(b, b.wi + 2u); }
    }
}
// 32 bits of the stream from bit q of the words in registers, q < 64
GZ_HD inline uint32_t gz_peek_at(const GzBits &b, uint32_t q)
{
    const bool second = q >= 32u;
    const uint32_t a = second ? b.hi : b.lo, c = second ? b.nxt : b.hi;
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(a, c, q);                                     // (shift amount modulo 32)
#else
    const uint32_t s = q & 31u;
    return s ? (a >> s) | (c << (32u - s)) : a;
#endif
}
// n bits (n < 32) of v from bit `from` (from < 32): a shift and one zero-extension (SGXT; PTX bfe clamps its byte-sized
// operands first - five instructions)
GZ_HD inline uint32_t gz_bfe(uint32_t v, uint32_t from, uint32_t n)
{
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("szext.clamp.u32 %0, %1, %2;" : "=r"(r) : "r"(v >> from), "r"(n));
    return r;
#else
    return (v >> from) & ((1u << n) - 1u);
#endif
}
// the markers in front of a symbol region: prefix[j] for j in [0, 32768) = 256 + j = "byte j of the unknown window"
GZ_HD inline void gz_marker_prefix(uint16_t *prefix, uint32_t tid, uint32_t n_threads)
{
    for (uint32_t j = tid; j < 32768u; j += n_threads) prefix[j] = (uint16_t)(256u + j);
}

GZ_HD inline uint32_t gz_bitrev(uint32_t v, int n)
{
#if defined(__CUDA_ARCH__)
    return __brev(v) >> (32 - n);
#else
    uint32_t r = 0;
    for (int i = 0; i < n; ++i) r |= ((v >> i) & 1u) << (n - 1 - i);
    return r;
#endif
}

// RFC 1951 3.2.5: base value and extra bits of length symbol s (257..285 -> 0..28) / distance symbol d (0..29)
GZ_HD inline uint32_t gz_len_entry(int s)
{
    if (s < 8) return (uint32_t)(3 + s) << 16 | GZ_F_BASE;
    if (s == 28) return 258u << 16 | GZ_F_BASE;
    const uint32_t e = (uint32_t)(s - 4) >> 2;
    return (3u + ((4u + ((uint32_t)s & 3u)) << e)) << 16 | GZ_F_BASE | e << 4;
}
GZ_HD inline uint32_t gz_dist_entry(int d)
{
    if (d < 4) return (uint32_t)(1 + d) << 16 | GZ_F_BASE;
    const uint32_t e = ((uint32_t)d >> 1) - 1u;
    return (1u + ((2u + ((uint32_t)d & 1u)) << e)) << 16 | GZ_F_BASE | e << 4;
}
GZ_HD inline int gz_clen_order(int i)
{
    const uint64_t lo = 16ull | 17ull << 5 | 18ull << 10 | 0ull << 15 | 8ull << 20 | 7ull << 25 | 9ull << 30 | 6ull << 35 | 10ull << 40 | 5ull << 45 |
                        11ull << 50 | 4ull << 55;
    const uint64_t hi = 12ull | 3ull << 5 | 13ull << 10 | 2ull << 15 | 14ull << 20 | 1ull << 25 | 15ull << 30;
    return (int)((i < 12 ? lo >> (5 * i) : hi >> (5 * (i - 12))) & 31u);
}

// ------------------------------------------------------------------------------------------------
// canonical Huffman code -> decoding table (root-bit first level + second-level tables for longer codes)
// ------------------------------------------------------------------------------------------------
// kind 0: literal/length alphabet, 1: distance alphabet, 2: code-length alphabet (value = symbol, flagged LIT).
// Kraft sum of lens[0..n): returns 0 complete, 1 incomplete, -1 over-subscribed; *max_len = longest code
GZ_HD inline int gz_kraft(const uint8_t *lens, int n, int *max_len, int *n_codes)
{
    uint32_t sum = 0;                      // in units of 2^-15
    int mx = 0, nc = 0;
    for (int s = 0; s < n; ++s) {
        const int l = lens[s];
        if (l) { sum += 32768u >> l; ++nc; if (l > mx) mx = l; }
    }
    *max_len = mx; *n_codes = nc;
    return sum == 32768u ? 0 : sum < 32768u ? 1 : -1;
}

GZ_HD inline uint32_t gz_symbol_entry(int kind, int s)
{
    if (kind == 2) return (uint32_t)s << 16 | GZ_F_LIT;
    if (kind == 1) return s < 30 ? gz_dist_entry(s) : 0u;
    if (s < 256) return (uint32_t)s << 16 | GZ_F_LIT;
    if (s == 256) return GZ_F_EOB;
    return s < 286 ? gz_len_entry(s - 257) : 0u;          // 286 / 287 may have codes (fixed blocks) but may not occur
}

// tab[0..cap): built from lens[0..n).  scratch: (1 << root) + 2 * n bytes.  Returns 0, GZ_ERR_HEADER (over-subscribed),
// GZ_ERR_TABLE.  Unused patterns of an incomplete code decode to an entry without flags (an error when met).
GZ_HD inline int gz_build(uint32_t *tab, uint32_t cap, int root, int kind, const uint8_t *lens, int n, uint8_t *scratch, int lane, int nl)
{
    uint8_t *submax = scratch;                                      // per first-level prefix: longest code below it, minus root
    uint16_t *code = reinterpret_cast<uint16_t *>(scratch + (1u << root));
    const uint32_t rsize = 1u << root;
    for (uint32_t i = (uint32_t)lane; i < cap; i += (uint32_t)nl) tab[i] = 0u;
    for (uint32_t i = (uint32_t)lane; i < rsize; i += (uint32_t)nl) submax[i] = 0;
    GZ_SYNC();
    int rc = 0;
    if (lane == 0) {
        uint32_t count[16], next[16];
        GZ_UNROLL
        for (int l = 0; l < 16; ++l) count[l] = 0;
        for (int s = 0; s < n; ++s) count[lens[s]]++;
        uint32_t c = 0;
        int left = 1;
        count[0] = 0;
        for (int l = 1; l < 16; ++l) {
            c = (c + count[l - 1]) << 1;
            next[l] = c;
            left = (left << 1) - (int)count[l];
            if (left < 0) rc = GZ_ERR_HEADER;
        }
        bool any_long = false;
        if (rc == 0)
            for (int s = 0; s < n; ++s) {
                const int l = lens[s];
                if (!l) continue;
                const uint32_t cd = next[l]++;
                code[s] = (uint16_t)gz_bitrev(cd, l);
                if (l > root) {
                    const uint32_t prefix = code[s] & (rsize - 1u);
                    if (submax[prefix] < l - root) submax[prefix] = (uint8_t)(l - root);
                    any_long = true;
                }
            }
        // second-level tables, in prefix order (none for a code without long words: the code-length code, most distance codes)
        uint32_t at = rsize;
        for (uint32_t p = 0; any_long && p < rsize && rc == 0; ++p)
            if (submax[p]) {
                if (at + (1u << submax[p]) > cap) { rc = GZ_ERR_TABLE; break; }
                tab[p] = at << 16 | GZ_F_SUB | (uint32_t)submax[p] << 4 | (uint32_t)root;
                at += 1u << submax[p];
            }
    }
    GZ_SYNC();
    rc = GZ_BCAST0(rc);
    if (rc) return rc;
    for (int s = lane; s < n; s += nl) {
        const int l = lens[s];
        if (!l) continue;
        const uint32_t e = gz_symbol_entry(kind, s);
        const uint32_t rev = code[s];
        if (l <= root) {
            for (uint32_t idx = rev; idx < rsize; idx += 1u << l) tab[idx] = e | (uint32_t)l;
        } else {
            const uint32_t sub = tab[rev & (rsize - 1u)];
            const uint32_t off = sub >> 16, sb = (sub >> 4) & 15u;
            for (uint32_t idx = rev >> root; idx < (1u << sb); idx += 1u << (l - root)) tab[off + idx] = e | (uint32_t)(l - root);
        }
    }
    GZ_SYNC();
    return 0;
}

// one symbol of a code: the entry, with its bits consumed (second-level lookup included).  At most 15 bits.
GZ_HD inline uint32_t gz_decode_sym(GzBits &b, const uint32_t *tab, int root)
{
    const uint32_t bits = gz_peek(b);
    uint32_t e = tab[bits & ((1u << root) - 1u)], used = 0;
    if (e & GZ_F_SUB) {
        used = (uint32_t)root;
        e = tab[(e >> 16) + ((bits >> root) & ((1u << ((e >> 4) & 15u)) - 1u))];
    }
    gz_skip(b, used + (e & 15u));
    return e;
}

// ------------------------------------------------------------------------------------------------
// block header (the bit reader stands behind BFINAL / BTYPE) -> tables.  check_only: code lengths are read and
// judged, no literal/length or distance table is built (the block finder).  strict: both codes must be complete
// (what every zlib-family compressor writes); otherwise zlib's own rule (incomplete only as a single 1-bit code).
// ------------------------------------------------------------------------------------------------
GZ_HD inline int gz_dynamic_header(GzBits &b, GzTables &t, bool check_only, bool strict, int lane, int nl)
{
    gz_refill(b);
    const int nlen = (int)gz_take(b, 5) + 257, ndist = (int)gz_take(b, 5) + 1, ncode = (int)gz_take(b, 4) + 4;
    if (nlen > 286 || ndist > 30) return GZ_ERR_HEADER;
    GZ_SYNC();
    if (lane == 0) {
        for (int i = 0; i < 19; ++i) t.lens[i] = 0;
    }
    GZ_SYNC();
    for (int i = 0; i < ncode; ++i) {
        gz_refill(b);
        const uint32_t v = gz_take(b, 3);
        if (lane == 0) t.lens[gz_clen_order(i)] = (uint8_t)v;
    }
    GZ_SYNC();
    {
        int mx, nc;
        const int k = gz_kraft(t.lens, 19, &mx, &nc);
        if (k < 0 || (k > 0 && (strict || mx != 1))) return GZ_ERR_HEADER;
    }
    // the code-length code lives in the distance table's storage (the build is done with lens[] before they are overwritten)
    int rc = gz_build(t.dist, 1u << GZ_PRE_ROOT, GZ_PRE_ROOT, 2, t.lens, 19, reinterpret_cast<uint8_t *>(t.dist + 256), lane, nl);
    if (rc) return rc;
    int i = 0, prev = 0;
    const int total = nlen + ndist;
    while (i < total) {
        gz_refill(b);
        const uint32_t e = gz_decode_sym(b, t.dist, GZ_PRE_ROOT);
        if (!(e & GZ_F_LIT)) return GZ_ERR_HEADER;
        const int sym = (int)(e >> 16);
        if (sym < 16) { if (lane == 0) t.lens[i] = (uint8_t)sym; prev = sym; ++i; continue; }
        int rep, val = 0;
        if (sym == 16) { if (i == 0) return GZ_ERR_HEADER; val = prev; rep = 3 + (int)gz_take(b, 2); }
        else if (sym == 17) rep = 3 + (int)gz_take(b, 3);
        else rep = 11 + (int)gz_take(b, 7);
        if (i + rep > total) return GZ_ERR_HEADER;
        for (int k = lane; k < rep; k += nl) t.lens[i + k] = (uint8_t)val;
        i += rep; prev = val;
    }
    GZ_SYNC();
    if (gz_bits_pos(b) > (uint64_t)b.n_words * 32u) return GZ_ERR_INPUT;
    if (t.lens[256] == 0) return GZ_ERR_HEADER;                         // no end-of-block code
    int mx, nc;
    int k = gz_kraft(t.lens, nlen, &mx, &nc);
    if (k < 0 || (k > 0 && (strict || mx != 1))) return GZ_ERR_HEADER;
    k = gz_kraft(t.lens + nlen, ndist, &mx, &nc);
    if (k < 0 || (k > 0 && (strict || mx > 1))) return GZ_ERR_HEADER;   // zlib's rule: incomplete only as no code at all or a single 1-bit code
    if (check_only) return 0;
    // distance lengths move out of the way (lit's build uses dist[] as scratch), then lit, then dist
    uint8_t *dl = t.scratch + 128;
    GZ_SYNC();
    for (int s = lane; s < ndist; s += nl) dl[s] = t.lens[nlen + s];
    GZ_SYNC();
    rc = gz_build(t.lit, GZ_LIT_ENTRIES, GZ_LIT_ROOT, 0, t.lens, nlen, reinterpret_cast<uint8_t *>(t.dist), lane, nl);
    if (rc) return rc;
    return gz_build(t.dist, GZ_DIST_ENTRIES, GZ_DIST_ROOT, 1, dl, ndist, t.scratch, lane, nl);
}

GZ_HD inline int gz_fixed_header(GzTables &t, int lane, int nl)
{
    GZ_SYNC();
    for (int s = lane; s < 288; s += nl) t.lens[s] = (uint8_t)(s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8);
    uint8_t *dl = t.scratch + 128;
    for (int s = lane; s < 30; s += nl) dl[s] = 5;
    GZ_SYNC();
    const int rc = gz_build(t.lit, GZ_LIT_ENTRIES, GZ_LIT_ROOT, 0, t.lens, 288, reinterpret_cast<uint8_t *>(t.dist), lane, nl);
    if (rc) return rc;
    return gz_build(t.dist, GZ_DIST_ENTRIES, GZ_DIST_ROOT, 1, dl, 30, t.scratch, lane, nl);
}

// ------------------------------------------------------------------------------------------------
// decode from a block start to the first block boundary at or after stop_bit (or the end of the stream)
// ------------------------------------------------------------------------------------------------
// out[-32768..cap]: 16-bit symbols - out[0..cap) count, out[cap] is a guard slot (literals and the pending copy store
// without a test), out[-j] holds the marker of the unknown window's j-th last byte (gz_marker_prefix: a match that reaches
// back before the sub-chunk copies its markers like any other symbol); window: how many bytes before out[0] a match may reach (32768 for a sub-chunk in the
// middle of a stream, 0 at the start of a member).  Returns GZ_OK (stopped at a boundary: *end_bit), GZ_FINAL (the last
// block ended at *end_bit) or an error.  *n_out = symbols written.
GZ_HD inline int gz_decode_blocks(GzBits &b, GzTables &t, uint16_t *out, uint32_t cap, uint32_t window, uint64_t stop_bit,
                                  uint32_t *n_out, uint64_t *end_bit, int lane, int nl)
{
    uint32_t o = 0;
    int rc = GZ_OK;
    uint32_t p_at = cap;                                                 // this lane's copied symbol that is still to be stored, and where
    uint16_t p_val = 0;                                                  // (nothing yet: the guard slot)
#if defined(__CUDA_ARCH__)
    const uint32_t gz_keep_slot = (uint32_t)__cvta_generic_to_shared(t.scratch + 64);       // 8 bytes of the tables' scratch area (free between table builds)
#else
    const uint32_t gz_keep_slot = 0;
#endif
    GZ_KEEP64(out);
    GZ_KEEP64(b.w);
    for (;;) {
        gz_refill(b);
        // a sub-chunk ends at the first boundary at or after the next cut whose block is one the finder can see (not
        // final, dynamic codes): final, stored and fixed-code blocks are decoded by whoever arrives at them
        if (gz_bits_pos(b) >= stop_bit && (gz_peek(b) & 7u) == 4u) break;
        const uint32_t last = gz_take(b, 1), type = gz_take(b, 2);
        if (type == 3) { rc = GZ_ERR_HEADER; break; }
        if (type == 0) {
            uint64_t pos = (gz_bits_pos(b) + 7u) & ~7ull;                 // stored: byte aligned LEN, ~LEN, bytes
            gz_bits_seek(b, pos);
            gz_refill(b);
            const uint32_t len = gz_take(b, 16), nlen = gz_take(b, 16);
            if ((len ^ 0xFFFFu) != nlen) { rc = GZ_ERR_HEADER; break; }
            pos += 32;
            if (pos + 8ull * len > (uint64_t)b.n_words * 32u) { rc = GZ_ERR_INPUT; break; }
            if (o + len > cap) { rc = GZ_ERR_OUTPUT; break; }
            const uint8_t *bytes = reinterpret_cast<const uint8_t *>(b.w) + (pos >> 3);
            for (uint32_t i = (uint32_t)lane; i < len; i += (uint32_t)nl) out[o + i] = bytes[i];
            o += len;
            gz_bits_seek(b, pos + 8ull * len);
        } else {
            rc = type == 1 ? gz_fixed_header(t, lane, nl) : gz_dynamic_header(b, t, false, false, lane, nl);
            if (rc) break;
            // The symbol loop.  The three words in registers hold 64 valid bits behind `pos`, a literal/length code with its
            // extra bits and the distance code with its extra bits are 48 at most: both are read from the words as they
            // stand (the second peek picks its pair of words), and the reader is brought back to pos < 32 once per symbol.
            // Everything is 32-bit arithmetic; the conditional load / store of the match copy are predicated instructions,
            // not branches (a warp's instruction stream is one dependent chain: the kernel is bound by instructions per symbol).
            const gz_tab_t lit = gz_tab(t.lit, gz_keep_slot), dtab = gz_tab(t.dist, gz_keep_slot);
            for (;;) {
                gz_norm1(b);                                              // the one place where the words move on
                uint32_t bits = gz_peek(b);
                uint32_t e = gz_tab_at(lit, bits & ((1u << GZ_LIT_ROOT) - 1u));
                uint32_t used = e & 15u;
                if (GZ_UNLIKELY(e & GZ_F_SUB)) {
                    // (written as a loop - second-level entries never point on - so that it stays a branch: as predicated
                    // instructions its eight issue slots would be spent on every symbol)
                    GZ_NOUNROLL
                    do e = gz_tab_at(lit, (e >> 16) + gz_bfe(bits, GZ_LIT_ROOT, (e >> 4) & 15u)); while (GZ_UNLIKELY(e & GZ_F_SUB));
                    used = GZ_LIT_ROOT + (e & 15u);
                }
                if (e & GZ_F_BASE) {
                    const uint32_t xl = (e >> 4) & 15u;
                    const uint32_t len = (e >> 16) + gz_bfe(bits, used, xl);
                    const uint32_t q = b.pos + used + xl;                 // <= 31 + 15 + 5
                    bits = gz_peek_at(b, q);
                    uint32_t d = gz_tab_at(dtab, bits & ((1u << GZ_DIST_ROOT) - 1u));
                    used = d & 15u;
                    if (GZ_UNLIKELY(d & GZ_F_SUB)) {
                        GZ_NOUNROLL
                        do d = gz_tab_at(dtab, (d >> 16) + gz_bfe(bits, GZ_DIST_ROOT, (d >> 4) & 15u)); while (GZ_UNLIKELY(d & GZ_F_SUB));
                        used = GZ_DIST_ROOT + (d & 15u);
                    }
                    const uint32_t xd = (d >> 4) & 15u;
                    // a pattern that is no distance code has an entry without a value and without extra bits: distance 0
                    const uint32_t dist = (d >> 16) + gz_bfe(bits, used, xd);
                    b.pos = q + used + xd;                                // <= 51 + 15 + 13
                    if (GZ_UNLIKELY(b.pos >= 64u)) { b.pos -= 32u; b.lo = b.hi; b.hi = b.nxt; ++b.wi; b.nxt = gz_word(b, b.wi + 2u); }
                    // A copied symbol is loaded now and stored when the NEXT match arrives (or the blocks end): the symbols
                    // were written past L1, so the load is an L2 round trip, and a store right behind it would hold the warp -
                    // in-order issue - for all of it.  Nothing reads the place in between: literals only store, and the next
                    // match stores the pending symbol before it loads.  Lanes behind the end of a match do what its last
                    // lane does (the same value to the same place); before the first match the pending store goes to the
                    // guard slot: no lane ever needs a predicate.  A source before out[0] is a place in the marker prefix.
                    out[p_at] = p_val;
                    GZ_FENCE();                                          // (the lanes run in step: a barrier would be a test and a no-op)
                    uint32_t k = len - 1u < (uint32_t)lane ? len - 1u : (uint32_t)lane;
                    const uint32_t at = o + k;
                    // ONE branch for everything that is not a plain short match: the three things that can be wrong with it
                    // (sorted out on the cold side), a match that overlaps itself (a repeating pattern of `dist` symbols) or is
                    // longer than the warp is wide
                    if (GZ_UNLIKELY((dist - 1u >= o + window) | (o + len >= cap) | (dist < len) | (len > (uint32_t)nl))) {
                        if ((dist - 1u >= o + window) | (o + len >= cap)) {
                            rc = dist == 0u ? GZ_ERR_SYMBOL : dist > o + window ? GZ_ERR_DISTANCE : GZ_ERR_OUTPUT;
                            p_at = cap;                                   // (nothing pending any more)
                            break;
                        }
                        const bool overlap = dist < len;
                        GZ_SYNC();
                        GZ_NOUNROLL
                        for (uint32_t i = (uint32_t)(lane + nl); i < len; i += (uint32_t)nl)
                            out[o + i] = out[(int32_t)(o - dist) + (int32_t)(overlap ? i % dist : i)];
                        if (overlap) k %= dist;
                    }
                    p_val = out[(int32_t)(o - dist) + (int32_t)k];        // >= -32768: before out[0] lies the unknown window
                    p_at = at;
                    o += len;
                    continue;
                }
                b.pos += used;
                if (e & GZ_F_LIT) {
                    // every lane stores the same value to the same place (one transaction, no branch around it); a full
                    // region keeps `o` at cap (the guard slot takes the stores) - reported when the block ends or a match arrives
                    out[o] = (uint16_t)(e >> 16);
                    o = o + 1u < cap ? o + 1u : cap;
                    continue;
                }
                gz_norm1(b);
                if (!(e & GZ_F_EOB)) rc = GZ_ERR_SYMBOL;
                else if (o >= cap) rc = GZ_ERR_OUTPUT;                 // (a region filled to the last symbol counts as overflowed)
                break;
            }
            if (rc) break;
            if (gz_bits_pos(b) > (uint64_t)b.n_words * 32u) { rc = GZ_ERR_INPUT; break; }
        }
        if (last) { rc = GZ_FINAL; break; }
    }
    out[p_at] = p_val;
    GZ_SYNC();
    *n_out = o;
    *end_bit = gz_bits_pos(b);
    return rc;
}

// ------------------------------------------------------------------------------------------------
// block finder: the first bit position >= from_bit (and < limit_bit) where a dynamic-code block plausibly starts
// ------------------------------------------------------------------------------------------------
// 96 bits of the stream starting at bit p
GZ_HD inline void gz_peek96(const GzBits &b, uint64_t p, uint64_t *lo, uint32_t *hi)
{
    const uint32_t i = (uint32_t)(p >> 5);
    const uint32_t s = (uint32_t)(p & 31u);
    const uint64_t w0 = gz_word(b, i), w1 = gz_word(b, i + 1), w2 = gz_word(b, i + 2), w3 = gz_word(b, i + 3);
    const uint64_t a = w0 | w1 << 32, c = w2 | w3 << 32;
    *lo = s ? (a >> s) | (c << (64 - s)) : a;
    *hi = (uint32_t)(c >> s);
}

// Kraft sum (in units of 1/128) of three 3-bit code lengths at once: kraft9[v] for the 9 bits v.  512 bytes, filled once per
// CTA / process; the finder's test of a position is then seven lookups instead of a loop over up to 19 lengths - with 32
// lanes on 32 positions nearly every warp iteration ran that loop in full, a tenth to a quarter of the decode kernel's work.
GZ_HD inline void gz_kraft9_fill(uint8_t *kraft9, uint32_t tid, uint32_t n_threads)
{
    for (uint32_t v = tid; v < 512u; v += n_threads) {
        uint32_t sum = 0;
        for (uint32_t i = 0; i < 3u; ++i) { const uint32_t l = (v >> (3u * i)) & 7u; sum += l ? 128u >> l : 0u; }
        kraft9[v] = (uint8_t)sum;                                        // <= 192
    }
}

// cheap test of one position: not-final dynamic block, sane code counts, complete code-length code
GZ_HD inline bool gz_candidate(const GzBits &b, uint64_t p, const uint8_t *kraft9)
{
    uint64_t lo; uint32_t hi;
    gz_peek96(b, p, &lo, &hi);
    const uint32_t h = (uint32_t)lo;
    if ((h & 7u) != 4u) return false;                                   // BFINAL = 0, BTYPE = 2
    if (((h >> 3) & 31u) > 29u || ((h >> 8) & 31u) > 29u) return false;
    const uint32_t ncode = ((h >> 13) & 15u) + 4u;
    uint64_t bits = lo >> 17 | (uint64_t)hi << 47;                      // 3 bits per code length (19 x 3 = 57 bits: all inside `bits`)
    bits &= (1ull << (3u * ncode)) - 1ull;                               // lengths that are not there count as 0
    const uint32_t x0 = (uint32_t)bits, x1 = (uint32_t)(bits >> 27), x2 = (uint32_t)(bits >> 54);
    const uint32_t sum = kraft9[x0 & 511u] + kraft9[(x0 >> 9) & 511u] + kraft9[(x0 >> 18) & 511u] +
                         kraft9[x1 & 511u] + kraft9[(x1 >> 9) & 511u] + kraft9[(x1 >> 18) & 511u] + kraft9[x2 & 511u];
    return sum == 128u;
}

// Two stages, because the stages' costs differ by an order of magnitude and one position in nine passes the first: every
// lane tests its own position for the 13 header bits alone (not final, dynamic codes, sane code counts) and the survivors
// are QUEUED; whenever the queue holds a warp's worth, every lane runs the Kraft test of the code-length code on one of
// them - all 32 lanes busy, instead of the three or four whose position happened to pass - and what survives that gets the
// whole header parsed, in position order, so the first hit is the first block start.  The queue (positions relative to
// from_bit, 2 x nl words) lives in the literal/length table, which nothing else uses before a block is decoded.
GZ_HD inline int gz_find_block(GzBits &b, GzTables &t, const uint8_t *kraft9, uint64_t from_bit, uint64_t limit_bit, uint64_t *found, int lane, int nl)
{
    uint32_t *queue = t.lit;
    uint32_t count = 0;
    // the scan runs on 32-bit offsets from from_bit (the search limit is a few megabytes) over the words from from_bit's own
    if (limit_bit <= from_bit) return GZ_ERR_NOT_FOUND;
    const uint32_t *wp = b.w + (from_bit >> 5);
    const uint32_t nw = b.n_words - (uint32_t)(from_bit >> 5), s0 = (uint32_t)from_bit & 31u;
    const uint32_t n_off = limit_bit - from_bit > 0xFFFFFF00ull ? 0xFFFFFF00u : (uint32_t)(limit_bit - from_bit);
    const uint32_t lt = (1u << lane) - 1u;                              // the lanes below this one
    for (uint32_t off = 0;; off += (uint32_t)nl) {
        const bool more = off < n_off;
        if (more) {
            const uint32_t q = s0 + off + (uint32_t)lane;
            const uint32_t wi = q >> 5;
            const uint32_t w0 = wi < nw ? wp[wi] : 0u, w1 = wi + 1u < nw ? wp[wi + 1u] : 0u;
#if defined(__CUDA_ARCH__)
            const uint32_t h = __funnelshift_r(w0, w1, q);
#else
            const uint32_t h = (q & 31u) ? (w0 >> (q & 31u)) | (w1 << (32u - (q & 31u))) : w0;
#endif
            const bool pass = off + (uint32_t)lane < n_off && (h & 7u) == 4u && ((h >> 3) & 31u) <= 29u && ((h >> 8) & 31u) <= 29u;
            const uint32_t m = GZ_BALLOT(pass);
            if (pass) queue[count + GZ_POPC(m & lt)] = off + (uint32_t)lane;
            count += GZ_POPC(m);
        }
        if (count >= (uint32_t)nl || (!more && count)) {
            GZ_SYNC();
            const uint32_t take = count < (uint32_t)nl ? count : (uint32_t)nl;
            const bool mine = (uint32_t)lane < take;
            const bool cand = mine && gz_candidate(b, from_bit + queue[mine ? lane : 0], kraft9);
            uint32_t c = GZ_BALLOT(cand);
            while (c) {                                                  // survivors in order, the whole warp on each
                const int k = GZ_CTZ(c);
                c &= c - 1u;
                const uint64_t at = from_bit + queue[k];
                gz_bits_seek(b, at + 3u);
                if (gz_dynamic_header(b, t, true, true, lane, nl) == 0) { *found = at; return 0; }
            }
            const uint32_t rest = count - take;                          // < nl: they move to the front
            const uint32_t v = (uint32_t)lane < rest ? queue[take + (uint32_t)lane] : 0u;
            GZ_SYNC();
            if ((uint32_t)lane < rest) queue[lane] = v;
            GZ_SYNC();
            count = rest;
        }
        if (!more && !count) break;
    }
    return GZ_ERR_NOT_FOUND;
}

// ------------------------------------------------------------------------------------------------
// one sub-chunk: find (unless the start is known) + decode
// ------------------------------------------------------------------------------------------------
struct GzSubResult {
    uint64_t start_bit, end_bit;       // where decoding started / ended (bit offsets inside the file's bytes)
    uint32_t n_out;
    int32_t status;                    // GZ_OK / GZ_FINAL / error
};

// words/n_words: the FILE's compressed bytes (4-byte aligned, zero padded).  known_start: bit offset of the member's
// first block for the sub-chunk that begins a member, else ~0.  cut_bit / next_cut_bit: this sub-chunk's and the next
// one's cut.  search_limit_bits: how far past the cut the finder looks.
GZ_HD inline void gz_subchunk(const uint32_t *words, uint64_t n_words, uint64_t known_start, uint64_t cut_bit, uint64_t next_cut_bit,
                              uint64_t search_limit_bits, uint16_t *out, uint32_t cap, GzTables &t, const uint8_t *kraft9, GzSubResult *res, int lane, int nl)
{
    GzBits b;
    b.w = words; b.n_words = (uint32_t)n_words;
    uint64_t start = known_start;
    int rc = 0;
    const bool find_only = (search_limit_bits >> 63) != 0;              // (profiling: the finder's share of the kernel - S2_GZ_FIND_ONLY)
    search_limit_bits &= ~(1ull << 63);
    if (known_start == ~0ull) {
        const uint64_t end_bits = (uint64_t)n_words * 32u;
        uint64_t limit = cut_bit + search_limit_bits;
        if (limit > end_bits) limit = end_bits;
        rc = gz_find_block(b, t, kraft9, cut_bit, limit, &start, lane, nl);
    }
    uint32_t n_out = 0;
    uint64_t end_bit = start;
    if (rc == 0 && !find_only) {
        gz_bits_seek(b, start);
        rc = gz_decode_blocks(b, t, out, cap, known_start == ~0ull ? GZ_WINDOW : 0u, next_cut_bit, &n_out, &end_bit, lane, nl);
    }
    if (lane == 0) { res->start_bit = rc == GZ_ERR_NOT_FOUND ? ~0ull : start; res->end_bit = end_bit; res->n_out = n_out; res->status = rc; }
}

// ------------------------------------------------------------------------------------------------
// chain pass and translate pass (shared by the kernels and the host harness)
// ------------------------------------------------------------------------------------------------
// window after a sub-chunk = the last 32 KB of text up to its end: from its own symbols (markers resolved through the
// previous window) and, where it produced fewer than 32 K symbols, the tail of the previous window
GZ_HD inline void gz_next_window(const uint8_t *prev_win, const uint16_t *out, uint32_t n_out, uint8_t *next_win, uint32_t tid, uint32_t n_threads)
{
    for (uint32_t j = tid; j < GZ_WINDOW; j += n_threads) {
        const int64_t k = (int64_t)n_out - (int64_t)GZ_WINDOW + (int64_t)j;
        uint8_t v;
        if (k >= 0) { const uint16_t sym = out[k]; v = sym < 256 ? (uint8_t)sym : prev_win[sym - 256]; }
        else v = prev_win[(int64_t)GZ_WINDOW + k];
        next_win[j] = v;
    }
}

GZ_HD inline void gz_translate(const uint8_t *prev_win, const uint16_t *out, uint32_t n_out, uint8_t *text, uint32_t tid, uint32_t n_threads)
{
    for (uint32_t j = tid; j < n_out; j += n_threads) {
        const uint16_t sym = out[j];
        text[j] = sym < 256 ? (uint8_t)sym : prev_win[sym - 256];
    }
}
