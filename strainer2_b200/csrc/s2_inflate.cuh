// s2_inflate.cuh - a DEFLATE (RFC 1951) / gzip (RFC 1952) decoder written once for host and device.
//
// STATUS: groundwork for SURVEY 8(f) rank 1, second half - NOT on the product path yet.  The decoder is unit-tested on
// the host against zlib (tests/test_host.py::test_inflate_*, through tests/sim/inflate_harness.cpp); the probe kernels
// that wrap it (tools/gunzip_probe.cu) decoded 2,048 .gz images correctly on a B200 at 8.5 - 14.3 GB/s of text
// (profiles/r1s_gunzip_kernel_v0_probe.txt); nothing in libstrainer2_b200.so calls it yet.  After those runs the match copy,
// the input refill and the length / distance tables were rewritten for the device (no load that waits for the thread's own
// store, loads issued in groups, no tables on the thread's stack): same results on the host tests, effect on the GPU not
// measured yet.
//
// Why it exists.  The reference reads every input through zlib's gzread (/root/reference/src/genome_compare.c:194,
// src/strain_detect.c:417-433), and the inputs it ships and documents are ORDINARY single-member .gz files
// (test/example.sh: *.fna.gz, *.fastq.gz).  The Blackwell decompression engine that s2_ingest.cu drives cannot take
// those: the end of the member's DEFLATE stream is unknown without inflating it, a wrong length costs the CUDA context,
// and 4 MB of text per stream is the engine's limit (profiles/r1s_hw_decompression_error_probe.txt).  A software decoder
// has no such limits.  One DEFLATE stream is sequential, so the parallelism is across files: config #2 is 2,000 genome
// files, and a decoder instance needs about 3.4 KB of tables, so thousands run side by side.  A single multi-GB FASTQ
// stream has no such parallelism and stays on host zlib (or becomes BGZF).
//
// Design of the decoder: a 64-bit bit buffer refilled eight loads at a time; canonical Huffman codes decoded through a first-level
// table (10 bits for literal/length codes, 8 bits for distance codes: one load for nearly every symbol of real data)
// with the canonical count/symbol arrays as the slow path for longer codes; every read and every write is bounds-
// checked, damaged input ends in an error code, never in an out-of-bounds access.
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define S2I_HD __host__ __device__
#else
#define S2I_HD
#endif

enum {
    S2I_OK = 0,
    S2I_ERR_TRUNCATED = -1,        // the input ended inside the stream
    S2I_ERR_BLOCK_TYPE = -2,       // reserved block type 3
    S2I_ERR_STORED_LEN = -3,       // LEN / NLEN of a stored block do not match
    S2I_ERR_CODE_LENGTHS = -4,     // over-subscribed or otherwise impossible set of code lengths
    S2I_ERR_BAD_SYMBOL = -5,       // a bit pattern that is no code, or a symbol that may not occur
    S2I_ERR_DISTANCE = -6,         // a match reaches back before the start of the output
    S2I_ERR_OUTPUT_FULL = -7,      // the text does not fit the destination
    S2I_ERR_GZIP_HEADER = -8,      // not a gzip member (magic, method, reserved flags)
    S2I_ERR_GZIP_SIZE = -9         // ISIZE of the trailer differs from the bytes produced
};

#define S2I_LIT_BITS 10
#define S2I_DIST_BITS 8
#define S2I_MAX_BITS 15

struct S2InfBits {
    const uint8_t *p;
    uint64_t n, pos;               // input length, next byte to load
    uint64_t buf;                  // bit buffer, next bit in bit 0
    unsigned cnt;                  // valid bits in buf
    int overrun;                   // bits were consumed that the input does not have
};

// top the bit buffer up to at least 57 bits (or to the end of the input).  All loads first, then the shifts: a device
// thread issues in order, so a load -> shift -> or chain per byte would wait out one cache latency per byte
S2I_HD inline void s2i_refill(S2InfBits &b)
{
    const unsigned want = (64u - b.cnt) >> 3;                       // whole bytes that fit
    const uint64_t avail = b.n - b.pos;
    const unsigned take = avail < want ? (unsigned)avail : want;
    uint8_t r[8];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (unsigned k = 0; k < 8; ++k) r[k] = k < take ? b.p[b.pos + k] : (uint8_t)0;
    uint64_t w = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (unsigned k = 0; k < 8; ++k) w |= (uint64_t)r[k] << (8 * k);
    if (take) b.buf |= w << b.cnt;                                  // cnt <= 56 whenever take > 0
    b.pos += take;
    b.cnt += 8 * take;
}
// the next k <= 32 bits without consuming them (missing bits read as 0; consuming them sets overrun)
S2I_HD inline uint32_t s2i_peek(S2InfBits &b, unsigned k)
{
    if (b.cnt < k) s2i_refill(b);
    return (uint32_t)(b.buf & ((1ull << k) - 1));
}
S2I_HD inline void s2i_drop(S2InfBits &b, unsigned k)
{
    if (k > b.cnt) { b.overrun = 1; b.buf = 0; b.cnt = 0; return; }
    b.buf >>= k; b.cnt -= k;
}
S2I_HD inline uint32_t s2i_bits(S2InfBits &b, unsigned k)
{
    const uint32_t v = s2i_peek(b, k);
    s2i_drop(b, k);
    return v;
}

// one Huffman code: first-level table + canonical arrays
template <int TBITS, int NSYM>
struct S2InfCode {
    uint16_t lut[1 << TBITS];      // (symbol << 4) | length for codes of <= TBITS bits, 0 = longer code or no code
    uint16_t count[S2I_MAX_BITS + 1];
    uint16_t symbol[NSYM];         // symbols ordered by (length, value)
};

// lengths[0..n) -> code.  Returns 0, or S2I_ERR_CODE_LENGTHS for an over-subscribed set.  An incomplete set is
// accepted (RFC 1951 allows a single distance code; zlib accepts incomplete sets only there - the unused patterns
// decode to S2I_ERR_BAD_SYMBOL, so damaged input is still caught when such a pattern occurs).
template <int TBITS, int NSYM>
S2I_HD inline int s2i_build(S2InfCode<TBITS, NSYM> &h, const uint8_t *lengths, int n)
{
    for (int i = 0; i <= S2I_MAX_BITS; ++i) h.count[i] = 0;
    for (int i = 0; i < n; ++i) h.count[lengths[i]]++;
    for (int i = 0; i < (1 << TBITS); ++i) h.lut[i] = 0;
    if (h.count[0] == n) return 0;                                   // no codes at all: every decode fails
    int left = 1;
    for (int len = 1; len <= S2I_MAX_BITS; ++len) {
        left <<= 1;
        left -= h.count[len];
        if (left < 0) return S2I_ERR_CODE_LENGTHS;
    }
    uint16_t offs[S2I_MAX_BITS + 2];
    offs[1] = 0;
    for (int len = 1; len <= S2I_MAX_BITS; ++len) offs[len + 1] = (uint16_t)(offs[len] + h.count[len]);
    for (int s = 0; s < n; ++s) if (lengths[s]) h.symbol[offs[lengths[s]]++] = (uint16_t)s;
    // first-level table: canonical code of every short symbol, bit-reversed (the stream carries codes MSB first in
    // LSB-first bit order), replicated over the bits that follow it
    unsigned code = 0, idx = 0;
    for (int len = 1; len <= TBITS; ++len) {
        for (unsigned k = 0; k < h.count[len]; ++k, ++idx, ++code) {
            unsigned rev = 0;
            for (int bit = 0; bit < len; ++bit) rev |= ((code >> bit) & 1u) << (len - 1 - bit);
            const uint16_t entry = (uint16_t)((h.symbol[idx] << 4) | len);
            for (unsigned fill = rev; fill < (1u << TBITS); fill += 1u << len) h.lut[fill] = entry;
        }
        code <<= 1;
    }
    return 0;
}

// next symbol of code h, or a negative error
template <int TBITS, int NSYM>
S2I_HD inline int s2i_decode(S2InfBits &b, const S2InfCode<TBITS, NSYM> &h)
{
    const uint32_t look = s2i_peek(b, S2I_MAX_BITS);
    const uint16_t e = h.lut[look & ((1u << TBITS) - 1)];
    if (e) { s2i_drop(b, e & 15u); return e >> 4; }
    // longer than the table (or no code): canonical decode, one bit at a time
    int code = 0, first = 0, index = 0;
    for (int len = 1; len <= S2I_MAX_BITS; ++len) {
        code |= (int)((look >> (len - 1)) & 1u);
        const int count = h.count[len];
        if (code - count < first) { s2i_drop(b, (unsigned)len); return h.symbol[index + (code - first)]; }
        index += count;
        first += count;
        first <<= 1;
        code <<= 1;
    }
    return S2I_ERR_BAD_SYMBOL;
}

// The match copy dst[i] = dst[i - dist], i in [0, len).  Written so that a thread never has to read back a byte it has
// just stored: on the device such a load goes to L2 (the stores write through L1) and costs some 300 cycles, and the
// naive byte loop pays that for EVERY byte - which is what the first kernel measurements showed (270 cycles per byte
// of FASTA, whose gzip -6 stream is mostly 6-8 byte matches).  dist >= 8: groups of 8 loads, then 8 stores (the source
// group lies wholly before the destination group, so the loads are independent and overlap).  dist < 8 (runs, e.g. a
// FASTQ quality line of one letter): the dist-byte pattern is loaded once into a register and replicated from there.
S2I_HD inline void s2i_copy_match(uint8_t *dst, uint64_t dist, uint32_t len)
{
    const uint8_t *src = dst - dist;
    if (dist >= 8) {
        uint32_t i = 0;
        for (; i + 8 <= len; i += 8) {
            uint8_t r[8];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int k = 0; k < 8; ++k) r[k] = src[i + k];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int k = 0; k < 8; ++k) dst[i + k] = r[k];
        }
        uint8_t r[8];
        const uint32_t tail = len - i;                              // < 8 <= dist: still no overlap inside the group
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (uint32_t k = 0; k < 8; ++k) if (k < tail) r[k] = src[i + k];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (uint32_t k = 0; k < 8; ++k) if (k < tail) dst[i + k] = r[k];
        return;
    }
    uint64_t pattern = 0;
    for (uint64_t k = 0; k < dist; ++k) pattern |= (uint64_t)src[k] << (8 * k);
    uint64_t word = pattern;
    uint32_t left = (uint32_t)dist;                                 // bytes of the pattern still in `word`
    for (uint32_t i = 0; i < len; ++i) {
        dst[i] = (uint8_t)word;
        word >>= 8;
        if (--left == 0) { word = pattern; left = (uint32_t)dist; }
    }
}

struct S2InfTables {
    S2InfCode<S2I_LIT_BITS, 288> lit;
    S2InfCode<S2I_DIST_BITS, 32> dist;
};

// RFC 1951 3.2.5 / 3.2.7 as arithmetic instead of tables: a table on a device thread's stack is local memory, which is
// interleaved over the lanes of the warp - with one decoding lane every entry sits in a cache line of its own
// length symbol s = 0..28 (codes 257..285): base length and number of extra bits
S2I_HD inline void s2i_len_code(int s, uint32_t *base, unsigned *extra)
{
    if (s < 8) { *base = 3u + (uint32_t)s; *extra = 0; return; }
    if (s == 28) { *base = 258; *extra = 0; return; }
    const unsigned e = (unsigned)(s - 4) >> 2;
    *base = 3u + ((4u + ((unsigned)s & 3u)) << e);
    *extra = e;
}
// distance symbol d = 0..29
S2I_HD inline void s2i_dist_code(int d, uint32_t *base, unsigned *extra)
{
    if (d < 4) { *base = 1u + (uint32_t)d; *extra = 0; return; }
    const unsigned e = ((unsigned)d >> 1) - 1u;
    *base = 1u + ((2u + ((unsigned)d & 1u)) << e);
    *extra = e;
}
// order in which the code lengths of the code-length code are sent: 16 17 18 0 8 7 9 6 10 5 11 4 12 3 13 2 14 1 15, 5 bits each
S2I_HD inline int s2i_clen_order(int i)
{
    const uint64_t lo = 16ull | 17ull << 5 | 18ull << 10 | 0ull << 15 | 8ull << 20 | 7ull << 25 | 9ull << 30 | 6ull << 35 | 10ull << 40 | 5ull << 45 |
                        11ull << 50 | 4ull << 55;
    const uint64_t hi = 12ull | 3ull << 5 | 13ull << 10 | 2ull << 15 | 14ull << 20 | 1ull << 25 | 15ull << 30;
    return (int)((i < 12 ? lo >> (5 * i) : hi >> (5 * (i - 12))) & 31u);
}

// One raw DEFLATE stream src[0..src_len) -> dst[0..dst_cap).  *out_len = bytes written (also on error: what was
// produced before it), *consumed = input bytes the stream occupied (rounded up to a whole byte).  `t` is scratch.
S2I_HD inline int s2_inflate_raw(const uint8_t *src, uint64_t src_len, uint8_t *dst, uint64_t dst_cap, uint64_t *out_len,
                                 uint64_t *consumed, S2InfTables &t)
{
    S2InfBits b = { src, src_len, 0, 0, 0, 0 };
    uint64_t out = 0;
    int rc = S2I_OK, last = 0;
    uint8_t lengths[320];
    while (!last && rc == S2I_OK) {
        last = (int)s2i_bits(b, 1);
        const unsigned type = s2i_bits(b, 2);
        if (b.overrun) { rc = S2I_ERR_TRUNCATED; break; }
        if (type == 0) {                                             // stored: byte aligned LEN, ~LEN, bytes
            s2i_drop(b, b.cnt & 7u);
            const uint32_t len = s2i_bits(b, 16), nlen = s2i_bits(b, 16);
            if (b.overrun) { rc = S2I_ERR_TRUNCATED; break; }
            if ((len ^ 0xFFFFu) != nlen) { rc = S2I_ERR_STORED_LEN; break; }
            uint32_t left = len;
            while (left && b.cnt >= 8) {                             // bytes already in the bit buffer
                if (out >= dst_cap) { rc = S2I_ERR_OUTPUT_FULL; break; }
                dst[out++] = (uint8_t)s2i_bits(b, 8);
                --left;
            }
            if (rc != S2I_OK) break;
            if (left) {
                if (b.n - b.pos < left) { rc = S2I_ERR_TRUNCATED; break; }
                if (dst_cap - out < left) { rc = S2I_ERR_OUTPUT_FULL; break; }
                for (uint32_t i = 0; i < left; ++i) dst[out + i] = b.p[b.pos + i];
                out += left; b.pos += left;
            }
            continue;
        }
        if (type == 3) { rc = S2I_ERR_BLOCK_TYPE; break; }
        if (type == 1) {                                             // fixed codes
            int s = 0;
            for (; s < 144; ++s) lengths[s] = 8;
            for (; s < 256; ++s) lengths[s] = 9;
            for (; s < 280; ++s) lengths[s] = 7;
            for (; s < 288; ++s) lengths[s] = 8;
            s2i_build(t.lit, lengths, 288);
            for (s = 0; s < 30; ++s) lengths[s] = 5;
            s2i_build(t.dist, lengths, 30);
        } else {                                                     // dynamic codes
            const int nlen = (int)s2i_bits(b, 5) + 257, ndist = (int)s2i_bits(b, 5) + 1, ncode = (int)s2i_bits(b, 4) + 4;
            if (b.overrun) { rc = S2I_ERR_TRUNCATED; break; }
            if (nlen > 286 || ndist > 30) { rc = S2I_ERR_CODE_LENGTHS; break; }
            int i = 0;
            for (; i < ncode; ++i) lengths[s2i_clen_order(i)] = (uint8_t)s2i_bits(b, 3);
            for (; i < 19; ++i) lengths[s2i_clen_order(i)] = 0;
            // the code-length code is decoded with the distance table's storage (8-bit first level, <= 7-bit codes)
            if (s2i_build(t.dist, lengths, 19)) { rc = S2I_ERR_CODE_LENGTHS; break; }
            i = 0;
            while (i < nlen + ndist) {
                const int sym = s2i_decode(b, t.dist);
                if (sym < 0) { rc = sym; break; }
                if (sym < 16) { lengths[i++] = (uint8_t)sym; continue; }
                uint8_t val = 0;
                int rep;
                if (sym == 16) {
                    if (i == 0) { rc = S2I_ERR_CODE_LENGTHS; break; }
                    val = lengths[i - 1];
                    rep = 3 + (int)s2i_bits(b, 2);
                } else if (sym == 17) rep = 3 + (int)s2i_bits(b, 3);
                else rep = 11 + (int)s2i_bits(b, 7);
                if (i + rep > nlen + ndist) { rc = S2I_ERR_CODE_LENGTHS; break; }
                while (rep--) lengths[i++] = val;
            }
            if (rc != S2I_OK) break;
            if (b.overrun) { rc = S2I_ERR_TRUNCATED; break; }
            if (lengths[256] == 0) { rc = S2I_ERR_CODE_LENGTHS; break; }          // no end-of-block code
            if (s2i_build(t.lit, lengths, nlen)) { rc = S2I_ERR_CODE_LENGTHS; break; }
            if (s2i_build(t.dist, lengths + nlen, ndist)) { rc = S2I_ERR_CODE_LENGTHS; break; }
        }
        // literals and matches until the end-of-block symbol
        for (;;) {
            int sym = s2i_decode(b, t.lit);
            if (sym < 0) { rc = sym; break; }
            if (sym < 256) {
                if (out >= dst_cap) { rc = S2I_ERR_OUTPUT_FULL; break; }
                dst[out++] = (uint8_t)sym;
                continue;
            }
            if (sym == 256) break;
            sym -= 257;
            if (sym >= 29) { rc = S2I_ERR_BAD_SYMBOL; break; }
            uint32_t base; unsigned extra;
            s2i_len_code(sym, &base, &extra);
            const uint32_t len = base + s2i_bits(b, extra);
            const int ds = s2i_decode(b, t.dist);
            if (ds < 0) { rc = ds; break; }
            if (ds >= 30) { rc = S2I_ERR_BAD_SYMBOL; break; }
            s2i_dist_code(ds, &base, &extra);
            const uint64_t dist = (uint64_t)base + s2i_bits(b, extra);
            if (b.overrun) { rc = S2I_ERR_TRUNCATED; break; }
            if (dist > out) { rc = S2I_ERR_DISTANCE; break; }
            if (dst_cap - out < len) { rc = S2I_ERR_OUTPUT_FULL; break; }
            s2i_copy_match(dst + out, dist, len);
            out += len;
        }
        if (rc == S2I_OK && b.overrun) rc = S2I_ERR_TRUNCATED;
    }
    if (out_len) *out_len = out;
    if (consumed) *consumed = b.pos - b.cnt / 8;                     // whole bytes still in the buffer were not used
    return rc;
}

// gzip header at p[0..n): length of the header (offset of the DEFLATE stream), or 0 if this is not a gzip member
S2I_HD inline uint64_t s2_gzip_header_len(const uint8_t *p, uint64_t n)
{
    if (n < 18 || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || (p[3] & 0xE0)) return 0;
    const unsigned flg = p[3];
    uint64_t o = 10;
    if (flg & 4) { if (o + 2 > n) return 0; o += 2 + (uint64_t)(p[o] | (p[o + 1] << 8)); }      // FEXTRA
    if (flg & 8) { while (o < n && p[o]) ++o; ++o; }                                             // FNAME
    if (flg & 16) { while (o < n && p[o]) ++o; ++o; }                                            // FCOMMENT
    if (flg & 2) o += 2;                                                                         // FHCRC
    return o + 8 <= n ? o : 0;
}

// A whole .gz file (one member or several, like gzread reads them) -> text.  *out_len = bytes of text.  Trailing
// zero bytes after the last member are ignored (gzip does the same); anything else there is an error.
// The CRC-32 of the trailer is not verified here (a separate, parallel pass); ISIZE is.
S2I_HD inline int s2_gunzip(const uint8_t *src, uint64_t src_len, uint8_t *dst, uint64_t dst_cap, uint64_t *out_len, S2InfTables &t)
{
    uint64_t in = 0, out = 0;
    int members = 0;
    while (in < src_len) {
        const uint64_t hl = s2_gzip_header_len(src + in, src_len - in);
        if (!hl) {
            bool zeros = members > 0;
            for (uint64_t i = in; zeros && i < src_len; ++i) zeros = src[i] == 0;
            if (zeros) break;
            if (out_len) *out_len = out;
            return S2I_ERR_GZIP_HEADER;
        }
        uint64_t got = 0, used = 0;
        const int rc = s2_inflate_raw(src + in + hl, src_len - in - hl, dst + out, dst_cap - out, &got, &used, t);
        out += got;
        if (rc != S2I_OK) { if (out_len) *out_len = out; return rc; }
        in += hl + used;
        if (src_len - in < 8) { if (out_len) *out_len = out; return S2I_ERR_TRUNCATED; }
        const uint32_t isize = (uint32_t)src[in + 4] | ((uint32_t)src[in + 5] << 8) | ((uint32_t)src[in + 6] << 16) | ((uint32_t)src[in + 7] << 24);
        if (isize != (uint32_t)got) { if (out_len) *out_len = out; return S2I_ERR_GZIP_SIZE; }
        in += 8;
        ++members;
    }
    if (out_len) *out_len = out;
    return members ? S2I_OK : S2I_ERR_GZIP_HEADER;
}
