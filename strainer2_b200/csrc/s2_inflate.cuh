// s2_inflate.cuh - a DEFLATE (RFC 1951) / gzip (RFC 1952) decoder written once for host and device.
//
// STATUS: groundwork for SURVEY 8(f) rank 1, second half - NOT on the product path yet.  The decoder is unit-tested on
// the host against zlib (tests/test_host.py::test_inflate_*, through tests/sim/inflate_harness.cpp); the probe kernels
// that wrap it (tools/gunzip_probe.cu) decoded 2,048 .gz images correctly on a B200 at 8.5 - 14.3 GB/s of text
// (profiles/r1s_gunzip_kernel_v0_probe.txt); nothing in libstrainer2_b200.so calls it yet.
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
// Design of the decoder: a 64-bit bit buffer refilled bytewise; canonical Huffman codes decoded through a first-level
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

S2I_HD inline void s2i_refill(S2InfBits &b)
{
    while (b.cnt <= 56 && b.pos < b.n) { b.buf |= (uint64_t)b.p[b.pos++] << b.cnt; b.cnt += 8; }
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

struct S2InfTables {
    S2InfCode<S2I_LIT_BITS, 288> lit;
    S2InfCode<S2I_DIST_BITS, 32> dist;
};

// One raw DEFLATE stream src[0..src_len) -> dst[0..dst_cap).  *out_len = bytes written (also on error: what was
// produced before it), *consumed = input bytes the stream occupied (rounded up to a whole byte).  `t` is scratch.
S2I_HD inline int s2_inflate_raw(const uint8_t *src, uint64_t src_len, uint8_t *dst, uint64_t dst_cap, uint64_t *out_len,
                                 uint64_t *consumed, S2InfTables &t)
{
    const uint16_t len_base[29] = { 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258 };
    const uint8_t len_extra[29] = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0 };
    const uint16_t dist_base[30] = { 1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145,
                                     8193, 12289, 16385, 24577 };
    const uint8_t dist_extra[30] = { 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13 };
    const uint8_t clen_order[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };

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
            for (; i < ncode; ++i) lengths[clen_order[i]] = (uint8_t)s2i_bits(b, 3);
            for (; i < 19; ++i) lengths[clen_order[i]] = 0;
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
            const uint32_t len = len_base[sym] + s2i_bits(b, len_extra[sym]);
            const int ds = s2i_decode(b, t.dist);
            if (ds < 0) { rc = ds; break; }
            if (ds >= 30) { rc = S2I_ERR_BAD_SYMBOL; break; }
            const uint64_t dist = dist_base[ds] + s2i_bits(b, dist_extra[ds]);
            if (b.overrun) { rc = S2I_ERR_TRUNCATED; break; }
            if (dist > out) { rc = S2I_ERR_DISTANCE; break; }
            if (dst_cap - out < len) { rc = S2I_ERR_OUTPUT_FULL; break; }
            for (uint32_t i = 0; i < len; ++i) dst[out + i] = dst[out + i - dist];     // overlapping by design
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
