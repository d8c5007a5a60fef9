// s2_gunzip.h - launch interface of the chunk-parallel gunzip kernels (s2_gunzip.cu; scheme in s2_gunzip.cuh)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// one .gz file of a batch - or one PIECE of a big file that is streamed through the stage in several batches; built by the host
struct GzFileDesc {
    uint64_t comp_off;     // where its bytes start in the batch's compressed buffer (multiple of 4; zero padded behind)
    uint64_t comp_len;     // bytes present (a piece: its own bytes plus a tail of the next piece's, for the last sub-chunk to run on)
    uint64_t first_bit;    // bit offset of the member's first DEFLATE block (8 x length of the gzip header); ~0: a later piece (block finder)
    uint64_t chain_bit;    // where the chain starts: first_bit, or (a later piece) where the previous piece ended, relative to this piece
    uint64_t text_off;     // where its text goes, relative to the text base handed to the translate launch
    uint64_t text_len;     // whole file: ISIZE of the trailer as the host read it (files below 4 GiB of text); piece: capacity of the text buffer
    uint64_t text_before;  // piece: bytes of text the earlier pieces produced
    uint32_t sub0, n_sub;  // its sub-chunks
    uint32_t piece;        // 0 whole file, 1 a piece that is not the last (the stream stays open), 2 the last piece
    uint32_t pad_;
};

struct GzSubResult;        // s2_gunzip.cuh

enum { GZ_CHAIN_BROKEN = -20, GZ_STREAM_OPEN = -21, GZ_SIZE_MISMATCH = -22, GZ_TRAILING_BYTES = -23, GZ_CRC_MISMATCH = -24 };

struct GzFileResult {
    uint64_t text_len;     // bytes the stream (piece) inflated to
    uint64_t end_bit;      // where decoding ended, relative to the file's (piece's) first byte
    int32_t status;        // 0 = whole file / last piece: a complete single member whose size (and, whole file, CRC-32) match the trailer;
                           //     piece: chain intact so far; else why not (host reader)
    uint32_t crc;          // CRC-32 of the trailer (whole file / last piece)
    uint32_t crc_ok;
    uint32_t crc_raw;      // piece: remainder of its text (no conditioning), for the host to combine
};

size_t gz_tables_bytes(void);
size_t gz_sub_result_bytes(void);
// The symbol area: cudaMalloc gz_sym_slots(n_sub, sub_cap) 16-bit slots, then gz_launch_sym_init once - it writes the window
// markers in front of every region and returns the pointer the other launches take as `sym` (region i = sym + i * sub_cap).
// sub_cap must exceed 32,769 + what a sub-chunk can produce: the last 32,768 slots of a region are its successor's markers.
size_t gz_sym_slots(size_t n_sub, uint32_t sub_cap);
uint16_t *gz_launch_sym_init(uint16_t *alloc, size_t n_sub, uint32_t sub_cap, cudaStream_t st);
void gz_launch_decode(const uint8_t *comp, const GzFileDesc *files, const uint32_t *sub_file, uint32_t n_sub, uint32_t sub_bytes, uint16_t *sym,
                      uint32_t sub_cap, GzSubResult *res, cudaStream_t st);
void gz_launch_chain(const uint8_t *comp, const GzFileDesc *files, uint32_t n_files, const uint16_t *sym, uint32_t sub_cap, const GzSubResult *res,
                     uint8_t *win, uint64_t *sub_off, GzFileResult *fres, cudaStream_t st);
// text of the files whose sub-chunks are [sub_lo, sub_hi) -> text + desc.text_off (files whose chain failed are skipped)
void gz_launch_translate(const GzFileDesc *files, const uint32_t *sub_file, uint32_t sub_lo, uint32_t sub_hi, const uint16_t *sym, uint32_t sub_cap,
                         const uint8_t *win, const uint64_t *sub_off, const GzFileResult *fres, uint8_t *text, cudaStream_t st);
// CRC-32 of the text of files [file0, file0 + n) (already translated at text + desc.text_off) against their trailers: sets
// crc_ok / status in fres, and act[i] = the size file file0 + i inflated to, or ~0 if anything about it is wrong.
// file_slice0: n + 1 ascending slice numbers (4 KB slices of each file's text), built by the host from the ISIZEs.
void gz_launch_crc(const GzFileDesc *files, uint32_t file0, uint32_t n_files, const uint32_t *file_slice0, uint32_t n_slices, const uint8_t *text,
                   GzFileResult *fres, uint32_t *crc_acc, unsigned *act, cudaStream_t st);

// CRC-32 of the members of a BGZF chunk against their trailers (isz / want_crc / toff: per member - inflated size, the trailer's CRC-32,
// where its text starts; xp128: 2,048 words of constants, filled once by gz_launch_xp128_init).  A mismatch sets act[m] to all ones (act != NULL) and / or
// *bad_flag to 1 (bad_flag != NULL).
void gz_launch_xp128_init(uint32_t *xp128, cudaStream_t st);
void gz_launch_member_crc(const uint8_t *text, const uint32_t *isz, const uint32_t *want_crc, const uint32_t *toff, uint32_t n_members, unsigned *act,
                          uint32_t *bad_flag, const uint32_t *xp128, cudaStream_t st);

// CRC-32 arithmetic for combining the pieces of a streamed file on the host (the same operations as the kernels':
// reflected polynomial 0xEDB88320, bit 31 = coefficient of x^0)
static inline uint32_t gz_crc_mulmod(uint32_t a, uint32_t b)
{
    uint32_t r = 0;
    for (int i = 0; i < 32; ++i) {
        if (a & 0x80000000u) r ^= b;
        a <<= 1;
        b = (b >> 1) ^ ((b & 1u) ? 0xEDB88320u : 0u);
    }
    return r;
}
static inline uint32_t gz_crc_xpow8(uint64_t n)
{
    uint32_t r = 0x80000000u, sq = 0x00800000u;
    while (n) { if (n & 1u) r = gz_crc_mulmod(r, sq); sq = gz_crc_mulmod(sq, sq); n >>= 1; }
    return r;
}
// remainder of A || B from the remainders of A and B
static inline uint32_t gz_crc_append(uint32_t raw_a, uint32_t raw_b, uint64_t len_b) { return gz_crc_mulmod(raw_a, gz_crc_xpow8(len_b)) ^ raw_b; }
// the gzip trailer's CRC-32 from the remainder of the whole text
static inline uint32_t gz_crc_finish(uint32_t raw, uint64_t len) { return raw ^ gz_crc_mulmod(0xFFFFFFFFu, gz_crc_xpow8(len)) ^ 0xFFFFFFFFu; }
