// s2_gunzip.h - launch interface of the chunk-parallel gunzip kernels (s2_gunzip.cu; scheme in s2_gunzip.cuh)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// one .gz file of a batch; built by the host
struct GzFileDesc {
    uint64_t comp_off;     // where its bytes start in the batch's compressed buffer (multiple of 4; zero padded behind)
    uint64_t comp_len;
    uint64_t first_bit;    // bit offset of the member's first DEFLATE block (8 x length of the gzip header)
    uint64_t text_off;     // where its text goes, relative to the text base handed to the translate launch
    uint64_t text_len;     // ISIZE of the trailer as the host read it (files below 4 GiB of text)
    uint32_t sub0, n_sub;  // its sub-chunks
};

struct GzSubResult;        // s2_gunzip.cuh

enum { GZ_CHAIN_BROKEN = -20, GZ_STREAM_OPEN = -21, GZ_SIZE_MISMATCH = -22, GZ_TRAILING_BYTES = -23, GZ_CRC_MISMATCH = -24 };

struct GzFileResult {
    uint64_t text_len;     // bytes the stream inflated to
    int32_t status;        // 0 = a complete single member whose size and CRC-32 match its trailer; else why not (host reader)
    uint32_t crc;          // CRC-32 of the trailer
    uint32_t crc_ok;
    uint32_t pad_;
};

size_t gz_tables_bytes(void);
size_t gz_sub_result_bytes(void);
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
