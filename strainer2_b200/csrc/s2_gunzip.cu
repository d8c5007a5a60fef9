// s2_gunzip.cu - kernels around s2_gunzip.cuh: ordinary .gz files decoded on the GPU, many warps per stream.
//
// Replaces zlib's gzread under the reference's reader (/root/reference/src/genome_compare.c:194-203, src/kseq.h:68-101)
// for single-member .gz inputs - the format of every file the reference ships (test/genomes_to_scrub.txt,
// metagenomes_to_scrub.txt, target_metagenomes.txt) and of BASELINE config #3.  Three launches per batch of files:
//
//   gz_decode_kernel     one warp per sub-chunk of compressed bytes: find the first block start behind the cut, decode
//                        to the first block start behind the next cut into 16-bit symbols (byte, or marker into the
//                        unknown 32 KB window before the sub-chunk)
//   gz_chain_kernel      one CTA per file, sub-chunks in order: every sub-chunk must start where its predecessor ended
//                        (else the file is NOT handled: host reader), window i = last 32 KB of text up to sub-chunk i,
//                        text offsets, the member's end (trailer, ISIZE, nothing behind it)
//   gz_translate_kernel  symbols -> text through window i-1, all sub-chunks in parallel, straight into the ingest
//                        pipeline's text buffer
//   gz_crc_kernel        CRC-32 of every file's text against its trailer (slices in parallel, combined by
//                        multiplication with x^(8 n) in GF(2)[x] / P)
//
// This is byte/bit work bounded by instruction issue (a Huffman symbol is a dependent chain of a shared-memory lookup,
// shifts and a branch) - no tensor cores, little HBM traffic (5 bytes per byte of text).
#include "s2_gunzip.h"
#include "s2_gunzip.cuh"
#include <cstdlib>

#define GZ_WARPS 4                       /* decoding warps per CTA: 4 x 6.4 KB of tables */
#define GZ_CHAIN_THREADS 1024

// Occupancy against registers (profiles/r2f_gz_decode_ncu.txt): a warp's instruction stream is one dependent chain - it issues
// about once in nine cycles - so the schedulers fill up only with eight or more warps each; 64 registers (8 CTAs per SM,
// which is also what the tables' shared memory allows) beat 93 registers without re-computed addresses at 5 CTAs.
template <int MIN_CTAS>
__global__ void __launch_bounds__(GZ_WARPS * 32, MIN_CTAS)
gz_decode_kernel(const uint8_t *__restrict__ comp, const GzFileDesc *__restrict__ files, const uint32_t *__restrict__ sub_file, uint32_t n_sub,
                 uint32_t sub_bytes, uint16_t *sym, uint32_t sub_cap, GzSubResult *res, uint64_t search_limit_bits)
{
    __shared__ GzTables tables[GZ_WARPS];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t sub = blockIdx.x * GZ_WARPS + warp;
    if (sub >= n_sub) return;
    const GzFileDesc f = files[sub_file[sub]];
    const uint32_t j = sub - f.sub0;                                            // sub-chunk j of its file
    const uint64_t n_words = (f.comp_len + 3) / 4;
    gz_subchunk(reinterpret_cast<const uint32_t *>(comp + f.comp_off), n_words, j == 0 ? f.first_bit : ~0ull, (uint64_t)j * sub_bytes * 8ull,
                (uint64_t)(j + 1) * sub_bytes * 8ull, search_limit_bits, sym + (uint64_t)sub * sub_cap, sub_cap, tables[warp], res + sub, (int)lane, 32);
}

// The window behind sub-chunk j (the last 32 KB of text up to its end) is, byte for byte, either a literal of the sub-chunk's
// own last symbols - known now, for every sub-chunk at once - or a copy of one byte of the window before it: a marker, or
// the tail of that window where the sub-chunk produced fewer than 32 K symbols.  gz_tail_kernel writes the literal bytes
// of every window and lists the copies (place << 16 | place in the previous window); the chain kernel, which must go
// through a file's sub-chunks in order, then only moves the listed bytes - a few per cent of a window - instead of
// building 32 KB per step (a 96 MB piece of a big file is 3,000 steps: 43 ms when each step built a window through global
// memory, 8.8 ms with the windows in shared memory, profiles/r2g_pgunzip_probe.txt).
#define GZ_TAIL_THREADS 256
__global__ void __launch_bounds__(GZ_TAIL_THREADS)
gz_tail_kernel(const GzFileDesc *__restrict__ files, const uint32_t *__restrict__ sub_file, const uint16_t *__restrict__ sym, uint32_t sub_cap,
               const GzSubResult *__restrict__ res, uint8_t *win, uint32_t *__restrict__ ml, uint32_t *__restrict__ ml_count)
{
    const uint32_t sub = blockIdx.x, fi = sub_file[sub];
    const uint32_t n_out = res[sub].n_out;
    const uint16_t *sy = sym + (uint64_t)sub * sub_cap;
    uint8_t *wn = win + ((uint64_t)sub + fi + 1u) * GZ_WINDOW;                  // window j + 1 of its file = behind sub-chunk j
    uint32_t *list = ml + (uint64_t)sub * GZ_WINDOW;
    __shared__ uint32_t s_n;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t k0 = threadIdx.x * 4u; k0 < GZ_WINDOW; k0 += GZ_TAIL_THREADS * 4u) {
        uint32_t bytes = 0, copies[4]; bool is_copy[4];
#pragma unroll
        for (uint32_t u = 0; u < 4; ++u) {
            const uint32_t k = k0 + u;
            const int32_t idx = (int32_t)n_out - (int32_t)GZ_WINDOW + (int32_t)k;
            const uint32_t sy_k = idx >= 0 ? sy[idx] : 0u;
            is_copy[u] = idx < 0 || sy_k >= 256u;
            copies[u] = k << 16 | (idx < 0 ? k + n_out : sy_k - 256u);
            bytes |= (is_copy[u] ? 0u : sy_k) << (8u * u);
        }
        *reinterpret_cast<uint32_t *>(wn + k0) = bytes;
#pragma unroll
        for (uint32_t u = 0; u < 4; ++u) {                                      // warp-aggregated append
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, is_copy[u]);
            if (m) {
                uint32_t base = 0;
                if (lane == (uint32_t)__ffs(m) - 1u) base = atomicAdd(&s_n, (uint32_t)__popc(m));
                base = __shfl_sync(0xFFFFFFFFu, base, __ffs(m) - 1);
                if (is_copy[u]) list[base + __popc(m & ((1u << lane) - 1u))] = copies[u];
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) ml_count[sub] = s_n;
}

// one CTA per file.  Phase 0, in parallel over the file's sub-chunks: the chain test (sub-chunk j must start where j-1
// ended), the first sub-chunk that ends the stream or breaks the chain, text offsets by a block scan.  Phase 1, in order:
// the copies of window j+1 (gz_tail_kernel's list) from window j; the next step's list is loaded while this one's bytes move.
#define GZ_CHAIN_PRE 4
__global__ void __launch_bounds__(GZ_CHAIN_THREADS)
gz_chain_kernel(const uint8_t *__restrict__ comp, const GzFileDesc *__restrict__ files, const GzSubResult *__restrict__ res, uint8_t *win,
                const uint32_t *__restrict__ ml, const uint32_t *__restrict__ ml_count, uint64_t *__restrict__ sub_off, GzFileResult *__restrict__ out)
{
    const GzFileDesc f = files[blockIdx.x];
    __shared__ uint32_t s_first;                                 // first sub-chunk that is not a plain link of the chain
    __shared__ uint64_t s_warp[GZ_CHAIN_THREADS / 32];
    __shared__ uint64_t s_carry;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    const GzSubResult *r = res + f.sub0;
    if (tid == 0) { s_first = f.n_sub; s_carry = 0; }
    __syncthreads();
    for (uint32_t j = tid; j < f.n_sub; j += GZ_CHAIN_THREADS) {
        const uint64_t expect = j ? r[j - 1].end_bit : f.chain_bit;
        if (r[j].start_bit != expect || r[j].status != GZ_OK) atomicMin(&s_first, j);
    }
    __syncthreads();
    const uint32_t first = s_first;
    int state = 0;                                               // 0 the stream goes on, 1 it ended, < 0 error
    uint32_t n_steps = first;                                    // sub-chunks that contribute text
    if (first < f.n_sub) {
        const uint64_t expect = first ? r[first - 1].end_bit : f.chain_bit;
        const int st = r[first].status;
        if (r[first].start_bit != expect) state = GZ_CHAIN_BROKEN;
        else if (st < 0) state = st;
        else { state = 1; n_steps = first + 1; }                 // GZ_FINAL: the stream's last block ended here
    }
    // text offsets: exclusive scan of the lengths (0 behind the end of the stream / the break)
    for (uint32_t j0 = 0; j0 < f.n_sub; j0 += GZ_CHAIN_THREADS) {
        const uint32_t j = j0 + tid;
        const uint64_t len = j < n_steps ? r[j].n_out : 0u;
        uint64_t inc = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint64_t n = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= (uint32_t)o) inc += n; }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        uint64_t base = s_carry;
        for (uint32_t w = 0; w < wid; ++w) base += s_warp[w];
        if (j < f.n_sub) sub_off[f.sub0 + j] = base + inc - len;
        __syncthreads();
        if (tid == GZ_CHAIN_THREADS - 1) s_carry = base + inc;
        __syncthreads();
    }
    const uint64_t total = s_carry;
    // windows, in order
    uint8_t *w0 = win + ((uint64_t)f.sub0 + blockIdx.x) * GZ_WINDOW;          // n_sub + 1 windows of this file
    if (state >= 0 && n_steps) {                                 // (a broken file's windows are nobody's business)
        const uint32_t *cnt = ml_count + f.sub0;
        const uint32_t *lists = ml + (uint64_t)f.sub0 * GZ_WINDOW;
        uint32_t n_cur = cnt[0], e_cur[GZ_CHAIN_PRE];
#pragma unroll
        for (int u = 0; u < GZ_CHAIN_PRE; ++u) e_cur[u] = lists[u * GZ_CHAIN_THREADS + tid];          // (a list has room for 32 K entries: always readable)
        for (uint32_t j = 0; j < n_steps; ++j) {
            const uint8_t *wp = w0 + (uint64_t)j * GZ_WINDOW;
            uint8_t *wn = w0 + (uint64_t)(j + 1u) * GZ_WINDOW;
            const uint32_t *list = lists + (uint64_t)j * GZ_WINDOW;
            uint32_t n_nxt = 0, e_nxt[GZ_CHAIN_PRE];
            if (j + 1u < n_steps) {
                n_nxt = cnt[j + 1u];
#pragma unroll
                for (int u = 0; u < GZ_CHAIN_PRE; ++u) e_nxt[u] = list[GZ_WINDOW + u * GZ_CHAIN_THREADS + tid];
            }
#pragma unroll
            for (int u = 0; u < GZ_CHAIN_PRE; ++u)
                if (u * GZ_CHAIN_THREADS + tid < n_cur) wn[e_cur[u] >> 16] = __ldcg(wp + (e_cur[u] & 0xFFFFu));
            for (uint32_t i = GZ_CHAIN_PRE * GZ_CHAIN_THREADS + tid; i < n_cur; i += GZ_CHAIN_THREADS) {
                const uint32_t e = list[i];
                wn[e >> 16] = __ldcg(wp + (e & 0xFFFFu));
            }
            __syncthreads();                                     // window j + 1 is complete (and visible to the block) before step j + 1 reads it
            n_cur = n_nxt;
#pragma unroll
            for (int u = 0; u < GZ_CHAIN_PRE; ++u) e_cur[u] = e_nxt[u];
        }
    }
    if (threadIdx.x == 0) {
        const uint64_t s_total = total;
        const uint64_t s_cur = n_steps ? r[n_steps - 1].end_bit : f.chain_bit;
        const int s_state = state;
        GzFileResult r;
        r.text_len = s_total; r.end_bit = s_cur; r.crc = 0; r.crc_ok = 0; r.crc_raw = 0;
        r.status = s_state == 1 ? 0 : (s_state == 0 ? (f.piece == 1 ? 0 : GZ_STREAM_OPEN) : s_state);
        if (s_state == 1 && f.piece == 1) r.status = GZ_TRAILING_BYTES;        // the stream ended inside a piece that is not the last: more members, or junk
        if (f.piece && s_total > f.text_len) r.status = GZ_SIZE_MISMATCH;       // more text than the piece buffer holds (text_len = its capacity): nothing is translated
        if (s_state == 1 && f.piece != 1) {                                                     // trailer: CRC-32, ISIZE; nothing but zeros behind it
            const uint64_t tr = (s_cur + 7) / 8;
            const uint8_t *p = comp + f.comp_off;
            if (tr + 8 > f.comp_len) r.status = GZ_ERR_INPUT;
            else {
                const uint32_t isize = (uint32_t)p[tr + 4] | (uint32_t)p[tr + 5] << 8 | (uint32_t)p[tr + 6] << 16 | (uint32_t)p[tr + 7] << 24;
                r.crc = (uint32_t)p[tr] | (uint32_t)p[tr + 1] << 8 | (uint32_t)p[tr + 2] << 16 | (uint32_t)p[tr + 3] << 24;
                if (isize != (uint32_t)(f.text_before + s_total) || (f.piece == 0 && s_total != f.text_len)) r.status = GZ_SIZE_MISMATCH;
                for (uint64_t i = tr + 8; i < f.comp_len && r.status == 0; ++i) if (p[i]) r.status = GZ_TRAILING_BYTES;     // another member, or junk: host reader
            }
        }
        out[blockIdx.x] = r;
    }
}

// symbols -> text.  One CTA per sub-chunk; files whose chain failed are skipped (their text area stays unwritten and the
// pipeline's act != isz check vetoes the chunk).
__global__ void __launch_bounds__(256)
gz_translate_kernel(const GzFileDesc *__restrict__ files, const uint32_t *__restrict__ sub_file, uint32_t sub_lo,
                    const uint16_t *__restrict__ sym, uint32_t sub_cap, const uint8_t *__restrict__ win,
                    const uint64_t *__restrict__ sub_off, const GzFileResult *__restrict__ fres, uint8_t *text)
{
    const uint32_t sub = sub_lo + blockIdx.x;
    const uint32_t fi = sub_file[sub];
    const GzFileDesc f = files[fi];
    const GzFileResult fr = fres[fi];
    if (fr.status != 0) return;
    const uint32_t j = sub - f.sub0;
    const uint64_t off = sub_off[sub];
    const uint64_t next = j + 1 < f.n_sub ? sub_off[sub + 1] : fr.text_len;
    const uint32_t n = (uint32_t)(next - off);                                  // 0 behind the end of the stream
    gz_translate(win + ((uint64_t)sub + fi) * GZ_WINDOW, sym + (uint64_t)sub * sub_cap, n, text + f.text_off + off, threadIdx.x, 256);
}

// ------------------------------------------------------------------------------------------------
// CRC-32 (the gzip trailer's, reflected polynomial 0xEDB88320) of every file's text, in parallel
// ------------------------------------------------------------------------------------------------
// In the reflected representation bit 31 of a register is the coefficient of x^0.  a * b mod P:
__device__ __forceinline__ uint32_t crc_mulmod(uint32_t a, uint32_t b)
{
    uint32_t r = 0;
#pragma unroll 4
    for (int i = 0; i < 32; ++i) {
        if (a & 0x80000000u) r ^= b;                    // coefficient of x^i in a
        a <<= 1;
        b = (b >> 1) ^ ((b & 1u) ? 0xEDB88320u : 0u);   // b *= x
    }
    return r;
}
// x^(8 n) mod P by square and multiply
__device__ __forceinline__ uint32_t crc_xpow8(uint64_t n)
{
    uint32_t r = 0x80000000u;                           // 1
    uint32_t sq = 0x00800000u;                          // x^8
    while (n) {
        if (n & 1u) r = crc_mulmod(r, sq);
        sq = crc_mulmod(sq, sq);
        n >>= 1;
    }
    return r;
}

#define GZ_CRC_SLICE 4096u
// one warp per 4 KB slice of a file's text, 128 bytes per lane: the remainder of each piece (byte-wise table), moved to its
// place by a multiplication with x^(8 x bytes behind it), XORed into the file's accumulator (the CRC is linear)
__global__ void __launch_bounds__(256)
gz_crc_kernel(const GzFileDesc *__restrict__ files, uint32_t file0, uint32_t n_files, const uint32_t *__restrict__ file_slice0, uint32_t n_slices,
              const uint8_t *__restrict__ text, const GzFileResult *__restrict__ fres, uint32_t *crc_acc)
{
    __shared__ uint32_t table[256];
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
        uint32_t c = i;
        for (int k = 0; k < 8; ++k) c = (c >> 1) ^ ((c & 1u) ? 0xEDB88320u : 0u);
        table[i] = c;
    }
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t slice = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (slice >= n_slices) return;                                             // (whole warps)
    uint32_t lo = 0, hi = n_files;                                             // file of this slice: last f with file_slice0[f] <= slice
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (file_slice0[mid] <= slice) lo = mid; else hi = mid; }
    const GzFileDesc f = files[file0 + lo];
    const GzFileResult fr = fres[file0 + lo];
    if (fr.status != 0) return;                                                // (the same for every lane of the warp)
    const uint64_t len = fr.text_len;
    const uint64_t s0 = (uint64_t)(slice - file_slice0[lo]) * GZ_CRC_SLICE + (uint64_t)lane * (GZ_CRC_SLICE / 32);
    uint32_t c = 0;
    if (s0 < len) {
        const uint64_t s1 = s0 + GZ_CRC_SLICE / 32 < len ? s0 + GZ_CRC_SLICE / 32 : len;
        const uint8_t *p = text + f.text_off;
        for (uint64_t i = s0; i < s1; ++i) c = table[(c ^ p[i]) & 0xFFu] ^ (c >> 8);
        c = crc_mulmod(c, crc_xpow8(len - s1));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c ^= __shfl_xor_sync(0xFFFFFFFFu, c, o);
    if (lane == 0 && c) atomicXor(&crc_acc[file0 + lo], c);
}

// the conditioning (initial value ~0, final complement) as one more term, the comparison with the trailer, and what the
// ingest pipeline's act == isz check will see for the file
__global__ void gz_crc_finish_kernel(const GzFileDesc *__restrict__ files, uint32_t file0, uint32_t n_files, GzFileResult *fres, uint32_t *crc_acc, unsigned *act)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_files) return;
    GzFileResult r = fres[file0 + i];
    if (files[file0 + i].piece) {                        // a piece of a streamed file: the host combines the pieces' remainders
        r.crc_raw = crc_acc[file0 + i];
        fres[file0 + i] = r;
        crc_acc[file0 + i] = 0;
        if (act) act[i] = r.status == 0 ? (unsigned)r.text_len : 0xFFFFFFFFu;
        return;
    }
    if (r.status == 0) {
        // crc(M) = rem(M) ^ rem(0xFFFFFFFF . x^(8 len)) ^ 0xFFFFFFFF
        const uint32_t got = crc_acc[file0 + i] ^ crc_mulmod(0xFFFFFFFFu, crc_xpow8(r.text_len)) ^ 0xFFFFFFFFu;
        r.crc_ok = got == r.crc ? 1u : 0u;
        if (!r.crc_ok) r.status = GZ_CRC_MISMATCH;
        fres[file0 + i] = r;
    }
    crc_acc[file0 + i] = 0;
    if (act) act[i] = r.status == 0 ? (unsigned)r.text_len : 0xFFFFFFFFu;
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
size_t gz_tables_bytes(void) { return sizeof(GzTables); }

void gz_launch_decode(const uint8_t *comp, const GzFileDesc *files, const uint32_t *sub_file, uint32_t n_sub, uint32_t sub_bytes, uint16_t *sym,
                      uint32_t sub_cap, GzSubResult *res, cudaStream_t st)
{
    if (!n_sub) return;
    static const int ctas = getenv("S2_GZ_DECODE_CTAS") ? atoi(getenv("S2_GZ_DECODE_CTAS")) : 8;      // (experiments; 8 is the measured best)
    const dim3 grid((n_sub + GZ_WARPS - 1) / GZ_WARPS);
    if (ctas <= 5) gz_decode_kernel<5><<<grid, GZ_WARPS * 32, 0, st>>>(comp, files, sub_file, n_sub, sub_bytes, sym, sub_cap, res, 8ull << 20);
    else if (ctas == 6) gz_decode_kernel<6><<<grid, GZ_WARPS * 32, 0, st>>>(comp, files, sub_file, n_sub, sub_bytes, sym, sub_cap, res, 8ull << 20);
    else gz_decode_kernel<8><<<grid, GZ_WARPS * 32, 0, st>>>(comp, files, sub_file, n_sub, sub_bytes, sym, sub_cap, res, 8ull << 20);
}

void gz_launch_chain(const uint8_t *comp, const GzFileDesc *files, uint32_t n_files, const uint32_t *sub_file, uint32_t n_sub, const uint16_t *sym,
                     uint32_t sub_cap, const GzSubResult *res, uint8_t *win, uint32_t *ml, uint32_t *ml_count, uint64_t *sub_off, GzFileResult *fres,
                     cudaStream_t st)
{
    if (!n_files) return;
    if (n_sub) gz_tail_kernel<<<n_sub, GZ_TAIL_THREADS, 0, st>>>(files, sub_file, sym, sub_cap, res, win, ml, ml_count);
    gz_chain_kernel<<<n_files, GZ_CHAIN_THREADS, 0, st>>>(comp, files, res, win, ml, ml_count, sub_off, fres);
}

size_t gz_sub_result_bytes(void) { return sizeof(GzSubResult); }

void gz_launch_translate(const GzFileDesc *files, const uint32_t *sub_file, uint32_t sub_lo, uint32_t sub_hi, const uint16_t *sym, uint32_t sub_cap,
                         const uint8_t *win, const uint64_t *sub_off, const GzFileResult *fres, uint8_t *text, cudaStream_t st)
{
    if (sub_hi <= sub_lo) return;
    gz_translate_kernel<<<sub_hi - sub_lo, 256, 0, st>>>(files, sub_file, sub_lo, sym, sub_cap, win, sub_off, fres, text);
}

void gz_launch_crc(const GzFileDesc *files, uint32_t file0, uint32_t n_files, const uint32_t *file_slice0, uint32_t n_slices, const uint8_t *text,
                   GzFileResult *fres, uint32_t *crc_acc, unsigned *act, cudaStream_t st)
{
    if (!n_files) return;
    if (n_slices) gz_crc_kernel<<<(n_slices + 7) / 8, 256, 0, st>>>(files, file0, n_files, file_slice0, n_slices, text, fres, crc_acc);
    gz_crc_finish_kernel<<<(n_files + 127) / 128, 128, 0, st>>>(files, file0, n_files, fres, crc_acc, act);
}
