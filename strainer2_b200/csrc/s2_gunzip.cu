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
#define GZ_CHAIN_LENS 8192u

// Occupancy against registers (profiles/r2f_gz_decode_ncu.txt): a warp's instruction stream is one dependent chain - it issues
// about once in nine cycles - so the schedulers fill up only with eight or more warps each; 64 registers (8 CTAs per SM,
// which is also what the tables' shared memory allows) beat 93 registers without re-computed addresses at 5 CTAs.
template <int MIN_CTAS>
__global__ void __launch_bounds__(GZ_WARPS * 32, MIN_CTAS)
gz_decode_kernel(const uint8_t *__restrict__ comp, const GzFileDesc *__restrict__ files, const uint32_t *__restrict__ sub_file, uint32_t n_sub,
                 uint32_t sub_bytes, uint16_t *sym, uint32_t sub_cap, uint32_t sub_syms, GzSubResult *res, uint64_t search_limit_bits)
{
    // sub_cap: slots per sub-chunk region; its last 32,768 hold the NEXT region's marker prefix (gz_launch_sym_init), the one
    // before them is the decoder's guard slot: sub_syms = sub_cap - 32,769 symbols may be produced
    __shared__ GzTables tables[GZ_WARPS];
    __shared__ uint8_t kraft9[512];
    gz_kraft9_fill(kraft9, threadIdx.x, GZ_WARPS * 32);
    __syncthreads();
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t sub = blockIdx.x * GZ_WARPS + warp;
    if (sub >= n_sub) return;
    const GzFileDesc f = files[sub_file[sub]];
    const uint32_t j = sub - f.sub0;                                            // sub-chunk j of its file
    const uint64_t n_words = (f.comp_len + 3) / 4;
    gz_subchunk(reinterpret_cast<const uint32_t *>(comp + f.comp_off), n_words, j == 0 ? f.first_bit : ~0ull, (uint64_t)j * sub_bytes * 8ull,
                (uint64_t)(j + 1) * sub_bytes * 8ull, search_limit_bits, sym + (uint64_t)sub * sub_cap, sub_syms, tables[warp], kraft9, res + sub, (int)lane, 32);
}

// one CTA per file.  Phase 0, in parallel over the file's sub-chunks: the chain test (sub-chunk j must start where j-1
// ended), the first sub-chunk that ends the stream or breaks the chain, text offsets by a block scan.  Phase 1, in order:
// window j+1 from the last 32 KB of sub-chunk j's symbols and window j.  That loop is the serial part of a big file
// (3,000 steps per 96 MB piece), so a step is one global round trip and one barrier: both windows live in shared memory
// (ping-pong), the symbols of step j+1 are pulled towards L2 while step j runs, and window j goes out to global memory
// (for the translate pass) with 128-bit stores during step j+1.
__global__ void __launch_bounds__(GZ_CHAIN_THREADS)
gz_chain_kernel(const uint8_t *__restrict__ comp, const GzFileDesc *__restrict__ files, const uint16_t *__restrict__ sym, uint32_t sub_cap,
                const GzSubResult *__restrict__ res, uint8_t *win, uint64_t *__restrict__ sub_off, GzFileResult *__restrict__ out)
{
    extern __shared__ __align__(16) uint8_t s_win[];            // 2 x GZ_WINDOW, then GZ_CHAIN_LENS lengths
    uint32_t *s_lens = reinterpret_cast<uint32_t *>(s_win + 2 * GZ_WINDOW);      // symbols per sub-chunk (the loop below must not wait for global memory to learn them)
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
        if (j < GZ_CHAIN_LENS) s_lens[j] = (uint32_t)len;
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
    if (state >= 0) {                                            // (a broken file's windows are nobody's business)
        for (uint32_t i = tid; i < GZ_WINDOW / 16; i += GZ_CHAIN_THREADS) reinterpret_cast<uint4 *>(s_win)[i] = reinterpret_cast<const uint4 *>(w0)[i];
        __syncthreads();
        for (uint32_t j = 0; j < n_steps; ++j) {
            const uint8_t *prev = s_win + (j & 1u) * GZ_WINDOW;
            uint8_t *next = s_win + ((j + 1u) & 1u) * GZ_WINDOW;
            const uint32_t len = j < GZ_CHAIN_LENS ? s_lens[j] : r[j].n_out;
            const uint16_t *sy = sym + (uint64_t)(f.sub0 + j) * sub_cap;
            // the symbols that matter: the last min(len, 32 K); aligned 32-bit loads of two symbols, 17 per thread in flight
            const uint32_t used = len < GZ_WINDOW ? len : GZ_WINDOW;
            const uint32_t s_lo = len - used;                                  // first symbol that lands in the window ...
            const uint32_t k_lo = GZ_WINDOW - used;                            // ... at this place
            const uint32_t a = s_lo & ~1u;
            const uint32_t n_words = (len - a + 1u) / 2u;                      // <= 16385
            uint32_t wv[17];
#pragma unroll
            for (int i = 0; i < 17; ++i) {
                const uint32_t w = (uint32_t)i * GZ_CHAIN_THREADS + tid;
                wv[i] = w < n_words ? *reinterpret_cast<const uint32_t *>(sy + a + 2u * w) : 0u;
            }
            if (j + 1 < n_steps) {                                             // next step's symbols -> L2
                const uint32_t len2 = j + 1 < GZ_CHAIN_LENS ? s_lens[j + 1] : r[j + 1].n_out, used2 = len2 < GZ_WINDOW ? len2 : GZ_WINDOW;
                const uint8_t *p2 = reinterpret_cast<const uint8_t *>(sym + (uint64_t)(f.sub0 + j + 1) * sub_cap + (len2 - used2));
                if (tid * 128u < used2 * 2u) asm volatile("prefetch.global.L2 [%0];" :: "l"(p2 + tid * 128u));
            }
            // window j (complete since the last barrier) -> global memory, for the translate pass
            {
                uint4 *g = reinterpret_cast<uint4 *>(w0 + (uint64_t)j * GZ_WINDOW);
                const uint4 *sp = reinterpret_cast<const uint4 *>(prev);
                g[tid] = sp[tid]; g[tid + GZ_CHAIN_THREADS] = sp[tid + GZ_CHAIN_THREADS];
            }
            // where this sub-chunk produced fewer than 32 K symbols the window begins with the tail of the previous one
            for (uint32_t k = tid; k < k_lo; k += GZ_CHAIN_THREADS) next[k] = prev[k + used];
#pragma unroll
            for (int i = 0; i < 17; ++i) {
                const uint32_t w = (uint32_t)i * GZ_CHAIN_THREADS + tid;
                if (w < n_words) {
                    const uint32_t s0 = a + 2u * w;                            // symbol index of the low half
                    const uint32_t lo = wv[i] & 0xFFFFu, hi = wv[i] >> 16;
                    if (s0 >= s_lo) next[k_lo + (s0 - s_lo)] = lo < 256u ? (uint8_t)lo : prev[lo - 256u];
                    if (s0 + 1u < len) next[k_lo + (s0 + 1u - s_lo)] = hi < 256u ? (uint8_t)hi : prev[hi - 256u];
                }
            }
            __syncthreads();
        }
        {                                                                      // the last window (a later piece of the file starts from it)
            uint4 *g = reinterpret_cast<uint4 *>(w0 + (uint64_t)n_steps * GZ_WINDOW);
            const uint4 *sp = reinterpret_cast<const uint4 *>(s_win + (n_steps & 1u) * GZ_WINDOW);
            g[tid] = sp[tid]; g[tid + GZ_CHAIN_THREADS] = sp[tid + GZ_CHAIN_THREADS];
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
// one warp per 4 KB slice of a file's text, 128 bytes per lane: the remainder of each lane's run - four bytes per step with
// four tables (slicing-by-4: the dependent chain is 32 table steps instead of 128), head and tail bytes one at a time -
// moved to its place by a multiplication with x^(8 x bytes behind it), XORed into the file's accumulator (the CRC is linear)
__global__ void __launch_bounds__(256)
gz_crc_kernel(const GzFileDesc *__restrict__ files, uint32_t file0, uint32_t n_files, const uint32_t *__restrict__ file_slice0, uint32_t n_slices,
              const uint8_t *__restrict__ text, const GzFileResult *__restrict__ fres, uint32_t *crc_acc)
{
    __shared__ uint32_t table[4][256];
    {
        uint32_t c = threadIdx.x;                                                // (256 threads)
        for (int k = 0; k < 8; ++k) c = (c >> 1) ^ ((c & 1u) ? 0xEDB88320u : 0u);
        table[0][threadIdx.x] = c;
        __syncthreads();
        uint32_t v = c;
        for (int t = 1; t < 4; ++t) { v = (v >> 8) ^ table[0][v & 0xFFu]; table[t][threadIdx.x] = v; }
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
        uint64_t i = s0;
        for (; i < s1 && ((uintptr_t)(p + i) & 3u); ++i) c = table[0][(c ^ p[i]) & 0xFFu] ^ (c >> 8);
        for (; i + 4 <= s1; i += 4) {
            c ^= *reinterpret_cast<const uint32_t *>(p + i);
            c = table[3][c & 0xFFu] ^ table[2][(c >> 8) & 0xFFu] ^ table[1][(c >> 16) & 0xFFu] ^ table[0][c >> 24];
        }
        for (; i < s1; ++i) c = table[0][(c ^ p[i]) & 0xFFu] ^ (c >> 8);
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
// CRC-32 of the members of a BGZF chunk (the hardware engine inflates them and checks nothing: a member that is damaged
// but still inflates to its stated size would be counted silently, while the same file through zlib ends the run -
// ADVICE r1).  One CTA per member (<= 64 KB of text): 128-byte runs with slicing-by-4, each moved into place by
// x^(8 x 128 x runs behind it) from a table and x^(8 x length of the last run); a mismatch turns the member's
// "bytes produced" into all ones, which ing_check_chunk's act == isz test turns into a veto of the chunk.
// ------------------------------------------------------------------------------------------------
// Constants of the member CRC kernel, filled once per pipeline (2,048 words):
//   xp[j]          j = 0 .. 512   x^(8 * 128 * j)     a distance of j rows of 128 bytes
//   xp[513 + j]    j = 0 .. 128   x^(8 * j)
//   xp[768 + j]    j = 0 .. 64    x^(32 * j)          a distance of j words
//   xp[1024 + 256 t + b]          (byte b in place t of a register) * x^1024: multiplication by x^1024 as four lookups
__global__ void gz_xp128_kernel(uint32_t *xp)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j <= 512u) xp[j] = crc_xpow8(128ull * j);
    else if (j <= 513u + 128u) xp[j] = crc_xpow8(j - 513u);
    else if (j >= 768u && j <= 768u + 64u) xp[j] = crc_xpow8(4ull * (j - 768u));
    else if (j >= 1024u && j < 2048u) xp[j] = crc_mulmod(((j - 1024u) & 255u) << (8u * ((j - 1024u) >> 8)), crc_xpow8(128));
}

// One CTA per member.  The member's words (at its own alignment: two aligned loads and a funnel shift) are taken COLUMN-wise:
// lane k of a warp runs Horner's rule over words k, k + 32, k + 64 ... of the warp's rows - Y = Y * x^1024 + W, the
// multiplication being four table lookups - so every load is coalesced, nothing is staged and there is one modular
// multiplication per lane at the end (to move the column to its place) instead of two per 128 bytes.  (Earlier forms,
// profiles/r2n_ingest_crc_modes.txt, r2q: per-lane 128-byte runs straight from global memory, then through padded
// shared-memory tiles - 16 us per member whatever the access pattern: 0.5 warp instructions per byte, a serial chain per
// row, two multiplications per row.  14 % of the end-to-end rate.)
#define GZ_MC_WARPS 8
__global__ void __launch_bounds__(GZ_MC_WARPS * 32)
gz_member_crc_kernel(const uint8_t *__restrict__ text, const uint32_t *__restrict__ isz, const uint32_t *__restrict__ want_crc,
                     const uint32_t *__restrict__ toff, unsigned *act, uint32_t *bad_flag, const uint32_t *__restrict__ xp)
{
    __shared__ uint32_t V[4][256];
    __shared__ uint32_t s_acc;
    for (uint32_t i = threadIdx.x; i < 1024u; i += GZ_MC_WARPS * 32) (&V[0][0])[i] = xp[1024u + i];
    if (threadIdx.x == 0) s_acc = 0;
    __syncthreads();
    const uint32_t m = blockIdx.x;
    const uint32_t len = isz[m];
    if (len == 0 || len > 65536u) return;                               // (uniform; the host lists no empty members)
    const uint8_t *p = text + toff[m];
    const uint32_t sh = (uint32_t)((uintptr_t)p & 3u) * 8u;              // the member's misalignment, in bits
    const uint32_t *pw = reinterpret_cast<const uint32_t *>((uintptr_t)p & ~(uintptr_t)3);
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const uint32_t nw = len >> 2, n_rows = nw >> 5, rw = nw & 31u;       // whole words, whole rows of 32 words, words of the last partial row
    const uint32_t rpw = (n_rows + GZ_MC_WARPS - 1u) / GZ_MC_WARPS;
    const uint32_t r0 = wid * rpw < n_rows ? wid * rpw : n_rows, r1 = r0 + rpw < n_rows ? r0 + rpw : n_rows;
    uint32_t acc = 0;
    if (r0 < r1) {
        uint32_t y = 0;
        for (uint32_t r = r0; r < r1; r += 8u) {
            // eight rows' words first - sixteen independent loads in flight when the member is misaligned - then the eight
            // Horner steps (left to itself the compiler issues each load one step ahead of its use: 64 memory latencies per warp)
            uint32_t wa[8], wb[8];
#pragma unroll
            for (uint32_t u = 0; u < 8u; ++u) {
                const uint32_t j = (r + u) * 32u + lane;
                wa[u] = r + u < r1 ? pw[j] : 0u;
                wb[u] = sh && r + u < r1 ? pw[j + 1u] : 0u;
            }
#pragma unroll
            for (uint32_t u = 0; u < 8u; ++u)
                if (r + u < r1) {
                    const uint32_t w = sh ? __funnelshift_r(wa[u], wb[u], sh) : wa[u];
                    y = V[0][y & 0xFFu] ^ V[1][(y >> 8) & 0xFFu] ^ V[2][(y >> 16) & 0xFFu] ^ V[3][y >> 24] ^ w;
                }
        }
        // the column's last word is word 32 (r1 - 1) + lane of nw: (rw + 32 - lane) words from the end of the words, plus the rows of the warps behind
        acc = crc_mulmod(y, xp[768u + rw + 32u - lane]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc ^= __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    if (r0 < r1 && r1 < n_rows) acc = crc_mulmod(acc, xp[n_rows - r1]);                        // (every lane: no divergence, lane 0's is used)
    if (wid == GZ_MC_WARPS - 1 && rw) {                                                        // the last, partial row
        uint32_t c = 0;
        if (lane < rw) {
            const uint32_t j = n_rows * 32u + lane;
            const uint32_t w = sh ? __funnelshift_r(pw[j], pw[j + 1u], sh) : pw[j];
            c = crc_mulmod(w, xp[768u + rw - lane]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c ^= __shfl_xor_sync(0xFFFFFFFFu, c, o);
        acc ^= c;
    }
    if (lane == 0 && acc) atomicXor(&s_acc, acc);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t c = s_acc;                                                                    // the state behind the last whole word
        for (uint32_t i = nw * 4u; i < len; ++i) {                                             // up to three bytes
            c ^= p[i];
            for (int k = 0; k < 8; ++k) c = (c >> 1) ^ ((c & 1u) ? 0xEDB88320u : 0u);
        }
        const uint32_t xlen = crc_mulmod(xp[len >> 7], xp[513u + (len & 127u)]);               // x^(8 len)
        const uint32_t got = c ^ crc_mulmod(0xFFFFFFFFu, xlen) ^ 0xFFFFFFFFu;
        if (got != want_crc[m]) { if (act) act[m] = 0xFFFFFFFFu; if (bad_flag) *bad_flag = 1u; }
    }
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
    static const uint64_t limit = (8ull << 20) | (getenv("S2_GZ_FIND_ONLY") && atoi(getenv("S2_GZ_FIND_ONLY")) ? 1ull << 63 : 0ull);   // (bit 63: find, do not decode - profiling)
    if (ctas <= 5) gz_decode_kernel<5><<<grid, GZ_WARPS * 32, 0, st>>>(comp, files, sub_file, n_sub, sub_bytes, sym, sub_cap, sub_cap - GZ_WINDOW - 1u, res, limit);
    else if (ctas == 6) gz_decode_kernel<6><<<grid, GZ_WARPS * 32, 0, st>>>(comp, files, sub_file, n_sub, sub_bytes, sym, sub_cap, sub_cap - GZ_WINDOW - 1u, res, limit);
    else gz_decode_kernel<8><<<grid, GZ_WARPS * 32, 0, st>>>(comp, files, sub_file, n_sub, sub_bytes, sym, sub_cap, sub_cap - GZ_WINDOW - 1u, res, limit);
}

void gz_launch_chain(const uint8_t *comp, const GzFileDesc *files, uint32_t n_files, const uint16_t *sym, uint32_t sub_cap, const GzSubResult *res,
                     uint8_t *win, uint64_t *sub_off, GzFileResult *fres, cudaStream_t st)
{
    if (!n_files) return;
    cudaFuncSetAttribute(gz_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * GZ_WINDOW + GZ_CHAIN_LENS * 4);       // (per device: every launch)
    gz_chain_kernel<<<n_files, GZ_CHAIN_THREADS, 2 * GZ_WINDOW + GZ_CHAIN_LENS * 4, st>>>(comp, files, sym, sub_cap, res, win, sub_off, fres);
}

size_t gz_sub_result_bytes(void) { return sizeof(GzSubResult); }

// The symbol area of n_sub regions of sub_cap slots: 32,768 more in front, and in front of EVERY region (= the tail of its
// predecessor, which the decoder leaves alone) the markers of the unknown window, written once here.
size_t gz_sym_slots(size_t n_sub, uint32_t sub_cap) { return n_sub * sub_cap + GZ_WINDOW; }

__global__ void gz_marker_kernel(uint16_t *alloc, uint32_t sub_cap) { gz_marker_prefix(alloc + (uint64_t)blockIdx.x * sub_cap, threadIdx.x, blockDim.x); }

uint16_t *gz_launch_sym_init(uint16_t *alloc, size_t n_sub, uint32_t sub_cap, cudaStream_t st)
{
    if (n_sub) gz_marker_kernel<<<(unsigned)n_sub, 256, 0, st>>>(alloc, sub_cap);
    return alloc + GZ_WINDOW;
}

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

void gz_launch_xp128_init(uint32_t *xp128, cudaStream_t st) { gz_xp128_kernel<<<8, 256, 0, st>>>(xp128); }        // 2,048 words

void gz_launch_member_crc(const uint8_t *text, const uint32_t *isz, const uint32_t *want_crc, const uint32_t *toff, uint32_t n_members, unsigned *act,
                          uint32_t *bad_flag, const uint32_t *xp128, cudaStream_t st)
{
    if (n_members) gz_member_crc_kernel<<<n_members, GZ_MC_WARPS * 32, 0, st>>>(text, isz, want_crc, toff, act, bad_flag, xp128);
}
