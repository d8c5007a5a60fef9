// s2_ingest.cu - GPU-side ingest of FASTQ files: hardware DEFLATE + record splitting on the device.
//
// SURVEY 8(f) rank 1.  The reference inflates and parses every input on one CPU thread (zlib gzread +
// the vendored line parser, /root/reference/src/genome_compare.c:194-203, src/kseq.h:171-211); zlib gives
// about 0.34 GB/s of text per core, three orders of magnitude below the scan kernel.  Blackwell has a
// hardware decompression engine, reachable through the CUDA driver's batch-decompress entry point: it inflates
// independent raw-DEFLATE streams of up to 4 MiB each (measured here: 240-320 GB/s of text for batches of
// 64 KB blocks).  A plain .gz file is ONE long stream and cannot be split, but BGZF (bgzip, the block-
// gzip flavour htslib writes; still a valid multi-member gzip file for the reference's zlib) is a sequence
// of independent <= 64 KB members whose sizes are in their headers.  For BGZF FASTQ - and for uncompressed
// FASTQ - this file does on the GPU what the reader threads do for everything else:
//
//   host   : read() the compressed bytes into pinned memory, walk the BGZF headers (no inflate)
//   engine : inflate all blocks of a chunk into one contiguous text buffer
//   kernels: index the newlines, check that the text is strict 4-line FASTQ, copy the sequence lines into
//            the flat batch format (sequence bytes + '\n'), carry the partial last record to the next chunk
//   scan   : the normal count kernel, with the batch length read from device memory
//
// Parity: the parser the reference vendors accepts many irregular layouts (multi-line FASTQ, CR LF, '>'
// records mixed in, truncated last record ...).  The kernels do not emulate those; they PROVE that a file is
// regular (every record is '@' line, sequence line, '+' line, quality line of the same length, no CR, no
// sequence starting with '>', '+' or '@', nothing left over at the end) in a first pass that scans nothing,
// and only then run the pass that counts.  Anything else is handed back to the host reader untouched
// (return value 1), so the counters are identical either way.
#include "s2_private.h"
#include "s2_kmer.cuh"

#include <cuda.h>
#include <fcntl.h>
#include <zlib.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>

#define ING_THREADS 256
#define ING_MAXCARRY (1u << 20)          /* longest partial record carried between chunks */
#define ING_COMP_CHUNK (8u << 20)        /* compressed bytes read per chunk */
#define ING_TEXT_CAP (96u << 20)         /* inflated text per chunk (BGZF: <= 64 KB per block) */
#define ING_MAX_LINES (ING_TEXT_CAP / 8) /* shorter average lines than 8 bytes: not a read file */

struct IngState {                 // lives in device memory, one per ingest object
    unsigned long long t0, t1;    // current text range inside the text buffer
    unsigned long long flat_len;  // bytes of the flat batch produced from this chunk
    unsigned long long carry_len; // bytes after the last complete record
    unsigned long long bases, lookups, records;   // totals over the file (pass 2)
    unsigned int n_lines, n_rec;
    unsigned int irregular;       // sticky: the file is not strict FASTQ
    unsigned int inf_overflow;    // a chunk produced more informative windows than the chunk list holds
    unsigned int last_chunk, first_chunk;
    unsigned int tail_len;        // FASTA: last bytes of the previous chunk's flat stream, re-scanned in front of this one
    unsigned char tail[32];
};

// ------------------------------------------------------------------------------------------------
// newline index
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned nl_mask16(const uint8_t *text, unsigned long long base, unsigned long long t0,
                                              unsigned long long t1, unsigned *cr)
{
    // bit i set <=> text[base+i] == '\n' and t0 <= base+i < t1   (base is 16-byte aligned)
    unsigned m = 0;
    *cr = 0;
    if (base + 16 <= t0 || base >= t1) return 0;
    const uint4 v = *reinterpret_cast<const uint4 *>(text + base);
    const unsigned w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const unsigned c = (w[k] >> (8 * b)) & 0xFFu;
            const unsigned long long pos = base + 4 * k + b;
            const bool in = pos >= t0 && pos < t1;
            if (in && c == '\n') m |= 1u << (4 * k + b);
            if (in && c == '\r') *cr = 1;
        }
    return m;
}

__device__ __forceinline__ unsigned ing_block_scan(unsigned v, unsigned &total)
{
    __shared__ unsigned warp_sums[ING_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned n = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += n; }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    unsigned base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < ING_THREADS / 32; ++i) { const unsigned s = warp_sums[i]; if (i < wid) base += s; tot += s; }
    __syncthreads();
    total = tot;
    return base + inc - v;
}

// one CTA covers 4096 bytes of the text buffer (absolute, aligned); CTAs outside [t0, t1) do nothing
__global__ void __launch_bounds__(ING_THREADS) ing_nl_count(const uint8_t *__restrict__ text, IngState *st, unsigned *__restrict__ block_nl)
{
    const unsigned long long t0 = st->t0, t1 = st->t1;
    const unsigned long long base = (unsigned long long)blockIdx.x * 4096 + threadIdx.x * 16;
    unsigned cr;
    const unsigned m = nl_mask16(text, base, t0, t1, &cr);
    if (cr) atomicOr(&st->irregular, 1u);
    unsigned total;
    ing_block_scan(__popc(m), total);
    if (threadIdx.x == 0) block_nl[blockIdx.x] = total;
}

// single CTA: exclusive scan of n counters in place, total to *total_out (n may come from device memory)
__global__ void __launch_bounds__(ING_THREADS) ing_scan_u32(unsigned *__restrict__ v, unsigned n_host, const unsigned *n_dev, unsigned *total_out)
{
    const unsigned n = n_dev ? *n_dev : n_host;
    __shared__ unsigned carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (unsigned b0 = 0; b0 < n; b0 += ING_THREADS * 4) {
        unsigned x[4], sum = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) { const unsigned i = b0 + threadIdx.x * 4 + k; x[k] = i < n ? v[i] : 0; sum += x[k]; }
        unsigned total;
        unsigned ex = ing_block_scan(sum, total) + carry_s;
#pragma unroll
        for (int k = 0; k < 4; ++k) { const unsigned i = b0 + threadIdx.x * 4 + k; if (i < n) v[i] = ex; ex += x[k]; }
        __syncthreads();
        if (threadIdx.x == 0) carry_s += total;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry_s;
}

__global__ void __launch_bounds__(ING_THREADS) ing_nl_scatter(const uint8_t *__restrict__ text, IngState *st, const unsigned *__restrict__ block_off,
                                                               unsigned *__restrict__ line_end)
{
    const unsigned long long t0 = st->t0, t1 = st->t1;
    const unsigned long long base = (unsigned long long)blockIdx.x * 4096 + threadIdx.x * 16;
    unsigned cr;
    unsigned m = nl_mask16(text, base, t0, t1, &cr);
    unsigned total;
    unsigned r = block_off[blockIdx.x] + ing_block_scan(__popc(m), total);
    while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        if (r < ING_MAX_LINES) line_end[r] = (unsigned)(base + b);     // the text buffer is < 4 GiB
        ++r;
    }
}

// ------------------------------------------------------------------------------------------------
// strict FASTQ: validate, measure, copy
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ING_THREADS) ing_fastq_prepare(IngState *st)
{
    // on the last chunk a final line without '\n' is completed (the buffer has room past t1)
    if (st->n_lines > ING_MAX_LINES) { st->irregular = 1; st->n_lines = 0; }
    st->n_rec = st->n_lines / 4;
}

__global__ void __launch_bounds__(ING_THREADS) ing_fastq_len(const uint8_t *__restrict__ text, IngState *st, const unsigned *__restrict__ line_end,
                                                              unsigned *__restrict__ out_len, int count_stats)
{
    const unsigned n_rec = st->n_rec;
    unsigned long long bases = 0, lookups = 0;
    for (unsigned r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rec; r += gridDim.x * blockDim.x) {
        const unsigned long long h0 = r ? (unsigned long long)line_end[4 * r - 1] + 1 : st->t0;   // '@' line
        const unsigned long long s0 = (unsigned long long)line_end[4 * r] + 1;                     // sequence line
        const unsigned long long p0 = (unsigned long long)line_end[4 * r + 1] + 1;                 // '+' line
        const unsigned long long q0 = (unsigned long long)line_end[4 * r + 2] + 1;                 // quality line
        const unsigned long long e0 = (unsigned long long)line_end[4 * r + 3];
        const unsigned long long len = p0 - 1 - s0, qlen = e0 - q0;
        bool ok = text[h0] == '@' && text[p0] == '+' && len == qlen;
        if (len) { const uint8_t c = text[s0]; ok = ok && c != '>' && c != '+' && c != '@'; }
        if (!ok) atomicOr(&st->irregular, 1u);
        out_len[r] = len >= S2_K ? (unsigned)len + 1 : 0;       // records without a window are not copied (genome_compare.c:204)
        bases += len;
        if (len >= S2_K) lookups += len - (S2_K - 1);
    }
    if (count_stats) {
        if (bases) atomicAdd(&st->bases, bases);
        if (lookups) atomicAdd(&st->lookups, lookups);
    }
}

__global__ void __launch_bounds__(ING_THREADS) ing_fastq_copy(const uint8_t *__restrict__ text, const IngState *st, const unsigned *__restrict__ line_end,
                                                               const unsigned *__restrict__ out_off, uint8_t *__restrict__ flat)
{
    const unsigned n_rec = st->n_rec;
    const int lane = threadIdx.x & 31;
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (unsigned r = warp; r < n_rec; r += n_warps) {
        const unsigned long long s0 = (unsigned long long)line_end[4 * r] + 1;
        const unsigned len = line_end[4 * r + 1] - (unsigned)s0;
        if (len < S2_K) continue;
        uint8_t *dst = flat + out_off[r];
        for (unsigned i = lane; i <= len; i += 32) dst[i] = text[s0 + i];       // includes the line's own '\n' = the separator
    }
}

// end of chunk: totals, carry of the partial last record to the front of the buffer for the next chunk
__global__ void __launch_bounds__(ING_THREADS) ing_fastq_finish(uint8_t *__restrict__ text, IngState *st, const unsigned *__restrict__ line_end,
                                                                 uint8_t *__restrict__ carry_tmp, int phase)
{
    __shared__ unsigned long long c_from, c_len;
    if (threadIdx.x == 0) {
        const unsigned n_rec = st->n_rec;
        c_from = n_rec ? (unsigned long long)line_end[4 * n_rec - 1] + 1 : st->t0;
        c_len = st->t1 - c_from;
        if (c_len > ING_MAXCARRY) { st->irregular = 1; c_len = 0; }
        if (st->last_chunk && c_len) { st->irregular = 1; c_len = 0; }     // truncated last record: the host parser's business
    }
    __syncthreads();
    if (phase == 0) {
        for (unsigned long long i = threadIdx.x; i < c_len; i += blockDim.x) carry_tmp[i] = text[c_from + i];
    } else {
        for (unsigned long long i = threadIdx.x; i < c_len; i += blockDim.x) text[ING_MAXCARRY - c_len + i] = carry_tmp[i];
        __syncthreads();
        if (threadIdx.x == 0) { st->carry_len = c_len; st->records += st->n_rec; }
    }
}

__global__ void ing_begin_chunk(IngState *st, unsigned long long new_bytes, unsigned last_chunk, unsigned first_chunk)
{
    if (first_chunk) { st->carry_len = 0; st->tail_len = 0; }
    st->first_chunk = first_chunk;
    st->t0 = ING_MAXCARRY - st->carry_len;
    st->t1 = ING_MAXCARRY + new_bytes;
    st->last_chunk = last_chunk;
    st->flat_len = 0;
}

// terminate a final line that lacks its '\n' (the parser the reference uses accepts that)
__global__ void ing_terminate_last_line(uint8_t *text, IngState *st)
{
    if (st->last_chunk && st->t1 > st->t0 && text[st->t1 - 1] != '\n') { text[st->t1] = '\n'; st->t1 += 1; }
}

__global__ void ing_set_flat_len(IngState *st, const unsigned *total) { st->flat_len = *total; }

// ------------------------------------------------------------------------------------------------
// strict FASTA: '>' lines are headers, every other line is sequence (joined), nothing else
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ING_THREADS) ing_fasta_len(const uint8_t *__restrict__ text, IngState *st, const unsigned *__restrict__ line_end,
                                                              unsigned *__restrict__ out_len, int count_stats)
{
    const unsigned n_lines = st->n_lines;
    unsigned long long bases = 0, recs = 0;
    for (unsigned L = blockIdx.x * blockDim.x + threadIdx.x; L < n_lines; L += gridDim.x * blockDim.x) {
        const unsigned long long s0 = L ? (unsigned long long)line_end[L - 1] + 1 : st->t0;
        const unsigned len = line_end[L] - (unsigned)s0;
        const uint8_t first = len ? text[s0] : 0;
        const bool header = first == '>';
        if (first == '@' || first == '+') atomicOr(&st->irregular, 1u);              // the reference's parser would switch to FASTQ rules
        if (L == 0 && st->first_chunk && !header) atomicOr(&st->irregular, 1u);       // text before the first record
        out_len[L] = header ? 1u : len;                                               // a header becomes the record separator
        if (header) ++recs; else bases += len;
    }
    if (count_stats) {
        if (bases) atomicAdd(&st->bases, bases);
        if (recs) atomicAdd(&st->records, recs);
    }
}

__global__ void __launch_bounds__(ING_THREADS) ing_fasta_copy(const uint8_t *__restrict__ text, const IngState *st, const unsigned *__restrict__ line_end,
                                                               const unsigned *__restrict__ out_off, uint8_t *__restrict__ flat)
{
    const unsigned n_lines = st->n_lines, tail = st->tail_len;
    const int lane = threadIdx.x & 31;
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    if (warp == 0) for (unsigned i = lane; i < tail; i += 32) flat[i] = st->tail[i];
    for (unsigned L = warp; L < n_lines; L += n_warps) {
        const unsigned long long s0 = L ? (unsigned long long)line_end[L - 1] + 1 : st->t0;
        const unsigned len = line_end[L] - (unsigned)s0;
        uint8_t *dst = flat + tail + out_off[L];
        if (len && text[s0] == '>') { if (lane == 0) dst[0] = '\n'; continue; }
        for (unsigned i = lane; i < len; i += 32) dst[i] = text[s0 + i];
    }
}

__global__ void ing_fasta_set_flat_len(IngState *st, const unsigned *total) { st->flat_len = (unsigned long long)st->tail_len + *total; }

// end of a FASTA chunk: carry the partial last line, remember the last 30 bytes of the flat stream
__global__ void __launch_bounds__(ING_THREADS) ing_fasta_finish(uint8_t *__restrict__ text, IngState *st, const unsigned *__restrict__ line_end,
                                                                 uint8_t *__restrict__ carry_tmp, const uint8_t *__restrict__ flat, int phase, int scan)
{
    __shared__ unsigned long long c_from, c_len;
    if (threadIdx.x == 0) {
        const unsigned n_lines = st->n_lines;
        c_from = n_lines ? (unsigned long long)line_end[n_lines - 1] + 1 : st->t0;
        c_len = st->t1 - c_from;
        if (c_len > ING_MAXCARRY) { st->irregular = 1; c_len = 0; }
    }
    __syncthreads();
    if (phase == 0) {
        for (unsigned long long i = threadIdx.x; i < c_len; i += blockDim.x) carry_tmp[i] = text[c_from + i];
    } else {
        for (unsigned long long i = threadIdx.x; i < c_len; i += blockDim.x) text[ING_MAXCARRY - c_len + i] = carry_tmp[i];
        __syncthreads();
        if (threadIdx.x == 0) {
            st->carry_len = c_len;
            if (scan) {
                const unsigned long long fl = st->flat_len;
                const unsigned keep = fl < (S2_K - 1) ? (unsigned)fl : (S2_K - 1);
                unsigned char tmp[32];
                for (unsigned i = 0; i < keep; ++i) tmp[i] = flat[fl - keep + i];
                for (unsigned i = 0; i < keep; ++i) st->tail[i] = tmp[i];
                st->tail_len = keep;
            }
        }
    }
}

// ---- detect mode (strain_detect pass 1 on ingested reads) ------------------------------------------
__global__ void __launch_bounds__(ING_THREADS) ing_make_rec_off(const IngState *st, const unsigned *__restrict__ out_off, unsigned long long *__restrict__ rec_off)
{
    const unsigned n_rec = st->n_rec;
    for (unsigned r = blockIdx.x * blockDim.x + threadIdx.x; r <= n_rec; r += gridDim.x * blockDim.x)
        rec_off[r] = r < n_rec ? out_off[r] : st->flat_len;
}

// per-record results of this chunk -> the file-level arrays (record numbering continues across chunks)
__global__ void __launch_bounds__(ING_THREADS) ing_store_records(const IngState *st, const unsigned *__restrict__ line_end, const unsigned *__restrict__ hits_c,
                                                                  const unsigned *__restrict__ inf_c, unsigned *__restrict__ len_all,
                                                                  unsigned *__restrict__ hits_all, unsigned *__restrict__ inf_all, unsigned long long cap)
{
    const unsigned n_rec = st->n_rec;
    const unsigned long long base = st->records;
    for (unsigned r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rec; r += gridDim.x * blockDim.x) {
        if (base + r >= cap) break;
        len_all[base + r] = line_end[4 * r + 1] - (line_end[4 * r] + 1);
        hits_all[base + r] = hits_c[r];
        inf_all[base + r] = inf_c[r];
    }
}

// informative windows of this chunk (byte offsets in the flat batch) -> (record number, offset, canonical k-mer)
__global__ void __launch_bounds__(ING_THREADS) ing_collect_inf(IngState *st, const uint8_t *__restrict__ flat, const unsigned long long *__restrict__ rec_off,
                                                                const unsigned long long *__restrict__ pos_c, const unsigned long long *__restrict__ cnt_c,
                                                                unsigned long long cap_c, unsigned *__restrict__ f_rec, unsigned *__restrict__ f_off,
                                                                unsigned long long *__restrict__ f_kmer, unsigned long long *__restrict__ f_cnt, unsigned long long cap_f)
{
    const unsigned long long n = *cnt_c;
    if (n > cap_c) { if (blockIdx.x == 0 && threadIdx.x == 0) st->inf_overflow = 1; }
    const unsigned n_rec = st->n_rec;
    const unsigned long long base = st->records;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n && i < cap_c; i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long pos = pos_c[i];
        unsigned lo = 0, hi = n_rec;
        while (hi - lo > 1) { const unsigned mid = (lo + hi) >> 1; if (rec_off[mid] <= pos) lo = mid; else hi = mid; }
        unsigned long long fwd = 0;
        for (int b = 0; b < S2_K; ++b) {
            const unsigned c = flat[pos + b];
            const unsigned x = (c >> 1) & 3u;
            fwd = (fwd << 2) | (x ^ (x >> 1));
        }
        const unsigned long long at = atomicAdd(f_cnt, 1ull);
        if (at < cap_f) { f_rec[at] = (unsigned)(base + lo); f_off[at] = (unsigned)(pos - rec_off[lo]); f_kmer[at] = s2_canonical(fwd, s2_revcomp31(fwd)); }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*decompress_fn)(CUmemDecompressParams *, size_t, unsigned int, size_t *, CUstream);
typedef CUresult (*devattr_fn)(int *, CUdevice_attribute, CUdevice);

struct s2_ingest {
    s2_ctx *ctx = nullptr;
    cudaStream_t stream = nullptr;
    uint8_t *h_comp[2] = { nullptr, nullptr }, *d_comp = nullptr, *d_text = nullptr, *d_flat = nullptr, *d_carry = nullptr;
    cudaEvent_t h_free[2] = { nullptr, nullptr };      // pinned buffer b may be overwritten once its H2D copy is done
    uint8_t *d_file = nullptr; size_t d_file_cap = 0;  // the whole compressed file, kept on the device between the two passes
    unsigned *d_block_nl = nullptr, *d_line_end = nullptr, *d_out = nullptr, *d_total = nullptr, *d_act = nullptr;
    IngState *d_state = nullptr, *h_state = nullptr;
    decompress_fn decompress = nullptr;
    bool hw_deflate = false;
    bool fasta = false;                // format of the file being ingested (strict FASTA instead of strict FASTQ)
    // detect mode: chunk-local and file-level result arrays
    unsigned *d_hits_c = nullptr, *d_inf_c = nullptr;
    unsigned long long *d_rec_off = nullptr, *d_pos_c = nullptr, *d_cnt_c = nullptr, *d_fcnt = nullptr;
    unsigned *d_len_all = nullptr, *d_hits_all = nullptr, *d_inf_all = nullptr; unsigned long long rec_cap = 0;
    unsigned *d_frec = nullptr, *d_foff = nullptr; unsigned long long *d_fkmer = nullptr; unsigned long long f_cap = 0;
    struct Chunk { std::vector<CUmemDecompressParams> params; size_t src_off = 0, src_len = 0, text_len = 0; bool eof = false; };
    std::vector<Chunk> chunks;                         // pass 1 records them, pass 2 replays them without touching the host
};

static void ingest_free(s2_ingest *g)
{
    if (!g) return;
    cudaSetDevice(g->ctx->device);
    if (g->stream) { cudaStreamSynchronize(g->stream); cudaStreamDestroy(g->stream); }
    for (int b = 0; b < 2; ++b) { cudaFreeHost(g->h_comp[b]); if (g->h_free[b]) cudaEventDestroy(g->h_free[b]); }
    cudaFree(g->d_file); cudaFree(g->d_comp); cudaFree(g->d_text); cudaFree(g->d_flat); cudaFree(g->d_carry);
    cudaFree(g->d_block_nl); cudaFree(g->d_line_end); cudaFree(g->d_out); cudaFree(g->d_total); cudaFree(g->d_act);
    cudaFree(g->d_state); cudaFreeHost(g->h_state);
    cudaFree(g->d_hits_c); cudaFree(g->d_inf_c); cudaFree(g->d_rec_off); cudaFree(g->d_pos_c); cudaFree(g->d_cnt_c); cudaFree(g->d_fcnt);
    cudaFree(g->d_len_all); cudaFree(g->d_hits_all); cudaFree(g->d_inf_all); cudaFree(g->d_frec); cudaFree(g->d_foff); cudaFree(g->d_fkmer);
    delete g;
}

static int ingest_init(s2_ingest *g, s2_ctx *c)
{
    g->ctx = c;
    CK(cudaSetDevice(c->device));
    CK(cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking));
    for (int b = 0; b < 2; ++b) {
        CK(cudaHostAlloc((void **)&g->h_comp[b], ING_COMP_CHUNK, cudaHostAllocDefault));
        CK(cudaEventCreateWithFlags(&g->h_free[b], cudaEventDisableTiming));
    }
    CK(cudaMalloc((void **)&g->d_comp, ING_COMP_CHUNK + 256));
    CK(cudaMalloc((void **)&g->d_text, (size_t)ING_MAXCARRY + ING_TEXT_CAP + 4096));
    CK(cudaMalloc((void **)&g->d_flat, (size_t)ING_TEXT_CAP + ING_MAXCARRY + 4096));
    CK(cudaMalloc((void **)&g->d_carry, ING_MAXCARRY));
    const size_t n_blocks = ((size_t)ING_MAXCARRY + ING_TEXT_CAP) / 4096 + 2;
    CK(cudaMalloc((void **)&g->d_block_nl, n_blocks * sizeof(unsigned)));
    CK(cudaMalloc((void **)&g->d_line_end, (size_t)ING_MAX_LINES * sizeof(unsigned)));
    CK(cudaMalloc((void **)&g->d_out, ((size_t)ING_MAX_LINES + 4) * sizeof(unsigned)));     // per record (FASTQ) or per line (FASTA)
    CK(cudaMalloc((void **)&g->d_total, 4 * sizeof(unsigned)));
    CK(cudaMalloc((void **)&g->d_act, (ING_TEXT_CAP / 256) * sizeof(unsigned)));
    CK(cudaMalloc((void **)&g->d_state, sizeof(IngState)));
    CK(cudaHostAlloc((void **)&g->h_state, sizeof(IngState), cudaHostAllocDefault));
    // the decompression engine is reached through the driver; no link-time dependency on libcuda
    cudaDriverEntryPointQueryResult q;
    void *fn = nullptr, *fa = nullptr;
    if (cudaGetDriverEntryPoint("cuMemBatchDecompressAsync", &fn, cudaEnableDefault, &q) == cudaSuccess && fn &&
        cudaGetDriverEntryPoint("cuDeviceGetAttribute", &fa, cudaEnableDefault, &q) == cudaSuccess && fa) {
        int mask = 0, maxlen = 0;
        ((devattr_fn)fa)(&mask, CU_DEVICE_ATTRIBUTE_MEM_DECOMPRESS_ALGORITHM_MASK, c->device);
        ((devattr_fn)fa)(&maxlen, CU_DEVICE_ATTRIBUTE_MEM_DECOMPRESS_MAXIMUM_LENGTH, c->device);
        g->hw_deflate = (mask & CU_MEM_DECOMPRESS_ALGORITHM_DEFLATE) && maxlen >= 65536;
        g->decompress = (decompress_fn)fn;
    }
    cudaGetLastError();
    return 0;
}

// a BGZF member header at p (RFC 1952 + the 'BC' extra subfield): total block size, offset/length of its deflate data, ISIZE
static bool bgzf_block(const uint8_t *p, size_t avail, size_t *block_size, size_t *data_off, size_t *data_len, uint32_t *isize)
{
    if (avail < 18 || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4)) return false;
    const size_t xlen = p[10] | (p[11] << 8);
    if (avail < 12 + xlen) return false;
    size_t bsize = 0;
    for (size_t o = 12; o + 4 <= 12 + xlen;) {
        const size_t slen = p[o + 2] | (p[o + 3] << 8);
        if (p[o] == 'B' && p[o + 1] == 'C' && slen == 2 && o + 6 <= 12 + xlen) bsize = (size_t)(p[o + 4] | (p[o + 5] << 8)) + 1;
        o += 4 + slen;
    }
    if (!bsize || (p[3] & ~4)) return false;                 // only FEXTRA set, as bgzip writes it
    if (avail < bsize) { *block_size = bsize; *data_len = (size_t)-1; return true; }      // incomplete in this buffer
    if (bsize < 12 + xlen + 8) return false;
    *block_size = bsize;
    *data_off = 12 + xlen;
    *data_len = bsize - 12 - xlen - 8;
    *isize = (uint32_t)p[bsize - 4] | ((uint32_t)p[bsize - 3] << 8) | ((uint32_t)p[bsize - 2] << 16) | ((uint32_t)p[bsize - 1] << 24);
    return true;
}

static bool is_bgzf_header(const uint8_t *p, ssize_t n)
{
    return n >= 18 && p[0] == 0x1f && p[1] == 0x8b && p[2] == 8 && p[3] == 4 && (p[10] | (p[11] << 8)) >= 6 &&
           p[12] == 'B' && p[13] == 'C' && p[14] == 2 && p[15] == 0;
}

// where the compressed (or plain) bytes come from: a file read chunk by chunk into pinned staging buffers, or a
// caller's host buffer (pinned for full PCIe rate) that is copied to the device directly
struct IngSource {
    int fd = -1;
    const uint8_t *mem = nullptr;
    size_t mem_len = 0;
    ssize_t size() const { if (mem) return (ssize_t)mem_len; struct stat sb; return fstat(fd, &sb) == 0 ? (ssize_t)sb.st_size : -1; }
    ssize_t peek(void *dst, size_t len, off_t off) const
    {
        if (!mem) return pread(fd, dst, len, off);
        if ((size_t)off >= mem_len) return 0;
        const size_t n = std::min(len, mem_len - (size_t)off);
        memcpy(dst, mem + off, n);
        return (ssize_t)n;
    }
};

// device work for one chunk whose bytes are already on the device: inflate (or copy), index, validate,
// and - when scan is set - copy the sequence lines out and count them
#define ING_CAP_C (8ull << 20)        /* informative windows one chunk may report */
enum { ING_VALIDATE = 0, ING_COUNT = 1, ING_DETECT = 2 };

static int ingest_chunk(s2_ingest *g, s2_table *t, int col, int mode, bool bgzf, const s2_ingest::Chunk &ch, const uint8_t *d_src, bool first)
{
    const bool scan = mode != ING_VALIDATE;
    const bool fasta = g->fasta;
    s2_ctx *c = g->ctx;
    cudaStream_t st = g->stream;
    const unsigned text_blocks = (unsigned)(((size_t)ING_MAXCARRY + ch.text_len) / 4096 + 1);      // covers [0, t1] of the text buffer
    ing_begin_chunk<<<1, 1, 0, st>>>(g->d_state, ch.text_len, ch.eof ? 1u : 0u, first ? 1u : 0u);
    if (bgzf) {
        if (!ch.params.empty()) {
            size_t err_index = 0;
            const CUresult r = g->decompress(const_cast<CUmemDecompressParams *>(ch.params.data()), ch.params.size(), 0, &err_index, (CUstream)st);
            if (r != CUDA_SUCCESS) { s2_set_error("hardware decompression failed (driver error %d at block %zu)", (int)r, err_index); return -1; }
        }
    } else if (ch.src_len) {
        CK(cudaMemcpyAsync(g->d_text + ING_MAXCARRY, d_src, ch.src_len, cudaMemcpyDeviceToDevice, st));
    }
    ing_terminate_last_line<<<1, 1, 0, st>>>(g->d_text, g->d_state);
    ing_nl_count<<<text_blocks, ING_THREADS, 0, st>>>(g->d_text, g->d_state, g->d_block_nl);
    ing_scan_u32<<<1, ING_THREADS, 0, st>>>(g->d_block_nl, text_blocks, nullptr, &g->d_state->n_lines);
    ing_nl_scatter<<<text_blocks, ING_THREADS, 0, st>>>(g->d_text, g->d_state, g->d_block_nl, g->d_line_end);
    ing_fastq_prepare<<<1, 1, 0, st>>>(g->d_state);
    if (fasta) {
        ing_fasta_len<<<c->n_sm * 4, ING_THREADS, 0, st>>>(g->d_text, g->d_state, g->d_line_end, g->d_out, scan ? 1 : 0);
        if (scan) {
            ing_scan_u32<<<1, ING_THREADS, 0, st>>>(g->d_out, 0, &g->d_state->n_lines, g->d_total);
            ing_fasta_set_flat_len<<<1, 1, 0, st>>>(g->d_state, g->d_total);
            ing_fasta_copy<<<c->n_sm * 8, ING_THREADS, 0, st>>>(g->d_text, g->d_state, g->d_line_end, g->d_out, g->d_flat);
            s2_launch_scan_count_devlen(g->d_flat, &g->d_state->flat_len, t->v, col, c->d_stats, c->grid_count, st);
        }
        ing_fasta_finish<<<1, ING_THREADS, 0, st>>>(g->d_text, g->d_state, g->d_line_end, g->d_carry, g->d_flat, 0, scan ? 1 : 0);
        ing_fasta_finish<<<1, ING_THREADS, 0, st>>>(g->d_text, g->d_state, g->d_line_end, g->d_carry, g->d_flat, 1, scan ? 1 : 0);
        CK(cudaGetLastError());
        return 0;
    }
    ing_fastq_len<<<c->n_sm * 4, ING_THREADS, 0, st>>>(g->d_text, g->d_state, g->d_line_end, g->d_out, scan ? 1 : 0);
    if (scan) {
        ing_scan_u32<<<1, ING_THREADS, 0, st>>>(g->d_out, 0, &g->d_state->n_rec, g->d_total);
        ing_set_flat_len<<<1, 1, 0, st>>>(g->d_state, g->d_total);
        ing_fastq_copy<<<c->n_sm * 8, ING_THREADS, 0, st>>>(g->d_text, g->d_state, g->d_line_end, g->d_out, g->d_flat);
        if (mode == ING_COUNT) {
            s2_launch_scan_count_devlen(g->d_flat, &g->d_state->flat_len, t->v, col, c->d_stats, c->grid_count, st);
        } else {
            const size_t max_rec = (size_t)ING_MAX_LINES / 4 + 4;
            ing_make_rec_off<<<c->n_sm * 2, ING_THREADS, 0, st>>>(g->d_state, g->d_out, g->d_rec_off);
            CK(cudaMemsetAsync(g->d_hits_c, 0, max_rec * sizeof(unsigned), st));
            CK(cudaMemsetAsync(g->d_inf_c, 0, max_rec * sizeof(unsigned), st));
            CK(cudaMemsetAsync(g->d_cnt_c, 0, sizeof(unsigned long long), st));
            S2DetectOut out;
            out.rec_off = (const uint64_t *)g->d_rec_off; out.n_rec = 0; out.n_rec_dev = &g->d_state->n_rec; out.n_bytes_dev = &g->d_state->flat_len;
            out.read_hits = g->d_hits_c; out.read_inf = g->d_inf_c; out.inf_pos = (uint64_t *)g->d_pos_c; out.inf_count = g->d_cnt_c; out.inf_cap = ING_CAP_C;
            s2_launch_scan_detect_dev(g->d_flat, t->v, out, c->d_stats, c->grid_detect, st);
            ing_collect_inf<<<c->n_sm * 2, ING_THREADS, 0, st>>>(g->d_state, g->d_flat, g->d_rec_off, g->d_pos_c, g->d_cnt_c, ING_CAP_C,
                                                                  g->d_frec, g->d_foff, g->d_fkmer, g->d_fcnt, g->f_cap);
            ing_store_records<<<c->n_sm * 2, ING_THREADS, 0, st>>>(g->d_state, g->d_line_end, g->d_hits_c, g->d_inf_c, g->d_len_all, g->d_hits_all,
                                                                    g->d_inf_all, g->rec_cap);
        }
    }
    ing_fastq_finish<<<1, ING_THREADS, 0, st>>>(g->d_text, g->d_state, g->d_line_end, g->d_carry, 0);
    ing_fastq_finish<<<1, ING_THREADS, 0, st>>>(g->d_text, g->d_state, g->d_line_end, g->d_carry, 1);
    CK(cudaGetLastError());
    return 0;
}

// Pass over the file FROM THE HOST: read() chunks of whole BGZF blocks (or of raw text) into two alternating
// pinned buffers, copy them to the device (into the file cache when the file fits, so that the second pass
// needs no I/O), and run the device stage.  Returns 0 ok, 1 irregular / unsupported, -1 error.
static int ingest_pass_host(s2_ingest *g, s2_table *t, const IngSource &src, bool bgzf, int col, int mode, bool cache)
{
    cudaStream_t st = g->stream;
    CK(cudaMemsetAsync(g->d_state, 0, sizeof(IngState), st));
    g->chunks.clear();
    off_t file_off = 0;
    bool first = true, eof = false;
    int buf = 0;
    while (!eof) {
        const uint8_t *h = g->h_comp[buf];
        ssize_t got;
        if (src.mem) {                                                           // caller's buffer: no staging copy
            h = src.mem + file_off;
            got = (size_t)file_off < src.mem_len ? (ssize_t)std::min<size_t>(ING_COMP_CHUNK, src.mem_len - (size_t)file_off) : 0;
        } else {
            CK(cudaEventSynchronize(g->h_free[buf]));                            // its previous H2D copy is done
            got = pread(src.fd, g->h_comp[buf], ING_COMP_CHUNK, file_off);
            if (got < 0) { s2_set_error("read failed"); return -1; }
        }
        uint8_t *d_dst = cache ? g->d_file + file_off : g->d_comp;
        s2_ingest::Chunk local, &ch = cache ? (g->chunks.emplace_back(), g->chunks.back()) : local;
        size_t used = 0, text_len = 0;
        if (bgzf) {
            while (used < (size_t)got) {
                size_t bs = 0, doff = 0, dlen = 0; uint32_t isz = 0;
                if (!bgzf_block(h + used, (size_t)got - used, &bs, &doff, &dlen, &isz)) return 1;        // not BGZF after all
                if (dlen == (size_t)-1) break;                                   // partial block: the next chunk starts here
                if (isz > 65536) return 1;
                if (text_len + isz > ING_TEXT_CAP) break;
                if (isz) {
                    CUmemDecompressParams p; memset(&p, 0, sizeof p);
                    p.srcNumBytes = dlen; p.dstNumBytes = isz;
                    p.dstActBytes = (cuuint32_t *)(g->d_act + ch.params.size());
                    p.src = d_dst + used + doff;
                    p.dst = g->d_text + ING_MAXCARRY + text_len;
                    p.algo = CU_MEM_DECOMPRESS_ALGORITHM_DEFLATE;
                    ch.params.push_back(p);
                    text_len += isz;
                }
                used += bs;
            }
            if (used == 0 && got > 0) return 1;                                  // a block larger than the chunk: not bgzip output
        } else {
            used = (size_t)got;
            text_len = used;
        }
        eof = (size_t)got < ING_COMP_CHUNK && used == (size_t)got;              // short read and everything consumed
        ch.src_off = (size_t)file_off; ch.src_len = used; ch.text_len = text_len; ch.eof = eof;
        file_off += (off_t)used;
        if (!cache) CK(cudaStreamSynchronize(st));                               // ch.params (a local) and d_comp are reused
        if (used) CK(cudaMemcpyAsync(d_dst, h, used, cudaMemcpyHostToDevice, st));
        CK(cudaEventRecord(g->h_free[buf], st));
        if (ingest_chunk(g, t, col, mode, bgzf, ch, d_dst, first)) return -1;
        if (!cache) CK(cudaStreamSynchronize(st));
        first = false;
        buf ^= 1;
        if (got == 0) break;
    }
    CK(cudaMemcpyAsync(g->h_state, g->d_state, sizeof(IngState), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return g->h_state->irregular ? 1 : 0;
}

// second pass when the compressed file is resident on the device: replay the recorded chunks, no host I/O
static int ingest_pass_cached(s2_ingest *g, s2_table *t, bool bgzf, int col, int mode)
{
    cudaStream_t st = g->stream;
    CK(cudaMemsetAsync(g->d_state, 0, sizeof(IngState), st));
    bool first = true;
    for (const auto &ch : g->chunks) {
        if (ingest_chunk(g, t, col, mode, bgzf, ch, g->d_file + ch.src_off, first)) return -1;
        first = false;
    }
    CK(cudaMemcpyAsync(g->h_state, g->d_state, sizeof(IngState), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return g->h_state->irregular ? 1 : 0;
}

static thread_local s2_ingest *tl_ingest = nullptr;

// first byte of the text inside a BGZF file ('@' FASTQ, '>' FASTA): inflate the first non-empty block on the host
static int bgzf_first_text_byte(const IngSource &src)
{
    std::vector<uint8_t> buf(1 << 17);
    const ssize_t got = src.peek(buf.data(), buf.size(), 0);
    size_t used = 0;
    while (got > 0 && used < (size_t)got) {
        size_t bs = 0, doff = 0, dlen = 0; uint32_t isz = 0;
        if (!bgzf_block(buf.data() + used, (size_t)got - used, &bs, &doff, &dlen, &isz) || dlen == (size_t)-1) return -1;
        if (isz) {
            z_stream z; memset(&z, 0, sizeof z);
            if (inflateInit2(&z, -15) != Z_OK) return -1;
            uint8_t out[16];
            z.next_in = buf.data() + used + doff; z.avail_in = (uInt)dlen;
            z.next_out = out; z.avail_out = sizeof out;
            inflate(&z, Z_SYNC_FLUSH);
            const int first = z.total_out ? out[0] : -1;
            inflateEnd(&z);
            return first;
        }
        used += bs;
    }
    return -1;
}

// classify a source for the GPU path and make sure this thread's pipeline (and its file cache) exists.
// Returns 0 ready, 1 not eligible, -1 error.
static int ingest_prepare(s2_ctx *c, const IngSource &src, bool *bgzf_out, bool *cache_out, s2_ingest **g_out)
{
    uint8_t head[32];
    const ssize_t hn = src.peek(head, sizeof head, 0);
    const bool bgzf = is_bgzf_header(head, hn);
    const int first = bgzf ? bgzf_first_text_byte(src) : (hn >= 1 ? head[0] : -1);
    if (first != '@' && first != '>') return 1;                           // neither FASTQ nor FASTA (or an ordinary .gz): host reader
    // uncompressed text gains nothing but PCIe from this path (the host parser does GB/s per thread): opt-in only
    if (!bgzf && !s2_env_int("S2_GPU_INGEST_PLAIN", 0)) return 1;
    if (tl_ingest && tl_ingest->ctx != c) { ingest_free(tl_ingest); tl_ingest = nullptr; }
    if (!tl_ingest) {
        tl_ingest = new s2_ingest();
        if (ingest_init(tl_ingest, c)) { ingest_free(tl_ingest); tl_ingest = nullptr; return -1; }
    }
    s2_ingest *g = tl_ingest;
    if (bgzf && !g->hw_deflate) return 1;
    // keep the compressed file on the device between the two passes when it fits (S2_INGEST_CACHE_MB, default 2048)
    const ssize_t size = src.size();
    bool cache = size >= 0 && (uint64_t)size <= (s2_env_u64("S2_INGEST_CACHE_MB", 2048) << 20);
    if (cache && (size_t)size + ING_COMP_CHUNK > g->d_file_cap) {
        cudaStreamSynchronize(g->stream);
        cudaFree(g->d_file); g->d_file = nullptr; g->d_file_cap = 0;
        const size_t want = (size_t)size + ING_COMP_CHUNK + ((size_t)size >> 2);
        if (cudaMalloc((void **)&g->d_file, want) == cudaSuccess) g->d_file_cap = want; else { cudaGetLastError(); cache = false; }
    }
    g->fasta = first == '>';
    *bgzf_out = bgzf; *cache_out = cache; *g_out = g;
    return 0;
}

static int ingest_open(s2_ctx *c, const char *path, IngSource *src, bool *bgzf_out, bool *cache_out, s2_ingest **g_out)
{
    src->fd = open(path, O_RDONLY);
    if (src->fd < 0) return 1;
    const int rc = ingest_prepare(c, *src, bgzf_out, cache_out, g_out);
    if (rc) { close(src->fd); src->fd = -1; }
    return rc;
}

static int ingest_count_source(s2_ingest *g, s2_table *t, const IngSource &src, bool bgzf, bool cache, int col, uint64_t *bases, uint64_t *lookups)
{
    int rc = ingest_pass_host(g, t, src, bgzf, col, ING_VALIDATE, cache);         // pass 1: prove the text is strict FASTQ / FASTA
    if (rc == 0) rc = cache ? ingest_pass_cached(g, t, bgzf, col, ING_COUNT)      // pass 2: count
                            : ingest_pass_host(g, t, src, bgzf, col, ING_COUNT, false);
    if (rc == 0) {
        if (bases) *bases = g->h_state->bases;
        // FASTA records are not measured one by one on the device: every record is assumed to have at least one window
        if (lookups) *lookups = g->fasta ? (g->h_state->bases > 30 * g->h_state->records ? g->h_state->bases - 30 * g->h_state->records : 0)
                                         : g->h_state->lookups;
    }
    return rc;
}

// GEN_calculate_kmer_count for one file, entirely on the GPU when the file is BGZF-compressed or plain strict
// FASTQ.  Returns 0 = done (counters updated, *bases / *lookups set), 1 = not handled (nothing was counted: use
// the host reader), -1 = error.
extern "C" int s2_ingest_count_file(s2_ctx *c, s2_table *t, const char *path, int col, uint64_t *bases, uint64_t *lookups)
{
    if (col < 0 || col >= t->v.n_cols) { s2_set_error("column out of range"); return -1; }
    if (t->partitioned) return 1;                                    // union tables keep the host reader + two-phase scan
    IngSource src; bool bgzf, cache; s2_ingest *g;
    int rc = ingest_open(c, path, &src, &bgzf, &cache, &g);
    if (rc) return rc;
    rc = ingest_count_source(g, t, src, bgzf, cache, col, bases, lookups);
    close(src.fd);
    return rc;
}

// The same for a file image that is already in host memory (the bytes of a BGZF or plain FASTA/FASTQ file; pinned
// memory gives the full PCIe rate).  Only the compressed bytes cross PCIe.
extern "C" int s2_ingest_count_mem(s2_ctx *c, s2_table *t, const void *image, uint64_t n_bytes, int col, uint64_t *bases, uint64_t *lookups)
{
    if (col < 0 || col >= t->v.n_cols) { s2_set_error("column out of range"); return -1; }
    if (t->partitioned) return 1;
    IngSource src; src.mem = (const uint8_t *)image; src.mem_len = (size_t)n_bytes;
    bool bgzf, cache; s2_ingest *g;
    const int rc = ingest_prepare(c, src, &bgzf, &cache, &g);
    if (rc) return rc;
    return ingest_count_source(g, t, src, bgzf, cache, col, bases, lookups);
}

// Pass 1 of quantify_hits_PE (src/strain_detect.c:465-491) for every read of one file, inflated and split on the
// GPU.  out->len / hits / inf are per record in file order (all records, also those shorter than 31);
// out->inf_* list the informative windows sorted by (record, offset) with their canonical k-mer.  The arrays are
// malloc()ed here and released by s2_ingest_detect_free.  Returns 0 / 1 (not handled) / -1 like the count form.
extern "C" int s2_ingest_detect_file(s2_ctx *c, s2_table *t, const char *path, s2_ingest_detect_result *out)
{
    memset(out, 0, sizeof *out);
    if (t->partitioned) return 1;
    IngSource src; bool bgzf, cache; s2_ingest *g;
    int rc = ingest_open(c, path, &src, &bgzf, &cache, &g);
    if (rc) return rc;
    const int fd = src.fd;
    if (g->fasta) { close(fd); return 1; }                            // per-read results are a FASTQ feature here
    rc = ingest_pass_host(g, t, src, bgzf, 0, ING_VALIDATE, cache);
    if (rc) { close(fd); return rc; }
    const unsigned long long n_rec = g->h_state->records;
    const size_t max_rec = (size_t)ING_MAX_LINES / 4 + 4;
    auto dev_alloc = [](void **p, size_t bytes) { return *p ? cudaSuccess : cudaMalloc(p, bytes); };
    if (dev_alloc((void **)&g->d_hits_c, max_rec * 4) || dev_alloc((void **)&g->d_inf_c, max_rec * 4) ||
        dev_alloc((void **)&g->d_rec_off, (max_rec + 1) * 8) || dev_alloc((void **)&g->d_pos_c, (ING_CAP_C + 1) * 8) ||
        dev_alloc((void **)&g->d_cnt_c, 8) || dev_alloc((void **)&g->d_fcnt, 8)) { s2_set_error("out of device memory"); close(fd); return -1; }
    if (n_rec + 1 > g->rec_cap) {
        cudaFree(g->d_len_all); cudaFree(g->d_hits_all); cudaFree(g->d_inf_all);
        g->d_len_all = g->d_hits_all = g->d_inf_all = nullptr;
        g->rec_cap = n_rec + n_rec / 8 + 1024;
        if (cudaMalloc((void **)&g->d_len_all, g->rec_cap * 4) || cudaMalloc((void **)&g->d_hits_all, g->rec_cap * 4) ||
            cudaMalloc((void **)&g->d_inf_all, g->rec_cap * 4)) { s2_set_error("out of device memory"); g->rec_cap = 0; close(fd); return -1; }
    }
    unsigned long long n_inf = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        if (!g->d_frec) {
            if (!g->f_cap) g->f_cap = 1ull << 20;
            if (cudaMalloc((void **)&g->d_frec, g->f_cap * 4) || cudaMalloc((void **)&g->d_foff, g->f_cap * 4) ||
                cudaMalloc((void **)&g->d_fkmer, g->f_cap * 8)) { s2_set_error("out of device memory"); close(fd); return -1; }
        }
        CK(cudaMemsetAsync(g->d_fcnt, 0, 8, g->stream));
        rc = cache ? ingest_pass_cached(g, t, bgzf, 0, ING_DETECT) : ingest_pass_host(g, t, src, bgzf, 0, ING_DETECT, false);
        if (rc) break;
        CK(cudaMemcpy(&n_inf, g->d_fcnt, 8, cudaMemcpyDeviceToHost));
        if (g->h_state->inf_overflow) { rc = 1; break; }                         // absurdly dense chunk: host path
        if (n_inf <= g->f_cap) break;
        cudaFree(g->d_frec); cudaFree(g->d_foff); cudaFree(g->d_fkmer);          // list too small: grow and run the pass again
        g->d_frec = g->d_foff = nullptr; g->d_fkmer = nullptr;
        g->f_cap = n_inf + n_inf / 8 + 1024;
        rc = 2;
    }
    close(fd);
    if (rc) return rc == 2 ? -1 : rc;
    out->n_records = n_rec; out->n_inf = n_inf; out->bases = g->h_state->bases;
    out->len = (uint32_t *)malloc((n_rec + 1) * 4); out->hits = (uint32_t *)malloc((n_rec + 1) * 4); out->inf = (uint32_t *)malloc((n_rec + 1) * 4);
    out->inf_rec = (uint32_t *)malloc((n_inf + 1) * 4); out->inf_off = (uint32_t *)malloc((n_inf + 1) * 4); out->inf_kmer = (uint64_t *)malloc((n_inf + 1) * 8);
    CK(cudaMemcpy(out->len, g->d_len_all, n_rec * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out->hits, g->d_hits_all, n_rec * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out->inf, g->d_inf_all, n_rec * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out->inf_rec, g->d_frec, n_inf * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out->inf_off, g->d_foff, n_inf * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out->inf_kmer, g->d_fkmer, n_inf * 8, cudaMemcpyDeviceToHost));
    // the kernels append in arbitrary order: sort by (record, offset) = the order pass 2 prints them
    std::vector<uint32_t> perm(n_inf);
    for (uint32_t i = 0; i < n_inf; ++i) perm[i] = i;
    std::sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) {
        return out->inf_rec[a] != out->inf_rec[b] ? out->inf_rec[a] < out->inf_rec[b] : out->inf_off[a] < out->inf_off[b];
    });
    std::vector<uint32_t> r2(n_inf), o2(n_inf); std::vector<uint64_t> k2(n_inf);
    for (uint64_t i = 0; i < n_inf; ++i) { r2[i] = out->inf_rec[perm[i]]; o2[i] = out->inf_off[perm[i]]; k2[i] = out->inf_kmer[perm[i]]; }
    if (n_inf) { memcpy(out->inf_rec, r2.data(), n_inf * 4); memcpy(out->inf_off, o2.data(), n_inf * 4); memcpy(out->inf_kmer, k2.data(), n_inf * 8); }
    return 0;
}

extern "C" void s2_ingest_detect_free(s2_ingest_detect_result *r)
{
    if (!r) return;
    free(r->len); free(r->hits); free(r->inf); free(r->inf_rec); free(r->inf_off); free(r->inf_kmer);
    memset(r, 0, sizeof *r);
}

extern "C" void s2_ingest_thread_cleanup(void)
{
    if (tl_ingest) { ingest_free(tl_ingest); tl_ingest = nullptr; }
}
