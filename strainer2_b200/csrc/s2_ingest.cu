// s2_ingest.cu - GPU-side ingest of FASTA / FASTQ files: hardware DEFLATE + record splitting on the device.
//
// SURVEY 8(f) rank 1.  The reference inflates and parses every input on one CPU thread (zlib gzread +
// the vendored line parser, /root/reference/src/genome_compare.c:194-203, src/kseq.h:171-211); zlib gives
// about 0.34 GB/s of text per core, three orders of magnitude below the scan kernel.  Blackwell has a
// hardware decompression engine, reachable through the CUDA driver's batch-decompress entry point: it inflates
// independent raw-DEFLATE streams of up to 4 MiB each (measured here: 240-320 GB/s of text for batches of
// 64 KB blocks).  A plain .gz file is ONE long stream and cannot be split, but BGZF (bgzip, the block-
// gzip flavour htslib writes; still a valid multi-member gzip file for the reference's zlib) is a sequence
// of independent <= 64 KB members whose sizes are in their headers.  For BGZF files - and for uncompressed
// text - this file does on the GPU what the reader threads do for everything else:
//
//   host   : hand the compressed bytes to the copy engine (from the caller's memory, or read() into pinned
//            staging), walk the BGZF headers (no inflate)
//   engine : inflate all blocks of a chunk into one contiguous text buffer
//   kernels: index the newlines (per 16 KB text block: count -> scan over blocks -> scatter), check that the text is
//            strict 4-line FASTQ / strict FASTA and measure it, copy the sequence lines into the flat batch format
//            (sequence bytes + '\n'; output offsets = scan over blocks + a scan inside each block), carry the partial
//            last record to the next chunk
//   scan   : the normal count / detect kernel; batch length, veto and increment are read from device memory
//
// A chunk is either a slice of one big file or a GROUP of whole small files whose texts lie back to back (2,000
// genomes of 5 Mb are 2,000 x 1.5 MB of BGZF: one launch sequence per file would be all latency).  Chunks travel
// through a three-slot ring on three streams, one per engine: the host -> device copy of chunk i+2, the inflate of
// chunk i+1 and the kernels of chunk i run at the same time.  Steps that need every block's result (the scans over
// the blocks, the chunk's verdict and carry) run in the last block to finish of the preceding kernel, so a chunk
// is five launches.
//
// Parity: the parser the reference vendors accepts many irregular layouts (multi-line FASTQ, CR LF, '>'
// records mixed in, truncated last record ...).  The kernels do not emulate those; they PROVE that a chunk is
// regular before its scan kernel starts (the scan reads the verdict from device memory and counts nothing after
// the first irregular chunk).  A file that fits one chunk is therefore either counted completely or not at all;
// a streamed file that turns irregular after some chunks were counted is replayed once with increment -1
// (uint32 wrap-around add: the counters return to their exact previous values).  Irregular files are handed back
// to the host reader (return value 1) with the counters untouched.
#include "s2_private.h"
#include "s2_kmer.cuh"
#include "s2_inflate.cuh"
#include "s2_gunzip.h"

#include <cuda.h>
#include <fcntl.h>
#include <zlib.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <thread>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <set>
#include <string>

#define ING_THREADS 256
#define ING_TILE 4096u                          /* bytes one pass of a CTA covers: 256 threads x 16 B */
#define ING_TILES 4u
#define ING_BLOCK (ING_TILE * ING_TILES)        /* text bytes per CTA of the indexing kernels */
#define ING_MAXCARRY (1u << 20)                 /* longest partial record carried between chunks (multiple of ING_BLOCK) */
#define ING_MAX_DBLOCKS 32768u                  /* DEFLATE blocks per chunk */
#define ING_MAX_FILES 4096u                     /* whole files grouped into one chunk */
#define ING_MAX_RESULTS 4096u                   /* chunks-with-a-verdict in flight between two host syncs */
#define ING_DECOMP_CALL 4096u                   /* blocks per driver call */
#define ING_CAP_C (8ull << 20)                  /* informative windows one chunk may report */

typedef unsigned long long ull;

struct IngState {                 // lives in device memory, one per ingest pipeline
    ull t0, t1;                   // current text range inside the text buffer
    ull flat_len;                 // \  S2DevBatch: bytes of the flat batch produced from this chunk,
    unsigned int skip;            //  | veto (the text seen so far is irregular),
    unsigned int inc;             // /  increment (1 or -1)
    ull carry_from, carry_len;    // bytes after the last complete record
    ull bases, lookups, records;  // totals over the file / group
    unsigned int n_lines, n_rec;
    unsigned int irregular;       // sticky: not strict FASTQ / FASTA
    unsigned int inf_overflow;    // a chunk produced more informative windows / records than the lists hold
    unsigned int last_chunk, first_chunk;
    unsigned int tail_len;        // FASTA: last bytes of the previous chunk's flat stream, re-scanned in front of this one
    unsigned int open_kind;       // FASTA, streamed: what this chunk ends in - 0 a line start, 1 inside a sequence line, 2 inside a header
    unsigned int open_kind_in;    // ... and what the previous chunk ended in (= what this chunk's first line continues)
    unsigned char tail[32];
};

static_assert(offsetof(IngState, skip) - offsetof(IngState, flat_len) == offsetof(S2DevBatch, skip) &&
              offsetof(IngState, inc) - offsetof(IngState, flat_len) == offsetof(S2DevBatch, inc) && offsetof(S2DevBatch, n_bytes) == 0,
              "the scan kernels read IngState::flat_len / skip / inc as an S2DevBatch");

struct IngResult { ull bases, lookups, records; unsigned int irregular, inf_overflow; };

struct IngChunkArgs {             // by value to the kernels of one chunk
    ull new_bytes;                // text bytes that arrived behind the carry
    unsigned int first_chunk, last_chunk, inc, fasta;
    unsigned int lsh;             // record formats: log2 of the lines per record - 2 strict FASTQ, 1 two-line FASTA reads (strain_detect)
    unsigned int n_files, n_dblocks;
    const ull *file_end;          // group: text offset (relative to the chunk's new bytes) where file i ends
    const unsigned int *isz;      // expected inflated size per DEFLATE block
    const unsigned int *act;      // what the engine reported
};

// ------------------------------------------------------------------------------------------------
// block-level helpers
// ------------------------------------------------------------------------------------------------
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

__device__ __forceinline__ ull ing_block_sum(ull v)
{
    __shared__ ull warp_sums64[ING_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
    if (lane == 0) warp_sums64[wid] = v;
    __syncthreads();
    ull tot = 0;
#pragma unroll
    for (int i = 0; i < ING_THREADS / 32; ++i) tot += warp_sums64[i];
    __syncthreads();
    return tot;
}

// 4 bits: which bytes of w equal c
__device__ __forceinline__ unsigned eq_bytes4(unsigned w, unsigned c)
{
    const unsigned x = w ^ (c * 0x01010101u);
    const unsigned y = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);      // 0x80 in every zero byte of x, exact
    return ((((y >> 7) & 0x01010101u) * 0x01020408u) >> 24) & 0xFu;
}

// nonzero iff some byte of w equals c (no positions: the borrow trick is exact for "any")
__device__ __forceinline__ unsigned any_byte4(unsigned w, unsigned c)
{
    const unsigned x = w ^ (c * 0x01010101u);
    return (x - 0x01010101u) & ~x & 0x80808080u;
}

// bit i set <=> text[base+i] == '\n' and t0 <= base+i < t1   (base is 16-byte aligned); *cr: a '\r' in that range
__device__ __forceinline__ unsigned nl_mask16(const uint8_t *text, ull base, ull t0, ull t1, unsigned *cr)
{
    *cr = 0;
    if (base + 16 <= t0 || base >= t1) return 0;
    const uint4 v = *reinterpret_cast<const uint4 *>(text + base);
    const unsigned nl = eq_bytes4(v.x, '\n') | (eq_bytes4(v.y, '\n') << 4) | (eq_bytes4(v.z, '\n') << 8) | (eq_bytes4(v.w, '\n') << 12);
    if (base >= t0 && base + 16 <= t1) {                                          // the common case: all 16 bytes are text
        *cr = (any_byte4(v.x, '\r') | any_byte4(v.y, '\r') | any_byte4(v.z, '\r') | any_byte4(v.w, '\r')) ? 1u : 0u;
        return nl;
    }
    const unsigned lo = t0 > base ? (unsigned)(t0 - base) : 0u;                  // < 16 here
    const unsigned hi = t1 - base >= 16 ? 16u : (unsigned)(t1 - base);
    const unsigned range = ((1u << hi) - 1u) & ~((1u << lo) - 1u);
    const unsigned r = eq_bytes4(v.x, '\r') | (eq_bytes4(v.y, '\r') << 4) | (eq_bytes4(v.z, '\r') << 8) | (eq_bytes4(v.w, '\r') << 12);
    *cr = (r & range) ? 1u : 0u;
    return nl & range;
}

// The per-block kernels end with a step that needs every block's result (the scan over the blocks' counts, the
// carry and verdict of the chunk): the block that finishes LAST does it (fence, then a ticket), which saves a
// kernel boundary per step - a chunk is a chain of dependent launches, so each boundary is pure latency.
__device__ __forceinline__ bool ing_last_block(unsigned *ticket)
{
    __shared__ bool last;
    __threadfence();                       // this thread's results before the ticket
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(ticket, 1u);
        last = t == gridDim.x - 1;
        if (last) *ticket = 0;             // ready for the next launch
    }
    __syncthreads();
    if (last) __threadfence();
    return last;
}

// single CTA: exclusive scan of v[0..n) in place, v[n] = total -> returned to every thread (reads bypass L1: the
// values were written by other blocks of the same launch)
__device__ __forceinline__ unsigned ing_scan_array(unsigned *v, unsigned n)
{
    __shared__ unsigned carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (unsigned b0 = 0; b0 < n; b0 += ING_THREADS * 4) {
        unsigned x[4], sum = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) { const unsigned i = b0 + threadIdx.x * 4 + k; x[k] = i < n ? __ldcg(v + i) : 0; sum += x[k]; }
        unsigned total;
        unsigned ex = ing_block_scan(sum, total) + carry_s;
#pragma unroll
        for (int k = 0; k < 4; ++k) { const unsigned i = b0 + threadIdx.x * 4 + k; if (i < n) v[i] = ex; ex += x[k]; }
        __syncthreads();
        if (threadIdx.x == 0) carry_s += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) v[n] = carry_s;
    return carry_s;
}

// ------------------------------------------------------------------------------------------------
// newline index: count per 16 KB block -> scan over blocks -> scatter
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ING_THREADS) ing_index_count(uint8_t *text, IngState *st, unsigned *block_nl,
                                                                unsigned short *__restrict__ masks, IngChunkArgs a, unsigned max_lines, unsigned *ticket)
{
    const ull carry = a.first_chunk ? 0ull : st->carry_len;
    const ull t0 = ING_MAXCARRY - carry, t1 = ING_MAXCARRY + a.new_bytes;
    // a final line without '\n' is completed (the parser the reference uses accepts that; the buffer has room past t1).
    // FASTA carries no bytes between chunks: a chunk of a streamed file always ends its last line here - the line's
    // bytes are emitted with this chunk and st->open_kind tells the next chunk what its first line continues - so
    // lines may be as long as they like.
    const bool term = a.last_chunk ? (t1 > t0 && text[t1 - 1] != '\n') : (a.fasta != 0);
    unsigned n = 0, any_cr = 0;
#pragma unroll
    for (unsigned k = 0; k < ING_TILES; ++k) {
        const ull base = (ull)blockIdx.x * ING_BLOCK + k * ING_TILE + threadIdx.x * 16;
        unsigned cr;
        unsigned m = nl_mask16(text, base, t0, t1, &cr);
        if (term && t1 >= base && t1 < base + 16) { m |= 1u << (unsigned)(t1 - base); text[t1] = '\n'; }
        masks[base >> 4] = (unsigned short)m;                      // the scatter pass reads these 2 bytes instead of the 16 of text
        n += __popc(m);
        any_cr |= cr;
    }
    if (any_cr) atomicOr(&st->irregular, 1u);
    unsigned total;
    ing_block_scan(n, total);
    if (threadIdx.x == 0) block_nl[blockIdx.x] = total;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->t0 = t0; st->t1 = t1 + (term ? 1 : 0);
        st->first_chunk = a.first_chunk; st->last_chunk = a.last_chunk;
        st->inc = a.inc; st->flat_len = 0;
        if (a.first_chunk) { st->tail_len = 0; st->open_kind = 0; }
        st->open_kind_in = a.first_chunk ? 0u : st->open_kind;
    }
    if (!ing_last_block(ticket)) return;
    // the last block: counts -> offsets; the total is the line count of the chunk
    unsigned lines = ing_scan_array(block_nl, gridDim.x);
    if (threadIdx.x == 0) {
        if (lines > max_lines) { atomicOr(&st->irregular, 1u); lines = 0; }       // shorter average lines than 8 bytes: not a sequence file
        st->n_lines = lines;
        st->n_rec = a.fasta ? 0u : lines >> a.lsh;
    }
}

__global__ void __launch_bounds__(ING_THREADS) ing_index_scatter(const unsigned short *__restrict__ masks, const unsigned *__restrict__ block_off,
                                                                  unsigned *__restrict__ line_end, unsigned max_lines)
{
    unsigned r = block_off[blockIdx.x];
    unsigned m[ING_TILES];
#pragma unroll
    for (unsigned k = 0; k < ING_TILES; ++k) m[k] = masks[((ull)blockIdx.x * ING_BLOCK + k * ING_TILE) / 16 + threadIdx.x];
#pragma unroll
    for (unsigned k = 0; k < ING_TILES; ++k) {
        const ull base = (ull)blockIdx.x * ING_BLOCK + k * ING_TILE + threadIdx.x * 16;
        unsigned total;
        unsigned at = r + ing_block_scan(__popc(m[k]), total);
        unsigned mm = m[k];
        while (mm) {
            const int b = __ffs(mm) - 1;
            mm &= mm - 1;
            if (at < max_lines) line_end[at] = (unsigned)(base + b);       // the text buffer is < 4 GiB
            ++at;
        }
        r += total;
    }
}

// ------------------------------------------------------------------------------------------------
// checks that do not belong to one record: what the engine inflated, where the files of a group meet
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ing_check_chunk(const uint8_t *__restrict__ text, IngState *st, const unsigned *__restrict__ line_end, const IngChunkArgs &a)
{
    const unsigned gtid = blockIdx.x * blockDim.x + threadIdx.x, n_threads = gridDim.x * blockDim.x;
    for (unsigned i = gtid; i < a.n_dblocks; i += n_threads)
        if (a.act[i] != a.isz[i]) atomicOr(&st->irregular, 1u);                       // damaged DEFLATE block
    // every file of a group but the last must end with '\n' (only the chunk's last line can be completed), and in
    // FASTQ on a record boundary; that the next file starts with '>' / '@' was checked by the host
    const unsigned n_lines = st->n_lines;
    for (unsigned i = gtid; i + 1 < a.n_files; i += n_threads) {
        const unsigned p = (unsigned)(ING_MAXCARRY + a.file_end[i] - 1);
        bool ok = text[p] == '\n';
        if (ok && !a.fasta) {
            unsigned lo = 0, hi = n_lines;                                            // first line_end >= p
            while (lo < hi) { const unsigned mid = (lo + hi) >> 1; if (line_end[mid] < p) lo = mid + 1; else hi = mid; }
            ok = lo < n_lines && line_end[lo] == p && ((lo + 1) & ((1u << a.lsh) - 1u)) == 0;
        }
        if (!ok) atomicOr(&st->irregular, 1u);
    }
}

// the chunk's verdict (everything that can veto the scan is known once all blocks are measured), the carry, and the
// batch length the scan kernel will read.  Runs in the last block of the measure kernels, after the scan over the
// blocks' output sizes.
// kind of FASTA line L: 1 sequence, 2 header (3 = a header that begins here: counts as a record and emits the separator)
__device__ __forceinline__ unsigned ing_fasta_kind(unsigned L, unsigned len, uint8_t first, unsigned open_prev)
{
    if (L == 0 && open_prev) return open_prev;                               // continues the previous chunk's last line
    return len && first == '>' ? 3u : 1u;
}

__device__ __forceinline__ void ing_chunk_verdict(const uint8_t *text, IngState *st, const unsigned *line_end, unsigned total, unsigned fasta, unsigned lsh, ull *rec_off)
{
    if (fasta && !st->last_chunk) {                                          // what does the next chunk's first line continue?
        const ull t1 = st->t1 - 1;                                           // (the chunk's text without the line end added at its end)
        const unsigned n_lines = st->n_lines;
        if (t1 > st->t0 && n_lines) {
            if (text[t1 - 1] == '\n') st->open_kind = 0;
            else {
                const unsigned L = n_lines - 1;
                const ull s0 = L ? (ull)__ldcg(line_end + L - 1) + 1 : st->t0;
                const unsigned k = ing_fasta_kind(L, (unsigned)(t1 - s0), text[s0], st->open_kind_in);
                st->open_kind = k == 1 ? 1u : 2u;
            }
        }
    }
    const unsigned irregular = atomicOr(&st->irregular, 0u);                 // what every block reported (L2)
    const unsigned n_done = fasta ? st->n_lines : st->n_rec << lsh;          // lines that belong to complete records
    const ull c_from = n_done ? (ull)__ldcg(line_end + n_done - 1) + 1 : st->t0;
    ull c_len = st->t1 - c_from;
    unsigned bad = irregular;
    if (c_len > ING_MAXCARRY) { bad = 1; c_len = 0; }
    if (st->last_chunk && c_len) { bad = 1; c_len = 0; }                     // truncated last record: the host parser's business
    if (bad) st->irregular = 1;
    st->carry_from = c_from; st->carry_len = c_len;
    st->flat_len = (fasta ? st->tail_len : 0u) + (ull)total;
    st->skip = bad;
    if (rec_off) rec_off[st->n_rec] = st->flat_len;
}

// ------------------------------------------------------------------------------------------------
// strict FASTQ: validate + measure per block, copy
// ------------------------------------------------------------------------------------------------
// The records of block b are those whose LAST line ends in it: lines [off[b], off[b+1]) -> records [off[b]/4, off[b+1]/4).
__global__ void __launch_bounds__(ING_THREADS) ing_fastq_measure(const uint8_t *__restrict__ text, IngState *st, const unsigned *__restrict__ block_off,
                                                                  const unsigned *__restrict__ line_end, unsigned *block_out, IngChunkArgs a,
                                                                  ull *rec_off, unsigned *ticket)
{
    ing_check_chunk(text, st, line_end, a);
    const unsigned n_lines = st->n_lines;
    const unsigned lsh = a.lsh, lmask = (1u << lsh) - 1u;
    const unsigned r_lo = min(block_off[blockIdx.x], n_lines) >> lsh, r_hi = min(block_off[blockIdx.x + 1], n_lines) >> lsh;
    const ull t0 = st->t0;
    ull bases = 0, lookups = 0, out = 0;
    bool bad = false;
    for (unsigned r = r_lo + threadIdx.x; r < r_hi; r += ING_THREADS) {
        const unsigned L0 = r << lsh;
        const ull h0 = r ? (ull)line_end[L0 - 1] + 1 : t0;         // '@' / '>' line
        const ull s0 = (ull)line_end[L0] + 1;                       // sequence line
        const ull p0 = (ull)line_end[L0 + 1] + 1;                   // FASTQ: '+' line; two-line FASTA: the next record
        const ull len = p0 - 1 - s0;
        bool ok;
        if (lsh == 2) {
            const ull q0 = (ull)line_end[L0 + 2] + 1;               // quality line
            const ull e0 = (ull)line_end[L0 + 3];
            ok = text[h0] == '@' && text[p0] == '+' && len == e0 - q0;
        } else {
            ok = text[h0] == '>';                                   // (that the NEXT line is a header again is the next record's check)
        }
        // a sequence line that begins with one of these would make the reference's parser see another record / the quality part
        if (len) { const uint8_t c = text[s0]; ok = ok && c != '>' && c != '+' && c != '@'; }
        bad |= !ok;
        bases += len;
        if (len >= S2_K) lookups += len - (S2_K - 1);
    }
    // output bytes are accounted to the block in which the SEQUENCE line ends (the copy kernel's ownership rule)
    {
        const unsigned l_lo = min(block_off[blockIdx.x], n_lines), l_hi = min(block_off[blockIdx.x + 1], n_lines), n_rec = st->n_rec;
        for (unsigned L = l_lo + threadIdx.x; L < l_hi; L += ING_THREADS)
            if ((L & lmask) == 1u && (L >> lsh) < n_rec) {
                const unsigned len = line_end[L] - (line_end[L - 1] + 1);
                if (len >= S2_K) out += len + 1;                    // records without a window are not copied (genome_compare.c:204)
            }
    }
    if (bad) atomicOr(&st->irregular, 1u);
    bases = ing_block_sum(bases); lookups = ing_block_sum(lookups); out = ing_block_sum(out);
    if (threadIdx.x == 0) {
        block_out[blockIdx.x] = (unsigned)out;
        if (bases) atomicAdd(&st->bases, bases);
        if (lookups) atomicAdd(&st->lookups, lookups);
    }
    if (!ing_last_block(ticket)) return;
    const unsigned total = ing_scan_array(block_out, gridDim.x);
    if (threadIdx.x == 0) ing_chunk_verdict(text, st, line_end, total, 0u, a.lsh, rec_off);
}

// ------------------------------------------------------------------------------------------------
// end of chunk (one block): copy the partial last record in front of where the next chunk's text lands (the next chunk
// of a streamed file is inflated into ANOTHER buffer of the ring, perhaps already, behind its carry region), FASTA
// tail, record count, the verdict of a finished file / group, and a clean state for the next one
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ing_finish_body(const uint8_t *text, uint8_t *text_next, IngState *st, const uint8_t *flat, unsigned fasta, IngResult *res)
{
    const ull c_from = st->carry_from, c_len = st->carry_len, dst = ING_MAXCARRY - c_len;
    for (ull i = threadIdx.x; i < c_len; i += ING_THREADS) text_next[dst + i] = text[c_from + i];
    if (threadIdx.x == 0) {
        if (!fasta) st->records += st->n_rec;
        if (fasta && !st->skip) {
            const ull fl = st->flat_len;
            const unsigned keep = fl < (S2_K - 1) ? (unsigned)fl : (S2_K - 1);
            unsigned char tmp[32];
            for (unsigned i = 0; i < keep; ++i) tmp[i] = __ldcg(flat + fl - keep + i);
            for (unsigned i = 0; i < keep; ++i) st->tail[i] = tmp[i];
            st->tail_len = keep;
        }
        if (res) { res->bases = st->bases; res->lookups = st->lookups; res->records = st->records; res->irregular = st->irregular; res->inf_overflow = st->inf_overflow; }
        if (st->last_chunk) { st->bases = 0; st->lookups = 0; st->records = 0; st->irregular = 0; st->inf_overflow = 0; }      // (skip stays: the scan reads it)
    }
}

// The copy kernels work by BYTE OWNERSHIP: a block copies the sequence bytes that lie in its own 16 KB of text,
// whichever line they belong to - lines that end in the block (output offset = the block's offset + a scan inside
// the block), the head part of a line that started in an earlier block, and the part of a line that is still open at
// the block's end (such a line is the first one to end in its last block, so its output offset is that block's
// offset).  Long lines - unwrapped genomes, long reads - are thereby spread over as many blocks as they span.  The
// block's text is staged in shared memory with 128-bit loads; pieces up to 512 bytes are copied by one warp each,
// longer ones by the whole block.
#define ING_BIG_PIECE 512u
#define ING_MAX_BIG 40
struct IngPieces {                      // shared-memory work list of one block
    unsigned src[ING_THREADS], len[ING_THREADS], dst[ING_THREADS];     // this round's small pieces (stage index, bytes, flat offset)
    unsigned big_src[ING_MAX_BIG], big_len[ING_MAX_BIG], big_dst[ING_MAX_BIG];
    unsigned n_big;
};

__device__ __forceinline__ void ing_stage_block(const uint8_t *__restrict__ text, uint8_t *stage)
{
    const ull b0 = (ull)blockIdx.x * ING_BLOCK;
    constexpr int N = ING_BLOCK / (ING_THREADS * 16);        // 4 loads in flight per thread
    uint4 v[N];
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] = __ldg(reinterpret_cast<const uint4 *>(text + b0 + (ull)k * ING_THREADS * 16 + threadIdx.x * 16));
#pragma unroll
    for (int k = 0; k < N; ++k) *reinterpret_cast<uint4 *>(stage + (ull)k * ING_THREADS * 16 + threadIdx.x * 16) = v[k];
}

// a piece [lo, hi) of text (inside this block) that goes to flat + dst: small ones into this thread's slot of the round,
// big ones onto the block's list
__device__ __forceinline__ void ing_add_piece(IngPieces &w, unsigned lo, unsigned hi, unsigned dst, bool to_slot)
{
    const unsigned b0 = blockIdx.x * ING_BLOCK;
    unsigned n = hi > lo ? hi - lo : 0u;
    if (n > ING_BIG_PIECE || (!to_slot && n)) {
        const unsigned at = atomicAdd(&w.n_big, 1u);
        if (at < ING_MAX_BIG) { w.big_src[at] = lo - b0; w.big_len[at] = n; w.big_dst[at] = dst; }
        n = 0;
    }
    if (to_slot) { w.src[threadIdx.x] = lo - b0; w.len[threadIdx.x] = n; w.dst[threadIdx.x] = dst; }
}

__device__ __forceinline__ void ing_copy_small(const IngPieces &w, const uint8_t *stage, uint8_t *__restrict__ flat, unsigned cnt)
{
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (unsigned e = wid; e < cnt; e += ING_THREADS / 32) {
        const unsigned n = w.len[e];
        const uint8_t *sp = stage + w.src[e] + lane;                 // one address pair per piece; the passes use immediate offsets
        uint8_t *dp = flat + w.dst[e] + lane;
        for (unsigned off = 0; off < n; off += 128, sp += 128, dp += 128) {
            const unsigned r = n - off;
            uint8_t b0 = 0, b1 = 0, b2 = 0, b3 = 0;
            if (lane < r) b0 = sp[0];
            if (lane + 32 < r) b1 = sp[32];
            if (lane + 64 < r) b2 = sp[64];
            if (lane + 96 < r) b3 = sp[96];
            if (lane < r) dp[0] = b0;
            if (lane + 32 < r) dp[32] = b1;
            if (lane + 64 < r) dp[64] = b2;
            if (lane + 96 < r) dp[96] = b3;
        }
    }
}

__device__ __forceinline__ void ing_copy_big(const IngPieces &w, const uint8_t *stage, uint8_t *__restrict__ flat)
{
    const unsigned nb = min(w.n_big, (unsigned)ING_MAX_BIG);
    for (unsigned e = 0; e < nb; ++e) {
        const unsigned n = w.big_len[e];
        const uint8_t *sp = stage + w.big_src[e];
        uint8_t *dp = flat + w.big_dst[e];
        for (unsigned i = threadIdx.x; i < n; i += ING_THREADS) dp[i] = sp[i];
    }
}

__global__ void __launch_bounds__(ING_THREADS) ing_fastq_copy(const uint8_t *text, uint8_t *text_next, IngState *st, const unsigned *__restrict__ block_off,
                                                               const unsigned *__restrict__ block_out, const unsigned *__restrict__ line_end,
                                                               uint8_t *flat, ull *__restrict__ rec_off, unsigned do_finish, unsigned lsh, IngResult *res, unsigned *ticket)
{
    __shared__ IngPieces w;
    __shared__ __align__(16) uint8_t stage[ING_BLOCK];
    const unsigned n_lines = st->n_lines, n_rec = st->n_rec;
    const unsigned l_lo = min(block_off[blockIdx.x], n_lines), l_hi = min(block_off[blockIdx.x + 1], n_lines);
    const unsigned b0 = blockIdx.x * ING_BLOCK, b1 = b0 + ING_BLOCK;
    // is a sequence line of a complete record still open at the end of this block?
    const unsigned lmask = (1u << lsh) - 1u;
    const bool open_seq = l_hi < n_lines && (l_hi & lmask) == 1u && (l_hi >> lsh) < n_rec;
    if (!st->skip && (l_lo < l_hi || open_seq)) {
        if (threadIdx.x == 0) w.n_big = 0;
        ing_stage_block(text, stage);
        __syncthreads();
        unsigned run = block_out[blockIdx.x];
        for (unsigned l0 = l_lo; l0 < l_hi; l0 += ING_THREADS) {
            const unsigned L = l0 + threadIdx.x;
            unsigned s0 = 0, e1 = 0, olen = 0;
            const bool seq = L < l_hi && (L & lmask) == 1u && (L >> lsh) < n_rec;
            if (seq) {
                s0 = line_end[L - 1] + 1;
                e1 = line_end[L] + 1;                                // one past the line's own '\n' = the separator
                olen = e1 - 1 - s0 >= S2_K ? e1 - s0 : 0;           // records without a window are not copied
            }
            unsigned total;
            const unsigned at = run + ing_block_scan(olen, total);
            if (seq && rec_off) rec_off[L >> lsh] = at;
            const unsigned lo = max(s0, b0);                         // the line may have started in an earlier block
            ing_add_piece(w, lo, olen ? e1 : lo, at + (lo - s0), true);
            __syncthreads();
            ing_copy_small(w, stage, flat, min((unsigned)ING_THREADS, l_hi - l0));
            __syncthreads();
            run += total;
        }
        if (open_seq && threadIdx.x == 0) {
            const unsigned s0 = line_end[l_hi - 1] + 1, len = line_end[l_hi] - s0;
            if (s0 < b1 && len >= S2_K) {
                const unsigned lo = max(s0, b0);
                ing_add_piece(w, lo, b1, block_out[line_end[l_hi] / ING_BLOCK] + (lo - s0), false);
            }
        }
        __syncthreads();
        ing_copy_big(w, stage, flat);
    }
    if (!do_finish || !ing_last_block(ticket)) return;
    ing_finish_body(text, text_next, st, flat, 0u, res);
}

// ------------------------------------------------------------------------------------------------
// strict FASTA: '>' lines are headers, every other line is sequence (joined), nothing else
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ING_THREADS) ing_fasta_measure(const uint8_t *__restrict__ text, IngState *st, const unsigned *__restrict__ block_off,
                                                                  const unsigned *__restrict__ line_end, unsigned *block_out, IngChunkArgs a, unsigned *ticket)
{
    ing_check_chunk(text, st, line_end, a);
    const unsigned n_lines = st->n_lines;
    const unsigned l_lo = min(block_off[blockIdx.x], n_lines), l_hi = min(block_off[blockIdx.x + 1], n_lines);
    const ull t0 = st->t0;
    const unsigned open_prev = st->open_kind_in;
    ull bases = 0, recs = 0, out = 0;
    bool bad = false;
    for (unsigned L = l_lo + threadIdx.x; L < l_hi; L += ING_THREADS) {
        const ull s0 = L ? (ull)line_end[L - 1] + 1 : t0;
        const unsigned len = line_end[L] - (unsigned)s0;
        const uint8_t first = len ? text[s0] : 0;
        const unsigned kind = ing_fasta_kind(L, len, first, open_prev);
        const bool line_start = !(L == 0 && open_prev);
        if (line_start && (first == '@' || first == '+')) bad = true;              // the reference's parser would switch to FASTQ rules
        if (L == 0 && a.first_chunk && kind != 3) bad = true;                      // text before the first record
        out += kind == 3 ? 1u : kind == 1 ? len : 0u;                              // a header becomes the record separator
        if (kind == 3) ++recs; else if (kind == 1) bases += len;
    }
    if (bad) atomicOr(&st->irregular, 1u);
    bases = ing_block_sum(bases); recs = ing_block_sum(recs); out = ing_block_sum(out);
    if (threadIdx.x == 0) {
        block_out[blockIdx.x] = (unsigned)out;
        if (bases) atomicAdd(&st->bases, bases);
        if (recs) atomicAdd(&st->records, recs);
    }
    if (!ing_last_block(ticket)) return;
    const unsigned total = ing_scan_array(block_out, gridDim.x);
    if (threadIdx.x == 0) ing_chunk_verdict(text, st, line_end, total, 1u, 0u, nullptr);
}

__global__ void __launch_bounds__(ING_THREADS) ing_fasta_copy(const uint8_t *text, uint8_t *text_next, IngState *st, const unsigned *__restrict__ block_off,
                                                               const unsigned *__restrict__ block_out, const unsigned *__restrict__ line_end,
                                                               uint8_t *flat, IngResult *res, unsigned *ticket)
{
    __shared__ IngPieces w;
    __shared__ __align__(16) uint8_t stage[ING_BLOCK];
    const unsigned n_lines = st->n_lines, tail = st->tail_len;
    const unsigned l_lo = min(block_off[blockIdx.x], n_lines), l_hi = min(block_off[blockIdx.x + 1], n_lines);
    const unsigned b0 = blockIdx.x * ING_BLOCK, b1 = b0 + ING_BLOCK;
    const bool skip = st->skip != 0;
    if (!skip && blockIdx.x == 0 && threadIdx.x < tail) flat[threadIdx.x] = st->tail[threadIdx.x];
    if (!skip && (l_lo < l_hi || l_hi < n_lines)) {
        const unsigned t0 = (unsigned)st->t0;
        const unsigned open_prev = st->open_kind_in;
        if (threadIdx.x == 0) w.n_big = 0;
        ing_stage_block(text, stage);
        __syncthreads();
        unsigned run = tail + block_out[blockIdx.x];
        for (unsigned l0 = l_lo; l0 < l_hi; l0 += ING_THREADS) {
            const unsigned L = l0 + threadIdx.x;
            unsigned s0 = 0, e = 0, olen = 0;
            bool header = false;
            if (L < l_hi) {
                s0 = L ? line_end[L - 1] + 1 : t0;
                e = line_end[L];
                olen = e - s0;
                const unsigned kind = ing_fasta_kind(L, olen, olen ? (s0 >= b0 ? stage[s0 - b0] : text[s0]) : 0, open_prev);
                header = kind != 1;
                if (header) olen = kind == 3 ? 1u : 0u;              // a header becomes the record separator, once
            }
            unsigned total;
            const unsigned at = run + ing_block_scan(olen, total);
            if (header && olen) flat[at] = '\n';
            const unsigned lo = max(s0, b0);                         // the line may have started in an earlier block
            ing_add_piece(w, lo, header ? lo : e, at + (lo - s0), true);
            __syncthreads();
            ing_copy_small(w, stage, flat, min((unsigned)ING_THREADS, l_hi - l0));
            __syncthreads();
            run += total;
        }
        if (l_hi < n_lines && threadIdx.x == 0) {                   // the line that is still open at the end of this block
            const unsigned s0 = l_hi ? line_end[l_hi - 1] + 1 : t0;
            if (s0 < b1 && line_end[l_hi] > s0 && ing_fasta_kind(l_hi, 1u, s0 >= b0 ? stage[s0 - b0] : text[s0], open_prev) == 1) {
                const unsigned lo = max(s0, b0);
                ing_add_piece(w, lo, b1, tail + block_out[line_end[l_hi] / ING_BLOCK] + (lo - s0), false);
            }
        }
        __syncthreads();
        ing_copy_big(w, stage, flat);
    }
    if (!ing_last_block(ticket)) return;
    ing_finish_body(text, text_next, st, flat, 1u, res);
}

// A BGZF member whose CRC-32 does not match (checked beside the kernels above, on another stream): the chunk is not
// scanned and the file counts as irregular.  `res` was already written by the copy kernel's last block (count mode);
// a file that goes on (not the last chunk) stays irregular for its later chunks, a finished one has left a clean state.
__global__ void ing_crc_veto(IngState *st, const uint32_t *bad, IngResult *res, unsigned last_chunk)
{
    if (!*bad) return;
    st->skip = 1;
    if (res) res->irregular = 1;
    if (!last_chunk || !res) st->irregular = 1;          // (detect mode: ing_finish reads and clears it after the scan)
}

// detect mode: the per-record results are stored after the scan, so the end-of-chunk step is a launch of its own
__global__ void __launch_bounds__(ING_THREADS) ing_finish(const uint8_t *text, uint8_t *text_next, IngState *st, const uint8_t *flat, unsigned fasta, IngResult *res)
{
    ing_finish_body(text, text_next, st, flat, fasta, res);
}

// ---- detect mode (strain_detect pass 1 on ingested reads) ------------------------------------------
// per-record results of this chunk -> the file-level arrays (record numbering continues across chunks)
__global__ void __launch_bounds__(ING_THREADS) ing_store_records(IngState *st, const unsigned *__restrict__ line_end, const unsigned *__restrict__ hits_c,
                                                                  const unsigned *__restrict__ inf_c, unsigned *__restrict__ len_all,
                                                                  unsigned *__restrict__ hits_all, unsigned *__restrict__ inf_all, ull cap, unsigned lsh)
{
    if (st->skip) return;
    const unsigned n_rec = st->n_rec;
    const ull base = st->records;
    if (base + n_rec > cap) { if (blockIdx.x == 0 && threadIdx.x == 0) st->inf_overflow = 1; return; }
    for (unsigned r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rec; r += gridDim.x * blockDim.x) {
        len_all[base + r] = line_end[(r << lsh) + 1] - (line_end[r << lsh] + 1);
        hits_all[base + r] = hits_c[r];
        inf_all[base + r] = inf_c[r];
    }
}

// informative windows of this chunk (byte offsets in the flat batch) -> (record number, offset, canonical k-mer)
__global__ void __launch_bounds__(ING_THREADS) ing_collect_inf(IngState *st, const uint8_t *__restrict__ flat, const ull *__restrict__ rec_off,
                                                                const ull *__restrict__ pos_c, const ull *__restrict__ cnt_c,
                                                                ull cap_c, unsigned *__restrict__ f_rec, unsigned *__restrict__ f_off,
                                                                ull *__restrict__ f_kmer, ull *__restrict__ f_cnt, ull cap_f)
{
    if (st->skip) return;
    const ull n = *cnt_c;
    if (n > cap_c) { if (blockIdx.x == 0 && threadIdx.x == 0) st->inf_overflow = 1; }
    const unsigned n_rec = st->n_rec;
    const ull base = st->records;
    for (ull i = (ull)blockIdx.x * blockDim.x + threadIdx.x; i < n && i < cap_c; i += (ull)gridDim.x * blockDim.x) {
        const ull pos = pos_c[i];
        unsigned lo = 0, hi = n_rec;
        while (hi - lo > 1) { const unsigned mid = (lo + hi) >> 1; if (rec_off[mid] <= pos) lo = mid; else hi = mid; }
        ull fwd = 0;
        for (int b = 0; b < S2_K; ++b) {
            const unsigned c = flat[pos + b];
            const unsigned x = (c >> 1) & 3u;
            fwd = (fwd << 2) | (x ^ (x >> 1));
        }
        const ull at = atomicAdd(f_cnt, 1ull);
        if (at < cap_f) { f_rec[at] = (unsigned)(base + lo); f_off[at] = (unsigned)(pos - rec_off[lo]); f_kmer[at] = s2_canonical(fwd, s2_revcomp31(fwd)); }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*decompress_fn)(CUmemDecompressParams *, size_t, unsigned int, size_t *, CUstream);
typedef CUresult (*devattr_fn)(int *, CUdevice_attribute, CUdevice);

enum { ING_COUNT = 1, ING_DETECT = 2 };

#define ING_SLOTS 3
// ---- ordinary .gz (single-member gzip: what the reference's own inputs are) -------------------------------------------
// The hardware engine cannot take it (DESIGN 4.4 "Damaged data"), so those files go through the chunk-parallel software
// gunzip of s2_gunzip.cu: a BATCH of whole files (up to S2_GZ_BATCH_MB of compressed bytes) is copied to the device and
// decoded by one launch - one warp per S2_GZ_SUB_KB of compressed bytes, thousands of warps - then chained per file; the
// files' texts are then translated straight into the text buffers of ordinary pipeline chunks (groups of whole files),
// CRC-checked, and from there on a gz group is a BGZF group whose "blocks" are whole files: isz[f] = ISIZE as the host read
// it, act[f] = what the file inflated to (all ones on any doubt: chain broken, size or CRC-32 mismatch, a second member),
// so ing_check_chunk's act == isz test vetoes the scan exactly as it does for a damaged BGZF member.
struct GzBufSet {                        // what one batch brings along: two sets, so that the next batch's bytes travel while this one is decoded
    uint8_t *h_comp = nullptr, *d_comp = nullptr;
    GzFileDesc *h_files = nullptr, *d_files = nullptr;
    uint32_t *h_sub_file = nullptr, *d_sub_file = nullptr, *h_slice0 = nullptr, *d_slice0 = nullptr;
    cudaEvent_t idle = nullptr;          // inflate stream: the last chunk of the batch that used this set has been translated
};
struct GzStage {
    size_t comp_cap = 0;
    uint32_t sub_bytes = 0, sub_cap = 0, max_sub = 0, max_files = 0;
    GzBufSet set[2];
    unsigned batch_no = 0;
    uint16_t *d_sym = nullptr, *d_sym_alloc = nullptr;                 // (d_sym = what gz_launch_sym_init made of the allocation)
    GzSubResult *d_res = nullptr;
    uint8_t *d_win = nullptr;
    uint8_t *d_carry = nullptr;          // streamed files: the window a piece leaves to the next one
    size_t sym_subs = 0, win_slots = 0;  // what d_sym / d_win hold now (they grow with the batches: gz_stage_reserve)
    uint64_t *d_sub_off = nullptr;
    GzFileResult *d_fres = nullptr;
    uint32_t *d_crc_acc = nullptr;
    uint8_t *d_piece_text = nullptr; size_t piece_text_cap = 0;      // streamed files: the text of one piece
    GzFileResult *h_fres = nullptr;      // pinned: a piece's result
    bool used = false;
};

struct IngSlot {                       // a chunk travels through one slot of the ring: copy engine -> decompression engine -> kernels
    uint8_t *h_comp = nullptr;         // pinned staging for file sources (allocated on first use)
    uint8_t *d_comp = nullptr;         // compressed bytes on the device
    uint8_t *d_text = nullptr;         // [carry region: ING_MAXCARRY][inflated text: text_cap]
    uint8_t *h_meta = nullptr, *d_meta = nullptr;      // [file_end: ull x ING_MAX_FILES][isz: u32 x ING_MAX_DBLOCKS]
    unsigned *d_act = nullptr;         // bytes the engine produced per DEFLATE block
    std::vector<CUmemDecompressParams> params;
    cudaEvent_t h2d_done = nullptr;    // copy stream: the chunk's bytes are on the device
    cudaEvent_t inflated = nullptr;    // inflate stream: the text is in d_text
    cudaEvent_t consumed = nullptr;    // kernel stream: the chunk's kernels are through with this slot
    cudaEvent_t crc_done = nullptr;    // crc stream: the BGZF members' CRC-32 have been compared
    uint32_t *d_crc_bad = nullptr;     // ... and this is 1 if one of them did not match
};

struct s2_ingest {
    s2_ctx *ctx = nullptr;
    int device = 0;
    // A pipeline belongs to its context's pool and is used by one thread at a time: whoever enqueues chunks (a job's
    // submit, its retries / streamed files, a detect call) holds `mu`; verdicts are read without it.
    std::mutex mu;
    std::mutex res_mu; std::set<uint64_t> res_live;      // verdict-ring entries handed out and not read yet
    // three streams, one per engine, so that the copy of chunk i+2, the inflate of chunk i+1 and the kernels of chunk i overlap
    cudaStream_t stream = nullptr, copy_stream = nullptr, inflate_stream = nullptr, crc_stream = nullptr;
    size_t comp_chunk = 0, text_cap = 0; unsigned max_lines = 0;
    IngSlot slot[ING_SLOTS];
    uint64_t n_chunks = 0;             // chunks enqueued so far (slot = n_chunks % ING_SLOTS)
    uint64_t call_chunk0 = 0;          // n_chunks at the start of the current call: the first chunks of a call are small
                                       // (2, 4, 8 ... MB) so that the inflate engine and the kernels start early
    uint8_t *d_flat = nullptr;
    unsigned *d_block_nl = nullptr, *d_block_out = nullptr, *d_line_end = nullptr;
    unsigned short *d_masks = nullptr;   // newline mask of every 16 text bytes
    unsigned *d_tickets = nullptr;       // last-block election of the three fused kernels
    uint64_t *part_pool = nullptr; unsigned long long *part_cursor = nullptr; uint32_t *part_overflow = nullptr;     // two-phase scan scratch
    IngState *d_state = nullptr;
    IngResult *h_results = nullptr, *d_results = nullptr;      // ring of verdicts in mapped pinned memory (host pointer, device alias)
    uint64_t res_seq = 0;                // verdicts handed out so far; slot = res_seq % ING_MAX_RESULTS
    decompress_fn decompress = nullptr;
    bool hw_deflate = false;
    int bgzf_crc = 1;                    // S2_BGZF_CRC: 0 trust the members' ISIZE alone (round 1), 1 (default) check on the kernel stream, in front of the chunk's other kernels, 2 on a stream of its own
    uint32_t *d_xp128 = nullptr;         // x^(8 * 128 * j) mod P for the member CRC kernel
    int grid_scan = 0;                   // CTAs of the count scan launched from this pipeline
    GzStage gz;                          // ordinary .gz batches (allocated on first use)
    // detect mode: chunk-local and file-level result arrays
    unsigned *d_hits_c = nullptr, *d_inf_c = nullptr;
    ull *d_rec_off = nullptr, *d_pos_c = nullptr, *d_cnt_c = nullptr, *d_fcnt = nullptr;
    unsigned *d_len_all = nullptr, *d_hits_all = nullptr, *d_inf_all = nullptr; ull rec_cap = 0;
    unsigned *d_frec = nullptr, *d_foff = nullptr; ull *d_fkmer = nullptr; ull f_cap = 0;
};

static void ingest_free(s2_ingest *g)
{
    if (!g) return;
    cudaSetDevice(g->device);
    if (g->stream) cudaStreamSynchronize(g->stream);
    if (g->copy_stream) { cudaStreamSynchronize(g->copy_stream); cudaStreamDestroy(g->copy_stream); }
    if (g->inflate_stream) { cudaStreamSynchronize(g->inflate_stream); cudaStreamDestroy(g->inflate_stream); }
    if (g->crc_stream) { cudaStreamSynchronize(g->crc_stream); cudaStreamDestroy(g->crc_stream); }
    if (g->stream) cudaStreamDestroy(g->stream);
    for (auto &s : g->slot) {
        cudaFreeHost(s.h_comp); cudaFree(s.d_comp); cudaFree(s.d_text); cudaFreeHost(s.h_meta); cudaFree(s.d_meta); cudaFree(s.d_act);
        if (s.h2d_done) cudaEventDestroy(s.h2d_done);
        if (s.inflated) cudaEventDestroy(s.inflated);
        if (s.consumed) cudaEventDestroy(s.consumed);
        if (s.crc_done) cudaEventDestroy(s.crc_done);
        cudaFree(s.d_crc_bad);
    }
    {
        GzStage &z = g->gz;
        cudaFree(z.d_sym_alloc); cudaFree(z.d_res); cudaFree(z.d_win); cudaFree(z.d_carry); cudaFree(z.d_sub_off);
        for (auto &zs : z.set) {
            cudaFreeHost(zs.h_comp); cudaFree(zs.d_comp);
            cudaFreeHost(zs.h_files); cudaFree(zs.d_files); cudaFreeHost(zs.h_sub_file); cudaFree(zs.d_sub_file); cudaFreeHost(zs.h_slice0); cudaFree(zs.d_slice0);
            if (zs.idle) cudaEventDestroy(zs.idle);
        }
        cudaFree(z.d_fres); cudaFree(z.d_crc_acc); cudaFree(z.d_piece_text); cudaFreeHost(z.h_fres);
    }
    cudaFree(g->d_flat);
    cudaFree(g->d_block_nl); cudaFree(g->d_block_out); cudaFree(g->d_line_end); cudaFree(g->d_masks); cudaFree(g->d_tickets);
    cudaFree(g->part_pool); cudaFree(g->part_cursor); cudaFree(g->part_overflow);
    cudaFree(g->d_state); cudaFreeHost(g->h_results); cudaFree(g->d_xp128);
    cudaFree(g->d_hits_c); cudaFree(g->d_inf_c); cudaFree(g->d_rec_off); cudaFree(g->d_pos_c); cudaFree(g->d_cnt_c); cudaFree(g->d_fcnt);
    cudaFree(g->d_len_all); cudaFree(g->d_hits_all); cudaFree(g->d_inf_all); cudaFree(g->d_frec); cudaFree(g->d_foff); cudaFree(g->d_fkmer);
    cudaGetLastError();
    delete g;
}

// per chunk: [file_end: u64 x MAX_FILES][isz: u32 x MAX_DBLOCKS][BGZF members' CRC-32: u32 x MAX_DBLOCKS][where their text starts: u32 x MAX_DBLOCKS]
#define ING_META_ISZ ((size_t)ING_MAX_FILES * 8)
#define ING_META_CRC (ING_META_ISZ + (size_t)ING_MAX_DBLOCKS * 4)
#define ING_META_TOFF (ING_META_CRC + (size_t)ING_MAX_DBLOCKS * 4)
#define ING_META_BYTES (ING_META_TOFF + (size_t)ING_MAX_DBLOCKS * 4)
static int ingest_init(s2_ingest *g, s2_ctx *c)
{
    g->ctx = c; g->device = c->device;
    // S2_INGEST_CHUNK_MB: compressed bytes per chunk; S2_INGEST_TEXT_MB: inflated text per chunk (BGZF: <= 64 KB per block)
    g->comp_chunk = (size_t)std::min<uint64_t>(std::max<uint64_t>(s2_env_u64("S2_INGEST_CHUNK_MB", 16), 1), 1024) << 20;
    g->text_cap = (size_t)std::min<uint64_t>(std::max<uint64_t>(s2_env_u64("S2_INGEST_TEXT_MB", 64), 4), 2048) << 20;
    g->max_lines = (unsigned)(g->text_cap / 8);
    g->grid_scan = c->grid_count;
    CK(cudaSetDevice(c->device));
    CK(cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&g->copy_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&g->inflate_stream, cudaStreamNonBlocking));
    for (auto &s : g->slot) {
        CK(cudaMalloc((void **)&s.d_comp, g->comp_chunk + 256));
        CK(cudaMalloc((void **)&s.d_text, (size_t)ING_MAXCARRY + g->text_cap + ING_BLOCK));
        CK(cudaHostAlloc((void **)&s.h_meta, ING_META_BYTES, cudaHostAllocDefault));
        CK(cudaMalloc((void **)&s.d_meta, ING_META_BYTES));
        CK(cudaMalloc((void **)&s.d_act, (size_t)ING_MAX_DBLOCKS * sizeof(unsigned)));
        CK(cudaEventCreateWithFlags(&s.h2d_done, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&s.inflated, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&s.consumed, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&s.crc_done, cudaEventDisableTiming));
        CK(cudaMalloc((void **)&s.d_crc_bad, sizeof(uint32_t)));
    }
    CK(cudaMalloc((void **)&g->d_flat, g->text_cap + ING_MAXCARRY + 4096));
    const size_t n_blocks = ((size_t)ING_MAXCARRY + g->text_cap) / ING_BLOCK + 4;
    CK(cudaMalloc((void **)&g->d_block_nl, n_blocks * sizeof(unsigned)));
    CK(cudaMalloc((void **)&g->d_block_out, n_blocks * sizeof(unsigned)));
    CK(cudaMalloc((void **)&g->d_masks, n_blocks * (ING_BLOCK / 16) * sizeof(unsigned short)));
    CK(cudaMalloc((void **)&g->d_line_end, (size_t)g->max_lines * sizeof(unsigned)));
    CK(cudaMalloc((void **)&g->d_state, sizeof(IngState)));
    CK(cudaMemset(g->d_state, 0, sizeof(IngState)));                   // afterwards every finished file leaves a clean state behind
    g->bgzf_crc = s2_env_int("S2_BGZF_CRC", 1);
    CK(cudaStreamCreateWithFlags(&g->crc_stream, cudaStreamNonBlocking));
    CK(cudaMalloc((void **)&g->d_xp128, 2048 * sizeof(uint32_t)));                   // constants of the member CRC kernel (s2_gunzip.cu)
    gz_launch_xp128_init(g->d_xp128, g->inflate_stream);
    CK(cudaMalloc((void **)&g->d_tickets, 4 * sizeof(unsigned)));
    CK(cudaMemset(g->d_tickets, 0, 4 * sizeof(unsigned)));
    CK(cudaHostAlloc((void **)&g->h_results, ING_MAX_RESULTS * sizeof(IngResult), cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer((void **)&g->d_results, g->h_results, 0));          // the end-of-chunk kernel writes the verdict straight to the host
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
// A buffer that ends inside the member - even inside its header - is reported as "incomplete" (true, *data_len =
// (size_t)-1, *block_size = the member's size where the header already tells it, else 0) as long as the bytes that are
// there fit a BGZF header: a streamed file's chunk may end anywhere (round 1 called a chunk that ended within the first
// 18 bytes of a header "not BGZF" and sent the whole file back to host zlib after un-counting it).
static bool bgzf_block(const uint8_t *p, size_t avail, size_t *block_size, size_t *data_off, size_t *data_len, uint32_t *isize)
{
    static const uint8_t magic[3] = { 0x1f, 0x8b, 8 };
    for (size_t i = 0; i < 3 && i < avail; ++i) if (p[i] != magic[i]) return false;
    if (avail > 3 && !(p[3] & 4)) return false;
    if (avail < 18) { *block_size = 0; *data_len = (size_t)-1; return avail > 0; }
    const size_t xlen = p[10] | (p[11] << 8);
    if (avail < 12 + xlen) { *block_size = 0; *data_len = (size_t)-1; return true; }
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

// ---- big reads on several threads ----------------------------------------------------------------------------
// One thread copies a file out of the page cache at 4-6 GB/s; a 16 MB chunk of a big BGZF file, or a 100 MB piece of a big
// .gz, read by the thread that also drives the pipeline left the GPU waiting for the file (strain_detect on one 60 MB
// file: 13 ms, nearly all of it pread - profiles/r2f_detect_bench.txt).  Reads of more than 4 MB are split over a small
// pool of helper threads (S2_READ_THREADS, default 6; 0 = none).
struct ReadPool {
    std::mutex mu; std::condition_variable cv;
    std::deque<std::function<void()>> q;
    std::vector<std::thread> th;
    bool stop = false;
    explicit ReadPool(int n)
    {
        for (int i = 0; i < n; ++i)
            th.emplace_back([this]() {
                for (;;) {
                    std::function<void()> f;
                    {
                        std::unique_lock<std::mutex> lk(mu);
                        cv.wait(lk, [this]() { return stop || !q.empty(); });
                        if (q.empty()) return;
                        f = std::move(q.front()); q.pop_front();
                    }
                    f();
                }
            });
    }
    ~ReadPool()
    {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv.notify_all();
        for (auto &t : th) t.join();
    }
};
static ReadPool *read_pool(void)
{
    static ReadPool *pool = []() -> ReadPool * {
        const int n = std::min(std::max(s2_env_int("S2_READ_THREADS", 6), 0), 32);
        return n ? new ReadPool(n) : nullptr;             // (lives as long as the process: its threads sleep when there is nothing to read)
    }();
    return pool;
}
// pread of exactly what is there: returns the bytes read (short only at the end of the file) or -1
static ssize_t ing_pread(int fd, uint8_t *dst, size_t len, off_t off)
{
    ReadPool *pool = len > (4u << 20) ? read_pool() : nullptr;
    auto read_all = [fd](uint8_t *d, size_t n, off_t o) -> ssize_t {
        size_t done = 0;
        while (done < n) {
            const ssize_t r = pread(fd, d + done, n - done, o + (off_t)done);
            if (r < 0) return -1;
            if (r == 0) break;
            done += (size_t)r;
        }
        return (ssize_t)done;
    };
    if (!pool) return read_all(dst, len, off);
    const size_t n_parts = std::min<size_t>(pool->th.size() + 1, (len + (2u << 20) - 1) / (2u << 20));
    const size_t part = (len / n_parts + 4095) & ~(size_t)4095;
    std::vector<ssize_t> got(n_parts, 0);
    std::mutex mu; std::condition_variable cv; size_t pending = 0;
    for (size_t k = 1; k < n_parts; ++k) {
        if (k * part >= len) break;
        { std::lock_guard<std::mutex> lk(mu); ++pending; }
        auto task = [&, k]() {
            got[k] = read_all(dst + k * part, std::min(part, len - k * part), off + (off_t)(k * part));
            std::lock_guard<std::mutex> lk(mu);
            if (--pending == 0) cv.notify_one();
        };
        { std::lock_guard<std::mutex> lk(pool->mu); pool->q.emplace_back(task); }
        pool->cv.notify_one();
    }
    got[0] = read_all(dst, std::min(part, len), off);
    { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&]() { return pending == 0; }); }
    ssize_t total = 0;
    for (size_t k = 0; k < n_parts; ++k) {
        if (got[k] < 0) return -1;
        total += got[k];
        if ((size_t)got[k] < std::min(part, len > k * part ? len - k * part : 0)) break;       // the file ends inside this part
    }
    return total;
}

// where the compressed (or plain) bytes come from: a file read chunk by chunk into pinned staging buffers, or a
// caller's host buffer (pinned for full PCIe rate) that is copied to the device directly
struct IngSource {
    int fd = -1;
    const uint8_t *mem = nullptr;
    size_t mem_len = 0;
    bool eligible = false, bgzf = false, fasta = false;      // classification
    bool gz = false; uint32_t gz_isize = 0;                  // ordinary .gz (S2_GPU_GUNZIP=1): ISIZE of the trailer
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

// first byte of the text inside a BGZF file ('@' FASTQ, '>' FASTA): inflate the start of the first non-empty member on
// the host (a few microseconds: one z_stream per thread, only the member's own bytes are looked at)
static int bgzf_first_text_byte(const IngSource &src)
{
    static thread_local z_stream z;
    static thread_local bool z_ready = false;
    static thread_local std::vector<uint8_t> buf;
    off_t off = 0;
    for (int guard = 0; guard < 64; ++guard) {                 // leading empty members
        uint8_t head[32];
        const uint8_t *p = head;
        ssize_t hn;
        if (src.mem) { hn = (size_t)off < src.mem_len ? (ssize_t)std::min<size_t>(src.mem_len - (size_t)off, 65536) : 0; p = src.mem + off; }
        else hn = src.peek(head, sizeof head, off);
        size_t bs = 0, doff = 0, dlen = 0; uint32_t isz = 0;
        if (hn <= 0 || !bgzf_block(p, (size_t)hn, &bs, &doff, &dlen, &isz)) return -1;
        if (dlen == (size_t)-1) {                              // the header was readable, the member is not (yet)
            if (src.mem || bs > 65536 || bs == 0) return -1;
            buf.resize(bs);
            if (src.peek(buf.data(), bs, off) != (ssize_t)bs) return -1;
            p = buf.data();
            if (!bgzf_block(p, bs, &bs, &doff, &dlen, &isz) || dlen == (size_t)-1) return -1;
        }
        if (isz) {
            if (!z_ready) { memset(&z, 0, sizeof z); if (inflateInit2(&z, -15) != Z_OK) return -1; z_ready = true; }
            else if (inflateReset(&z) != Z_OK) return -1;
            uint8_t out[16];
            z.next_in = const_cast<uint8_t *>(p) + doff; z.avail_in = (uInt)dlen;
            z.next_out = out; z.avail_out = sizeof out;
            inflate(&z, Z_SYNC_FLUSH);
            return z.total_out ? out[0] : -1;
        }
        off += (off_t)bs;
    }
    return -1;
}

// first byte of the text of an ordinary .gz file (its header may carry a name: up to 4 KB are looked at), -1 if unknown
static int gz_first_text_byte(const IngSource &src)
{
    static thread_local std::vector<uint8_t> buf;
    buf.resize(4096);
    const ssize_t hn = src.peek(buf.data(), buf.size(), 0);
    if (hn < 20) return -1;
    const uint64_t hl = s2_gzip_header_len(buf.data(), (uint64_t)hn);
    if (!hl) return -1;
    z_stream z; memset(&z, 0, sizeof z);
    if (inflateInit2(&z, -15) != Z_OK) return -1;
    uint8_t out[16];
    z.next_in = buf.data() + hl; z.avail_in = (uInt)((size_t)hn - hl);
    z.next_out = out; z.avail_out = sizeof out;
    inflate(&z, Z_SYNC_FLUSH);
    const int first = z.total_out ? out[0] : -1;
    inflateEnd(&z);
    return first;
}

static void ingest_classify(IngSource &src)
{
    uint8_t head[32];
    const ssize_t hn = src.peek(head, sizeof head, 0);
    src.bgzf = is_bgzf_header(head, hn);
    src.gz = !src.bgzf && hn >= 18 && head[0] == 0x1f && head[1] == 0x8b && head[2] == 8 && s2_env_int("S2_GPU_GUNZIP", 1) != 0;
    if (src.gz) {
        uint8_t tail[4];
        const ssize_t size = src.size();
        if (size < 18 || src.peek(tail, 4, (off_t)size - 4) != 4) src.gz = false;
        else src.gz_isize = (uint32_t)tail[0] | ((uint32_t)tail[1] << 8) | ((uint32_t)tail[2] << 16) | ((uint32_t)tail[3] << 24);
    }
    const int first = src.bgzf ? bgzf_first_text_byte(src) : src.gz ? gz_first_text_byte(src) : (hn >= 1 ? head[0] : -1);
    src.fasta = first == '>';
    src.eligible = first == '@' || first == '>';             // neither FASTQ nor FASTA (or an ordinary .gz): host reader
    // uncompressed text takes this path too (read() + PCIe against a host parser at about 1 GB/s per thread);
    // S2_GPU_INGEST_PLAIN=0 keeps it on the host
    if (!src.bgzf && !src.gz && !s2_env_int("S2_GPU_INGEST_PLAIN", 1)) src.eligible = false;
}

// The pipelines of a context: a small pool (S2_INGEST_PIPES, default 2; strain_detect asks for 3) shared by every thread that
// calls in.  One pipeline already overlaps copy engine, inflate engine and kernels; the second exists so that one thread can
// build its launch lists while another one's are being enqueued (reading the files is the callers' business: the executables'
// reader threads fill pinned arenas of their own).  (Round 1 had one pipeline per calling thread:
// 16 reader threads allocated 16 rings - 5 GB of HBM, 0.8 GB of pinned staging - and the allocation alone made the
// executables 18x slower than with one thread, profiles/r1n_e2e_cli.txt.)
struct IngPool {
    std::mutex mu;                       // guards `pipes` growing
    std::vector<s2_ingest *> pipes;
    std::atomic<unsigned> rr{0};
    unsigned max_pipes = 3;
};
static std::mutex g_pool_mu;             // guards c->ingest_pool coming into being

static IngPool *ingest_pool(s2_ctx *c)
{
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (!c->ingest_pool) {
        IngPool *p = new IngPool();
        p->max_pipes = (unsigned)std::min(std::max(s2_env_int("S2_INGEST_PIPES", 2), 1), 16);
        c->ingest_pool = p;
    }
    return (IngPool *)c->ingest_pool;
}

// host-side time accounting (S2_INGEST_TRACE=1): where the submitting thread spends its time
static thread_local double tr_wait = 0, tr_h2d = 0, tr_decomp = 0, tr_launch = 0;
struct IngTraceEv { cudaEvent_t e[6]; size_t comp = 0, text = 0; };     // copy begin/end, inflate begin/end, kernels begin/end
static thread_local std::vector<IngTraceEv> tr_events;
static thread_local bool tr_on = false;
static void tr_record(int which, cudaStream_t st) { if (tr_on && !tr_events.empty()) cudaEventRecord(tr_events.back().e[which], st); }
static inline double ing_now() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// a pipeline of the context, LOCKED (release with g->mu.unlock()): a free one, else a new one while the pool may grow,
// else wait for one
static s2_ingest *ingest_acquire(s2_ctx *c)
{
    // pipelines are shared by threads: the calling thread's current device may be another context's
    if (cudaSetDevice(c->device) != cudaSuccess) { s2_set_error("cannot select device %d", c->device); return nullptr; }
    IngPool *pool = ingest_pool(c);
    for (int round = 0; round < 2; ++round) {
        std::lock_guard<std::mutex> lk(pool->mu);
        for (s2_ingest *g : pool->pipes) if (g->mu.try_lock()) return g;
        if (pool->pipes.size() < pool->max_pipes) {
            s2_ingest *g = new s2_ingest();
            if (ingest_init(g, c)) { ingest_free(g); return nullptr; }
            g->mu.lock();
            pool->pipes.push_back(g);
            return g;
        }
    }
    s2_ingest *g;
    { std::lock_guard<std::mutex> lk(pool->mu); g = pool->pipes[pool->rr.fetch_add(1) % pool->pipes.size()]; }
    g->mu.lock();
    return g;
}

// creates the context's pipelines ahead of their first use (the executables do this beside the table build: a
// pipeline is 350 MB of device memory and, for file sources, 48 MB of pinned staging - tens of milliseconds each)
extern "C" int s2_ingest_warm(s2_ctx *c, int n_pipes)
{
    CK(cudaSetDevice(c->device));
    IngPool *pool = ingest_pool(c);
    std::lock_guard<std::mutex> lk(pool->mu);
    while (pool->pipes.size() < std::min<size_t>((size_t)std::max(n_pipes, 0), pool->max_pipes)) {
        s2_ingest *g = new s2_ingest();
        if (ingest_init(g, c)) { ingest_free(g); return -1; }
        // (the slots' pinned staging - for sources that are files, not memory images - comes with the first such source:
        // the executables hand over images from their reader threads' arenas)
        pool->pipes.push_back(g);
    }
    return 0;
}

static int ingest_gz_stage_init(s2_ingest *g);
static int gz_stage_reserve(s2_ingest *g, uint32_t n_sub, uint32_t n_files);
extern "C" int s2_ingest_warm_gz(s2_ctx *c, int n_pipes, uint64_t batch_comp_bytes)
{
    if (s2_ingest_warm(c, n_pipes)) return -1;
    IngPool *pool = ingest_pool(c);
    std::vector<s2_ingest *> pipes;
    { std::lock_guard<std::mutex> lk(pool->mu); pipes = pool->pipes; }
    for (size_t i = 0; i < pipes.size() && (int)i < n_pipes; ++i) {
        s2_ingest *g = pipes[i];
        std::lock_guard<std::mutex> lk(g->mu);
        if (ingest_gz_stage_init(g)) return -1;
        const uint32_t files = 64;
        const uint32_t subs = (uint32_t)std::min<uint64_t>(batch_comp_bytes / g->gz.sub_bytes + files, g->gz.max_sub);
        if (gz_stage_reserve(g, subs, files)) return -1;
    }
    return 0;
}

// a verdict-ring entry for the chunk about to be enqueued (caller holds g->mu); fails instead of overwriting a verdict
// that nobody has read yet (thousands of jobs submitted and never waited for)
static int ingest_result_claim(s2_ingest *g)
{
    std::lock_guard<std::mutex> lk(g->res_mu);
    if (!g->res_live.empty() && g->res_seq - *g->res_live.begin() + 2 >= ING_MAX_RESULTS) {
        s2_set_error("too many ingest jobs in flight on one pipeline: wait for earlier jobs first");
        return -1;
    }
    g->res_live.insert(g->res_seq);
    return 0;
}
static IngResult ingest_result_take(s2_ingest *g, uint64_t seq)
{
    const IngResult r = g->h_results[seq % ING_MAX_RESULTS];
    std::lock_guard<std::mutex> lk(g->res_mu);
    g->res_live.erase(seq);
    return r;
}

// compressed bytes the next chunk may hold: ramps up over the first chunks of a call
static size_t ingest_chunk_cap(const s2_ingest *g)
{
    const uint64_t k = g->n_chunks - g->call_chunk0;
    return k >= 6 ? g->comp_chunk : std::min<size_t>(g->comp_chunk, (size_t)(2u << 20) << k);
}

// ---- one chunk ------------------------------------------------------------------------------------------
struct IngChunk {
    size_t comp_len = 0;          // bytes staged in the slot's d_comp
    size_t text_len = 0;
    bool first = true, last = true;
    unsigned n_files = 0;         // > 0: a group of whole files (their ends are in the slot's meta)
    bool gz = false;              // the group's files are ordinary .gz, decoded by the pipeline's gz stage: its files [gz_file0, +n_files)
    uint32_t gz_file0 = 0, gz_sub_lo = 0, gz_sub_hi = 0, gz_slice0 = 0, gz_slices = 0;      // and their sub-chunks / CRC slices
    unsigned gz_set = 0;          // which of the stage's two buffer sets the batch sits in
    const uint8_t *dev_src = nullptr;   // the chunk's text is already on the device (a slice of a streamed .gz piece's text)
};

// wait until the slot's previous chunk has left d_comp / params / meta, and return the slot
static int ingest_slot_begin(s2_ingest *g, IngSlot **out)
{
    IngSlot &s = g->slot[g->n_chunks % ING_SLOTS];
    const double t0 = ing_now();
    CK(cudaEventSynchronize(s.consumed));
    tr_wait += ing_now() - t0;
    if (tr_on) {
        tr_events.emplace_back();
        for (auto &e : tr_events.back().e) cudaEventCreate(&e);
        tr_record(0, g->copy_stream);
    }
    s.params.clear();
    *out = &s;
    return 0;
}

static int ingest_staging(IngSlot &s, size_t bytes)
{
    if (!s.h_comp) CK(cudaHostAlloc((void **)&s.h_comp, bytes, cudaHostAllocDefault));
    return 0;
}

// walk the BGZF members in h[0..avail) and append their DEFLATE streams to the slot's decompress list; the bytes will
// sit at d_comp + comp_off, the text goes to text offset *text_len.  Stops at a partial member or when a cap would be
// exceeded (then *full is set).  Returns the bytes consumed, or (size_t)-1 if this is not BGZF.
static size_t ingest_walk_bgzf(s2_ingest *g, IngSlot &s, const uint8_t *h, size_t avail, size_t comp_off, size_t *text_len, bool *full)
{
    size_t used = 0;
    unsigned *isz_list = (unsigned *)(s.h_meta + ING_META_ISZ), *crc_list = (unsigned *)(s.h_meta + ING_META_CRC), *toff_list = (unsigned *)(s.h_meta + ING_META_TOFF);
    *full = false;
    while (used < avail) {
        size_t bs = 0, doff = 0, dlen = 0; uint32_t isz = 0;
        if (!bgzf_block(h + used, avail - used, &bs, &doff, &dlen, &isz)) return (size_t)-1;
        if (dlen == (size_t)-1) break;                                   // partial block: the next chunk starts here
        if (isz > 65536) return (size_t)-1;
        if (*text_len + isz > g->text_cap || s.params.size() >= ING_MAX_DBLOCKS) { *full = true; break; }
        if (isz) {
            CUmemDecompressParams p; memset(&p, 0, sizeof p);
            p.srcNumBytes = dlen; p.dstNumBytes = isz;
            p.dstActBytes = (cuuint32_t *)(s.d_act + s.params.size());
            p.src = s.d_comp + comp_off + used + doff;
            p.dst = s.d_text + ING_MAXCARRY + *text_len;
            p.algo = CU_MEM_DECOMPRESS_ALGORITHM_DEFLATE;
            isz_list[s.params.size()] = isz;
            const uint8_t *tr = h + used + bs - 8;                           // the member's trailer: CRC-32, ISIZE
            crc_list[s.params.size()] = (uint32_t)tr[0] | (uint32_t)tr[1] << 8 | (uint32_t)tr[2] << 16 | (uint32_t)tr[3] << 24;
            toff_list[s.params.size()] = (unsigned)*text_len;
            s.params.push_back(p);
            *text_len += isz;
        }
        used += bs;
    }
    return used;
}

// the count scan of the chunk's flat batch: direct kernel, or the two-phase scan for tables larger than L2
static int ingest_launch_count(s2_ingest *g, s2_table *t, const S2DevBatch *dev, int col)
{
    s2_ctx *c = g->ctx;
    if (!t->partitioned) {
        s2_launch_scan_count_dev(g->d_flat, dev, t->v, col, c->d_stats, g->grid_scan, g->stream);
        return 0;
    }
    const uint64_t n_max = g->text_cap + ING_MAXCARRY;                                     // the flat batch cannot be longer
    const uint64_t region_cap = n_max / S2_NPART + n_max / (2 * S2_NPART) + 8192;          // 1.5x the even share
    if (!g->part_pool) {
        CK(cudaMalloc((void **)&g->part_pool, region_cap * S2_NPART * sizeof(uint64_t)));
        CK(cudaMalloc((void **)&g->part_cursor, (S2_NPART + 1) * sizeof(unsigned long long)));
        CK(cudaMalloc((void **)&g->part_overflow, 2 * sizeof(uint32_t)));
    }
    s2_launch_scan_count_partitioned(g->d_flat, 0, t->v, col, c->d_stats, g->part_pool, region_cap, g->part_cursor, g->part_overflow,
                                     c->n_sm, g->grid_scan, g->stream, dev);
    return 0;
}

// the device work of one chunk whose compressed bytes are being copied into the slot by the copy stream
static int ingest_enqueue(s2_ingest *g, s2_table *t, IngSlot &s, const IngChunk &ch, bool bgzf, bool fasta, int mode, int col, unsigned inc,
                          bool want_result)
{
    s2_ctx *c = g->ctx;
    cudaStream_t st = g->stream;
    // meta (group file ends, expected block sizes) travels with the chunk
    const size_t n_db = bgzf ? s.params.size() : ch.gz ? (size_t)ch.n_files : 0;
    if (ch.n_files) CK(cudaMemcpyAsync(s.d_meta, s.h_meta, (size_t)ch.n_files * 8, cudaMemcpyHostToDevice, g->copy_stream));
    if (n_db) CK(cudaMemcpyAsync(s.d_meta + ING_META_ISZ, s.h_meta + ING_META_ISZ, n_db * 4, cudaMemcpyHostToDevice, g->copy_stream));
    const bool member_crc = bgzf && n_db && g->bgzf_crc;
    if (member_crc) {
        CK(cudaMemcpyAsync(s.d_meta + ING_META_CRC, s.h_meta + ING_META_CRC, n_db * 4, cudaMemcpyHostToDevice, g->copy_stream));
        CK(cudaMemcpyAsync(s.d_meta + ING_META_TOFF, s.h_meta + ING_META_TOFF, n_db * 4, cudaMemcpyHostToDevice, g->copy_stream));
    }
    tr_record(1, g->copy_stream);
    if (tr_on && !tr_events.empty()) { tr_events.back().comp = ch.comp_len; tr_events.back().text = ch.text_len; }
    CK(cudaEventRecord(s.h2d_done, g->copy_stream));
    CK(cudaStreamWaitEvent(g->inflate_stream, s.h2d_done, 0));
    tr_record(2, g->inflate_stream);
    const double t_dec = ing_now();
    if (bgzf) {
        for (size_t i = 0; i < s.params.size(); i += ING_DECOMP_CALL) {
            size_t err_index = 0;
            const size_t n = std::min<size_t>(ING_DECOMP_CALL, s.params.size() - i);
            const CUresult r = g->decompress(s.params.data() + i, n, 0, &err_index, (CUstream)g->inflate_stream);
            if (r != CUDA_SUCCESS) { s2_set_error("hardware decompression failed (driver error %d at block %zu)", (int)r, i + err_index); return -1; }
        }

    } else if (ch.gz) {
        // the batch was decoded and chained on this stream already: symbols -> text of this chunk's files, then their CRC-32
        GzStage &z = g->gz;
        const GzBufSet &zs = z.set[ch.gz_set];
        gz_launch_translate(zs.d_files, zs.d_sub_file, ch.gz_sub_lo, ch.gz_sub_hi, z.d_sym, z.sub_cap, z.d_win, z.d_sub_off, z.d_fres,
                            s.d_text + ING_MAXCARRY, g->inflate_stream);
        gz_launch_crc(zs.d_files, ch.gz_file0, ch.n_files, zs.d_slice0 + ch.gz_slice0, ch.gz_slices, s.d_text + ING_MAXCARRY, z.d_fres, z.d_crc_acc, s.d_act,
                      g->inflate_stream);
    } else if (ch.dev_src) {
        if (ch.text_len) CK(cudaMemcpyAsync(s.d_text + ING_MAXCARRY, ch.dev_src, ch.text_len, cudaMemcpyDeviceToDevice, g->inflate_stream));
    } else if (ch.comp_len) {
        CK(cudaMemcpyAsync(s.d_text + ING_MAXCARRY, s.d_comp, ch.comp_len, cudaMemcpyDeviceToDevice, g->inflate_stream));
    }
    tr_record(3, g->inflate_stream);
    CK(cudaEventRecord(s.inflated, g->inflate_stream));
    CK(cudaStreamWaitEvent(st, s.inflated, 0));
    // The engine checks no CRC: every member's text against its trailer, before the chunk's verdict is formed.  WHERE the
    // kernel runs matters more than how fast it is (profiles/r2n_ingest_crc_modes.txt, r2o): the scan kernel of the chunk
    // before is a persistent grid that holds every register of every SM for 250 us, so a kernel on another stream - the
    // inflate stream, or one of its own - waits for it; that wait, not the 20 us of CRC work, made the inflate stage the
    // slowest of the three (132 -> 100-112 Gbases/s end to end).  On the kernel stream it queues behind that scan anyway
    // and costs its own duration only.  (S2_BGZF_CRC=2 keeps the separate-stream form with its veto kernel.)
    if (member_crc && g->bgzf_crc == 1)
        gz_launch_member_crc(s.d_text + ING_MAXCARRY, (const uint32_t *)(s.d_meta + ING_META_ISZ), (const uint32_t *)(s.d_meta + ING_META_CRC),
                             (const uint32_t *)(s.d_meta + ING_META_TOFF), (uint32_t)n_db, s.d_act, nullptr, g->d_xp128, st);
    const bool crc_async = member_crc && g->bgzf_crc >= 2;
    if (crc_async) {
        CK(cudaStreamWaitEvent(g->crc_stream, s.inflated, 0));
        CK(cudaMemsetAsync(s.d_crc_bad, 0, sizeof(uint32_t), g->crc_stream));
        gz_launch_member_crc(s.d_text + ING_MAXCARRY, (const uint32_t *)(s.d_meta + ING_META_ISZ), (const uint32_t *)(s.d_meta + ING_META_CRC),
                             (const uint32_t *)(s.d_meta + ING_META_TOFF), (uint32_t)n_db, nullptr, s.d_crc_bad, g->d_xp128, g->crc_stream);
        CK(cudaEventRecord(s.crc_done, g->crc_stream));
    }
    tr_record(4, st);
    const double t_launch = ing_now();
    tr_decomp += t_launch - t_dec;
    uint8_t *d_text = s.d_text;
    uint8_t *d_text_next = g->slot[(g->n_chunks + 1) % ING_SLOTS].d_text;      // where a streamed file's next chunk will be inflated
    const bool detect = mode == ING_DETECT;
    // strain_detect wants per-read results: FASTA there means READS, two lines per record (what the reference's own target
    // metagenomes are, test/target_metagenomes.txt) and goes through the record kernels; anything else in a FASTA file
    // (wrapped sequences) is irregular to them.  The count path joins the lines of a FASTA record (genomes).
    const bool reads2 = fasta && detect;
    if (reads2) fasta = false;
    IngChunkArgs a;
    a.new_bytes = ch.text_len; a.first_chunk = ch.first ? 1u : 0u; a.last_chunk = ch.last ? 1u : 0u; a.inc = inc; a.fasta = fasta ? 1u : 0u;
    a.lsh = reads2 ? 1u : 2u;
    a.n_files = ch.n_files; a.n_dblocks = (unsigned)n_db;
    a.file_end = (const ull *)s.d_meta; a.isz = (const unsigned *)(s.d_meta + ING_META_ISZ); a.act = s.d_act;
    const unsigned n_blocks = (unsigned)(((size_t)ING_MAXCARRY + ch.text_len) / ING_BLOCK + 1);      // covers [0, t1] of the text buffer
    ing_index_count<<<n_blocks, ING_THREADS, 0, st>>>(d_text, g->d_state, g->d_block_nl, g->d_masks, a, g->max_lines, g->d_tickets + 0);
    ing_index_scatter<<<n_blocks, ING_THREADS, 0, st>>>(g->d_masks, g->d_block_nl, g->d_line_end, g->max_lines);
    if (want_result && ingest_result_claim(g)) return -1;
    IngResult *res = want_result ? g->d_results + g->res_seq % ING_MAX_RESULTS : nullptr;
    const S2DevBatch *dev = reinterpret_cast<const S2DevBatch *>(&g->d_state->flat_len);
    if (fasta) {
        ing_fasta_measure<<<n_blocks, ING_THREADS, 0, st>>>(d_text, g->d_state, g->d_block_nl, g->d_line_end, g->d_block_out, a, g->d_tickets + 1);
        ing_fasta_copy<<<n_blocks, ING_THREADS, 0, st>>>(d_text, d_text_next, g->d_state, g->d_block_nl, g->d_block_out, g->d_line_end, g->d_flat, res, g->d_tickets + 2);
        if (crc_async) { CK(cudaStreamWaitEvent(st, s.crc_done, 0)); ing_crc_veto<<<1, 1, 0, st>>>(g->d_state, s.d_crc_bad, res, a.last_chunk); }
        if (ingest_launch_count(g, t, dev, col)) return -1;
    } else {
        ing_fastq_measure<<<n_blocks, ING_THREADS, 0, st>>>(d_text, g->d_state, g->d_block_nl, g->d_line_end, g->d_block_out, a,
                                                             detect ? g->d_rec_off : nullptr, g->d_tickets + 1);
        ing_fastq_copy<<<n_blocks, ING_THREADS, 0, st>>>(d_text, d_text_next, g->d_state, g->d_block_nl, g->d_block_out, g->d_line_end, g->d_flat,
                                                          detect ? g->d_rec_off : nullptr, detect ? 0u : 1u, a.lsh, res, g->d_tickets + 2);
        if (crc_async) { CK(cudaStreamWaitEvent(st, s.crc_done, 0)); ing_crc_veto<<<1, 1, 0, st>>>(g->d_state, s.d_crc_bad, detect ? nullptr : res, a.last_chunk); }
        if (!detect) {
            if (ingest_launch_count(g, t, dev, col)) return -1;
        } else {
            const size_t max_rec = (size_t)g->max_lines / 2 + 4;
            CK(cudaMemsetAsync(g->d_hits_c, 0, max_rec * sizeof(unsigned), st));
            CK(cudaMemsetAsync(g->d_inf_c, 0, max_rec * sizeof(unsigned), st));
            CK(cudaMemsetAsync(g->d_cnt_c, 0, sizeof(ull), st));
            S2DetectOut out;
            out.rec_off = (const uint64_t *)g->d_rec_off; out.n_rec = 0; out.n_rec_dev = &g->d_state->n_rec;
            out.read_hits = g->d_hits_c; out.read_inf = g->d_inf_c; out.inf_pos = (uint64_t *)g->d_pos_c; out.inf_count = g->d_cnt_c; out.inf_cap = ING_CAP_C;
            s2_launch_scan_detect_dev(g->d_flat, dev, t->v, out, c->d_stats, c->grid_detect, st);
            ing_collect_inf<<<c->n_sm * 2, ING_THREADS, 0, st>>>(g->d_state, g->d_flat, g->d_rec_off, g->d_pos_c, g->d_cnt_c, ING_CAP_C,
                                                                  g->d_frec, g->d_foff, g->d_fkmer, g->d_fcnt, g->f_cap);
            ing_store_records<<<c->n_sm * 2, ING_THREADS, 0, st>>>(g->d_state, g->d_line_end, g->d_hits_c, g->d_inf_c, g->d_len_all, g->d_hits_all,
                                                                    g->d_inf_all, g->rec_cap, a.lsh);
            ing_finish<<<1, ING_THREADS, 0, st>>>(d_text, d_text_next, g->d_state, g->d_flat, 0u, res);
        }
    }
    tr_record(5, st);
    CK(cudaEventRecord(s.consumed, st));                   // the slot (compressed bytes, decompress list, meta, text) may be reused
    if (want_result) ++g->res_seq;
    CK(cudaGetLastError());
    tr_launch += ing_now() - t_launch;
    ++g->n_chunks;
    return 0;
}

// detect mode: the file-level per-record arrays grow with the text seen so far
static int ingest_grow_records(s2_ingest *g, const IngSource &src, ull text_total)
{
    if (text_total / 64 + 1024 <= g->rec_cap) return 0;      // records shorter than 64 bytes of text on average overflow the lists (-> host path)
    // sized once from the file's size where possible (FASTQ deflates about 5 : 1) instead of growing step by step
    const ssize_t fsize = src.size();
    const ull guess = fsize > 0 ? (ull)fsize * (src.bgzf || src.gz ? 6 : 1) / 64 + 4096 : 0;
    const ull want = std::max<ull>(std::max<ull>(text_total / 32 + 4096, g->rec_cap * 2), std::min<ull>(guess, 1ull << 28));
    unsigned *nl = nullptr, *nh = nullptr, *ni = nullptr;
    if (cudaMalloc((void **)&nl, want * 4) || cudaMalloc((void **)&nh, want * 4) || cudaMalloc((void **)&ni, want * 4)) { s2_set_error("out of device memory"); return -1; }
    CK(cudaStreamSynchronize(g->stream));
    if (g->rec_cap) {
        CK(cudaMemcpy(nl, g->d_len_all, g->rec_cap * 4, cudaMemcpyDeviceToDevice));
        CK(cudaMemcpy(nh, g->d_hits_all, g->rec_cap * 4, cudaMemcpyDeviceToDevice));
        CK(cudaMemcpy(ni, g->d_inf_all, g->rec_cap * 4, cudaMemcpyDeviceToDevice));
    }
    cudaFree(g->d_len_all); cudaFree(g->d_hits_all); cudaFree(g->d_inf_all);
    g->d_len_all = nl; g->d_hits_all = nh; g->d_inf_all = ni; g->rec_cap = want;
    return 0;
}

static int ingest_gz_stage_init(s2_ingest *g)
{
    GzStage &z = g->gz;
    if (z.d_res) return 0;
    z.comp_cap = (size_t)std::min<uint64_t>(std::max<uint64_t>(s2_env_u64("S2_GZ_BATCH_MB", 256), 1), 2048) << 20;
    z.sub_bytes = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(s2_env_u64("S2_GZ_SUB_KB", 64), 4), 4096) << 10;
    // symbols one sub-chunk may produce: S2_GZ_RATIO x its compressed bytes (FASTQ deflates 4-6 : 1, FASTA 3.5 : 1) plus the
    // run-on to the first block boundary behind the next cut; a sub-chunk that needs more makes its file the host reader's
    // (the run-on is one DEFLATE block: 16 K symbols, which are 30-100 KB of text for real reads and genomes but 300 KB for
    // FASTQ whose quality lines are all alike - every match 258 bytes long; hence the half million symbols of slack)
    z.sub_cap = z.sub_bytes * (uint32_t)std::min<uint64_t>(std::max<uint64_t>(s2_env_u64("S2_GZ_RATIO", 8), 2), 64) + (512u << 10) + 32768u;      // (+ the next region's marker prefix)
    z.max_files = ING_MAX_FILES;
    z.max_sub = (uint32_t)(z.comp_cap / z.sub_bytes) + z.max_files;
    for (auto &zs : z.set) {
        CK(cudaMalloc((void **)&zs.d_comp, z.comp_cap + (size_t)z.max_files * 32 + 256));
        CK(cudaHostAlloc((void **)&zs.h_files, (size_t)z.max_files * sizeof(GzFileDesc), cudaHostAllocDefault));
        CK(cudaMalloc((void **)&zs.d_files, (size_t)z.max_files * sizeof(GzFileDesc)));
        CK(cudaHostAlloc((void **)&zs.h_sub_file, (size_t)z.max_sub * sizeof(uint32_t), cudaHostAllocDefault));
        CK(cudaMalloc((void **)&zs.d_sub_file, (size_t)z.max_sub * sizeof(uint32_t)));
        CK(cudaHostAlloc((void **)&zs.h_slice0, ((size_t)z.max_files * 2 + 2) * sizeof(uint32_t), cudaHostAllocDefault));
        CK(cudaMalloc((void **)&zs.d_slice0, ((size_t)z.max_files * 2 + 2) * sizeof(uint32_t)));
        CK(cudaEventCreateWithFlags(&zs.idle, cudaEventDisableTiming));
    }
    // (the symbol area - 1 MB per sub-chunk - and the windows are sized by the batches that come: gz_stage_reserve.  Round 2's
    // first version took them for the largest batch there could be, 12 GB per pipeline, and three strain_detect workers
    // spent 0.8 s each in cudaMalloc - profiles/r2g_detect_bench.txt)
    CK(cudaMalloc((void **)&z.d_res, (size_t)z.max_sub * gz_sub_result_bytes()));
    CK(cudaMalloc((void **)&z.d_carry, 32768));
    CK(cudaMalloc((void **)&z.d_sub_off, (size_t)z.max_sub * sizeof(uint64_t)));
    CK(cudaMalloc((void **)&z.d_fres, (size_t)z.max_files * sizeof(GzFileResult)));
    CK(cudaMalloc((void **)&z.d_crc_acc, (size_t)z.max_files * sizeof(uint32_t)));
    CK(cudaMemset(z.d_crc_acc, 0, (size_t)z.max_files * sizeof(uint32_t)));
    return 0;
}

static int ingest_enqueue(s2_ingest *g, s2_table *t, IngSlot &s, const IngChunk &ch, bool bgzf, bool fasta, int mode, int col, unsigned inc, bool want_result);

// room for a batch of n_sub sub-chunks in n_files files (growing waits for whatever the stage still has in flight)
static int gz_stage_reserve(s2_ingest *g, uint32_t n_sub, uint32_t n_files)
{
    GzStage &z = g->gz;
    const size_t want_win = (size_t)n_sub + n_files + 1;
    if (n_sub <= z.sym_subs && want_win <= z.win_slots) return 0;
    CK(cudaStreamSynchronize(g->inflate_stream));
    if (n_sub > z.sym_subs) {
        const size_t subs = std::min<size_t>(z.max_sub, (size_t)n_sub + n_sub / 4 + 64);
        cudaFree(z.d_sym_alloc); z.d_sym = z.d_sym_alloc = nullptr; z.sym_subs = 0;
        if (cudaMalloc((void **)&z.d_sym_alloc, gz_sym_slots(subs, z.sub_cap) * sizeof(uint16_t)) != cudaSuccess) { cudaGetLastError(); s2_set_error("out of device memory (gunzip symbols)"); return -1; }
        z.d_sym = gz_launch_sym_init(z.d_sym_alloc, subs, z.sub_cap, g->inflate_stream);
        z.sym_subs = subs;
    }
    if (want_win > z.win_slots) {
        const size_t slots = std::min<size_t>((size_t)z.max_sub + z.max_files + 1, want_win + want_win / 4 + 64);
        cudaFree(z.d_win); z.d_win = nullptr; z.win_slots = 0;
        if (cudaMalloc((void **)&z.d_win, slots * 32768) != cudaSuccess) { cudaGetLastError(); s2_set_error("out of device memory (gunzip windows)"); return -1; }
        CK(cudaMemsetAsync(z.d_win, 0, slots * 32768, g->inflate_stream));
        z.win_slots = slots;
    }
    return 0;
}

// ---- one big ordinary .gz file, streamed: PIECES of S2_GZ_BATCH_MB through the gz stage, each piece's text in slices
// through the ring.  A piece is decoded like a batch of one file; its first sub-chunk finds its own block start, the
// chain continues where the previous piece ended (with that piece's last window), the stream's size and CRC-32 are
// checked when the last piece is decoded - by then the earlier pieces' text has been counted, so a mismatch there makes the
// caller take them back out with increment -1, as for any streamed file that turns irregular.  Same returns as ingest_stream.
static int ingest_stream_gz(s2_ingest *g, s2_table *t, const IngSource &src, int mode, int col, unsigned inc, IngResult *res, uint64_t *chunks_done)
{
    if (ingest_gz_stage_init(g)) return -1;
    GzStage &z = g->gz;
    GzBufSet &zs = z.set[0];                                                     // pieces go one at a time: one set is enough
    g->call_chunk0 = g->n_chunks;
    const size_t tail = std::min<size_t>(4u << 20, z.comp_cap / 4);             // bytes of the next piece the last sub-chunk may run on into
    const size_t piece_bytes = (z.comp_cap - tail) / z.sub_bytes * z.sub_bytes;
    const ssize_t size = src.size();
    {
        // the text of one piece: S2_GZ_RATIO x its compressed bytes (more makes the file the host reader's), for the pieces this file will have
        const uint64_t ratio = std::min<uint64_t>(std::max<uint64_t>(s2_env_u64("S2_GZ_RATIO", 8), 2), 64);
        const size_t want_cap = (size_t)std::min<uint64_t>((uint64_t)std::min<size_t>(piece_bytes, (size_t)std::max<ssize_t>(size, 0)) * ratio + (1u << 20), 4ull << 30);
        if (want_cap > z.piece_text_cap) {
            CK(cudaEventSynchronize(z.set[0].idle)); CK(cudaEventSynchronize(z.set[1].idle));
            CK(cudaStreamSynchronize(g->inflate_stream));
            cudaFree(z.d_piece_text); z.d_piece_text = nullptr; z.piece_text_cap = 0;
            if (cudaMalloc((void **)&z.d_piece_text, want_cap + 64) != cudaSuccess) { cudaGetLastError(); s2_set_error("out of device memory (gunzip text)"); return -1; }
            z.piece_text_cap = want_cap;
        }
        if (!z.h_fres) CK(cudaHostAlloc((void **)&z.h_fres, sizeof(GzFileResult), cudaHostAllocDefault));
    }
    uint8_t head[4096];
    const ssize_t hn = src.peek(head, sizeof head, 0);
    const uint64_t hl = hn > 0 ? s2_gzip_header_len(head, (uint64_t)hn) : 0;
    uint64_t n = 0;
    bool broken = size < 18 || hl == 0, first_chunk = true, finished = false;
    ull text_total = 0;
    uint64_t chain_abs = hl * 8;
    uint32_t crc_raw = 0;
    uint8_t *d_carry = z.d_carry;
    for (off_t base = 0; !broken && !finished && base < size; base += (off_t)piece_bytes) {
        const bool last_piece = (size_t)base + piece_bytes >= (size_t)size;
        const size_t want = std::min<size_t>(piece_bytes + tail, (size_t)size - (size_t)base);
        CK(cudaEventSynchronize(z.set[0].idle)); CK(cudaEventSynchronize(z.set[1].idle));      // the previous piece's (or batch's) text has left the stage
        if (gz_stage_reserve(g, (uint32_t)((std::min<size_t>(piece_bytes, (size_t)size - (size_t)base) + z.sub_bytes - 1) / z.sub_bytes), 1)) return -1;
        if (src.mem) {
            CK(cudaMemcpyAsync(zs.d_comp, src.mem + base, want, cudaMemcpyHostToDevice, g->copy_stream));
        } else {
            if (!zs.h_comp) CK(cudaHostAlloc((void **)&zs.h_comp, z.comp_cap + (size_t)z.max_files * 32 + 256, cudaHostAllocDefault));
            if (ing_pread(src.fd, zs.h_comp, want, base) != (ssize_t)want) { s2_set_error("read failed"); return -1; }
            CK(cudaMemcpyAsync(zs.d_comp, zs.h_comp, want, cudaMemcpyHostToDevice, g->copy_stream));
        }
        CK(cudaMemsetAsync(zs.d_comp + want, 0, 64, g->copy_stream));
        GzFileDesc &d = zs.h_files[0];
        d.comp_off = 0; d.comp_len = want; d.first_bit = base == 0 ? hl * 8 : ~0ull; d.chain_bit = chain_abs - (uint64_t)base * 8;
        d.text_off = 0; d.text_len = z.piece_text_cap; d.text_before = text_total; d.sub0 = 0;
        d.n_sub = (uint32_t)((std::min<size_t>(piece_bytes, (size_t)size - (size_t)base) + z.sub_bytes - 1) / z.sub_bytes);
        d.piece = last_piece ? 2u : 1u; d.pad_ = 0;
        for (uint32_t k = 0; k < d.n_sub; ++k) zs.h_sub_file[k] = 0;
        zs.h_slice0[0] = 0; zs.h_slice0[1] = 0xFFFFFFFFu;
        CK(cudaMemcpyAsync(zs.d_files, zs.h_files, sizeof(GzFileDesc), cudaMemcpyHostToDevice, g->copy_stream));
        CK(cudaMemcpyAsync(zs.d_sub_file, zs.h_sub_file, (size_t)d.n_sub * sizeof(uint32_t), cudaMemcpyHostToDevice, g->copy_stream));
        CK(cudaMemcpyAsync(zs.d_slice0, zs.h_slice0, 2 * sizeof(uint32_t), cudaMemcpyHostToDevice, g->copy_stream));
        cudaEvent_t up = g->slot[0].h2d_done;
        CK(cudaEventRecord(up, g->copy_stream));
        CK(cudaStreamWaitEvent(g->inflate_stream, up, 0));
        if (base) CK(cudaMemcpyAsync(z.d_win, d_carry, 32768, cudaMemcpyDeviceToDevice, g->inflate_stream));      // the window the previous piece left
        gz_launch_decode(zs.d_comp, zs.d_files, zs.d_sub_file, d.n_sub, z.sub_bytes, z.d_sym, z.sub_cap, z.d_res, g->inflate_stream);
        gz_launch_chain(zs.d_comp, zs.d_files, 1, z.d_sym, z.sub_cap, z.d_res, z.d_win, z.d_sub_off, z.d_fres, g->inflate_stream);
        gz_launch_translate(zs.d_files, zs.d_sub_file, 0, d.n_sub, z.d_sym, z.sub_cap, z.d_win, z.d_sub_off, z.d_fres, z.d_piece_text, g->inflate_stream);
        gz_launch_crc(zs.d_files, 0, 1, zs.d_slice0, (uint32_t)(z.piece_text_cap / 4096) + 1, z.d_piece_text, z.d_fres, z.d_crc_acc, nullptr, g->inflate_stream);
        CK(cudaMemcpyAsync(d_carry, z.d_win + (size_t)d.n_sub * 32768, 32768, cudaMemcpyDeviceToDevice, g->inflate_stream));
        CK(cudaMemcpyAsync(z.h_fres, z.d_fres, sizeof(GzFileResult), cudaMemcpyDeviceToHost, g->inflate_stream));
        CK(cudaStreamSynchronize(g->inflate_stream));
        const GzFileResult fr = *z.h_fres;
        if (fr.status != 0 || fr.text_len > z.piece_text_cap) { broken = true; break; }       // (the translate kernel wrote nothing past the symbols' own count)
        chain_abs = (uint64_t)base * 8 + fr.end_bit;
        crc_raw = gz_crc_append(crc_raw, fr.crc_raw, fr.text_len);
        const ull L = fr.text_len;
        if (last_piece) {
            // size and CRC-32 of the whole stream: a mismatch is known before the last piece's text is counted - the earlier
            // pieces were, and the caller takes them back out (same pieces, increment -1)
            if (gz_crc_finish(crc_raw, text_total + L) != fr.crc) { broken = true; break; }
            finished = true;
        }
        // the piece's text, in slices, through the ring
        ull off = 0;
        do {
            IngSlot *sp;
            if (ingest_slot_begin(g, &sp)) return -1;
            IngChunk ch;
            ch.dev_src = z.d_piece_text + off;
            ch.text_len = (size_t)std::min<ull>(g->text_cap, L - off);
            ch.comp_len = 0; ch.first = first_chunk; ch.last = last_piece && off + ch.text_len == L; ch.n_files = 0;
            if (!ch.text_len && !ch.last) break;
            if (mode == ING_DETECT && ingest_grow_records(g, src, text_total + off + ch.text_len)) return -1;
            if (ingest_enqueue(g, t, *sp, ch, false, src.fasta, mode, col, inc, ch.last)) return -1;
            first_chunk = false;
            ++n;
            off += ch.text_len;
        } while (off < L);
        CK(cudaEventRecord(zs.idle, g->inflate_stream));
        text_total += L;
    }
    if (chunks_done) *chunks_done = n;
    if (broken || !finished) {
        CK(cudaMemsetAsync(g->d_state, 0, sizeof(IngState), g->stream));        // no last chunk came to tidy up
        CK(cudaStreamSynchronize(g->stream));
        memset(res, 0, sizeof *res);
        res->irregular = 1;
        return 1;
    }
    CK(cudaStreamSynchronize(g->stream));
    *res = ingest_result_take(g, g->res_seq - 1);                   // the last chunk's verdict
    return 0;
}

// ---- a file too big for one chunk: streamed through the ring ----------------------------------------------
// Returns 0 ok (verdict in *res), 1 not BGZF after all (nothing enqueued for the offending chunk; the verdict then
// says irregular), -1 error.  Synchronous at the end.  *chunks_done: chunks that went to the device.
static int ingest_stream(s2_ingest *g, s2_table *t, const IngSource &src, int mode, int col, unsigned inc, IngResult *res, uint64_t *chunks_done)
{
    if (src.gz) return ingest_stream_gz(g, t, src, mode, col, inc, res, chunks_done);
    g->call_chunk0 = g->n_chunks;
    off_t file_off = 0;
    bool first = true, eof = false, broken = false;
    uint64_t n = 0;
    ull text_total = 0;
    while (!eof) {
        IngSlot *sp;
        if (ingest_slot_begin(g, &sp)) return -1;
        IngSlot &s = *sp;
        const uint8_t *h;
        ssize_t got;
        const size_t want_bytes = ingest_chunk_cap(g);
        if (src.mem) {                                                           // caller's buffer: no staging copy
            h = src.mem + file_off;
            got = (size_t)file_off < src.mem_len ? (ssize_t)std::min<size_t>(want_bytes, src.mem_len - (size_t)file_off) : 0;
        } else {
            if (ingest_staging(s, g->comp_chunk)) return -1;
            got = ing_pread(src.fd, s.h_comp, want_bytes, file_off);
            if (got < 0) { s2_set_error("read failed"); return -1; }
            h = s.h_comp;
        }
        IngChunk ch;
        size_t used;
        bool full = false;
        if (src.bgzf) {
            used = ingest_walk_bgzf(g, s, h, (size_t)got, 0, &ch.text_len, &full);
            if (used == (size_t)-1 || (used == 0 && got > 0)) { broken = true; break; }       // not BGZF after all / a member larger than a chunk
        } else {
            used = std::min<size_t>((size_t)got, g->text_cap);
            full = used < (size_t)got;
            ch.text_len = used;
        }
        eof = (size_t)got < want_bytes && used == (size_t)got && !full;        // short read and everything consumed
        ch.comp_len = used; ch.first = first; ch.last = eof; ch.n_files = 0;
        file_off += (off_t)used;
        text_total += ch.text_len;
        if (mode == ING_DETECT && ingest_grow_records(g, src, text_total)) return -1;
        if (used) CK(cudaMemcpyAsync(s.d_comp, h, used, cudaMemcpyHostToDevice, g->copy_stream));
        if (ingest_enqueue(g, t, s, ch, src.bgzf, src.fasta, mode, col, inc, eof)) return -1;
        first = false;
        ++n;
        if (got == 0) break;
    }
    if (chunks_done) *chunks_done = n;
    if (broken) {
        CK(cudaMemsetAsync(g->d_state, 0, sizeof(IngState), g->stream));        // no last chunk came to tidy up
        CK(cudaStreamSynchronize(g->stream));
        memset(res, 0, sizeof *res);
        res->irregular = 1;
        return 1;
    }
    CK(cudaStreamSynchronize(g->stream));
    *res = ingest_result_take(g, g->res_seq - 1);                   // the last chunk's verdict
    return 0;
}

// ---- count: any number of sources, small ones grouped ------------------------------------------------------
// A job is one call's worth of sources.  submit() classifies them, packs the small ones into groups and enqueues the
// groups (asynchronous: it returns while the device works); finish() waits for the job's last group, reads the
// verdicts, runs the members of irregular groups one by one and streams the files that are too big for a chunk.
// Several jobs of one thread may be in flight: the next one is submitted before the previous one is finished, which
// hides the pipeline's fill and drain.
struct IngGroup { std::vector<int> members; uint64_t result = 0; };

struct s2_ingest_job {
    s2_ctx *c = nullptr; s2_table *t = nullptr; s2_ingest *g = nullptr;
    int col = 0;
    std::vector<IngSource> srcs;
    std::vector<int> own_fds;                       // opened by the job (file form): closed when the job ends
    std::vector<int> rc;                            // per source: 0 handled / 1 not handled
    std::vector<IngGroup> groups;                   // enqueued, verdict not read yet
    std::vector<int> streamed, retry;
    uint64_t tot_bases = 0, tot_lookups = 0;
    cudaEvent_t done = nullptr;                     // after the last chunk submit() enqueued
    // the group being assembled in the current slot
    IngSlot *s = nullptr;
    IngGroup cur;
    IngChunk ch;
    bool cur_fasta = false, cur_bgzf = false;
    std::vector<int> gz_list;                       // ordinary .gz sources: decoded in batches by the pipeline's gz stage (gz_run)
    struct { const uint8_t *h = nullptr; size_t d_off = 0, len = 0; } pend;     // host -> device copies of neighbouring sources are merged

    ~s2_ingest_job()
    {
        if (done) cudaEventDestroy(done);
        for (int fd : own_fds) close(fd);
    }
    void add_totals(const IngResult &r, bool fasta)
    {
        tot_bases += r.bases;
        // FASTA records are not measured one by one on the device: every record is assumed to have at least one window
        tot_lookups += fasta ? (r.bases > 30 * r.records ? r.bases - 30 * r.records : 0) : r.lookups;
    }
    int push_copy()
    {
        if (pend.len) {
            const double t_cp = ing_now();
            CK(cudaMemcpyAsync(s->d_comp + pend.d_off, pend.h, pend.len, cudaMemcpyHostToDevice, g->copy_stream));
            tr_h2d += ing_now() - t_cp;
        }
        pend.len = 0;
        return 0;
    }
    int flush()
    {
        if (!s || cur.members.empty()) return 0;
        if (push_copy()) return -1;
        ch.first = true; ch.last = true; ch.n_files = (unsigned)cur.members.size(); ch.gz = false;
        cur.result = g->res_seq;
        if (ingest_enqueue(g, t, *s, ch, cur_bgzf, cur_fasta, ING_COUNT, col, 1u, true)) return -1;
        groups.push_back(cur);
        cur = IngGroup(); ch = IngChunk(); s = nullptr;
        return 0;
    }
    // verdicts of everything enqueued so far (the whole pipeline is idle afterwards)
    int harvest()
    {
        if (flush()) return -1;
        CK(cudaStreamSynchronize(g->stream));
        return read_verdicts();
    }
    int read_verdicts()
    {
        for (auto &gr : groups) {
            const IngResult r = ingest_result_take(g, gr.result);
            if (!r.irregular) add_totals(r, srcs[gr.members[0]].fasta);
            else if (gr.members.size() == 1) rc[gr.members[0]] = 1;
            else retry.insert(retry.end(), gr.members.begin(), gr.members.end());
        }
        groups.clear();
        return 0;
    }
    // ---- ordinary .gz sources: batches through the gz stage, then ordinary groups of whole files ---------------------
    // the listed sources (all ordinary .gz, classified).  Files that cannot go this way (larger than a batch or than a
    // chunk's text, no readable gzip header) are left to the host reader (rc stays 1).  Returns 0 / -1.
    int gz_run(const std::vector<int> &list)
    {
        if (flush()) return -1;                                     // the group being assembled goes first
        if (ingest_gz_stage_init(g)) return -1;
        GzStage &z = g->gz;
        size_t at = 0;
        while (at < list.size()) {
            // ---- plan one batch: files [at, end) ------------------------------------------------------------------------
            size_t comp_used = 0, end = at;
            uint32_t n_sub = 0, n_files = 0;
            std::vector<int> members;
            for (; end < list.size(); ++end) {
                IngSource &src = srcs[list[end]];
                const ssize_t size = src.size();
                const uint32_t subs = size > 0 ? (uint32_t)(((size_t)size + z.sub_bytes - 1) / z.sub_bytes) : 0;
                if (size < 18) { rc[list[end]] = 1; continue; }                               // host reader
                // too big for a group of whole files (ISIZE is the size modulo 4 GiB: trusted only for small files): streamed in pieces
                if ((size_t)size > z.comp_cap / 2 || (size_t)src.gz_isize > g->text_cap) { streamed.push_back(list[end]); continue; }
                const size_t need = ((size_t)size + 15) / 16 * 16 + 16;
                if (n_files && (comp_used + need > z.comp_cap || n_sub + subs > z.max_sub || n_files + 1 > z.max_files)) break;
                comp_used += need; n_sub += subs; ++n_files;
                members.push_back(list[end]);
            }
            at = end;
            if (members.empty()) continue;
            // ---- bytes: caller's memory goes straight to the device, files through the stage's pinned buffer ----------
            GzBufSet &zs = z.set[z.batch_no & 1u];                  // (the other set belongs to the batch before this one, which may still be decoding)
            const unsigned set_no = z.batch_no & 1u;
            ++z.batch_no;
            CK(cudaEventSynchronize(zs.idle));                      // the batch before the previous one is out of this set's buffers
            if (gz_stage_reserve(g, n_sub, n_files)) return -1;
            if (!srcs[members[0]].mem && !zs.h_comp) CK(cudaHostAlloc((void **)&zs.h_comp, z.comp_cap + (size_t)z.max_files * 32 + 256, cudaHostAllocDefault));
            CK(cudaMemsetAsync(zs.d_comp, 0, comp_used + 64, g->copy_stream));              // zero padding behind every file
            size_t off = 0;
            uint32_t sub0 = 0, nf = 0;
            std::vector<int> good;
            for (int i : members) {
                IngSource &src = srcs[i];
                const size_t size = (size_t)src.size();
                const uint8_t *h = src.mem;
                if (!src.mem) {
                    if (ing_pread(src.fd, zs.h_comp + off, size, 0) != (ssize_t)size) { s2_set_error("read failed"); return -1; }
                    memset(zs.h_comp + off + size, 0, (size + 15) / 16 * 16 + 16 - size);
                    h = zs.h_comp + off;
                }
                const uint64_t hl = s2_gzip_header_len(h, size);
                if (!hl) { rc[i] = 1; continue; }                   // not a gzip member after all
                const uint32_t subs = (uint32_t)((size + z.sub_bytes - 1) / z.sub_bytes);
                GzFileDesc &d = zs.h_files[nf];
                d.comp_off = off; d.comp_len = size; d.first_bit = hl * 8; d.chain_bit = hl * 8; d.text_off = 0; d.text_len = src.gz_isize; d.text_before = 0;
                d.sub0 = sub0; d.n_sub = subs; d.piece = 0; d.pad_ = 0;
                for (uint32_t k = 0; k < subs; ++k) zs.h_sub_file[sub0 + k] = nf;
                if (src.mem) CK(cudaMemcpyAsync(zs.d_comp + off, h, size, cudaMemcpyHostToDevice, g->copy_stream));
                off += (size + 15) / 16 * 16 + 16;
                sub0 += subs; ++nf;
                good.push_back(i);
            }
            if (good.empty()) continue;
            if (!srcs[good[0]].mem) CK(cudaMemcpyAsync(zs.d_comp, zs.h_comp, off, cudaMemcpyHostToDevice, g->copy_stream));
            // ---- plan the pipeline chunks: groups of whole files of one kind whose texts fit a chunk --------------------
            struct Plan { uint32_t f0, f1, sub_lo, sub_hi, slice0, slices; size_t text; bool fasta; };
            std::vector<Plan> plans;
            uint32_t slice_at = 0;
            for (uint32_t f = 0; f < nf;) {
                Plan p; p.f0 = f; p.sub_lo = zs.h_files[f].sub0; p.text = 0; p.fasta = srcs[good[f]].fasta; p.slice0 = slice_at + (uint32_t)plans.size();
                uint32_t local = 0;
                while (f < nf && srcs[good[f]].fasta == p.fasta && p.text + zs.h_files[f].text_len <= g->text_cap && f - p.f0 < ING_MAX_FILES) {
                    zs.h_files[f].text_off = p.text;
                    zs.h_slice0[p.slice0 + (f - p.f0)] = local;
                    local += (uint32_t)((zs.h_files[f].text_len + 4095) / 4096);
                    p.text += zs.h_files[f].text_len;
                    ++f;
                }
                zs.h_slice0[p.slice0 + (f - p.f0)] = local;
                p.f1 = f; p.sub_hi = f < nf ? zs.h_files[f].sub0 : sub0; p.slices = local;
                slice_at += f - p.f0;
                plans.push_back(p);
            }
            CK(cudaMemcpyAsync(zs.d_files, zs.h_files, (size_t)nf * sizeof(GzFileDesc), cudaMemcpyHostToDevice, g->copy_stream));
            CK(cudaMemcpyAsync(zs.d_sub_file, zs.h_sub_file, (size_t)sub0 * sizeof(uint32_t), cudaMemcpyHostToDevice, g->copy_stream));
            CK(cudaMemcpyAsync(zs.d_slice0, zs.h_slice0, ((size_t)nf + plans.size()) * sizeof(uint32_t), cudaMemcpyHostToDevice, g->copy_stream));
            // ---- decode + chain of the whole batch, on the inflate stream -------------------------------------------------
            cudaEvent_t up = g->slot[0].h2d_done;                   // (any event will do: recorded and waited for right here)
            CK(cudaEventRecord(up, g->copy_stream));
            CK(cudaStreamWaitEvent(g->inflate_stream, up, 0));
            gz_launch_decode(zs.d_comp, zs.d_files, zs.d_sub_file, sub0, z.sub_bytes, z.d_sym, z.sub_cap, z.d_res, g->inflate_stream);
            gz_launch_chain(zs.d_comp, zs.d_files, nf, z.d_sym, z.sub_cap, z.d_res, z.d_win, z.d_sub_off, z.d_fres, g->inflate_stream);
            CK(cudaGetLastError());
            // ---- one ordinary group per plan ------------------------------------------------------------------------------
            for (const Plan &p : plans) {
                if (!groups.empty() && g->res_seq - groups.front().result + 4 >= ING_MAX_RESULTS && harvest()) return -1;
                IngSlot *sl;
                if (ingest_slot_begin(g, &sl)) return -1;
                IngChunk c2;
                c2.comp_len = 0; c2.text_len = p.text; c2.first = true; c2.last = true; c2.n_files = p.f1 - p.f0; c2.gz = true;
                c2.gz_set = set_no; c2.gz_file0 = p.f0; c2.gz_sub_lo = p.sub_lo; c2.gz_sub_hi = p.sub_hi; c2.gz_slice0 = p.slice0; c2.gz_slices = p.slices;
                IngGroup gr;
                for (uint32_t f = p.f0; f < p.f1; ++f) {
                    ((ull *)sl->h_meta)[f - p.f0] = zs.h_files[f].text_off + zs.h_files[f].text_len;                    // where file f's text ends
                    ((unsigned *)(sl->h_meta + ING_META_ISZ))[f - p.f0] = (unsigned)zs.h_files[f].text_len;       // isz
                    gr.members.push_back(good[f]);
                }
                gr.result = g->res_seq;
                if (ingest_enqueue(g, t, *sl, c2, false, p.fasta, ING_COUNT, col, 1u, true)) return -1;
                groups.push_back(gr);
            }
            CK(cudaEventRecord(zs.idle, g->inflate_stream));
        }
        return 0;
    }
    // 0 added, 1 does not fit one chunk (stream it), 2 not BGZF after all, -1 error
    int add_to_group(int i, bool alone)
    {
        IngSource &src = srcs[i];
        const ssize_t size = src.size();
        if (size < 0) return 2;
        if (src.gz) return gz_run(std::vector<int>(1, i));                                            // (a retry of one member of an irregular gz group)
        const size_t cap_max = src.bgzf ? g->comp_chunk : std::min(g->comp_chunk, g->text_cap);
        if ((size_t)size > cap_max) return 1;
        for (int attempt = 0; attempt < 2; ++attempt) {
            // a group closes when the next file would push it over the (ramping) chunk size; a single file may exceed the ramp
            const size_t cap = std::min(cap_max, std::max(ingest_chunk_cap(g), (size_t)size));
            if (s && (alone || cur_fasta != src.fasta || cur_bgzf != src.bgzf || ch.comp_len + (size_t)size > cap || cur.members.size() >= ING_MAX_FILES)) { if (flush()) return -1; }
            if (!s) {
                // the verdict ring must not wrap onto verdicts this job has not read yet
                if (!groups.empty() && g->res_seq - groups.front().result + 4 >= ING_MAX_RESULTS && harvest()) return -1;
                if (ingest_slot_begin(g, &s)) return -1;
                cur_fasta = src.fasta; cur_bgzf = src.bgzf;
            }
            const uint8_t *h = src.mem;
            if (!src.mem) {
                if (ingest_staging(*s, g->comp_chunk)) return -1;
                if (ing_pread(src.fd, s->h_comp + ch.comp_len, (size_t)size, 0) != size) { s2_set_error("read failed"); return -1; }
                h = s->h_comp + ch.comp_len;
            }
            size_t text_len = ch.text_len;
            const size_t n_params = s->params.size();
            if (src.bgzf) {
                bool full = false;
                const size_t used = ingest_walk_bgzf(g, *s, h, (size_t)size, ch.comp_len, &text_len, &full);
                if (full) {                                      // the group's text or block list is full: close it and retry in a fresh one
                    s->params.resize(n_params);
                    if (cur.members.empty()) { s = nullptr; return 1; }          // does not fit a chunk on its own (nothing was enqueued on this slot)
                    if (flush()) return -1;
                    continue;
                }
                if (used != (size_t)size) { s->params.resize(n_params); if (cur.members.empty()) s = nullptr; return 2; }      // trailing garbage / truncated member
            } else {
                if (text_len + (size_t)size > g->text_cap) { if (cur.members.empty()) { s = nullptr; return 1; } if (flush()) return -1; continue; }
                text_len += (size_t)size;
            }
            if (pend.len && pend.h + pend.len == h && pend.d_off + pend.len == ch.comp_len && pend.len < (8u << 20)) pend.len += (size_t)size;
            else { if (push_copy()) return -1; pend.h = h; pend.d_off = ch.comp_len; pend.len = (size_t)size; }
            ch.comp_len += (size_t)size;
            ch.text_len = text_len;
            ((ull *)s->h_meta)[cur.members.size()] = text_len;
            cur.members.push_back(i);
            if (alone && flush()) return -1;
            return 0;
        }
        return 1;
    }
};

// whatever way a job fails, nothing it enqueued may still read the caller's memory afterwards
static void ingest_quiesce(s2_ingest *g)
{
    if (!g) return;
    cudaStreamSynchronize(g->copy_stream); cudaStreamSynchronize(g->inflate_stream); cudaStreamSynchronize(g->stream);
    cudaGetLastError();
}

// Measured (profiles/r1s_hw_decompression_error_probe.txt): a DEFLATE stream the engine cannot decode - not DEFLATE at
// all, cut short, or followed by other bytes inside the stated length - is not reported per stream: it surfaces as a
// sticky cudaErrorLaunchFailure and the CUDA context is lost.  The host walk only hands over complete BGZF members,
// so this takes a member whose payload is damaged.  Say so (the callers' generic message would be a bare CUDA error)
// and remember it: the executables then start over with host inflate (main_*.c: execv with S2_GPU_INGEST=0), where zlib names the file.
static std::atomic<int> g_engine_failed{0};
static void ingest_engine_failed_message(void)
{
    s2_set_error("the hardware decompression engine met a DEFLATE block it cannot decode (damaged data inside a BGZF member) and the CUDA "
                 "context is lost; S2_GPU_INGEST=0 inflates with zlib on the host instead, which names the damaged file");
}
static void ingest_explain_failure(void)
{
    if (cudaDeviceSynchronize() != cudaErrorLaunchFailure) { cudaGetLastError(); return; }
    g_engine_failed.store(1);
    ingest_engine_failed_message();
}

// also puts the explanation back into the calling thread's error string (other threads fail with bare CUDA errors once
// the context is gone; the executables report this cause instead of whichever error came first)
extern "C" int s2_ingest_engine_failed(void)
{
    if (!g_engine_failed.load()) return 0;
    ingest_engine_failed_message();
    return 1;
}

static int ingest_job_submit(s2_ingest_job *job)
{
    s2_ingest *g = job->g;
    const int n = (int)job->srcs.size();
    // small first chunks only when the pipeline is idle: behind a job that is still running there is nothing to start early
    const bool idle = cudaStreamQuery(g->stream) == cudaSuccess;
    cudaGetLastError();
    g->call_chunk0 = idle || g->n_chunks < 6 ? g->n_chunks : g->n_chunks - 6;
    const bool trace = s2_env_int("S2_INGEST_TRACE", 0) != 0;
    tr_wait = tr_h2d = tr_decomp = tr_launch = 0;
    tr_on = trace && s2_env_int("S2_INGEST_TRACE", 0) >= 2;
    double us_classify = 0, us_group = 0;
    const double t_begin = ing_now();
    for (int i = 0; i < n; ++i) {
        const double ta = ing_now();
        ingest_classify(job->srcs[i]);               // one file at a time, so that the first group is on its way while the rest is looked at
        if (job->srcs[i].bgzf && !g->hw_deflate) job->srcs[i].eligible = false;
        job->rc[i] = job->srcs[i].eligible ? 0 : 1;
        const double tb = ing_now();
        us_classify += tb - ta;
        if (!job->srcs[i].eligible) continue;
        if (job->srcs[i].gz) { job->gz_list.push_back(i); continue; }
        const int rc = job->add_to_group(i, false);
        us_group += ing_now() - tb;
        if (rc < 0) return -1;
        if (rc == 1) job->streamed.push_back(i);
        if (rc == 2) job->rc[i] = 1;
    }
    if (job->flush()) return -1;
    if (!job->gz_list.empty() && job->gz_run(job->gz_list)) return -1;
    CK(cudaEventRecord(job->done, g->stream));
    if (trace) fprintf(stderr, "[s2 ingest] %d sources submitted: classify %.0f us, group+enqueue %.0f us (ring wait %.0f, H2D calls %.0f, inflate calls %.0f, launches %.0f), "
                       "enqueue done at %.0f us, chunks so far %llu\n",
                       n, us_classify, us_group, tr_wait, tr_h2d, tr_decomp, tr_launch, ing_now() - t_begin, (unsigned long long)g->n_chunks);
    return 0;
}

static void ingest_trace_timeline(void)
{
    if (!tr_on) return;                              // device timeline of every chunk, relative to the first copy
    for (size_t k = 0; k < tr_events.size(); ++k) {
        float t[6] = { 0, 0, 0, 0, 0, 0 };
        for (int j = 0; j < 6; ++j) {
            const cudaError_t e = cudaEventElapsedTime(&t[j], tr_events[0].e[0], tr_events[k].e[j]);
            if (e != cudaSuccess) { fprintf(stderr, "[s2 ingest]   chunk %zu event %d: %s\n", k, j, cudaGetErrorName(e)); cudaGetLastError(); }
        }
        fprintf(stderr, "[s2 ingest]   chunk %zu (%5.1f MB -> %5.1f MB): copy %7.0f..%7.0f  inflate %7.0f..%7.0f  kernels %7.0f..%7.0f us\n", k,
                        tr_events[k].comp / 1e6, tr_events[k].text / 1e6, t[0] * 1e3, t[1] * 1e3, t[2] * 1e3, t[3] * 1e3, t[4] * 1e3, t[5] * 1e3);
    }
    for (auto &ev : tr_events) for (auto &e : ev.e) cudaEventDestroy(e);
    cudaGetLastError();
    tr_events.clear();
    tr_on = false;
}

static int ingest_job_finish(s2_ingest_job *job)
{
    s2_ingest *g = job->g;
    s2_table *t = job->t;
    const double t0 = ing_now();
    CK(cudaSetDevice(g->device));
    CK(cudaEventSynchronize(job->done));             // the job's last group is through (later jobs may still be running)
    if (job->read_verdicts()) return -1;
    if (s2_env_int("S2_INGEST_TRACE", 0)) fprintf(stderr, "[s2 ingest] verdicts after another %.0f us of waiting\n", ing_now() - t0);
    ingest_trace_timeline();
    // members of irregular groups, one by one (a group's verdict precedes its scan: nothing of it was counted)
    std::vector<int> again;
    again.swap(job->retry);
    if (again.empty() && job->streamed.empty()) return 0;
    std::lock_guard<std::mutex> pipeline_lock(g->mu);              // more chunks to enqueue: the pipeline is ours again
    for (int i : again) {
        const int rc = job->add_to_group(i, true);
        if (rc < 0) return -1;
        if (rc) job->rc[i] = 1;
    }
    if (!again.empty() && job->harvest()) return -1;
    // big files
    for (int i : job->streamed) {
        IngResult r; uint64_t done = 0;
        const int rc = ingest_stream(g, t, job->srcs[i], ING_COUNT, job->col, 1u, &r, &done);
        if (rc < 0) return -1;
        if (r.irregular) {
            // chunks before the first irregular one were counted: replay with increment -1 (same verdicts, same chunks)
            if (done > 1 || (rc == 1 && done > 0)) {
                IngResult r2; uint64_t d2 = 0;
                if (ingest_stream(g, t, job->srcs[i], ING_COUNT, job->col, 0xFFFFFFFFu, &r2, &d2) < 0) return -1;
            }
            job->rc[i] = 1;
        } else {
            job->add_totals(r, job->srcs[i].fasta);
        }
    }
    return 0;
}

static s2_ingest_job *ingest_job_new(s2_ctx *c, s2_table *t, int col, size_t n)
{
    if (col < 0 || col >= t->v.n_cols) { s2_set_error("column out of range"); return nullptr; }
    s2_ingest *g = ingest_acquire(c);                   // locked; the submit calls unlock it when everything is enqueued
    if (!g) return nullptr;
    s2_ingest_job *job = new s2_ingest_job();
    job->c = c; job->t = t; job->g = g; job->col = col;
    job->srcs.resize(n); job->rc.assign(n, 1);
    if (cudaEventCreateWithFlags(&job->done, cudaEventDisableTiming) != cudaSuccess) { s2_set_error("cannot create an event"); g->mu.unlock(); delete job; return nullptr; }
    return job;
}

// Asynchronous form: submit returns as soon as everything that fits a chunk is enqueued (the images must stay valid
// until the job has been waited for); wait blocks until the job's files are counted (or handed back) and frees the job.
extern "C" s2_ingest_job *s2_ingest_submit_mem_batch(s2_ctx *c, s2_table *t, const void *const *images, const uint64_t *n_bytes, int n, int col)
{
    s2_ingest_job *job = ingest_job_new(c, t, col, (size_t)std::max(n, 0));
    if (!job) return nullptr;
    for (int i = 0; i < n; ++i) { job->srcs[i].mem = (const uint8_t *)images[i]; job->srcs[i].mem_len = (size_t)n_bytes[i]; }
    if (ingest_job_submit(job)) { ingest_quiesce(job->g); job->g->mu.unlock(); ingest_explain_failure(); delete job; return nullptr; }
    job->g->mu.unlock();
    return job;
}

extern "C" s2_ingest_job *s2_ingest_submit_files(s2_ctx *c, s2_table *t, const char *const *paths, int n, int col)
{
    s2_ingest_job *job = ingest_job_new(c, t, col, (size_t)std::max(n, 0));
    if (!job) return nullptr;
    for (int i = 0; i < n; ++i) {
        const int fd = open(paths[i], O_RDONLY);
        job->srcs[i].fd = fd;                        // -1: not eligible (classification reads nothing)
        if (fd >= 0) job->own_fds.push_back(fd);
    }
    if (ingest_job_submit(job)) { ingest_quiesce(job->g); job->g->mu.unlock(); ingest_explain_failure(); delete job; return nullptr; }
    job->g->mu.unlock();
    return job;
}

extern "C" int s2_ingest_wait(s2_ingest_job *job, int *rc_each, uint64_t *bases, uint64_t *lookups)
{
    if (!job) { s2_set_error("no job"); return -1; }
    const int rc = ingest_job_finish(job);
    if (rc) { ingest_quiesce(job->g); ingest_explain_failure(); }
    for (size_t i = 0; i < job->rc.size(); ++i) if (rc_each) rc_each[i] = rc ? 1 : job->rc[i];
    if (bases) *bases = job->tot_bases;
    if (lookups) *lookups = job->tot_lookups;
    delete job;
    return rc;
}

// GEN_calculate_kmer_count for one file, entirely on the GPU when the file is BGZF-compressed (or uncompressed)
// strict FASTQ / FASTA.  Returns 0 = done (counters updated, *bases / *lookups set), 1 = not handled (nothing was
// counted: use the host reader), -1 = error.
extern "C" int s2_ingest_count_file(s2_ctx *c, s2_table *t, const char *path, int col, uint64_t *bases, uint64_t *lookups)
{
    int rc_each = 1;
    const int rc = s2_ingest_count_files(c, t, &path, 1, col, &rc_each, bases, lookups);
    return rc < 0 ? rc : rc_each;
}

// The same for a file image that is already in host memory (the bytes of a BGZF or plain FASTA/FASTQ file; pinned
// memory gives the full PCIe rate).  Only the compressed bytes cross PCIe.
extern "C" int s2_ingest_count_mem(s2_ctx *c, s2_table *t, const void *image, uint64_t n_bytes, int col, uint64_t *bases, uint64_t *lookups)
{
    int rc_each = 1;
    const int rc = s2_ingest_count_mem_batch(c, t, &image, &n_bytes, 1, col, &rc_each, bases, lookups);
    return rc < 0 ? rc : rc_each;
}

// Many files in one call: small files travel in groups (one launch sequence per group), verdicts are collected at
// the end.  rc_each[i] = 0 done / 1 not handled (nothing of file i was counted).  *bases / *lookups: totals of the
// files that were handled.  Returns 0, or -1 on error (then the counters are undefined).
extern "C" int s2_ingest_count_mem_batch(s2_ctx *c, s2_table *t, const void *const *images, const uint64_t *n_bytes, int n, int col,
                                         int *rc_each, uint64_t *bases, uint64_t *lookups)
{
    for (int i = 0; i < n; ++i) rc_each[i] = 1;
    if (bases) *bases = 0;
    if (lookups) *lookups = 0;
    if (n <= 0) return col < 0 || col >= t->v.n_cols ? -1 : 0;
    s2_ingest_job *job = s2_ingest_submit_mem_batch(c, t, images, n_bytes, n, col);
    return job ? s2_ingest_wait(job, rc_each, bases, lookups) : -1;
}

extern "C" int s2_ingest_count_files(s2_ctx *c, s2_table *t, const char *const *paths, int n, int col, int *rc_each, uint64_t *bases, uint64_t *lookups)
{
    for (int i = 0; i < n; ++i) rc_each[i] = 1;
    if (bases) *bases = 0;
    if (lookups) *lookups = 0;
    if (n <= 0) return col < 0 || col >= t->v.n_cols ? -1 : 0;
    s2_ingest_job *job = s2_ingest_submit_files(c, t, paths, n, col);
    return job ? s2_ingest_wait(job, rc_each, bases, lookups) : -1;
}

// Per-read results come back into pinned host buffers (one device -> host copy at link rate instead of six staged ones
// into pageable memory) that are kept for the next file: taken here, given back by s2_ingest_detect_free.
struct DetHostBuf { void *p = nullptr; size_t cap = 0; };
static std::mutex g_det_mu;
static std::vector<DetHostBuf> g_det_free;                          // idle buffers
static std::vector<DetHostBuf> g_det_out;                           // handed out (base pointer = result.len)
static void *det_buf_take(size_t bytes)
{
    {
        std::lock_guard<std::mutex> lk(g_det_mu);
        size_t best = g_det_free.size();
        for (size_t i = 0; i < g_det_free.size(); ++i)
            if (g_det_free[i].cap >= bytes && (best == g_det_free.size() || g_det_free[i].cap < g_det_free[best].cap)) best = i;
        if (best < g_det_free.size()) {
            DetHostBuf b = g_det_free[best];
            g_det_free.erase(g_det_free.begin() + (long)best);
            g_det_out.push_back(b);
            return b.p;
        }
    }
    DetHostBuf b;
    b.cap = (bytes + bytes / 4 + (1u << 20)) & ~(size_t)4095;
    if (cudaHostAlloc(&b.p, b.cap, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    std::lock_guard<std::mutex> lk(g_det_mu);
    g_det_out.push_back(b);
    return b.p;
}
static void det_buf_give(void *p)
{
    std::lock_guard<std::mutex> lk(g_det_mu);
    for (size_t i = 0; i < g_det_out.size(); ++i)
        if (g_det_out[i].p == p) {
            g_det_free.push_back(g_det_out[i]);
            g_det_out.erase(g_det_out.begin() + (long)i);
            // a handful of idle buffers is plenty (one per worker thread and mate file)
            while (g_det_free.size() > 40) { cudaFreeHost(g_det_free.front().p); g_det_free.erase(g_det_free.begin()); }
            return;
        }
}

// Pass 1 of quantify_hits_PE (src/strain_detect.c:465-491) for every read of one file, inflated and split on the
// GPU.  out->len / hits / inf are per record in file order (all records, also those shorter than 31);
// out->inf_* list the informative windows sorted by (record, offset) with their canonical k-mer.  The arrays are
// malloc()ed here and released by s2_ingest_detect_free.  Returns 0 / 1 (not handled) / -1 like the count form.
extern "C" int s2_ingest_detect_file(s2_ctx *c, s2_table *t, const char *path, s2_ingest_detect_result *out)
{
    memset(out, 0, sizeof *out);
    // (a table too large for L2 - a -r genome beyond 16 Mb - is probed by the direct detect kernel: there is no two-phase
    // form of the detect scan, so the probes are DRAM bound there, which still leaves the host parser far behind)
    const bool trace = s2_env_int("S2_INGEST_TRACE", 0) != 0;
    const double t_begin = ing_now();
    s2_ingest *g = ingest_acquire(c);
    if (!g) return -1;
    std::lock_guard<std::mutex> pipeline_lock(g->mu, std::adopt_lock);     // ours for the whole call
    IngSource src;
    src.fd = open(path, O_RDONLY);
    if (src.fd < 0) return 1;
    struct FdGuard { int fd; ~FdGuard() { close(fd); } } guard{ src.fd };          // closed on every way out
    ingest_classify(src);
    if (!src.eligible || (src.bgzf && !g->hw_deflate)) return 1;
    const size_t max_rec = (size_t)g->max_lines / 2 + 4;
    auto dev_alloc = [](void **p, size_t bytes) { return *p ? cudaSuccess : cudaMalloc(p, bytes); };
    if (dev_alloc((void **)&g->d_hits_c, max_rec * 4) || dev_alloc((void **)&g->d_inf_c, max_rec * 4) ||
        dev_alloc((void **)&g->d_rec_off, (max_rec + 1) * 8) || dev_alloc((void **)&g->d_pos_c, (ING_CAP_C + 1) * 8) ||
        dev_alloc((void **)&g->d_cnt_c, 8) || dev_alloc((void **)&g->d_fcnt, 8)) { s2_set_error("out of device memory"); return -1; }
    ull n_inf = 0;
    IngResult r;
    memset(&r, 0, sizeof r);
    int rc = 0;
    const double t_ready = ing_now();
    tr_wait = tr_h2d = tr_decomp = tr_launch = 0;
    tr_on = s2_env_int("S2_INGEST_TRACE", 0) >= 2;
    for (int attempt = 0; attempt < 2; ++attempt) {
        if (!g->d_frec) {
            if (!g->f_cap) g->f_cap = 1ull << 20;
            if (cudaMalloc((void **)&g->d_frec, g->f_cap * 4) || cudaMalloc((void **)&g->d_foff, g->f_cap * 4) ||
                cudaMalloc((void **)&g->d_fkmer, g->f_cap * 8)) { s2_set_error("out of device memory"); return -1; }
        }
        CK(cudaMemsetAsync(g->d_fcnt, 0, 8, g->stream));
        rc = ingest_stream(g, t, src, ING_DETECT, 0, 1u, &r, nullptr);
        if (rc < 0) { ingest_quiesce(g); ingest_explain_failure(); break; }
        if (rc || r.irregular || r.inf_overflow) { rc = 1; break; }              // irregular text / absurdly dense chunk: host path
        CK(cudaMemcpy(&n_inf, g->d_fcnt, 8, cudaMemcpyDeviceToHost));
        if (n_inf <= g->f_cap) break;
        cudaFree(g->d_frec); cudaFree(g->d_foff); cudaFree(g->d_fkmer);          // list too small: grow and run the pass again
        g->d_frec = g->d_foff = nullptr; g->d_fkmer = nullptr;
        g->f_cap = n_inf + n_inf / 8 + 1024;
        rc = 2;
    }
    if (rc == 2) s2_set_error("the list of informative windows kept growing");
    if (rc) return rc == 2 ? -1 : rc;
    const ull n_rec = r.records;
    const double t_streamed = ing_now();
    if (tr_on) fprintf(stderr, "[s2 ingest] detect %s: host time in ring wait %.0f us, inflate calls %.0f us, launches %.0f us\n", path, tr_wait, tr_decomp, tr_launch);
    ingest_trace_timeline();
    out->n_records = n_rec; out->n_inf = n_inf; out->bases = r.bases; out->fasta = src.fasta ? 1u : 0u;
    // one pinned buffer: [len | hits | inf | inf_rec | inf_off | inf_kmer], six asynchronous copies, one wait
    const size_t r4 = ((size_t)n_rec + 4) & ~(size_t)3, i4 = ((size_t)n_inf + 4) & ~(size_t)3;          // (keeps inf_kmer 8-byte aligned)
    uint8_t *hb = (uint8_t *)det_buf_take(r4 * 12 + i4 * 16);
    if (!hb) { s2_set_error("out of pinned host memory"); return -1; }
    out->len = (uint32_t *)hb; out->hits = out->len + r4; out->inf = out->hits + r4;
    out->inf_rec = out->inf + r4; out->inf_off = out->inf_rec + i4; out->inf_kmer = (uint64_t *)(out->inf_off + i4);
    auto d2h = [&](void *dst, const void *from, size_t bytes) { return !bytes || cudaMemcpyAsync(dst, from, bytes, cudaMemcpyDeviceToHost, g->stream) == cudaSuccess; };
    if (!d2h(out->len, g->d_len_all, n_rec * 4) || !d2h(out->hits, g->d_hits_all, n_rec * 4) || !d2h(out->inf, g->d_inf_all, n_rec * 4) ||
        !d2h(out->inf_rec, g->d_frec, n_inf * 4) || !d2h(out->inf_off, g->d_foff, n_inf * 4) || !d2h(out->inf_kmer, g->d_fkmer, n_inf * 8) ||
        cudaStreamSynchronize(g->stream) != cudaSuccess) {
        s2_set_error("reading the per-read results back failed: %s", cudaGetErrorString(cudaGetLastError()));
        s2_ingest_detect_free(out);
        return -1;
    }
    const double t_fetched = ing_now();
    // the kernels append in arbitrary order: sort by (record, offset) = the order pass 2 prints them
    std::vector<uint32_t> perm(n_inf);
    for (uint32_t i = 0; i < n_inf; ++i) perm[i] = i;
    std::sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) {
        return out->inf_rec[a] != out->inf_rec[b] ? out->inf_rec[a] < out->inf_rec[b] : out->inf_off[a] < out->inf_off[b];
    });
    std::vector<uint32_t> r2(n_inf), o2(n_inf); std::vector<uint64_t> k2(n_inf);
    for (uint64_t i = 0; i < n_inf; ++i) { r2[i] = out->inf_rec[perm[i]]; o2[i] = out->inf_off[perm[i]]; k2[i] = out->inf_kmer[perm[i]]; }
    if (n_inf) { memcpy(out->inf_rec, r2.data(), n_inf * 4); memcpy(out->inf_off, o2.data(), n_inf * 4); memcpy(out->inf_kmer, k2.data(), n_inf * 8); }
    if (trace) fprintf(stderr, "[s2 ingest] detect %s: %llu records, %llu informative windows, %.0f Mbases; pipeline %.0f us, stream %.0f us, results to the host %.0f us, sort %.0f us\n",
                       path, (unsigned long long)n_rec, (unsigned long long)n_inf, r.bases / 1e6, t_ready - t_begin, t_streamed - t_ready, t_fetched - t_streamed, ing_now() - t_fetched);
    return 0;
}

extern "C" void s2_ingest_detect_free(s2_ingest_detect_result *r)
{
    if (!r) return;
    if (r->len) det_buf_give(r->len);                    // (the six arrays are one pinned buffer)
    memset(r, 0, sizeof *r);
}

// kept for callers written against round 1's one-pipeline-per-thread design: the pipelines now belong to the context
extern "C" void s2_ingest_thread_cleanup(void) {}

extern "C" void s2_ingest_reset(s2_ctx *c) { if (c) s2_ingest_ctx_closing(c); }

void s2_ingest_ctx_closing(s2_ctx *c)
{
    IngPool *pool;
    { std::lock_guard<std::mutex> lk(g_pool_mu); pool = (IngPool *)c->ingest_pool; c->ingest_pool = nullptr; }
    if (!pool) return;
    for (s2_ingest *g : pool->pipes) ingest_free(g);
    delete pool;
}
