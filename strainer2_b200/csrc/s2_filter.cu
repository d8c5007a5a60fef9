// s2_filter.cu - the selection step of kmer_scrub_filter on the GPU.
//
// Reference: /root/reference/scripts/kmer_scrub_filter.py.  Its joint scrub (:88-143) turns the pangenome and
// metagenome counts of every k-mer into fractions of their column sums, takes the larger of the two, sorts all k-mers
// by that value (descending, stable: ties keep table order) and removes k-mers from the top until only min_fraction
// of the table would be left.  The sort is only a means: what matters is WHICH n_scrub rows are on top.  So instead
// of sorting 6.7 M (value, row) pairs this file selects:
//
//   fil_values : value bits per row - IEEE double division, the same two roundings Python does; non-negative
//                doubles order like their bit patterns
//   fil_hist / fil_pick x 8 : radix select, 8 bits per pass from the top: a 256-bin histogram of the rows that
//                still match the prefix (shared-memory atomics, one global flush per block), then one thread walks the
//                bins from the top, fixes the next byte of the n_scrub-th largest value v* and adds the rows in the
//                bins above to n_greater - all on the device, no host round trip between passes
//   fil_tie_count / scan / fil_mark : rows equal to v* are removed in table order until the quota is used up (an
//                exclusive scan of the tie flags gives every tie its rank)
//
// Integer / bit-exact work on 17 bytes per row; HBM-bound (10 passes over 8-byte values); no tensor cores.
// The independent scrub (:31-58, :72-84) needs "how many counts exceed t" for t = 0, 1, 2 ...: one histogram pass
// (counts clipped to 65536 bins) answers every t the script can reasonably reach; fil_count_above covers the rest.
#include "s2_private.h"

typedef unsigned long long ull;
#define FIL_THREADS 256
#define FIL_ROWS_PER_BLOCK (FIL_THREADS * 4)

struct FilSel {
    ull prefix, mask;      // bytes of v* fixed so far
    ull k;                 // rank (0 = largest) of v* among the rows that match the prefix
    ull n_greater;         // rows with a value above every value that matches the prefix
    ull quota;             // ties (value == v*) to remove, in table order
    unsigned hist[256];
};

__global__ void __launch_bounds__(FIL_THREADS) fil_values(const ull *__restrict__ pan, const ull *__restrict__ meta, const uint8_t *__restrict__ alive,
                                                           ull n, ull pan_sum, ull meta_sum, ull *__restrict__ vals)
{
    const double ps = (double)pan_sum, ms = (double)meta_sum;
    for (ull i = (ull)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (ull)gridDim.x * blockDim.x) {
        double v = 0.0;                                               // kmer_scrub_filter.py:112-117
        if (alive[i]) {
            if (meta[i] > 0) { const double f = (double)meta[i] / ms; if (f > v) v = f; }
            if (pan[i] > 0) { const double f = (double)pan[i] / ps; if (f > v) v = f; }
        }
        vals[i] = (ull)__double_as_longlong(v);
    }
}

__global__ void __launch_bounds__(FIL_THREADS) fil_hist(const ull *__restrict__ vals, const uint8_t *__restrict__ alive, ull n, FilSel *sel, int shift)
{
    __shared__ unsigned sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const ull prefix = sel->prefix, mask = sel->mask;
    for (ull i = (ull)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (ull)gridDim.x * blockDim.x) {
        const ull v = vals[i];
        if (alive[i] && (v & mask) == prefix) atomicAdd(&sh[(unsigned)(v >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&sel->hist[threadIdx.x], sh[threadIdx.x]);
}

__global__ void fil_pick(FilSel *sel, int shift)
{
    ull k = sel->k, cum = 0;
    for (int b = 255; b >= 0; --b) {
        const ull h = sel->hist[b];
        if (cum + h > k) {
            sel->n_greater += cum;
            sel->k = k - cum;
            sel->prefix |= (ull)b << shift;
            sel->mask |= 0xFFull << shift;
            break;
        }
        cum += h;
    }
    for (int b = 0; b < 256; ++b) sel->hist[b] = 0;
}

__device__ __forceinline__ unsigned fil_block_scan(unsigned v, unsigned &total)
{
    __shared__ unsigned warp_sums[FIL_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned x = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += x; }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    unsigned base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < FIL_THREADS / 32; ++i) { const unsigned s = warp_sums[i]; if (i < wid) base += s; tot += s; }
    __syncthreads();
    total = tot;
    return base + inc - v;
}

// ties per block of 1024 rows
__global__ void __launch_bounds__(FIL_THREADS) fil_tie_count(const ull *__restrict__ vals, const uint8_t *__restrict__ alive, ull n, const FilSel *sel,
                                                              ull *__restrict__ block_ties)
{
    const ull vstar = sel->prefix;
    unsigned c = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const ull i = (ull)blockIdx.x * FIL_ROWS_PER_BLOCK + threadIdx.x * 4 + k;
        if (i < n && alive[i] && vals[i] == vstar) ++c;
    }
    unsigned total;
    fil_block_scan(c, total);
    if (threadIdx.x == 0) block_ties[blockIdx.x] = total;
}

// single block: exclusive scan of the blocks' tie counts, and the quota
__global__ void __launch_bounds__(FIL_THREADS) fil_tie_scan(ull *__restrict__ block_ties, unsigned n_blocks, FilSel *sel, ull n_scrub)
{
    __shared__ ull carry;
    if (threadIdx.x == 0) { carry = 0; sel->quota = n_scrub - sel->n_greater; }
    __syncthreads();
    for (unsigned b0 = 0; b0 < n_blocks; b0 += FIL_THREADS) {
        const unsigned i = b0 + threadIdx.x;
        const unsigned x = i < n_blocks ? (unsigned)block_ties[i] : 0u;       // <= 1024
        unsigned total;
        const unsigned ex = fil_block_scan(x, total);
        if (i < n_blocks) block_ties[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(FIL_THREADS) fil_mark(const ull *__restrict__ vals, const uint8_t *__restrict__ alive, ull n, const FilSel *sel,
                                                         const ull *__restrict__ block_ties, uint8_t *__restrict__ keep)
{
    const ull vstar = sel->prefix, quota = sel->quota;
    unsigned tie[4], c = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const ull i = (ull)blockIdx.x * FIL_ROWS_PER_BLOCK + threadIdx.x * 4 + k;
        tie[k] = (i < n && alive[i] && vals[i] == vstar) ? 1u : 0u;
        c += tie[k];
    }
    unsigned total;
    ull rank = block_ties[blockIdx.x] + fil_block_scan(c, total);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const ull i = (ull)blockIdx.x * FIL_ROWS_PER_BLOCK + threadIdx.x * 4 + k;
        if (i < n) {
            uint8_t kp = 0;
            if (alive[i]) {
                const ull v = vals[i];
                kp = v < vstar || (v == vstar && rank >= quota);      // the first `quota` ties in table order go
            }
            keep[i] = kp;
        }
        rank += tie[k];
    }
}

__global__ void __launch_bounds__(FIL_THREADS) fil_histogram(const ull *__restrict__ vals, ull n, ull *__restrict__ hist)
{
    for (ull i = (ull)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (ull)gridDim.x * blockDim.x) {
        const ull v = vals[i];
        atomicAdd(&hist[v < 65536 ? v : 65536], 1ull);
    }
}

__global__ void __launch_bounds__(FIL_THREADS) fil_count_above(const ull *__restrict__ vals, ull n, ull t, ull *__restrict__ count)
{
    ull c = 0;
    for (ull i = (ull)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (ull)gridDim.x * blockDim.x) c += vals[i] > t;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xFFFFFFFFu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes) { CK(cudaMalloc(&p, bytes ? bytes : 16)); return 0; }
};

// joint scrub: keep_out[i] = 1 for the rows that survive.  n_scrub = how many of the alive rows go (the host
// evaluates the script's floating-point stopping rule, :127-131); 0 <= n_scrub <= number of alive rows.
extern "C" int s2_scrub_joint(s2_ctx *c, const uint64_t *pan, const uint64_t *meta, const uint8_t *alive, uint64_t n,
                              uint64_t pan_sum, uint64_t meta_sum, uint64_t n_scrub, uint8_t *keep_out)
{
    if (!c) { s2_set_error("no context"); return -1; }
    if (n == 0) return 0;
    CK(cudaSetDevice(c->device));
    DevBuf d_pan, d_meta, d_alive, d_vals, d_keep, d_sel, d_ties;
    const unsigned n_blocks = (unsigned)((n + FIL_ROWS_PER_BLOCK - 1) / FIL_ROWS_PER_BLOCK);
    if (d_pan.alloc(n * 8) || d_meta.alloc(n * 8) || d_alive.alloc(n) || d_vals.alloc(n * 8) || d_keep.alloc(n) || d_sel.alloc(sizeof(FilSel)) ||
        d_ties.alloc((size_t)n_blocks * 8)) return -1;
    cudaStream_t st = c->lanes[0].stream;
    CK(cudaMemcpyAsync(d_pan.p, pan, n * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_meta.p, meta, n * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_alive.p, alive, n, cudaMemcpyHostToDevice, st));
    if (n_scrub == 0) {                                               // nothing goes: the survivors are the alive rows
        CK(cudaMemcpyAsync(keep_out, d_alive.p, n, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        for (uint64_t i = 0; i < n; ++i) keep_out[i] = keep_out[i] ? 1 : 0;
        return 0;
    }
    const int grid = c->n_sm * 8;
    FilSel init; memset(&init, 0, sizeof init);
    init.k = n_scrub - 1;
    CK(cudaMemcpyAsync(d_sel.p, &init, sizeof init, cudaMemcpyHostToDevice, st));
    fil_values<<<grid, FIL_THREADS, 0, st>>>((const ull *)d_pan.p, (const ull *)d_meta.p, (const uint8_t *)d_alive.p, n, pan_sum, meta_sum, (ull *)d_vals.p);
    for (int shift = 56; shift >= 0; shift -= 8) {
        fil_hist<<<grid, FIL_THREADS, 0, st>>>((const ull *)d_vals.p, (const uint8_t *)d_alive.p, n, (FilSel *)d_sel.p, shift);
        fil_pick<<<1, 1, 0, st>>>((FilSel *)d_sel.p, shift);
    }
    fil_tie_count<<<n_blocks, FIL_THREADS, 0, st>>>((const ull *)d_vals.p, (const uint8_t *)d_alive.p, n, (const FilSel *)d_sel.p, (ull *)d_ties.p);
    fil_tie_scan<<<1, FIL_THREADS, 0, st>>>((ull *)d_ties.p, n_blocks, (FilSel *)d_sel.p, n_scrub);
    fil_mark<<<n_blocks, FIL_THREADS, 0, st>>>((const ull *)d_vals.p, (const uint8_t *)d_alive.p, n, (const FilSel *)d_sel.p, (const ull *)d_ties.p,
                                                (uint8_t *)d_keep.p);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(keep_out, d_keep.p, n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}

// independent scrub: hist[v] = number of entries equal to v for v < 65536, hist[65536] = entries >= 65536
extern "C" int s2_scrub_histogram(s2_ctx *c, const uint64_t *vals, uint64_t n, uint64_t *hist65537)
{
    if (!c) { s2_set_error("no context"); return -1; }
    CK(cudaSetDevice(c->device));
    DevBuf d_vals, d_hist;
    if (d_vals.alloc(n * 8) || d_hist.alloc(65537 * 8)) return -1;
    cudaStream_t st = c->lanes[0].stream;
    CK(cudaMemsetAsync(d_hist.p, 0, 65537 * 8, st));
    if (n) {
        CK(cudaMemcpyAsync(d_vals.p, vals, n * 8, cudaMemcpyHostToDevice, st));
        fil_histogram<<<c->n_sm * 8, FIL_THREADS, 0, st>>>((const ull *)d_vals.p, n, (ull *)d_hist.p);
        CK(cudaGetLastError());
    }
    CK(cudaMemcpyAsync(hist65537, d_hist.p, 65537 * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int s2_scrub_count_above(s2_ctx *c, const uint64_t *vals, uint64_t n, uint64_t t, uint64_t *count)
{
    if (!c) { s2_set_error("no context"); return -1; }
    CK(cudaSetDevice(c->device));
    DevBuf d_vals, d_cnt;
    if (d_vals.alloc(n * 8) || d_cnt.alloc(8)) return -1;
    cudaStream_t st = c->lanes[0].stream;
    CK(cudaMemsetAsync(d_cnt.p, 0, 8, st));
    if (n) {
        CK(cudaMemcpyAsync(d_vals.p, vals, n * 8, cudaMemcpyHostToDevice, st));
        fil_count_above<<<c->n_sm * 8, FIL_THREADS, 0, st>>>((const ull *)d_vals.p, n, t, (ull *)d_cnt.p);
        CK(cudaGetLastError());
    }
    CK(cudaMemcpyAsync(count, d_cnt.p, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}
