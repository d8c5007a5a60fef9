// s2_capi.cu - device half of the C ABI declared in include/strainer2_b200.h:
// contexts (streams + pinned batch ring), strain tables, count / detect scans, codecs.
// There is no CPU fallback anywhere in this file: every path ends in a kernel launch or an error.
#include "../../include/strainer2_b200.h"
#include "s2_kernels.cuh"
#include "s2_kmer.cuh"
#include "s2_internal.h"

#include <dlfcn.h>
#include <fcntl.h>
#include <nccl.h>
#include <sys/mman.h>
#include <unistd.h>

#include <algorithm>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

void s2_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

extern "C" const char *s2_last_error(void) { return g_err; }
extern "C" int s2_abi_version(void) { return S2_ABI_VERSION; }

#include "s2_private.h"

extern "C" int s2_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { s2_set_error("no usable CUDA runtime / driver"); return -1; }
    return n;
}

extern "C" s2_ctx *s2_init(int device, uint64_t batch_bytes, int n_lanes)
{
    int n = s2_device_count();
    if (n <= 0) { s2_set_error("no CUDA device: this library has no CPU path"); return nullptr; }
    if (device < 0 || device >= n) { s2_set_error("device %d out of range (0..%d)", device, n - 1); return nullptr; }
    cudaDeviceProp prop;
    CKN(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        s2_set_error("device %d (%s) is sm_%d%d; this library is built for sm_100a only", device, prop.name,
                     prop.major, prop.minor);
        return nullptr;
    }
    CKN(cudaSetDevice(device));
    static std::atomic<uint64_t> next_serial{1};
    s2_ctx *c = new s2_ctx();
    c->device = device;
    c->serial = next_serial.fetch_add(1);
    c->n_sm = prop.multiProcessorCount;
    c->batch_bytes = batch_bytes ? batch_bytes : (64ull << 20);
    c->batch_bytes = (c->batch_bytes + 511) & ~511ull;
    c->n_lanes = n_lanes > 0 ? n_lanes : 4;
    c->lanes.resize(c->n_lanes);
    for (auto &l : c->lanes) {                       // batch buffers are allocated on first use (lane_buffers)
        CKN(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
        CKN(cudaEventCreate(&l.k0));
        CKN(cudaEventCreate(&l.k1));
    }
    CKN(cudaMalloc((void **)&c->d_stats, 2 * sizeof(unsigned long long)));
    CKN(cudaMemset(c->d_stats, 0, 2 * sizeof(unsigned long long)));
    CKN(cudaHostAlloc((void **)&c->h_stats, 2 * sizeof(unsigned long long), cudaHostAllocDefault));
    for (auto &e : c->user_ev) CKN(cudaEventCreate(&e));
    if (getenv("S2_SCAN_VARIANT") && s2_scan_variant_set(atoi(getenv("S2_SCAN_VARIANT")))) {
        s2_set_error("S2_SCAN_VARIANT out of range");
        return nullptr;
    }
    c->grid_count = c->n_sm * s2_scan_blocks_per_sm(S2_MODE_COUNT);
    c->grid_detect = c->n_sm * s2_scan_blocks_per_sm(S2_MODE_DETECT);
    return c;
}

extern "C" void s2_shutdown(s2_ctx *c)
{
    if (!c) return;
    s2_ingest_ctx_closing(c);
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (auto &l : c->lanes) {
        if (l.k0) cudaEventDestroy(l.k0);
        if (l.k1) cudaEventDestroy(l.k1);
        if (l.d_buf) cudaFree(l.d_buf);
        if (l.h_buf) cudaFreeHost(l.h_buf);
        if (l.part_pool) cudaFree(l.part_pool);
        if (l.part_cursor) cudaFree(l.part_cursor);
        if (l.part_overflow) cudaFree(l.part_overflow);
        if (l.stream) cudaStreamDestroy(l.stream);
    }
    for (auto &e : c->ev_pending) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    for (auto &e : c->ev_free) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    for (auto &e : c->user_ev) if (e) cudaEventDestroy(e);
    for (auto &s : c->det) if (s.p) cudaFree(s.p);
    if (c->d_stats) cudaFree(c->d_stats);
    if (c->h_stats) cudaFreeHost(c->h_stats);
    delete c;
}

// kernel-shape selection for the sweep tool (tools/scan_sweep.py) and S2_SCAN_VARIANT
extern "C" int s2_tune_scan_variant(s2_ctx *c, int v)
{
    if (v < 0) return s2_scan_variant_count();
    if (s2_scan_variant_set(v)) { s2_set_error("scan variant %d out of range", v); return -1; }
    CK(cudaSetDevice(c->device));
    c->grid_count = c->n_sm * s2_scan_blocks_per_sm(S2_MODE_COUNT);
    c->grid_detect = c->n_sm * s2_scan_blocks_per_sm(S2_MODE_DETECT);
    return s2_scan_variant_count();
}

extern "C" const char *s2_tune_scan_variant_name(int v) { return s2_scan_variant_name(v); }

extern "C" int s2_ctx_device(const s2_ctx *c) { return c->device; }
extern "C" int s2_ctx_sm_count(const s2_ctx *c) { return c->n_sm; }

// Page-locked allocation is slow (~0.5 GB/s), so a lane gets its device buffer, and its pinned host
// buffer only if the caller fills batches through s2_batch_acquire, the first time it is used.
static int lane_buffers(s2_ctx *c, Lane &l, bool need_host)
{
    if (!l.d_buf) CK(cudaMalloc((void **)&l.d_buf, c->batch_bytes + 64));
    if (need_host && !l.h_buf) CK(cudaHostAlloc((void **)&l.h_buf, c->batch_bytes, cudaHostAllocDefault));
    return 0;
}

// wait for a lane's work and fold its kernel time into the context (caller holds c->mu)
static int lane_retire(s2_ctx *c, Lane &l)
{
    if (l.state != LANE_INFLIGHT) return 0;
    CK(cudaEventSynchronize(l.k1));
    if (l.timed) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, l.k0, l.k1));
        c->kernel_ms += ms;
        c->kernel_launches += 1;
        l.timed = false;
    }
    l.state = LANE_FREE;
    return 0;
}

// fold the event pairs of finished enqueue-only launches into the kernel time (caller holds c->mu;
// lane 0's stream must have been synchronised)
static int harvest_events(s2_ctx *c)
{
    for (auto &e : c->ev_pending) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e.first, e.second));
        c->kernel_ms += ms;
        c->kernel_launches += 1;
        c->ev_free.push_back(e);
    }
    c->ev_pending.clear();
    return 0;
}

extern "C" int s2_kernel_time(s2_ctx *c, double *ms, uint64_t *launches, int reset)
{
    std::lock_guard<std::mutex> g(c->mu);
    if (ms) *ms = c->kernel_ms;
    if (launches) *launches = c->kernel_launches;
    if (reset) { c->kernel_ms = 0.0; c->kernel_launches = 0; }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// strain table
// ------------------------------------------------------------------------------------------------
// the count scan of one batch on one lane's stream: direct kernel, or radix-partition + per-partition
// probes when the table's fingerprint array is too large to stay L2 resident (caller holds c->mu)
static int launch_count(s2_ctx *c, Lane &l, const uint8_t *d_bases, uint64_t n_bytes, s2_table *t, int col)
{
    return s2_launch_count_on_lane(c, l, d_bases, n_bytes, t, col);
}

int s2_launch_count_on_lane(s2_ctx *c, Lane &l, const uint8_t *d_bases, uint64_t n_bytes, s2_table *t, int col)
{
    if (!t->partitioned || n_bytes < (s2_env_u64("S2_PARTITION_MIN_BATCH_KB", 1024) << 10)) {
        s2_launch_scan_count(d_bases, n_bytes, t->v, col, c->d_stats, c->grid_count, l.stream);
        return 0;
    }
    const uint64_t region_cap = n_bytes / S2_NPART + n_bytes / (2 * S2_NPART) + 8192;     // 1.5x the even share
    const uint64_t need = region_cap * S2_NPART;
    if (need > l.part_entries) {
        if (l.part_pool) { CK(cudaStreamSynchronize(l.stream)); cudaFree(l.part_pool); l.part_pool = nullptr; l.part_entries = 0; }
        CK(cudaMalloc((void **)&l.part_pool, need * sizeof(uint64_t)));
        l.part_entries = need;
    }
    if (!l.part_cursor) CK(cudaMalloc((void **)&l.part_cursor, (S2_NPART + 1) * sizeof(unsigned long long)));
    if (!l.part_overflow) CK(cudaMalloc((void **)&l.part_overflow, 2 * sizeof(uint32_t)));
    s2_launch_scan_count_partitioned(d_bases, n_bytes, t->v, col, c->d_stats, l.part_pool, l.part_entries / S2_NPART,
                                     l.part_cursor, l.part_overflow, c->n_sm, c->grid_count, l.stream);
    return 0;
}

extern "C" uint64_t s2_table_n_keys(const s2_table *t) { return t->n_keys; }
extern "C" uint64_t s2_table_n_slots(const s2_table *t) { return t->v.n_slots; }
extern "C" uint64_t s2_table_probe_bytes(const s2_table *t) { return t->v.n_slots * sizeof(uint16_t); }
extern "C" uint64_t s2_table_hbm_bytes(const s2_table *t)
{
    return t->v.n_slots * (sizeof(uint16_t) + sizeof(uint64_t) + sizeof(uint32_t) * (uint64_t)t->v.n_cols) +
           t->n_keys * 2 * sizeof(uint32_t);
}

extern "C" void s2_table_free(s2_table *t)
{
    if (!t) return;
    cudaSetDevice(t->ctx->device);
    cudaFree(t->v.fp); cudaFree(t->v.keys); cudaFree(t->v.counts); cudaFree(t->rank_slot); cudaFree(t->rank_pos); cudaFree(t->scratch);
    delete t;
}

// temporaries of one call: released on every way out (the CK() macro returns from the middle of a function)
struct DevTemps {
    std::vector<void *> dev, host;
    template <typename T> cudaError_t alloc(T **p, size_t bytes) { const cudaError_t e = cudaMalloc((void **)p, bytes); if (e == cudaSuccess) dev.push_back(*p); return e; }
    template <typename T> cudaError_t alloc_host(T **p, size_t bytes) { const cudaError_t e = cudaHostAlloc((void **)p, bytes, cudaHostAllocDefault); if (e == cudaSuccess) host.push_back(*p); return e; }
    ~DevTemps() { for (void *p : dev) cudaFree(p); for (void *p : host) cudaFreeHost(p); }
};

static int table_build_impl(s2_ctx *c, s2_table *t, const void *bases, uint64_t n_bytes, int n_cols,
                            double load, int on_device)
{
    DevTemps tmp;
    CK(cudaSetDevice(c->device));
    if (n_bytes >= 0xFFFFFFF0ull) { s2_set_error("reference genome of %llu bytes exceeds the 4 GiB build limit", (unsigned long long)n_bytes); return -1; }
    if (n_cols < 1 || n_cols > 8) { s2_set_error("n_cols must be 1..8"); return -1; }
    if (load <= 0.0) load = 0.5;
    if (load > 0.9) load = 0.9;
    cudaStream_t st = c->lanes[0].stream;

    const uint8_t *d_bases = (const uint8_t *)bases;
    uint8_t *tmp_bases = nullptr;
    if (!on_device && n_bytes) {
        CK(tmp.alloc(&tmp_bases, n_bytes + 64));
        CK(cudaMemcpyAsync(tmp_bases, bases, n_bytes, cudaMemcpyHostToDevice, st));
        d_bases = tmp_bases;
    }
    const uint64_t upper = n_bytes >= S2_K ? n_bytes - S2_K + 1 : 0;          // windows, hence keys, at most
    uint64_t n_buckets = (uint64_t)((double)upper / (S2_BUCKET_SLOTS * load)) + 2;
    if (n_buckets * S2_BUCKET_SLOTS >= 0xFFFFFFF0ull) { s2_set_error("table too large"); return -1; }
    t->ctx = c;
    t->v.n_buckets = (uint32_t)n_buckets;
    t->v.n_slots = n_buckets * S2_BUCKET_SLOTS;
    t->v.n_cols = n_cols;
    CK(cudaMalloc((void **)&t->v.fp, t->v.n_slots * sizeof(uint16_t)));
    CK(cudaMalloc((void **)&t->v.keys, t->v.n_slots * sizeof(uint64_t)));
    CK(cudaMalloc((void **)&t->v.counts, t->v.n_slots * sizeof(uint32_t) * n_cols));
    CK(cudaMemsetAsync(t->v.fp, 0, t->v.n_slots * sizeof(uint16_t), st));
    CK(cudaMemsetAsync(t->v.keys, 0xFF, t->v.n_slots * sizeof(uint64_t), st));
    CK(cudaMemsetAsync(t->v.counts, 0, t->v.n_slots * sizeof(uint32_t) * n_cols, st));

    uint32_t *first_pos = nullptr, *slot_of_pos = nullptr, *block_sums = nullptr, *rank_tmp = nullptr, *pos_tmp = nullptr;
    unsigned long long *d_n = nullptr;
    const uint32_t n_blocks = (uint32_t)((n_bytes + 1023) / 1024);
    CK(tmp.alloc(&first_pos, t->v.n_slots * sizeof(uint32_t)));
    CK(tmp.alloc(&slot_of_pos, (n_bytes + 1) * sizeof(uint32_t)));
    CK(tmp.alloc(&block_sums, (n_blocks + 1) * sizeof(uint32_t)));
    CK(tmp.alloc(&rank_tmp, (upper + 1) * sizeof(uint32_t)));
    CK(tmp.alloc(&pos_tmp, (upper + 1) * sizeof(uint32_t)));
    CK(tmp.alloc(&d_n, sizeof(unsigned long long)));
    CK(cudaMemsetAsync(first_pos, 0xFF, t->v.n_slots * sizeof(uint32_t), st));

    s2_launch_build_insert(d_bases, n_bytes, t->v, first_pos, slot_of_pos, st);
    s2_launch_build_rank(n_bytes, first_pos, slot_of_pos, block_sums, n_blocks, rank_tmp, pos_tmp, d_n, st);
    CK(cudaGetLastError());
    unsigned long long n_keys = 0;
    CK(cudaMemcpyAsync(&n_keys, d_n, sizeof n_keys, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    t->n_keys = n_keys;
    t->partitioned = t->v.n_slots * sizeof(uint16_t) > (s2_env_u64("S2_PARTITION_MIN_MB", 64) << 20);
    CK(cudaMalloc((void **)&t->rank_slot, (n_keys + 1) * sizeof(uint32_t)));
    CK(cudaMalloc((void **)&t->scratch, (n_keys + 1) * sizeof(uint32_t)));
    CK(cudaMalloc((void **)&t->rank_pos, (n_keys + 1) * sizeof(uint32_t)));
    CK(cudaMemcpyAsync(t->rank_slot, rank_tmp, n_keys * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(t->rank_pos, pos_tmp, n_keys * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    CK(cudaStreamSynchronize(st));
    return 0;                                        // (the temporaries go with `tmp`)
}

extern "C" s2_table *s2_table_build(s2_ctx *c, const void *bases, uint64_t n_bytes, int n_cols,
                                    double load_factor, int on_device)
{
    if (!c) { s2_set_error("s2_table_build: NULL context"); return nullptr; }
    s2_table *t = new s2_table();
    if (table_build_impl(c, t, bases, n_bytes, n_cols, load_factor, on_device) != 0) {
        t->ctx = c;
        s2_table_free(t);
        return nullptr;
    }
    return t;
}

extern "C" int s2_table_export(s2_table *t, uint64_t *keys, uint32_t *djb2, uint32_t *first_pos)
{
    s2_ctx *c = t->ctx;
    CK(cudaSetDevice(c->device));
    if (t->n_keys == 0) return 0;
    cudaStream_t st = c->lanes[0].stream;
    uint64_t *d_keys = nullptr; uint32_t *d_h = nullptr;
    DevTemps tmp;
    CK(tmp.alloc(&d_keys, t->n_keys * sizeof(uint64_t)));
    CK(tmp.alloc(&d_h, t->n_keys * sizeof(uint32_t)));
    s2_launch_export(t->v, t->rank_slot, t->n_keys, d_keys, d_h, st);
    CK(cudaGetLastError());
    if (keys) CK(cudaMemcpyAsync(keys, d_keys, t->n_keys * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    if (djb2) CK(cudaMemcpyAsync(djb2, d_h, t->n_keys * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    if (first_pos) CK(cudaMemcpyAsync(first_pos, t->rank_pos, t->n_keys * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}

static int check_col(const s2_table *t, int col);

// print_hash_counts() on the device: rows in `order` (insertion indices, from s2_roworder_emulate), formatted by
// one thread per row, copied back through a pinned staging buffer and written to `out`.
extern "C" int s2_table_format(s2_table *t, const uint32_t *order, int n_print_cols, FILE *out)
{
    static const char header[] = "#kmer\treference_count\tpangenome_count\tmetagenome_count\tdrug_count\n";
    s2_ctx *c = t->ctx;
    CK(cudaSetDevice(c->device));
    if (n_print_cols < 1 || n_print_cols > 4 || n_print_cols > t->v.n_cols) { s2_set_error("n_print_cols out of range"); return -1; }
    if (fwrite(header, 1, sizeof header - 1, out) != sizeof header - 1) { s2_set_error("write failed"); return -1; }
    const uint64_t n = t->n_keys;
    if (n == 0) return 0;
    cudaStream_t st = c->lanes[0].stream;
    uint64_t *d_keys = nullptr; uint32_t *d_djb2 = nullptr, *d_order = nullptr, *d_cols[4] = { nullptr, nullptr, nullptr, nullptr };
    unsigned long long *d_sums = nullptr, *d_total = nullptr;
    char *d_text = nullptr, *h_stage = nullptr;
    const uint32_t n_blocks = (uint32_t)((n + 1023) / 1024);
    DevTemps tmp;
    CK(tmp.alloc(&d_keys, n * sizeof(uint64_t)));
    CK(tmp.alloc(&d_djb2, n * sizeof(uint32_t)));
    CK(tmp.alloc(&d_order, n * sizeof(uint32_t)));
    CK(tmp.alloc(&d_sums, (n_blocks + 1) * sizeof(unsigned long long)));
    CK(tmp.alloc(&d_total, sizeof(unsigned long long)));
    CK(cudaMemcpyAsync(d_order, order, n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    s2_launch_export(t->v, t->rank_slot, n, d_keys, d_djb2, st);
    for (int k = 0; k < n_print_cols; ++k) {
        CK(tmp.alloc(&d_cols[k], n * sizeof(uint32_t)));
        s2_launch_gather_counts(t->v, k, t->rank_slot, n, d_cols[k], st);
    }
    s2_launch_format(d_keys, d_order, n, d_cols, n_print_cols, d_sums, d_total, nullptr, 0, st);
    CK(cudaGetLastError());
    unsigned long long total = 0;
    CK(cudaMemcpyAsync(&total, d_total, sizeof total, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(tmp.alloc(&d_text, total + 1));
    s2_launch_format(d_keys, d_order, n, d_cols, n_print_cols, d_sums, d_total, d_text, 1, st);
    CK(cudaGetLastError());
    const size_t stage = 32u << 20;
    CK(tmp.alloc_host(&h_stage, stage));
    int rc = 0;
    for (unsigned long long off = 0; off < total && rc == 0; off += stage) {
        const size_t take = (size_t)std::min<unsigned long long>(stage, total - off);
        CK(cudaMemcpyAsync(h_stage, d_text + off, take, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (fwrite(h_stage, 1, take, out) != take) { s2_set_error("write failed"); rc = -1; }
    }
    return rc;
}

static int check_col(const s2_table *t, int col)
{
    if (col < 0 || col >= t->v.n_cols) { s2_set_error("column %d out of range (table has %d)", col, t->v.n_cols); return -1; }
    return 0;
}

extern "C" int s2_table_counts_gather_dev(s2_table *t, int col, void *dev_out)
{
    if (check_col(t, col)) return -1;
    CK(cudaSetDevice(t->ctx->device));
    cudaStream_t st = t->ctx->lanes[0].stream;
    s2_launch_gather_counts(t->v, col, t->rank_slot, t->n_keys, (uint32_t *)dev_out, st);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int s2_table_counts_scatter_dev(s2_table *t, int col, const void *dev_in)
{
    if (check_col(t, col)) return -1;
    CK(cudaSetDevice(t->ctx->device));
    cudaStream_t st = t->ctx->lanes[0].stream;
    s2_launch_scatter_counts(t->v, col, t->rank_slot, t->n_keys, (const uint32_t *)dev_in, st);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int s2_table_counts_fetch(s2_table *t, int col, uint32_t *host_out)
{
    if (s2_table_counts_gather_dev(t, col, t->scratch)) return -1;
    if (t->n_keys) CK(cudaMemcpy(host_out, t->scratch, t->n_keys * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int s2_table_counts_store(s2_table *t, int col, const uint32_t *host_in)
{
    if (check_col(t, col)) return -1;
    CK(cudaSetDevice(t->ctx->device));
    if (t->n_keys) CK(cudaMemcpy(t->scratch, host_in, t->n_keys * sizeof(uint32_t), cudaMemcpyHostToDevice));
    return s2_table_counts_scatter_dev(t, col, t->scratch);
}

extern "C" int s2_table_counts_clear(s2_table *t, int col)
{
    if (check_col(t, col)) return -1;
    CK(cudaSetDevice(t->ctx->device));
    CK(cudaMemset(t->v.counts + (uint64_t)col * t->v.n_slots, 0, t->v.n_slots * sizeof(uint32_t)));
    return 0;
}

static int table_query(s2_table *t, const uint64_t *kmers, uint64_t n, uint8_t *found, uint32_t *slots, bool flag, int set = 1)
{
    s2_ctx *c = t->ctx;
    CK(cudaSetDevice(c->device));
    if (n == 0) return 0;
    cudaStream_t st = c->lanes[0].stream;
    uint64_t *d_k = nullptr; void *d_o = nullptr;
    const size_t osz = flag ? n : n * sizeof(uint32_t);
    CK(cudaMalloc((void **)&d_k, n * sizeof(uint64_t)));
    CK(cudaMalloc(&d_o, osz));
    CK(cudaMemcpyAsync(d_k, kmers, n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    if (flag) s2_launch_flag(t->v, d_k, n, (uint8_t *)d_o, set, st);
    else s2_launch_lookup(t->v, d_k, n, (uint32_t *)d_o, st);
    CK(cudaGetLastError());
    void *dst = flag ? (void *)found : (void *)slots;
    if (dst) CK(cudaMemcpyAsync(dst, d_o, osz, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    cudaFree(d_k); cudaFree(d_o);
    return 0;
}

extern "C" int s2_table_flag(s2_table *t, const uint64_t *kmers, uint64_t n, uint8_t *found)
{
    return table_query(t, kmers, n, found, nullptr, true);
}

extern "C" int s2_table_counts_by_key(s2_table *t, int col, const uint64_t *kmers, uint64_t n, uint32_t *host_out)
{
    if (check_col(t, col)) return -1;
    s2_ctx *c = t->ctx;
    CK(cudaSetDevice(c->device));
    if (n == 0) return 0;
    cudaStream_t st = c->lanes[0].stream;
    uint64_t *d_k = nullptr; uint32_t *d_o = nullptr;
    CK(cudaMalloc((void **)&d_k, n * sizeof(uint64_t)));
    CK(cudaMalloc((void **)&d_o, n * sizeof(uint32_t)));
    CK(cudaMemcpyAsync(d_k, kmers, n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    s2_launch_counts_by_key(t->v, col, d_k, n, d_o, st);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(host_out, d_o, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    cudaFree(d_k); cudaFree(d_o);
    return 0;
}

extern "C" int s2_table_unflag(s2_table *t, const uint64_t *kmers, uint64_t n)
{
    return table_query(t, kmers, n, nullptr, nullptr, true, 0);
}

extern "C" int s2_table_lookup(s2_table *t, const uint64_t *kmers, uint64_t n, uint32_t *slot_out)
{
    return table_query(t, kmers, n, nullptr, slot_out, false);
}

// ------------------------------------------------------------------------------------------------
// the one collective of the path, for a single process that drives several GPUs (the executables)
// ------------------------------------------------------------------------------------------------
// NCCL is dlopen()ed on first use: the library itself has no link-time dependency on it, so loading
// it next to a framework that bundles its own NCCL (torch) is safe, and single-GPU users need none.
namespace {
struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool load()
    {
        if (h) return true;
        h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
        if (!h) { s2_set_error("cannot load libnccl.so.2: %s", dlerror()); return false; }
#define S2_NCCL_SYM(field, name) field = (decltype(field))dlsym(h, name); if (!field) { s2_set_error("libnccl lacks %s", name); return false; }
        S2_NCCL_SYM(CommInitAll, "ncclCommInitAll") S2_NCCL_SYM(CommDestroy, "ncclCommDestroy") S2_NCCL_SYM(AllReduce, "ncclAllReduce")
        S2_NCCL_SYM(GroupStart, "ncclGroupStart") S2_NCCL_SYM(GroupEnd, "ncclGroupEnd") S2_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef S2_NCCL_SYM
        return true;
    }
} g_nccl;
}

#define CKNCCL(call)                                                                               \
    do {                                                                                           \
        ncclResult_t r_ = (call);                                                                  \
        if (r_ != ncclSuccess) { s2_set_error("NCCL error at %s:%d: %s", __FILE__, __LINE__, g_nccl.GetErrorString(r_)); return -1; } \
    } while (0)

extern "C" int s2_tables_allreduce(s2_table **tabs, int n, int col)
{
    if (n <= 1) return 0;
    if (!g_nccl.load()) return -1;
    const uint64_t n_keys = tabs[0]->n_keys;
    std::vector<int> devs(n);
    for (int i = 0; i < n; ++i) {
        if (tabs[i]->n_keys != n_keys) { s2_set_error("replicas differ in key count (%llu vs %llu)", (unsigned long long)tabs[i]->n_keys, (unsigned long long)n_keys); return -1; }
        if (check_col(tabs[i], col)) return -1;
        devs[i] = tabs[i]->ctx->device;
    }
    if (n_keys == 0) return 0;
    for (int i = 0; i < n; ++i) if (s2_table_counts_gather_dev(tabs[i], col, tabs[i]->scratch)) return -1;
    // One process, n GPUs: the sum is taken straight out of peer memory over NVLink by one kernel per GPU that also
    // scatters it (s2_peer_sum_scatter_kernel) - no communicator to create (ncclCommInitAll cost 12 s of a 16 s run in
    // round 1, profiles/r1n_multigpu_cli_check.txt).  S2_ALLREDUCE=nccl, or GPUs without peer access, take ncclAllReduce.
    const char *how = getenv("S2_ALLREDUCE");
    bool p2p = n <= S2_MAX_PEERS && !(how && !strcmp(how, "nccl"));
    for (int i = 0; i < n && p2p; ++i)
        for (int j = 0; j < n && p2p; ++j) {
            int can = 0;
            if (i != j && (cudaDeviceCanAccessPeer(&can, devs[i], devs[j]) != cudaSuccess || !can)) p2p = false;
        }
    if (p2p) {
        S2PeerVecs pv;
        pv.n = n;
        for (int i = 0; i < n; ++i) pv.v[i] = tabs[i]->scratch;
        for (int i = 0; i < n; ++i) {                        // every gather is done before any GPU starts reading
            CK(cudaSetDevice(devs[i]));
            CK(cudaStreamSynchronize(tabs[i]->ctx->lanes[0].stream));
            for (int j = 0; j < n; ++j)
                if (i != j) {
                    const cudaError_t e = cudaDeviceEnablePeerAccess(devs[j], 0);
                    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
                    cudaGetLastError();
                }
        }
        for (int i = 0; i < n; ++i) {
            CK(cudaSetDevice(devs[i]));
            s2_launch_peer_sum_scatter(tabs[i]->v, col, tabs[i]->rank_slot, n_keys, pv, tabs[i]->ctx->lanes[0].stream);
            CK(cudaGetLastError());
        }
        for (int i = 0; i < n; ++i) {                        // nobody's vector is reused before everybody has read it
            CK(cudaSetDevice(devs[i]));
            CK(cudaStreamSynchronize(tabs[i]->ctx->lanes[0].stream));
        }
        return 0;
    }
    // communicators are created once per set of devices (creating them costs seconds) and kept for the life of the
    // process: the executables call this once per counter column
    static std::mutex comm_mu;
    static std::vector<int> comm_devs;
    static std::vector<ncclComm_t> comm_cache;
    std::lock_guard<std::mutex> comm_lock(comm_mu);
    std::vector<ncclComm_t> &comms = comm_cache;
    if (comm_devs != devs) {
        for (ncclComm_t cm : comm_cache) g_nccl.CommDestroy(cm);
        comm_cache.assign(n, nullptr);
        comm_devs.clear();
        // NCCL prints its version banner on stdout when NCCL_DEBUG=VERSION/WARN; stdout is the count table
        // of the drop-in executables, so it is parked on /dev/null while the communicators are created
        fflush(stdout);
        const int saved = dup(1), nul = open("/dev/null", O_WRONLY);
        if (saved >= 0 && nul >= 0) dup2(nul, 1);
        const ncclResult_t r = g_nccl.CommInitAll(comms.data(), n, devs.data());
        fflush(stdout);
        if (saved >= 0) { dup2(saved, 1); close(saved); }
        if (nul >= 0) close(nul);
        if (r != ncclSuccess) comm_cache.clear();
        CKNCCL(r);
        comm_devs = devs;
    }
    CKNCCL(g_nccl.GroupStart());
    for (int i = 0; i < n; ++i) {
        CK(cudaSetDevice(devs[i]));
        CKNCCL(g_nccl.AllReduce(tabs[i]->scratch, tabs[i]->scratch, n_keys, ncclUint32, ncclSum, comms[i], tabs[i]->ctx->lanes[0].stream));
    }
    CKNCCL(g_nccl.GroupEnd());
    for (int i = 0; i < n; ++i) {
        CK(cudaSetDevice(devs[i]));
        CK(cudaStreamSynchronize(tabs[i]->ctx->lanes[0].stream));
    }
    for (int i = 0; i < n; ++i) if (s2_table_counts_scatter_dev(tabs[i], col, tabs[i]->scratch)) return -1;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// count scan
// ------------------------------------------------------------------------------------------------
static int fetch_stats(s2_ctx *c, cudaStream_t st, s2_scan_stats *out)
{
    CK(cudaMemcpyAsync(c->h_stats, c->d_stats, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CK(cudaMemsetAsync(c->d_stats, 0, 2 * sizeof(unsigned long long), st));
    CK(cudaStreamSynchronize(st));
    if (out) { out->hits = c->h_stats[0]; out->valid_windows = c->h_stats[1]; }
    return 0;
}

extern "C" int s2_scan_count(s2_ctx *c, s2_table *t, const void *bases, uint64_t n_bytes, int col,
                             int on_device, s2_scan_stats *stats)
{
    if (check_col(t, col)) return -1;
    CK(cudaSetDevice(c->device));
    if (on_device) {
        if (((uintptr_t)bases & 15) != 0) { s2_set_error("device batch must be 16-byte aligned"); return -1; }
        std::lock_guard<std::mutex> g(c->mu);
        Lane &l = c->lanes[0];
        if (lane_retire(c, l)) return -1;
        CK(cudaEventRecord(l.k0, l.stream));
        if (launch_count(c, l, (const uint8_t *)bases, n_bytes, t, col)) return -1;
        CK(cudaGetLastError());
        CK(cudaEventRecord(l.k1, l.stream));
        l.state = LANE_INFLIGHT; l.timed = true; l.seq = c->next_seq++;
        if (lane_retire(c, l)) return -1;
        return fetch_stats(c, l.stream, stats);
    }
    // host input: H2D straight from the caller's buffer (pinned memory from s2_pinned_alloc gives
    // full PCIe rate) in batch_bytes pieces, one lane per piece, copy and kernel overlapped across
    // lanes.  Pieces overlap by 30 bytes so every 31-byte window lies in exactly one piece.
    s2_scan_stats dummy;
    if (s2_sync(c, &dummy)) return -1;
    const uint8_t *src = (const uint8_t *)bases;
    uint64_t off = 0;
    while (off < n_bytes) {
        const uint64_t take = std::min<uint64_t>(c->batch_bytes, n_bytes - off);
        {
            std::lock_guard<std::mutex> g(c->mu);
            Lane *l = nullptr;
            for (auto &x : c->lanes) if (x.state == LANE_FREE) { l = &x; break; }
            if (!l) {
                for (auto &x : c->lanes) if (x.state == LANE_INFLIGHT && (!l || x.seq < l->seq)) l = &x;
                if (!l) { s2_set_error("s2_scan_count: no lane available"); return -1; }
                if (lane_retire(c, *l)) return -1;
            }
            if (lane_buffers(c, *l, false)) return -1;
            CK(cudaMemcpyAsync(l->d_buf, src + off, take, cudaMemcpyHostToDevice, l->stream));
            CK(cudaEventRecord(l->k0, l->stream));
            if (launch_count(c, *l, l->d_buf, take, t, col)) return -1;
            CK(cudaGetLastError());
            CK(cudaEventRecord(l->k1, l->stream));
            l->state = LANE_INFLIGHT; l->timed = true; l->seq = c->next_seq++;
        }
        if (off + take >= n_bytes) break;
        off += take - (S2_K - 1);
    }
    return s2_sync(c, stats);
}

// enqueue-only form for device-resident batches: launches on lane 0's stream and returns at once.
extern "C" int s2_scan_count_enqueue(s2_ctx *c, s2_table *t, const void *dev_bases, uint64_t n_bytes, int col)
{
    if (check_col(t, col)) return -1;
    if (((uintptr_t)dev_bases & 15) != 0) { s2_set_error("device batch must be 16-byte aligned"); return -1; }
    std::lock_guard<std::mutex> g(c->mu);
    CK(cudaSetDevice(c->device));
    Lane &l = c->lanes[0];
    if (l.state == LANE_INFLIGHT && lane_retire(c, l)) return -1;
    std::pair<cudaEvent_t, cudaEvent_t> ev;
    if (!c->ev_free.empty()) { ev = c->ev_free.back(); c->ev_free.pop_back(); }
    else { CK(cudaEventCreate(&ev.first)); CK(cudaEventCreate(&ev.second)); }
    CK(cudaEventRecord(ev.first, l.stream));
    if (launch_count(c, l, (const uint8_t *)dev_bases, n_bytes, t, col)) return -1;
    CK(cudaGetLastError());
    CK(cudaEventRecord(ev.second, l.stream));
    c->ev_pending.push_back(ev);
    return 0;
}

// user timing events on lane 0's stream (the stream the enqueue-only scans run on)
extern "C" int s2_event_record(s2_ctx *c, int which)
{
    if (which < 0 || which >= 4) { s2_set_error("event index out of range"); return -1; }
    std::lock_guard<std::mutex> g(c->mu);
    CK(cudaSetDevice(c->device));
    CK(cudaEventRecord(c->user_ev[which], c->lanes[0].stream));
    return 0;
}

extern "C" int s2_event_elapsed_ms(s2_ctx *c, int from, int to, double *ms)
{
    if (from < 0 || from >= 4 || to < 0 || to >= 4) { s2_set_error("event index out of range"); return -1; }
    CK(cudaSetDevice(c->device));
    CK(cudaEventSynchronize(c->user_ev[to]));
    float f = 0.f;
    CK(cudaEventElapsedTime(&f, c->user_ev[from], c->user_ev[to]));
    *ms = f;
    return 0;
}

// Pinned host memory.  Measured on the pool's boxes (profiles/r2g_pinned_probe.txt): cudaHostAlloc pins 4 KB pages at
// 2.7 GB/s (512 MB: 192 ms); the same bytes on transparent huge pages (madvise), touched by a few threads and then
// registered, take 20 + 50 ms, and copy to the device at the same 55 GB/s.  Buffers of 8 MB and more go that way when
// the kernel allows it (S2_PINNED_HUGE=0 turns it off); s2_pinned_free tells the two kinds apart.
static std::mutex g_pin_mu;
static std::vector<std::pair<void *, size_t>> g_pin_registered;

extern "C" void *s2_pinned_alloc(uint64_t n_bytes)
{
    void *p = nullptr;
    static const bool huge_ok = s2_env_int("S2_PINNED_HUGE", 1) != 0;
    if (huge_ok && n_bytes >= (8u << 20)) {
        const size_t sz = ((size_t)n_bytes + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1);
        p = aligned_alloc(2u << 20, sz);
        if (p) {
            madvise(p, sz, MADV_HUGEPAGE);
            const int nt = (int)std::min<size_t>(4, sz >> 23);            // first touch on a few threads (one per 8 MB at least)
            std::vector<std::thread> th;
            for (int k = 1; k < nt; ++k) th.emplace_back([=]() { for (size_t o = sz / nt * k; o < (k + 1 == nt ? sz : sz / nt * (k + 1)); o += 4096) ((volatile uint8_t *)p)[o] = 0; });
            for (size_t o = 0; o < (nt > 1 ? sz / nt : sz); o += 4096) ((volatile uint8_t *)p)[o] = 0;
            for (auto &t : th) t.join();
            if (cudaHostRegister(p, sz, cudaHostRegisterPortable) == cudaSuccess) {
                std::lock_guard<std::mutex> g(g_pin_mu);
                g_pin_registered.emplace_back(p, sz);
                return p;
            }
            cudaGetLastError();
            free(p);
            p = nullptr;
        }
    }
    if (cudaHostAlloc(&p, n_bytes ? n_bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
        s2_set_error("cudaHostAlloc(%llu) failed", (unsigned long long)n_bytes);
        return nullptr;
    }
    return p;
}

extern "C" void s2_pinned_free(void *p)
{
    if (!p) return;
    {
        std::lock_guard<std::mutex> g(g_pin_mu);
        for (size_t i = 0; i < g_pin_registered.size(); ++i)
            if (g_pin_registered[i].first == p) {
                g_pin_registered.erase(g_pin_registered.begin() + (long)i);
                cudaHostUnregister(p);
                free(p);
                return;
            }
    }
    cudaFreeHost(p);
}

extern "C" uint8_t *s2_batch_acquire(s2_ctx *c, uint64_t *capacity)
{
    std::lock_guard<std::mutex> g(c->mu);
    if (cudaSetDevice(c->device) != cudaSuccess) { s2_set_error("cudaSetDevice failed"); return nullptr; }
    if (capacity) *capacity = c->batch_bytes;
    Lane *pick = nullptr;
    for (auto &l : c->lanes) if (l.state == LANE_FREE && l.h_buf) { pick = &l; break; }
    if (!pick) for (auto &l : c->lanes) if (l.state == LANE_FREE) { pick = &l; break; }
    if (!pick) {
        for (auto &l : c->lanes)
            if (l.state == LANE_INFLIGHT && (!pick || l.seq < pick->seq)) pick = &l;
        if (!pick) { s2_set_error("s2_batch_acquire: every lane is held by a caller (n_lanes=%d)", c->n_lanes); return nullptr; }
        if (lane_retire(c, *pick)) return nullptr;
    }
    if (lane_buffers(c, *pick, true)) return nullptr;
    pick->state = LANE_HELD;
    return pick->h_buf;
}

static Lane *find_lane(s2_ctx *c, const uint8_t *buf)
{
    if (!buf) return nullptr;
    for (auto &l : c->lanes) if (l.h_buf == buf) return &l;
    return nullptr;
}

extern "C" int s2_batch_release(s2_ctx *c, uint8_t *batch)
{
    std::lock_guard<std::mutex> g(c->mu);
    Lane *l = find_lane(c, batch);
    if (!l || l->state != LANE_HELD) { s2_set_error("s2_batch_release: not an acquired batch"); return -1; }
    l->state = LANE_FREE;
    return 0;
}

extern "C" int s2_batch_submit_count(s2_ctx *c, s2_table *t, uint8_t *batch, uint64_t n_bytes, int col)
{
    if (check_col(t, col)) return -1;
    std::lock_guard<std::mutex> g(c->mu);
    CK(cudaSetDevice(c->device));
    Lane *l = find_lane(c, batch);
    if (!l || l->state != LANE_HELD) { s2_set_error("s2_batch_submit_count: not an acquired batch"); return -1; }
    if (n_bytes > c->batch_bytes) { s2_set_error("batch of %llu bytes exceeds capacity", (unsigned long long)n_bytes); return -1; }
    if (n_bytes == 0) { l->state = LANE_FREE; return 0; }
    CK(cudaMemcpyAsync(l->d_buf, l->h_buf, n_bytes, cudaMemcpyHostToDevice, l->stream));
    CK(cudaEventRecord(l->k0, l->stream));
    if (launch_count(c, *l, l->d_buf, n_bytes, t, col)) return -1;
    CK(cudaGetLastError());
    CK(cudaEventRecord(l->k1, l->stream));
    l->state = LANE_INFLIGHT; l->timed = true; l->seq = c->next_seq++;
    return 0;
}

extern "C" int s2_sync(s2_ctx *c, s2_scan_stats *totals)
{
    std::lock_guard<std::mutex> g(c->mu);
    CK(cudaSetDevice(c->device));
    for (auto &l : c->lanes) if (lane_retire(c, l)) return -1;
    CK(cudaStreamSynchronize(c->lanes[0].stream));
    if (harvest_events(c)) return -1;
    return fetch_stats(c, c->lanes[0].stream, totals);
}

// ------------------------------------------------------------------------------------------------
// detect scan
// ------------------------------------------------------------------------------------------------
extern "C" int s2_scan_detect(s2_ctx *c, s2_table *t, const void *bases, uint64_t n_bytes,
                              const uint64_t *rec_off, uint32_t n_rec, uint32_t *read_hits,
                              uint32_t *read_inf, uint64_t *inf_pos, uint64_t inf_cap, uint64_t *n_inf,
                              int on_device, s2_scan_stats *stats)
{
    CK(cudaSetDevice(c->device));
    std::lock_guard<std::mutex> g(c->mu);
    Lane &l = c->lanes[0];
    if (lane_retire(c, l)) return -1;
    cudaStream_t st = l.stream;
    if (n_inf) *n_inf = 0;
    if (n_rec == 0 || n_bytes == 0) { if (stats) { stats->hits = 0; stats->valid_windows = 0; } return 0; }

    const uint8_t *d_bases = (const uint8_t *)bases;
    if (!on_device) {
        if (scratch_reserve(c->det[0], n_bytes + 64)) return -1;
        CK(cudaMemcpyAsync(c->det[0].p, bases, n_bytes, cudaMemcpyHostToDevice, st));
        d_bases = (const uint8_t *)c->det[0].p;
    } else if (((uintptr_t)bases & 15) != 0) { s2_set_error("device batch must be 16-byte aligned"); return -1; }

    if (scratch_reserve(c->det[1], (n_rec + 1ull) * sizeof(uint64_t)) || scratch_reserve(c->det[2], n_rec * sizeof(uint32_t)) ||
        scratch_reserve(c->det[3], n_rec * sizeof(uint32_t)) || scratch_reserve(c->det[4], (inf_cap + 1) * sizeof(uint64_t)) ||
        scratch_reserve(c->det[5], sizeof(unsigned long long))) return -1;
    uint64_t *d_off = (uint64_t *)c->det[1].p, *d_pos = (uint64_t *)c->det[4].p;
    uint32_t *d_hits = (uint32_t *)c->det[2].p, *d_inf = (uint32_t *)c->det[3].p;
    unsigned long long *d_cnt = (unsigned long long *)c->det[5].p;
    CK(cudaMemcpyAsync(d_off, rec_off, (n_rec + 1ull) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(d_hits, 0, n_rec * sizeof(uint32_t), st));
    CK(cudaMemsetAsync(d_inf, 0, n_rec * sizeof(uint32_t), st));
    CK(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), st));

    S2DetectOut out;
    out.rec_off = d_off; out.n_rec = n_rec; out.n_rec_dev = nullptr; out.read_hits = d_hits; out.read_inf = d_inf;
    out.inf_pos = d_pos; out.inf_count = d_cnt; out.inf_cap = inf_cap;
    CK(cudaEventRecord(l.k0, st));
    s2_launch_scan_detect(d_bases, n_bytes, t->v, out, c->d_stats, c->grid_detect, st);
    CK(cudaGetLastError());
    CK(cudaEventRecord(l.k1, st));
    l.state = LANE_INFLIGHT; l.timed = true; l.seq = c->next_seq++;
    if (lane_retire(c, l)) return -1;

    unsigned long long cnt = 0;
    CK(cudaMemcpyAsync(&cnt, d_cnt, sizeof cnt, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(read_hits, d_hits, n_rec * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(read_inf, d_inf, n_rec * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const uint64_t have = std::min<uint64_t>(cnt, inf_cap);
    if (have && inf_pos) {
        CK(cudaMemcpy(inf_pos, d_pos, have * sizeof(uint64_t), cudaMemcpyDeviceToHost));
        std::sort(inf_pos, inf_pos + have);          // the kernel appends in arbitrary order
    }
    if (n_inf) *n_inf = cnt;
    return fetch_stats(c, st, stats);
}

// ------------------------------------------------------------------------------------------------
// codecs
// ------------------------------------------------------------------------------------------------
extern "C" int s2_pack_2bit(s2_ctx *c, const void *bases, uint64_t n_bytes, int on_device,
                            uint32_t *words, uint16_t *masks)
{
    CK(cudaSetDevice(c->device));
    const uint64_t n_chunks = (n_bytes + 15) / 16;
    if (!n_chunks) return 0;
    cudaStream_t st = c->lanes[0].stream;
    const uint8_t *d_bases = (const uint8_t *)bases;
    uint8_t *tmp = nullptr;
    if (!on_device) {
        CK(cudaMalloc((void **)&tmp, n_chunks * 16));
        CK(cudaMemcpyAsync(tmp, bases, n_bytes, cudaMemcpyHostToDevice, st));
        d_bases = tmp;
    }
    uint32_t *d_w = nullptr; uint16_t *d_m = nullptr;
    CK(cudaMalloc((void **)&d_w, n_chunks * sizeof(uint32_t)));
    CK(cudaMalloc((void **)&d_m, n_chunks * sizeof(uint16_t)));
    s2_launch_pack(d_bases, n_bytes, d_w, d_m, st);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(words, d_w, n_chunks * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(masks, d_m, n_chunks * sizeof(uint16_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    cudaFree(d_w); cudaFree(d_m);
    if (tmp) cudaFree(tmp);
    return 0;
}
