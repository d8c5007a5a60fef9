// s2_cli_count.cpp - the drop-in `kmer_scrub_count` program (/root/reference/src/kmer_scrub_count.c:29-156,
// list drivers src/genome_compare.c:115-177) on top of the C ABI.  Same getopt string, same required
// arguments, usage text, progress file, stderr notes and count-table bytes; the scanning itself is
// s2_batch_submit_count() on the GPU.  Host threads only inflate + parse files into pinned batches.
//
// Extra controls come from the environment so argv stays drop-in:
//   S2_GPUS (1: GPUs to shard the input files over)  S2_DEVICE (0)  S2_THREADS (min(nproc,16) reader threads)  S2_BATCH_MB (16)  S2_LOAD (0.5)
//   S2_GPU_INGEST (1: BGZF-compressed strict FASTQ / FASTA are inflated by the hardware engine and split into
//                  records on the GPU, uncompressed files too unless S2_GPU_INGEST_PLAIN=0)
//   S2_STATS=1 prints a one-line throughput summary on stderr.
#include "../../include/strainer2_b200.h"
#include "s2_internal.h"

#include <getopt.h>
#include <fcntl.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <mutex>
#include <string>
#include <thread>
#include <sys/stat.h>
#include <vector>

static void count_usage()
{
    fprintf(stderr, "Usage: kmer_scrub_count -r <reference genome>  -A <file with multiple genome filenames> -B <file with multiple "
                    "metagenome filenames> -C <(optional) file with multiple genome filenames of drug strains> -p [progress output "
                    "file, optional]\n");
}

// One flat byte stream for s2_table_build: sequence bytes of every record, '\n' between records.
// Returns -1 if the file cannot be opened or its gzip data is damaged (the reference never returns from such a file).
int s2_load_flat(const char *path, std::vector<uint8_t> &flat)
{
    s2_reader *r = s2_reader_open(path);
    if (!r) return -1;
    const char *seq;
    int64_t l;
    while ((l = s2_reader_next(r, &seq)) >= 0) {
        flat.insert(flat.end(), (const uint8_t *)seq, (const uint8_t *)seq + l);
        flat.push_back('\n');
    }
    const bool damaged = s2_reader_damaged(r) != 0;
    s2_reader_close(r);
    if (damaged) { s2_set_error("damaged gzip data in %s", path); return -1; }
    return 0;
}


// reads the newline separated list exactly like the reference: getline, cut at the first '\n' only
int s2_read_list(const char *list_file, int col, const char *skip_file, std::vector<S2WorkItem> &out)
{
    FILE *fp = fopen(list_file, "r");
    if (!fp) {
        fprintf(stderr, "could not read file %s in GEN_all_kmer_counts()\n", list_file);   // src/genome_compare.c:125,159
        return -1;
    }
    char *line = nullptr; size_t cap = 0;
    while (getline(&line, &cap, fp) != -1) {
        char *pos = strchr(line, '\n');
        if (pos) *pos = '\0';
        out.push_back({ line, col, skip_file && strcmp(skip_file, line) == 0 });
    }
    free(line);
    fclose(fp);
    return 0;
}

struct BatchWriter {
    s2_ctx *ctx; s2_table *table;
    uint8_t *buf = nullptr; uint64_t cap = 0, used = 0; int col = -1;
    bool failed = false;

    bool flush()
    {
        if (!buf) return true;
        int rc = used ? s2_batch_submit_count(ctx, table, buf, used, col) : s2_batch_release(ctx, buf);
        buf = nullptr; used = 0;
        if (rc) failed = true;
        return rc == 0;
    }
    bool ensure(int want_col)
    {
        if (buf && col != want_col && !flush()) return false;
        if (!buf) {
            buf = s2_batch_acquire(ctx, &cap);
            if (!buf) { failed = true; return false; }
            used = 0; col = want_col;
        }
        return true;
    }
    // append one record (+ separator); records longer than the space left are split with a 30 byte
    // overlap so that every 31-byte window lands in exactly one batch
    bool append(const char *seq, uint64_t len, int want_col)
    {
        if (len < S2_K) return true;                       // src/genome_compare.c:204: no window fits
        while (len) {
            if (!ensure(want_col)) return false;
            if (cap - used < 4096 && used) { if (!flush()) return false; continue; }
            const uint64_t space = cap - used - 1;
            const uint64_t take = len < space ? len : space;
            memcpy(buf + used, seq, take);
            used += take;
            if (take < len) {
                if (!flush()) return false;
                seq += take - (S2_K - 1); len -= take - (S2_K - 1);
            } else {
                buf[used++] = '\n';
                len = 0;
            }
        }
        return true;
    }
};

int s2_default_reader_threads()
{
    int n = s2_env_int("S2_THREADS", 0);
    if (n <= 0) { n = (int)std::thread::hardware_concurrency(); if (n > 16) n = 16; }
    return n < 1 ? 1 : n;
}

// The list drivers GEN_all_kmer_counts / GEN_all_kmer_counts_skip_file (src/genome_compare.c:115-177) for a
// Size of a reader thread's arena.  Pinning memory costs (profiles/r2g_pinned_probe.txt: 7 GB/s on huge pages), and sixteen
// threads pin two arenas each while the scan is running: 16 MB is one full pipeline chunk of BGZF.  Ordinary .gz wants
// larger batches for its decode kernel (one warp per 64 KB of compressed bytes) and runs long enough to pay for them.
static uint64_t g_arena_default_mb = 16;
// Arenas per reader thread (S2_READ_ARENAS): two let a thread fill one while the job on the other is in flight; with one, the
// thread waits for its job before it reads on - half the memory to pin, which is what the big arenas of ordinary .gz cost.
static int g_arenas_default = 2;

// does the first file a list names begin like an ordinary (not block-) gzip file?  Nothing is reported here: the list is
// read again, with the reference's messages, when its turn comes
bool s2_list_starts_with_plain_gz(const char *list_file)
{
    if (!list_file) return false;
    FILE *fp = fopen(list_file, "r");
    if (!fp) return false;
    char *line = nullptr; size_t cap = 0;
    bool gz = false;
    if (getline(&line, &cap, fp) > 0) {
        char *pos = strchr(line, '\n');
        if (pos) *pos = '\0';
        const char *path = line;
        if (const char *tab = strrchr(line, '\t')) path = tab + 1;          // (strain_detect batch lines: the last column is a file)
        FILE *f = fopen(path, "rb");
        if (f) {
            unsigned char h[16];
            const size_t n = fread(h, 1, sizeof h, f);
            gz = n >= 16 && h[0] == 0x1f && h[1] == 0x8b && h[2] == 8 && !((h[3] & 4) && h[12] == 'B' && h[13] == 'C');
            fclose(f);
        }
    }
    free(line);
    fclose(fp);
    return gz;
}

// whole work list: reader threads take files in list order (progress line and "skipping" note written
// at dispatch, under the lock, so their order is the reference's), inflate + parse them straight into
// pinned batches and submit those to the GPU.  Returns false after a failure; open_error carries the
// reference's message when a file could not be opened (dispatch stops there, like the reference's exit).
bool s2_scan_work_items(s2_ctx *ctx, s2_table *table, s2_exotic *exotic, std::vector<S2WorkItem> &work, int n_threads,
                        FILE *progress, std::string &open_error, uint64_t *bases_out, uint64_t *lookups_out)
{
    std::vector<s2_ctx *> c(1, ctx);
    std::vector<s2_table *> t(1, table);
    return s2_scan_work_items_multi(c, t, exotic, work, n_threads, progress, open_error, bases_out, lookups_out);
}

// same with one (context, table replica) per GPU: reader thread i feeds GPU i % n_gpus, files are still
// taken from one queue in list order, so any GPU may end up scanning any file (file sharding)
bool s2_scan_work_items_multi(std::vector<s2_ctx *> &ctxs, std::vector<s2_table *> &tables, s2_exotic *exotic,
                              std::vector<S2WorkItem> &work, int n_threads, FILE *progress, std::string &open_error,
                              uint64_t *bases_out, uint64_t *lookups_out)
{
    std::mutex mu;                      // guards next / progress / stderr ordering
    const bool gpu_ingest = s2_env_int("S2_GPU_INGEST", 1) != 0;
    size_t next = 0;
    std::atomic<bool> stop(false);
    std::atomic<uint64_t> total_bases(0), total_lookups(0);
    std::atomic<uint64_t> n_gpu_files(0), n_host_files(0);        // S2_STATS: which way the files went

    // GPU ingest takes runs of files (same counter column, S2_INGEST_BATCH files / S2_INGEST_BATCH_MB compressed bytes
    // at most) so that small files share a chunk; everything it does not handle goes through the host reader below
    const size_t max_run = gpu_ingest && !exotic ? (size_t)std::max(1, s2_env_int("S2_INGEST_BATCH", 256)) : 1;
    const uint64_t run_bytes = s2_env_u64("S2_INGEST_BATCH_MB", std::max<uint64_t>(g_arena_default_mb, 16)) << 20;      // a run fills an arena
    struct Taken { std::string path; s2_reader *r; uint64_t size; };
    // Reader threads read whole files into pinned ARENAS of their own (two per thread: one being filled while the job
    // on the other is in flight) and hand the images to the ingest pipeline, which copies them to the device from
    // there.  (Handing over paths made the pipeline read the files itself, under its lock: three pipelines = three
    // threads reading, 15 GB/s for all sixteen reader threads - profiles/r2d_bench_n1.json, cli leg.)
    const uint64_t arena_bytes = gpu_ingest && !exotic ? std::max<uint64_t>(s2_env_u64("S2_READ_ARENA_MB", g_arena_default_mb), 1) << 20 : 0;
    const bool one_arena = s2_env_int("S2_READ_ARENAS", g_arenas_default) < 2;

    auto reader = [&](int tid) {
        BatchWriter w{ ctxs[tid % ctxs.size()], tables[tid % tables.size()] };
        // a run handed to the GPU ingest stays pending while the next one is taken and submitted (the ingest pipeline never
        // drains between runs); whatever the ingest hands back is then read by the host parser below
        struct Pending { std::vector<Taken> run; int col = 0; s2_ingest_job *job = nullptr; };
        Pending pending;
        uint8_t *arena[2] = { nullptr, nullptr };
        int arena_at = 0;
        struct ArenaGuard { uint8_t **a; ~ArenaGuard() { s2_pinned_free(a[0]); s2_pinned_free(a[1]); } } arena_guard{ arena };
        // the run's files, back to back in the arena; false: something did not fit or changed size (the run goes by path)
        auto read_run = [&](std::vector<Taken> &run, uint8_t *dst, std::vector<const void *> &images, std::vector<uint64_t> &sizes) -> bool {
            uint64_t at = 0;
            images.clear(); sizes.clear();
            for (auto &x : run) {
                if (at + x.size > arena_bytes) return false;
                const int fd = open(x.path.c_str(), O_RDONLY);
                if (fd < 0) return false;
                uint64_t done = 0;
                while (done < x.size) {
                    const ssize_t r = pread(fd, dst + at + done, x.size - done, (off_t)done);
                    if (r <= 0) break;
                    done += (uint64_t)r;
                }
                uint8_t extra;
                const bool grown = done == x.size && pread(fd, &extra, 1, (off_t)done) > 0;
                close(fd);
                if (done != x.size || grown) return false;
                images.push_back(dst + at); sizes.push_back(x.size);
                at += x.size;
            }
            return true;
        };
        auto host_read = [&](std::vector<Taken> &run, const std::vector<int> &handled, int col) {
            for (size_t k = 0; k < run.size(); ++k) {
                s2_reader *r = run[k].r;
                if (handled[k] == 0) ++n_gpu_files; else ++n_host_files;
                if (handled[k] == 0 || w.failed) { s2_reader_close(r); continue; }
                const char *seq; int64_t l; uint64_t bases = 0, lookups = 0;
                while ((l = s2_reader_next(r, &seq)) >= 0) {
                    bases += (uint64_t)l;
                    if (l >= S2_K) lookups += (uint64_t)l - (S2_K - 1);
                    if (!w.append(seq, (uint64_t)l, col)) {
                        std::lock_guard<std::mutex> g(mu);
                        if (open_error.empty()) open_error = s2_last_error();
                        stop.store(true);
                        break;
                    }
                    if (exotic) s2_exotic_count_record(exotic, seq, (uint64_t)l, col);
                }
                if (s2_reader_damaged(r)) {
                    // zlib met corrupt DEFLATE data or a CRC mismatch.  The reference spins for ever on this (s2_reader_damaged in
                    // include/strainer2_b200.h); counting the part before the damage without a word would be worse than either
                    std::lock_guard<std::mutex> g(mu);
                    if (open_error.empty()) open_error = "damaged gzip data in " + run[k].path + " (gzread error); nothing is printed";
                    stop.store(true);
                }
                s2_reader_close(r);
                total_bases += bases; total_lookups += lookups;
            }
        };
        auto finish_pending = [&]() -> bool {
            if (!pending.job) return true;
            std::vector<int> handled(pending.run.size(), 1);
            uint64_t gb = 0, gl = 0;
            const int rc = s2_ingest_wait(pending.job, handled.data(), &gb, &gl);
            pending.job = nullptr;
            if (rc < 0) {
                std::lock_guard<std::mutex> g(mu);
                if (open_error.empty()) open_error = s2_last_error();
                stop.store(true);
                for (auto &x : pending.run) s2_reader_close(x.r);
                pending.run.clear();
                return false;
            }
            total_bases += gb; total_lookups += gl;
            host_read(pending.run, handled, pending.col);
            pending.run.clear();
            return !w.failed;
        };
        for (;;) {
            std::vector<Taken> run;
            int col = 0;
            {
                std::lock_guard<std::mutex> g(mu);
                uint64_t bytes = 0;
                while (run.size() < max_run && bytes < run_bytes) {
                    if (stop.load() || next >= work.size()) break;
                    if (!run.empty() && work[next].col != col) break;
                    struct stat sb_next;
                    const uint64_t size_next = stat(work[next].path.c_str(), &sb_next) == 0 ? (uint64_t)sb_next.st_size : run_bytes;
                    if (!run.empty() && arena_bytes && !work[next].skip && bytes + size_next > arena_bytes) break;      // the run must fit an arena
                    S2WorkItem &it = work[next++];
                    if (progress) {
                        time_t now = time(nullptr);
                        fprintf(progress, "%s\t%s", it.path.c_str(), asctime(localtime(&now)));   // src/genome_compare.c:167-170
                    }
                    if (it.skip) { fprintf(stderr, "skipping %s (identical match)\n", it.path.c_str()); continue; }   // :141
                    s2_reader *r = s2_reader_open(it.path.c_str());
                    if (!r) {
                        open_error = "could not read file " + it.path + " in GEN_calculate_kmer_count()";   // :196
                        stop.store(true);
                        break;
                    }
                    col = it.col;
                    run.push_back({ it.path, r, size_next });
                    bytes += size_next;
                }
            }
            if (run.empty()) break;
            if (gpu_ingest && !exotic) {
                // BGZF / plain strict FASTQ + FASTA: hardware inflate + record splitting on the GPU, nothing parsed here
                std::vector<const char *> paths;
                for (auto &x : run) paths.push_back(x.path.c_str());
                std::vector<const void *> images; std::vector<uint64_t> sizes;
                if (arena_bytes && !arena[arena_at]) arena[arena_at] = (uint8_t *)s2_pinned_alloc(arena_bytes);
                s2_ingest_job *job;
                bool from_arena = false;
                if (arena_bytes && arena[arena_at] && read_run(run, arena[arena_at], images, sizes)) {
                    job = s2_ingest_submit_mem_batch(w.ctx, w.table, images.data(), sizes.data(), (int)images.size(), col);
                    from_arena = true;
                    if (!one_arena) arena_at ^= 1;           // the job before this one is finished below, before its arena is filled again
                } else {
                    job = s2_ingest_submit_files(w.ctx, w.table, paths.data(), (int)paths.size(), col);      // (a file bigger than an arena: streamed from the file)
                }
                if (!job) {
                    std::lock_guard<std::mutex> g(mu);
                    if (open_error.empty()) open_error = s2_last_error();
                    stop.store(true);
                    for (auto &x : run) s2_reader_close(x.r);
                    break;
                }
                if (!finish_pending()) { pending.run.swap(run); pending.col = col; pending.job = job; break; }
                pending.run.swap(run); pending.col = col; pending.job = job;
                if (one_arena && from_arena && !finish_pending()) break;      // the one arena is free again when its job is done
            } else {
                host_read(run, std::vector<int>(run.size(), 1), col);
            }
            if (w.failed) break;
        }
        finish_pending();
        if (!w.flush()) {
            std::lock_guard<std::mutex> g(mu);
            if (open_error.empty()) open_error = s2_last_error();
            stop.store(true);
        }
        s2_ingest_thread_cleanup();
    };
    std::vector<std::thread> pool;
    for (int i = 1; i < n_threads; ++i) pool.emplace_back(reader, i);
    reader(0);
    for (auto &t : pool) t.join();

    if (bases_out) *bases_out = total_bases.load();
    if (lookups_out) *lookups_out = total_lookups.load();
    if (s2_ingest_engine_failed()) { open_error = s2_last_error(); stop.store(true); }      // the cause, not whichever thread noticed first
    if (s2_env_int("S2_STATS", 0))
        fprintf(stderr, "[s2] files: %llu inflated + split on the GPU (BGZF / uncompressed), %llu through host zlib + parser (ordinary .gz, irregular text)\n",
                (unsigned long long)n_gpu_files.load(), (unsigned long long)n_host_files.load());
    return !stop.load() || !open_error.empty();
}

extern "C" int s2_kmer_scrub_count_main(int argc, char **argv)
{
    char *A_file = nullptr, *B_file = nullptr, *C_file = nullptr, *r_file = nullptr, *p_file = nullptr;
    int c;
    optind = 1;
    while ((c = getopt(argc, argv, "A:B:C:r:p:Hhud")) != EOF)     // src/kmer_scrub_count.c:52
        switch (c) {
        case 'A': A_file = optarg; break;
        case 'B': B_file = optarg; break;
        case 'C': C_file = optarg; break;
        case 'r': r_file = optarg; break;
        case 'p': p_file = optarg; break;
        case 'd': break;                                         // parsed and ignored (:63)
        case 'u': count_usage(); break;                          // usage, but no exit (:64-66)
        case 'h': count_usage(); break;
        default: count_usage(); break;
        }
    if (!r_file || !A_file || !B_file) { count_usage(); return 1; }   // :72-75

    FILE *progress = nullptr;
    if (p_file) {
        progress = fopen(p_file, "w");
        if (!progress) { fprintf(stderr, "could not open progress file %s\n", p_file); return EXIT_FAILURE; }
        fprintf(progress, "adding kmer counts for:\n");          // :84
    }
    auto fail = [&](const char *msg) {
        if (msg) fprintf(stderr, "%s\n", msg);
        if (progress) fclose(progress);                           // exit() would flush it the same way
        return EXIT_FAILURE;
    };

    const auto t_start = std::chrono::steady_clock::now();
    // S2_GPUS > 1: one context + table replica per GPU, files sharded over them, ONE all-reduce at the end
    int n_gpus = s2_env_int("S2_GPUS", 1);
    if (n_gpus < 1) n_gpus = 1;
    if (n_gpus > 1 && n_gpus > s2_device_count()) return fail("S2_GPUS exceeds the number of visible GPUs");
    const int n_threads = std::max(s2_default_reader_threads(), n_gpus);
    std::vector<s2_ctx *> ctxs(n_gpus, nullptr);
    std::vector<s2_table *> tables(n_gpus, nullptr);
    // The CUDA contexts come up side by side (each takes the better part of a second) while this thread inflates and
    // parses the -r genome; each starter then creates its context's ingest pipelines, which goes on beside the table build.
    const bool gpu_ingest_on = s2_env_int("S2_GPU_INGEST", 1) != 0;
    const int warm_pipes = gpu_ingest_on ? std::min((n_threads + n_gpus - 1) / n_gpus, std::max(1, s2_env_int("S2_INGEST_PIPES", 2))) : 0;
    std::vector<std::thread> starters, warmers;
    std::vector<std::string> start_errs(n_gpus);
    for (int g = 0; g < n_gpus; ++g)
        starters.emplace_back([&, g]() {
            ctxs[g] = s2_init(s2_env_int("S2_DEVICE", 0) + g, s2_env_u64("S2_BATCH_MB", 16) << 20, (n_threads + n_gpus - 1) / n_gpus + 2);
            if (!ctxs[g]) start_errs[g] = s2_last_error();
        });
    auto join_all = [](std::vector<std::thread> &v) { for (auto &t : v) if (t.joinable()) t.join(); };

    // ---- table from -r (GEN_hash_sequences_set_count_vec, default 1 / increment 1 / column 0 / 4 wide)
    std::vector<uint8_t> flat;
    const int load_rc = s2_load_flat(r_file, flat);
    const auto t_loaded = std::chrono::steady_clock::now();
    join_all(starters);
    const auto t_ctx = std::chrono::steady_clock::now();
    for (int g = 0; g < n_gpus; ++g) if (!ctxs[g]) return fail(start_errs[g].c_str());
    if (load_rc != 0) {
        fprintf(stderr, "could not read file %s GEN_hash_sequences_set_count_vec()\n", r_file);   // src/genome_compare.c:986
        return fail(nullptr);
    }
    // (an ordinary .gz at the head of a list: the pipelines' gunzip stages are made ready as well - s2_ingest_warm_gz)
    const bool gz_inputs = warm_pipes && (s2_list_starts_with_plain_gz(A_file) || s2_list_starts_with_plain_gz(B_file));
    if (gz_inputs) g_arena_default_mb = 192;
    const uint64_t arena_mb = std::max<uint64_t>(s2_env_u64("S2_READ_ARENA_MB", g_arena_default_mb), 1);
    for (int g = 0; g < n_gpus && warm_pipes; ++g)
        warmers.emplace_back([&, g]() { if (gz_inputs) s2_ingest_warm_gz(ctxs[g], warm_pipes, arena_mb << 20); else s2_ingest_warm(ctxs[g], warm_pipes); });
    struct JoinGuard { std::vector<std::thread> &v; ~JoinGuard() { for (auto &t : v) if (t.joinable()) t.join(); } } warm_guard{ warmers };
    // windows of the -r genome with a byte outside ACGTN become string keys on the host;
    // nullptr (the normal case) means no such window exists and the host never looks at a window again
    s2_exotic *exotic = s2_exotic_build(flat.data(), flat.size(), 4);
    const char *load_env = getenv("S2_LOAD");
    {
        std::vector<std::thread> builders;
        std::vector<std::string> errs(n_gpus);
        for (int g = 0; g < n_gpus; ++g)
            builders.emplace_back([&, g]() {
                tables[g] = s2_table_build(ctxs[g], flat.data(), flat.size(), 4, load_env ? atof(load_env) : 0.0, 0);
                if (!tables[g]) errs[g] = s2_last_error();
            });
        for (auto &b : builders) b.join();
        for (int g = 0; g < n_gpus; ++g) if (!tables[g]) return fail(errs[g].c_str());
    }
    s2_table *table = tables[0];
    std::vector<uint8_t>().swap(flat);
    const auto t_built = std::chrono::steady_clock::now();

    // ---- work list: -A into column 1, -B into column 2, -C (skipping -r itself) into column 3
    // The reference opens each list only when it gets to it, so an unreadable later list is reported
    // after the earlier lists have been scanned; the output (nothing on stdout, EXIT_FAILURE) is the same.
    std::vector<S2WorkItem> work;
    if (s2_read_list(A_file, 1, nullptr, work)) return fail(nullptr);
    if (s2_read_list(B_file, 2, nullptr, work)) return fail(nullptr);
    if (C_file && s2_read_list(C_file, 3, r_file, work)) return fail(nullptr);

    // The order of the output rows (the reference table's slot order, replayed from the keys' djb2 values: 0.2 s of
    // sequential host work for 5 M keys) depends on the -r genome alone: it is worked out beside the scan.
    const bool host_format = exotic != nullptr || s2_env_int("S2_HOST_FORMAT", 0);
    std::vector<uint32_t> early_order;
    int early_rc = 0; std::string early_err;
    std::thread order_thread;
    struct OrderGuard { std::thread &t; ~OrderGuard() { if (t.joinable()) t.join(); } } order_guard{ order_thread };
    if (!exotic && !host_format)
        order_thread = std::thread([&]() {
            const uint64_t nk = s2_table_n_keys(table);
            std::vector<uint32_t> h(nk);
            early_order.resize(nk);
            if (s2_table_export(table, nullptr, h.data(), nullptr) || s2_roworder_emulate(h.data(), nk, 0, early_order.data(), nullptr)) { early_rc = -1; early_err = s2_last_error(); }
        });

    std::string open_error;
    uint64_t total_bases = 0, total_lookups = 0;
    join_all(warmers);
    // (ordinary .gz: fewer readers with larger arenas - a 192 MB run is 3,000 decoding warps, and eight threads read faster than the GPU decodes)
    const int scan_threads = gz_inputs ? std::max(n_gpus, std::min(n_threads, 8 * n_gpus)) : n_threads;
    const bool pool_ok = s2_scan_work_items_multi(ctxs, tables, exotic, work, scan_threads, progress, open_error, &total_bases, &total_lookups);
    s2_scan_stats st = {};
    for (int g = 0; g < n_gpus; ++g) {
        s2_scan_stats sg = {};
        if (s2_sync(ctxs[g], &sg)) return fail(open_error.empty() ? s2_last_error() : open_error.c_str());
        st.hits += sg.hits; st.valid_windows += sg.valid_windows;
    }
    if (!open_error.empty()) return fail(open_error.c_str());
    if (!pool_ok) return fail(s2_last_error());
    const auto t_scan_end = std::chrono::steady_clock::now();
    for (int k = 1; k < 4 && n_gpus > 1; ++k)                          // sum the replicas' counters over NVLink
        if (s2_tables_allreduce(tables.data(), n_gpus, k)) return fail(s2_last_error());
    const auto t_scanned = std::chrono::steady_clock::now();

    // ---- print_hash_counts: rows in the reference table's slot order
    const uint64_t n = s2_table_n_keys(table);
    const int n_print = C_file ? 4 : 3;
    std::vector<uint64_t> keys;
    std::vector<uint32_t> djb2(n), pos, cols[4];
    const uint32_t *colp[4] = { nullptr, nullptr, nullptr, nullptr };
    if (order_thread.joinable()) order_thread.join();
    if (early_rc) return fail(early_err.c_str());
    if (host_format) {
        keys.resize(n);
        if (exotic) pos.resize(n);
        if (s2_table_export(table, keys.data(), djb2.data(), exotic ? pos.data() : nullptr)) return fail(s2_last_error());
        for (int k = 0; k < n_print; ++k) {
            cols[k].resize(n);
            if (s2_table_counts_fetch(table, k, cols[k].data())) return fail(s2_last_error());
            colp[k] = cols[k].data();
        }
    }
    if (!exotic) {
        // only the djb2 values leave the device for the row-order replay; the text is formatted on the GPU
        std::vector<uint32_t> order;
        if (host_format) {
            order.resize(n);
            if (s2_roworder_emulate(djb2.data(), n, 0, order.data(), nullptr)) return fail(s2_last_error());
        } else order.swap(early_order);
        if (host_format ? s2_format_count_table(stdout, keys.data(), order.data(), n, colp, n_print, n_threads)
                        : s2_table_format(table, order.data(), n_print, stdout)) return fail(s2_last_error());
    } else {
        // merge the device keys and the host string keys by first occurrence = the reference's insertion order
        std::vector<S2ExoRow> xr;
        s2_exotic_rows(exotic, xr);
        const uint64_t total = n + xr.size();
        std::vector<uint32_t> mdjb2(total), src(total), order(total);      // src: < n device key, else n + exotic index
        uint64_t a = 0, b = 0, o = 0;
        while (a < n || b < xr.size()) {
            if (b >= xr.size() || (a < n && pos[a] < xr[b].first_pos)) { mdjb2[o] = djb2[a]; src[o++] = (uint32_t)a++; }
            else { mdjb2[o] = xr[b].djb2; src[o++] = (uint32_t)(n + b++); }
        }
        if (s2_roworder_emulate(mdjb2.data(), total, 0, order.data(), nullptr)) return fail(s2_last_error());
        fputs("#kmer\treference_count\tpangenome_count\tmetagenome_count\tdrug_count\n", stdout);
        char spell[S2_K + 1];
        for (uint64_t r = 0; r < total; ++r) {
            const uint32_t id = src[order[r]];
            if (id < n) {
                s2_kmer_to_ascii(keys[id], spell);
                fputs(spell, stdout);
                for (int k = 0; k < n_print; ++k) printf("\t%d", (int)cols[k][id]);
            } else {
                const S2ExoRow &x = xr[id - n];
                fwrite(x.key.data(), 1, x.key.size(), stdout);
                for (int k = 0; k < n_print; ++k) printf("\t%d", (int)x.counts[k]);
            }
            fputc('\n', stdout);
        }
    }
    fflush(stdout);
    const auto t_done = std::chrono::steady_clock::now();

    if (s2_env_int("S2_STATS", 0)) {
        auto sec = [](auto a, auto b) { return std::chrono::duration<double>(b - a).count(); };
        double kms = 0; uint64_t kl = 0;
        for (int g = 0; g < n_gpus; ++g) { double m = 0; uint64_t l = 0; s2_kernel_time(ctxs[g], &m, &l, 0); kms += m; kl += l; }
        fprintf(stderr, "[s2] phases: read -r %.3fs | CUDA contexts up after %.3fs | table build %.3fs | scan %.3fs | all-reduce %.3fs | row order + format + write %.3fs\n",
                sec(t_start, t_loaded), sec(t_start, t_ctx), sec(t_ctx, t_built), sec(t_built, t_scan_end), sec(t_scan_end, t_scanned), sec(t_scanned, t_done));
        fprintf(stderr, "[s2] gpus=%d keys=%llu build=%.3fs scan=%.3fs print=%.3fs bases=%llu lookups=%llu hits=%llu "
                        "kernel_ms=%.3f launches=%llu scan_Gbases_per_s=%.3f\n",
                n_gpus, (unsigned long long)n, sec(t_start, t_built), sec(t_built, t_scanned), sec(t_scanned, t_done),
                (unsigned long long)total_bases, (unsigned long long)total_lookups,
                (unsigned long long)st.hits, kms, (unsigned long long)kl,
                total_bases / 1e9 / std::max(1e-9, sec(t_built, t_scanned)));
    }
    s2_exotic_free(exotic);
    for (int g = 0; g < n_gpus; ++g) { s2_table_free(tables[g]); s2_shutdown(ctxs[g]); }
    if (progress) fclose(progress);
    return 0;
}
