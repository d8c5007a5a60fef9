// s2_cli_batch.cpp - `kmer_scrub_count_batch`: many strains against the same scrub lists in ONE pass
// (BASELINE config #5: multi-strain FMT donor batch).
//
// The reference has no such mode: README.md:47 runs one kmer_scrub_count process per strain, i.e. it
// re-reads and re-scans every genome / metagenome once per strain.  Here every strain's 31-mers go into
// one union table, the -A / -B / -C lists are inflated, parsed and scanned ONCE, and each strain's table
// is then read back out of the union counters.  Every output file is byte-identical to
//     kmer_scrub_count -r <strain> -A <listA> -B <listB> [-C <listC>]
// because a key's pangenome / metagenome count does not depend on which strain asks, and the one
// strain-dependent rule - "-C skips the file whose path equals -r" (src/genome_compare.c:138-141) - is
// the union count minus that strain's own occurrences: scanning the -r file adds exactly
// reference_count(key) to every key, so drug_count = union_drug_count - m * reference_count, with m = how
// often the strain's path appears in the -C list.
//
//   kmer_scrub_count_batch -R <file with one strain genome path per line> -A <listA> -B <listB>
//                          [-C <listC>] -O <output directory> [-p <progress file>]
// writes <output directory>/<basename of the strain path>.scrub_kmer_counts for every strain.
#include "../../include/strainer2_b200.h"
#include "s2_internal.h"

#include <getopt.h>
#include <sys/stat.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <vector>

int s2_load_flat(const char *path, std::vector<uint8_t> &flat);   // s2_cli_count.cpp

static void batch_usage()
{
    fprintf(stderr, "Usage: kmer_scrub_count_batch -R <file with strain genome filenames> -A <file with multiple genome filenames> "
                    "-B <file with multiple metagenome filenames> -C <(optional) file with drug strain genome filenames> "
                    "-O <output directory> -p [progress output file, optional]\n");
}

struct Strain {
    std::string path, out_path;
    std::vector<uint64_t> keys;           // first-occurrence order
    std::vector<uint32_t> djb2, ref;
    unsigned self_in_C = 0;               // how often the strain's own path is listed in -C
};

extern "C" int s2_kmer_scrub_count_batch_main(int argc, char **argv)
{
    char *R_file = nullptr, *A_file = nullptr, *B_file = nullptr, *C_file = nullptr, *O_dir = nullptr, *p_file = nullptr;
    int c;
    optind = 1;
    while ((c = getopt(argc, argv, "R:A:B:C:O:p:h")) != EOF)
        switch (c) {
        case 'R': R_file = optarg; break;
        case 'A': A_file = optarg; break;
        case 'B': B_file = optarg; break;
        case 'C': C_file = optarg; break;
        case 'O': O_dir = optarg; break;
        case 'p': p_file = optarg; break;
        default: batch_usage(); break;
        }
    if (!R_file || !A_file || !B_file || !O_dir) { batch_usage(); return 1; }
    auto die = [](const char *msg) { fprintf(stderr, "%s\n", msg); return EXIT_FAILURE; };
    const auto t0 = std::chrono::steady_clock::now();

    // ---- the strains ----------------------------------------------------------------------------
    std::vector<S2WorkItem> strain_items;
    if (s2_read_list(R_file, 0, nullptr, strain_items)) return EXIT_FAILURE;
    if (strain_items.empty()) return die("no strain genomes listed in -R");
    mkdir(O_dir, 0777);
    std::vector<Strain> strains(strain_items.size());
    std::set<std::string> names;
    for (size_t i = 0; i < strains.size(); ++i) {
        strains[i].path = strain_items[i].path;
        const size_t slash = strains[i].path.find_last_of('/');
        const std::string base = slash == std::string::npos ? strains[i].path : strains[i].path.substr(slash + 1);
        if (!names.insert(base).second) return die(("two strains share the file name " + base).c_str());
        strains[i].out_path = std::string(O_dir) + "/" + base + ".scrub_kmer_counts";
    }

    const int n_threads = s2_default_reader_threads();
    s2_ctx *ctx = s2_init(s2_env_int("S2_DEVICE", 0), s2_env_u64("S2_BATCH_MB", 16) << 20, n_threads + 2);
    if (!ctx) return die(s2_last_error());

    // per strain: its own small table gives the keys in first-occurrence order, their djb2 and reference counts
    // (GEN_hash_sequences_set_count_vec); the bytes are also appended to the union stream.
    std::vector<uint8_t> union_flat, flat;
    for (auto &st : strains) {
        flat.clear();
        if (s2_load_flat(st.path.c_str(), flat) != 0) {
            fprintf(stderr, "could not read file %s GEN_hash_sequences_set_count_vec()\n", st.path.c_str());
            return EXIT_FAILURE;
        }
        if (s2_exotic *ex = s2_exotic_build(flat.data(), flat.size(), 4)) {
            s2_exotic_free(ex);
            return die(("strain " + st.path + " contains bytes other than ACGTN: run kmer_scrub_count on it instead").c_str());
        }
        s2_table *t = s2_table_build(ctx, flat.data(), flat.size(), 1, 0.0, 0);
        if (!t) return die(s2_last_error());
        const uint64_t n = s2_table_n_keys(t);
        st.keys.resize(n); st.djb2.resize(n); st.ref.resize(n);
        if (s2_table_export(t, st.keys.data(), st.djb2.data(), nullptr) || s2_table_counts_fetch(t, 0, st.ref.data())) return die(s2_last_error());
        s2_table_free(t);
        union_flat.insert(union_flat.end(), flat.begin(), flat.end());
    }
    s2_table *U = s2_table_build(ctx, union_flat.data(), union_flat.size(), 4, 0.0, 0);
    if (!U) return die(s2_last_error());
    const uint64_t union_keys = s2_table_n_keys(U);
    // S2_GPUS > 1: a replica of the union table per GPU, the input files sharded over them, one all-reduce
    // per counter column at the end; every strain's counters are then read from replica 0 as before
    int n_gpus = s2_env_int("S2_GPUS", 1);
    if (n_gpus < 1) n_gpus = 1;
    if (n_gpus > 1 && n_gpus > s2_device_count() - s2_env_int("S2_DEVICE", 0)) return die("S2_GPUS exceeds the number of visible GPUs");
    std::vector<s2_ctx *> ctxs(1, ctx);
    std::vector<s2_table *> tables(1, U);
    for (int g = 1; g < n_gpus; ++g) {
        s2_ctx *cg = s2_init(s2_env_int("S2_DEVICE", 0) + g, s2_env_u64("S2_BATCH_MB", 16) << 20, n_threads / n_gpus + 2);
        s2_table *tg = cg ? s2_table_build(cg, union_flat.data(), union_flat.size(), 4, 0.0, 0) : nullptr;
        if (!tg) return die(s2_last_error());
        ctxs.push_back(cg); tables.push_back(tg);
    }
    std::vector<uint8_t>().swap(union_flat);
    const auto t1 = std::chrono::steady_clock::now();

    // ---- ONE pass over -A, -B and (all of) -C ----------------------------------------------------------
    FILE *progress = nullptr;
    if (p_file) {
        progress = fopen(p_file, "w");
        if (!progress) { fprintf(stderr, "could not open progress file %s\n", p_file); return EXIT_FAILURE; }
        fprintf(progress, "adding kmer counts for:\n");
    }
    std::vector<S2WorkItem> work;
    if (s2_read_list(A_file, 1, nullptr, work)) return EXIT_FAILURE;
    if (s2_read_list(B_file, 2, nullptr, work)) return EXIT_FAILURE;
    if (C_file) {
        const size_t before = work.size();
        if (s2_read_list(C_file, 3, nullptr, work)) return EXIT_FAILURE;          // nothing is skipped in the union pass
        for (size_t i = before; i < work.size(); ++i)
            for (auto &st : strains) if (work[i].path == st.path) st.self_in_C += 1;
    }
    std::string open_error;
    uint64_t bases = 0, lookups = 0;
    const bool ok = s2_scan_work_items_multi(ctxs, tables, nullptr, work, std::max(n_threads, n_gpus), progress, open_error, &bases, &lookups);
    s2_scan_stats stats = {};
    for (s2_ctx *cg : ctxs) {
        s2_scan_stats sg = {};
        if (s2_sync(cg, &sg)) return die(open_error.empty() ? s2_last_error() : open_error.c_str());
        stats.hits += sg.hits; stats.valid_windows += sg.valid_windows;
    }
    for (int k = 1; k < 4 && n_gpus > 1 && ok && open_error.empty(); ++k)
        if (s2_tables_allreduce(tables.data(), n_gpus, k)) return die(s2_last_error());
    if (progress) fclose(progress);
    if (!open_error.empty()) return die(open_error.c_str());
    if (!ok) return die(s2_last_error());
    const auto t2 = std::chrono::steady_clock::now();

    // ---- every strain's table: counters by key out of the union table, row order replayed per strain -----
    const int n_print = C_file ? 4 : 3;
    std::atomic<size_t> next(0);
    std::atomic<bool> failed(false);
    std::string fail_msg;
    std::mutex mu;
    std::vector<std::vector<uint32_t>> cols_of(strains.size() * 3);
    for (size_t i = 0; i < strains.size(); ++i)            // device reads first (one thread owns the context)
        for (int k = 1; k < n_print; ++k) {
            auto &v = cols_of[i * 3 + (k - 1)];
            v.resize(strains[i].keys.size());
            if (s2_table_counts_by_key(U, k, strains[i].keys.data(), strains[i].keys.size(), v.data())) return die(s2_last_error());
        }
    auto writer = [&]() {
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= strains.size() || failed.load()) break;
            Strain &st = strains[i];
            const uint64_t n = st.keys.size();
            std::vector<uint32_t> order(n);
            if (n_print == 4 && st.self_in_C) {
                auto &drug = cols_of[i * 3 + 2];
                for (uint64_t r = 0; r < n; ++r) drug[r] -= st.self_in_C * st.ref[r];      // the skipped self scan, uint32 wrap-around
            }
            const uint32_t *colp[4] = { st.ref.data(), cols_of[i * 3].data(), cols_of[i * 3 + 1].data(), n_print == 4 ? cols_of[i * 3 + 2].data() : nullptr };
            FILE *out = fopen(st.out_path.c_str(), "w");
            bool good = out && s2_roworder_emulate(st.djb2.data(), n, 0, order.data(), nullptr) == 0 &&
                        s2_format_count_table(out, st.keys.data(), order.data(), n, colp, n_print, 2) == 0;
            if (out) good = (fclose(out) == 0) && good;
            if (!good) {
                std::lock_guard<std::mutex> g(mu);
                fail_msg = "could not write " + st.out_path;
                failed.store(true);
            }
        }
    };
    std::vector<std::thread> pool;
    for (int w = 0; w < std::max(1, std::min<int>(n_threads, (int)strains.size())); ++w) pool.emplace_back(writer);
    for (auto &t : pool) t.join();
    if (failed.load()) return die(fail_msg.c_str());
    const auto t3 = std::chrono::steady_clock::now();

    if (s2_env_int("S2_STATS", 0)) {
        auto sec = [](auto a, auto b) { return std::chrono::duration<double>(b - a).count(); };
        double kms = 0; uint64_t kl = 0;
        s2_kernel_time(ctx, &kms, &kl, 0);
        fprintf(stderr, "[s2 batch] strains=%zu union_keys=%llu build=%.3fs scan=%.3fs write=%.3fs bases=%llu lookups=%llu hits=%llu "
                        "kernel_ms=%.3f launches=%llu\n", strains.size(), (unsigned long long)union_keys, sec(t0, t1), sec(t1, t2), sec(t2, t3),
                (unsigned long long)bases, (unsigned long long)lookups, (unsigned long long)stats.hits, kms, (unsigned long long)kl);
    }
    for (size_t g = 0; g < ctxs.size(); ++g) { s2_table_free(tables[g]); s2_shutdown(ctxs[g]); }
    return 0;
}
