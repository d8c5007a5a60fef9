// s2_cli_detect.cpp - the drop-in `strain_detect` program (/root/reference/src/strain_detect.c) on top of
// the C ABI.  Same getopt string, validations, stdout chatter, stderr errors and kmer_hits text/.gz.
//
// Split of work: the GPU does pass 1 of quantify_hits_PE for every read of a batch (table hits,
// informative hits, positions of informative windows: s2_scan_detect).  The host keeps what is
// inherently sequential: the reference's read-pairing loop with its stale-state behaviour for reads
// shorter than 31 (src/strain_detect.c:444-448, :497-504) and the ordered emission of
// pass 2 (:547-623) into one gzip stream ("wb9", :299).
//
// Environment: S2_DEVICE (0), S2_DETECT_BATCH_MB (32), S2_GPU_INGEST (1), S2_THREADS (worker threads over batch lines).
#include "../../include/strainer2_b200.h"
#include "s2_internal.h"

#include <getopt.h>
#include <zlib.h>

#include <algorithm>
#include <cerrno>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_set>
#include <vector>

int s2_load_flat(const char *path, std::vector<uint8_t> &flat);   // s2_cli_count.cpp

enum { NOT_PAIRED_END = 0, IS_PAIRED_END = 1, IS_PAIRED_END_INTERLEAVE = 2, UNKNOWN_FILE_TYPE = -1 };

static void detect_usage()                                          // src/strain_detect.c:750-764
{
    fprintf(stderr, "Usage paired end with 2 files:\n\tstrain_detect -r <reference_genome.fna> -a <informative_kmer_file.txt> -b <paired-end-file1> -c <paired-end-file1> -t PE -o <kmer outfile>\n");
    fprintf(stderr, "Usage paired end interleaved 1 file:\n\tstrain_detect -r <reference_genome.fna> -a <informative_kmer_file.txt> -b <paired-end-file1>  -t PEI -o <kmer outfile>\n");
    fprintf(stderr, "Usage single end 1 file:\n\tstrain_detect -r <reference_genome.fna> -a <informative_kmer_file.txt> -b <single-end-file1>  -t SE -o <kmer outfile>\n");
    fprintf(stderr, "Usage single end 1 file:\n\tstrain_detect -r <reference_genome.fna> -a <informative_kmer_file.txt> -B <batch-list-of-metagenomes> -o <kmer outfile>\n\n");
    fprintf(stderr, "format for metagenomics batch file is:\n");
    fprintf(stderr, "PE\tfile1_PE1.fasta\tfile1_PE2.fasta\n");
    fprintf(stderr, "SE\tfile1_PE1.fasta\n");
    fprintf(stderr, "PEI\tfile1_PE1.fasta\n");
    fprintf(stderr, "\nlines that begin with # are considered comments and ignored\n");
    fprintf(stderr, "\ninformative kmer file is a list of all of the kmers left in the reference genome post scrubbing\n");
}

static int get_file_type(const char *s)                             // src/strain_detect.c:728-747
{
    if (!strcmp(s, "SE") || !strcmp(s, "se")) return NOT_PAIRED_END;
    if (!strcmp(s, "PE") || !strcmp(s, "pe")) return IS_PAIRED_END;
    if (!strcmp(s, "PEI") || !strcmp(s, "pei") || !strcmp(s, "IPE") || !strcmp(s, "ipe")) return IS_PAIRED_END_INTERLEAVE;
    return UNKNOWN_FILE_TYPE;
}

struct Detect {
    s2_ctx *ctx = nullptr;
    s2_table *table = nullptr;
    s2_exotic *exotic = nullptr;        // string keys for -r windows with bytes outside ACGTN (normally nullptr)
    s2_gz_writer *gzout = nullptr;      // one zlib stream (byte-identical to the reference) or, S2_GZ_THREADS=n, the parallel writer
    unsigned genome_kmers = 0, genome_informative = 0;
    uint64_t batch_bytes = 32ull << 20;
    std::unordered_set<uint64_t> informative;   // device keys currently labelled INFORMATIVE
    bool gpu_ingest = true;             // S2_GPU_INGEST: BGZF / plain strict FASTQ are inflated + split on the GPU
    std::mutex stat_mu;
    double t_read = 0, t_gpu = 0, t_emit = 0;   // S2_STATS (summed over worker threads)
    uint64_t n_bases = 0;
    uint64_t files_gpu = 0, files_host = 0;     // input files inflated + split on the GPU / read by the host parser

    void write_out(const std::string &out)
    {
        s2_gz_writer_write(gzout, out.data(), out.size());
    }
};

// one batch line (or the -b/-c pair): processed by a worker thread, written to the gzip stream in order
struct Job {
    std::string f1, f2;
    bool has_f2 = false;
    int pe = 0;
    std::string out, err;
    int rc = 0;
    bool done = false;
};

// hash_scrubbed_kmers (src/strain_detect.c:668-726): label the listed k-mers INFORMATIVE.
// Lines are read with gzgets into a 100-byte buffer exactly like the reference (longer lines split).
static int label_informative(Detect &d, const char *a_file, unsigned *num_lines_found)
{
    gzFile fp = gzopen(a_file, "r");
    if (!fp) {
        fprintf(stderr, "could not read file %s in hash_scrubbed_kmers()\n", a_file);
        return -1;
    }
    struct Line { std::string text; bool right_len; bool acgt; uint64_t kmer; };
    std::vector<Line> lines;
    char line[100];
    while (gzgets(fp, line, 100)) {
        if (line[0] == '#') continue;
        char *pos = strchr(line, '\n');
        if (pos) *pos = '\0';
        Line L; L.text = line; L.right_len = strlen(line) == S2_K; L.acgt = false; L.kmer = 0;
        if (L.right_len) {
            // the reference does NOT upper-case these lines: only upper-case ACGT spellings can equal a key
            bool upper = true;
            for (int i = 0; i < S2_K; ++i) upper = upper && (line[i] == 'A' || line[i] == 'C' || line[i] == 'G' || line[i] == 'T');
            if (upper && s2_kmer_from_ascii(line, &L.kmer) == 0) L.acgt = true;
        }
        lines.push_back(L);
    }
    gzclose(fp);
    std::vector<uint64_t> q;
    for (auto &L : lines) if (L.acgt) q.push_back(L.kmer);
    std::vector<uint8_t> found(q.size() + 1);
    if (s2_table_flag(d.table, q.data(), q.size(), found.data())) return -2;
    size_t qi = 0; unsigned n_found = 0;
    std::unordered_set<uint64_t> &distinct = d.informative;
    for (auto &L : lines) {
        if (!L.right_len) {
            printf("error string length in the scrubbed kmer file (%s) must be the same size as the kmer length (scrubbed kmer, "
                   "scrubbed kmer len, seed len): %s, %d, %d\n", a_file, L.text.c_str(), (int)L.text.size(), S2_K);
            continue;
        }
        bool hit = L.acgt && found[qi++];
        if (!L.acgt && d.exotic) hit = s2_exotic_flag(d.exotic, L.text.c_str());   // raw spelling, string semantics
        if (hit) { ++n_found; if (L.acgt) distinct.insert(L.kmer); }
        else printf("error could not find informative kmer %s in the total kmer list\n", L.text.c_str());
    }
    *num_lines_found = n_found;
    d.genome_informative = (unsigned)(distinct.size() + s2_exotic_n_informative(d.exotic));   // src/strain_detect.c:285-290
    return 0;
}

// background_filter (src/strain_detect.c:160-240): count the informative k-mers in background metagenomes
// (the count scan into column 5, on the GPU), then demote those at or above a threshold chosen so that at
// most half of them go.  num_inform = matched informative-list LINES (duplicates count), as in the reference.
static int background_filter(Detect &d, const char *background_file, unsigned num_inform, int n_threads)
{
    const double fraction = 0.5;                                                           // :82
    const unsigned keep = (unsigned)(int)(num_inform * fraction);
    printf("#removing %f proportion of %s kmers; informative %d keep at least %d\n", fraction, background_file, num_inform, keep);
    std::vector<S2WorkItem> work;
    if (s2_read_list(background_file, 5, nullptr, work)) return EXIT_FAILURE;               // GEN_all_kmer_counts(..., 5, NULL)
    std::string open_error;
    const bool ok = s2_scan_work_items(d.ctx, d.table, d.exotic, work, n_threads, nullptr, open_error, nullptr, nullptr);
    if (s2_sync(d.ctx, nullptr)) { fprintf(stderr, "%s\n", open_error.empty() ? s2_last_error() : open_error.c_str()); return EXIT_FAILURE; }
    if (!open_error.empty()) { fprintf(stderr, "%s\n", open_error.c_str()); return EXIT_FAILURE; }
    if (!ok) { fprintf(stderr, "%s\n", s2_last_error()); return EXIT_FAILURE; }

    // background counts of the informative keys (device keys by first-occurrence rank + string keys)
    const uint64_t n = s2_table_n_keys(d.table);
    std::vector<uint64_t> keys(n);
    std::vector<uint32_t> bg(n);
    if (s2_table_export(d.table, keys.data(), nullptr, nullptr) || s2_table_counts_fetch(d.table, 5, bg.data())) {
        fprintf(stderr, "%s\n", s2_last_error());
        return EXIT_FAILURE;
    }
    std::vector<unsigned> v;
    std::vector<uint64_t> inf_keys; std::vector<uint32_t> inf_bg;
    for (uint64_t i = 0; i < n; ++i)
        if (d.informative.count(keys[i])) { inf_keys.push_back(keys[i]); inf_bg.push_back(bg[i]); v.push_back(bg[i]); }
    std::vector<S2ExoRow> xr;
    s2_exotic_rows(d.exotic, xr);
    std::vector<std::string> x_inf;
    for (auto &x : xr) if (s2_exotic_is_informative(d.exotic, x.key.c_str())) { v.push_back(x.counts[5]); x_inf.push_back(x.key); }
    if (v.size() > num_inform) { fprintf(stderr, "Error: too many background kmers\n"); return 1; }   // :187-190
    v.resize(num_inform, 0u);                                                              // calloc'd tail
    std::sort(v.begin(), v.end(), [](unsigned a, unsigned b) { return a > b; });            // qsort, descending (:196)
    unsigned thr = 1;
    if (keep >= 1 && v[keep - 1] > thr) thr = v[keep - 1];                                  // :207-208
    auto removed = [&](unsigned t) { unsigned c = 0; for (unsigned x : v) c += x >= t ? 1 : 0; return c; };
    while (removed(thr) > keep) ++thr;                                                     // :212-214
    std::vector<uint64_t> demote;
    unsigned n_demoted = 0;
    for (size_t i = 0; i < inf_keys.size(); ++i)
        if (inf_bg[i] >= thr) { demote.push_back(inf_keys[i]); d.informative.erase(inf_keys[i]); ++n_demoted; }
    for (auto &x : xr)
        if (s2_exotic_is_informative(d.exotic, x.key.c_str()) && x.counts[5] >= thr) { s2_exotic_set_informative(d.exotic, x.key.c_str(), false); ++n_demoted; }
    if (s2_table_unflag(d.table, demote.data(), demote.size())) { fprintf(stderr, "%s\n", s2_last_error()); return EXIT_FAILURE; }
    printf("#final_threshold %d removes %d background kmers %d removed\n", thr, removed(thr), n_demoted);   // :230
    d.genome_informative = (unsigned)(d.informative.size() + s2_exotic_n_informative(d.exotic));
    return 0;
}

// one record handed to the GPU
struct Rec { uint64_t off; uint32_t len; };

// quantify_hits_PE for files that went through the GPU ingest (s2_ingest_detect_file): every record's length,
// hits and informative hits are already known, so the reference's pairing loop (src/strain_detect.c:443-627)
// is replayed directly, including what happens when PE2 runs out (stale length, :496-504).
static int replay_ingested(Detect &d, Job &job, const s2_ingest_detect_result &A, const s2_ingest_detect_result *B, int is_pe)
{
    const char *pe1 = job.f1.c_str();
    const char *pe2 = job.has_f2 ? job.f2.c_str() : nullptr;
    int h1 = 0, i1 = 0, h2 = 0, i2 = 0;
    std::vector<std::string> copy_kmers, pe2_kmers;
    bool have_copy = false;
    unsigned long long evaluated = 0, reads = 0;
    uint64_t a = 0, b = 0, pa = 0, pb = 0;
    uint32_t stale_a = 0, stale_b = 0;                    // kseq's seq.l after the last successful read of each reader
    char kbuf[S2_K + 1], head[96];
    auto kmers_of = [&](const s2_ingest_detect_result &X, uint64_t rec, uint64_t &p, std::vector<std::string> &out) {
        out.clear();
        while (p < X.n_inf && X.inf_rec[p] < rec) ++p;
        for (; p < X.n_inf && X.inf_rec[p] == rec; ++p) { s2_kmer_to_ascii(X.inf_kmer[p], kbuf); out.push_back(kbuf); }
    };
    auto emit = [&](const std::string &kmer) {
        job.out += pe1;
        const int n = snprintf(head, sizeof head, "\t%d\t%d\t%d\t%d\t", h1, i1, h2, i2);
        job.out.append(head, n);
        job.out += kmer;
        job.out += '\n';
    };
    int rc = 0;
    while (a < A.n_records) {                                                        // :443
        const uint64_t rec1 = a++;
        const uint32_t len1 = A.len[rec1];
        stale_a = len1;                                                              // (an interleaved FASTA file that ends on PE1: reset by the failed PE2 read below)
        if (len1 >= S2_K) {                                                          // :444-449
            ++reads; h1 = (int)A.hits[rec1]; i1 = (int)A.inf[rec1];
            evaluated += len1 - (S2_K - 1);
            have_copy = true;
            kmers_of(A, rec1, pa, copy_kmers);
        }
        bool pe2_valid = false; uint64_t rec2 = 0;
        if (is_pe) {
            const bool shared = is_pe == IS_PAIRED_END_INTERLEAVE;
            const s2_ingest_detect_result &S = shared ? A : *B;
            uint64_t &cur = shared ? a : b;
            uint32_t &stale = shared ? stale_a : stale_b;
            int64_t l2 = -1;
            if (cur < S.n_records) { rec2 = cur++; stale = S.len[rec2]; l2 = stale; }
            else if (S.fasta) stale = 0;                                             // kseq.h:179 resets the length before it meets the end of a FASTA file
            if (stale >= S2_K) {                                                     // :497
                if (l2 < 0) {                                                        // :501-504
                    char msg[1024];
                    snprintf(msg, sizeof msg, "reached end of PE2 (%s) before end of PE1 (%s), check that file names are correct\n",
                             pe2 ? pe2 : "(null)", pe1);
                    job.err = msg;
                    rc = EXIT_FAILURE;
                    break;
                }
                h2 = (int)S.hits[rec2]; i2 = (int)S.inf[rec2];
                evaluated += stale - (S2_K - 1);
                pe2_valid = true;
            }
        }
        if (h1 + h2 >= 1 && i1 + i2 >= 1) {                                          // :547
            if (have_copy) for (const std::string &k : copy_kmers) emit(k);
            if (pe2_valid) {
                const bool shared = is_pe == IS_PAIRED_END_INTERLEAVE;
                kmers_of(shared ? A : *B, rec2, shared ? pa : pb, pe2_kmers);
                for (const std::string &k : pe2_kmers) emit(k);
            }
        }
    }
    if (rc == 0) {
        char foot[4][512];
        snprintf(foot[0], sizeof foot[0], "#%s\ttotal_kmer_evaluated\t%lld\n", pe1, (long long)evaluated);
        snprintf(foot[1], sizeof foot[1], "#%s\ttotal_reads_evaluated\t%lld\n", pe1, (long long)reads);
        snprintf(foot[2], sizeof foot[2], "#%s\ttotal_genome_kmers\t%lld\n", pe1, (long long)d.genome_kmers);
        snprintf(foot[3], sizeof foot[3], "#%s\ttotal_genome_informative_kmers\t%lld\n", pe1, (long long)d.genome_informative);
        for (auto &f : foot) job.out += f;
    }
    return rc;
}

// quantify_hits_PE (src/strain_detect.c:387-663).  Returns 0 or EXIT_FAILURE (message already printed).
static int quantify_hits(Detect &d, Job &job, s2_ctx *ctx, s2_table *table)
{
    const char *pe1 = job.f1.c_str();
    const char *pe2 = job.has_f2 ? job.f2.c_str() : nullptr;
    const int is_pe = job.pe;
    char msg[1024];
    double t_read = 0, t_gpu = 0, t_emit = 0; uint64_t n_bases = 0;
    // ---- GPU ingest: BGZF / plain strict FASTQ never reach the host parser ------------------------------
    if (d.gpu_ingest && !d.exotic) {
        const auto tI = std::chrono::steady_clock::now();
        s2_ingest_detect_result A, B;
        memset(&A, 0, sizeof A); memset(&B, 0, sizeof B);
        int ra = s2_ingest_detect_file(ctx, table, pe1, &A), rb = 0;
        if (ra == 0 && is_pe == IS_PAIRED_END) rb = s2_ingest_detect_file(ctx, table, pe2, &B);
        if (ra < 0 || rb < 0) { job.err = std::string(s2_last_error()) + "\n"; s2_ingest_detect_free(&A); s2_ingest_detect_free(&B); return EXIT_FAILURE; }
        if (ra == 0 && rb == 0) {
            const auto tJ = std::chrono::steady_clock::now();
            const int rc = replay_ingested(d, job, A, is_pe == IS_PAIRED_END ? &B : nullptr, is_pe);
            {
                std::lock_guard<std::mutex> g(d.stat_mu);
                d.t_gpu += std::chrono::duration<double>(tJ - tI).count();
                d.t_emit += std::chrono::duration<double>(std::chrono::steady_clock::now() - tJ).count();
                d.n_bases += A.bases + B.bases;
                d.files_gpu += is_pe == IS_PAIRED_END ? 2 : 1;
            }
            s2_ingest_detect_free(&A); s2_ingest_detect_free(&B);
            return rc;
        }
        s2_ingest_detect_free(&A); s2_ingest_detect_free(&B);          // not strict FASTQ / not BGZF: the host parser below
    }
    s2_reader *r1 = s2_reader_open(pe1), *r2 = nullptr;
    if (!r1) {
        snprintf(msg, sizeof msg, "could not read file (read1) %s in quantify_hits_PE() (error: %s)\n", pe1, strerror(errno));
        job.err = msg;
        return EXIT_FAILURE;
    }
    if (is_pe == IS_PAIRED_END) {
        r2 = s2_reader_open(pe2);
        if (!r2) {
            snprintf(msg, sizeof msg, "could not read file (read2) is_PE %s in quantify_hits_PE() (error: %s)\n", pe2, strerror(errno));
            job.err = msg;
            s2_reader_close(r1);
            return EXIT_FAILURE;
        }
    } else if (is_pe == IS_PAIRED_END_INTERLEAVE) r2 = r1;

    // sequential state of the reference loop
    int h1 = 0, i1 = 0, h2 = 0, i2 = 0;
    std::vector<std::string> copy_kmers;       // informative k-mers (canonical spelling, in order) of the PE1 copy
    bool have_copy = false;
    unsigned long long evaluated = 0, reads = 0;
    int rc = 0;
    bool eof = false;

    std::vector<uint8_t> batch;
    std::vector<uint64_t> rec_off;
    struct Iter { int32_t r1, r2; bool pe2_valid; bool fatal; };   // record indices in the batch (-1: shorter than 31)
    std::vector<Iter> iters;
    std::vector<uint32_t> hits, inf;
    std::vector<uint64_t> pos;
    char kbuf[S2_K + 1];

    auto now = []() { return std::chrono::steady_clock::now(); };
    auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    while (!eof && rc == 0) {
        const auto tA = now();
        // ---- phase A: run the pairing loop ahead, collecting records >= 31 into one batch -------------
        batch.clear(); rec_off.assign(1, 0); iters.clear();
        auto add = [&](const char *seq, uint64_t len) -> int32_t {
            batch.insert(batch.end(), (const uint8_t *)seq, (const uint8_t *)seq + len);
            batch.push_back('\n');
            rec_off.push_back(batch.size());
            return (int32_t)rec_off.size() - 2;
        };
        do {                                  // at least one loop iteration per batch
            const char *seq;
            if (s2_reader_next(r1, &seq) < 0) { eof = true; break; }                       // :443
            Iter it = { -1, -1, false, false };
            if (s2_reader_len(r1) >= S2_K) it.r1 = add(seq, s2_reader_len(r1));             // :444
            if (is_pe) {
                const char *seq2;
                const int64_t l2 = s2_reader_next(r2, &seq2);                              // :496
                if (s2_reader_len(r2) >= S2_K) {                                           // :497
                    if (l2 < 0) it.fatal = true;                                           // :501-504
                    else { it.r2 = add(seq2, s2_reader_len(r2)); it.pe2_valid = true; }
                }
            }
            iters.push_back(it);
            if (it.fatal) { eof = true; break; }
        } while (batch.size() < d.batch_bytes);
        // ---- phase B: pass 1 of every collected read on the GPU -----------------------------------
        const auto tB = now();
        t_read += secs(tA, tB);
        n_bases += batch.size();
        const uint32_t n_rec = (uint32_t)rec_off.size() - 1;
        hits.assign(n_rec + 1, 0); inf.assign(n_rec + 1, 0);
        uint64_t n_inf = 0;
        if (n_rec) {
            uint64_t cap = std::max<uint64_t>(4096, batch.size() / 32);
            for (;;) {
                pos.resize(cap);
                if (s2_scan_detect(ctx, table, batch.data(), batch.size(), rec_off.data(), n_rec, hits.data(), inf.data(),
                                   pos.data(), cap, &n_inf, 0, nullptr)) {
                    job.err = std::string(s2_last_error()) + "\n";
                    rc = EXIT_FAILURE;
                    break;
                }
                if (n_inf <= cap) break;
                cap = n_inf;
            }
            if (rc) break;
        }
        if (d.exotic)                                   // foreign-byte windows: host string path, pass-1 orientation
            for (uint32_t r = 0; r < n_rec; ++r) {
                int xh = 0, xi = 0;
                s2_exotic_pass1(d.exotic, (const char *)batch.data() + rec_off[r], rec_off[r + 1] - rec_off[r] - 1, &xh, &xi);
                hits[r] += (uint32_t)xh; inf[r] += (uint32_t)xi;
            }
        const auto tC = now();
        t_gpu += secs(tB, tC);
        // informative windows per record: pos is ascending, so each record owns a contiguous range
        std::vector<uint64_t> first(n_rec + 1, 0);
        {
            uint64_t p = 0;
            for (uint32_t r = 0; r < n_rec; ++r) {
                first[r] = p;
                while (p < n_inf && pos[p] < rec_off[r + 1]) ++p;
            }
            first[n_rec] = p;
        }
        auto kmer_at = [&](uint64_t byte_off) -> const char * {
            uint64_t k = 0;
            s2_kmer_from_ascii((const char *)batch.data() + byte_off, &k);
            s2_kmer_to_ascii(k, kbuf);
            return kbuf;
        };
        // informative k-mers of record r in window order (device positions merged with the string path's)
        std::vector<std::pair<uint64_t, std::string>> xlist;
        auto record_kmers = [&](int32_t r, std::vector<std::string> &out) {
            out.clear();
            xlist.clear();
            if (d.exotic) s2_exotic_pass2(d.exotic, (const char *)batch.data() + rec_off[r], rec_off[r + 1] - rec_off[r] - 1, xlist);
            size_t xi = 0;
            for (uint64_t p = first[r]; p < first[r + 1]; ++p) {
                while (xi < xlist.size() && rec_off[r] + xlist[xi].first < pos[p]) out.push_back(xlist[xi++].second);
                out.push_back(kmer_at(pos[p]));
            }
            while (xi < xlist.size()) out.push_back(xlist[xi++].second);
        };
        std::vector<std::string> pe2_kmers;
        auto emit = [&](const char *kmer) {
            char head[96];
            job.out += pe1;
            const int n = snprintf(head, sizeof head, "\t%d\t%d\t%d\t%d\t", h1, i1, h2, i2);      // :567, :608
            job.out.append(head, n);
            job.out += kmer;
            job.out += '\n';
        };
        // ---- phase C: replay the loop in order with the counts filled in ---------------------------------
        for (const Iter &it : iters) {
            if (it.r1 >= 0) {                                                              // :444-449
                ++reads;
                h1 = (int)hits[it.r1]; i1 = (int)inf[it.r1];
                evaluated += (rec_off[it.r1 + 1] - rec_off[it.r1] - 1) - (S2_K - 1);
                copy_kmers.clear();
                have_copy = true;
                if (i1) record_kmers(it.r1, copy_kmers);
            }
            if (it.fatal) {
                snprintf(msg, sizeof msg, "reached end of PE2 (%s) before end of PE1 (%s), check that file names are correct\n",
                         pe2 ? pe2 : "(null)", pe1);
                job.err = msg;
                rc = EXIT_FAILURE;
                break;
            }
            if (it.r2 >= 0) {                                                              // :497-540
                h2 = (int)hits[it.r2]; i2 = (int)inf[it.r2];
                evaluated += (rec_off[it.r2 + 1] - rec_off[it.r2] - 1) - (S2_K - 1);
            }
            if (h1 + h2 >= 1 && i1 + i2 >= 1) {                                            // :547
                if (have_copy) for (const std::string &k : copy_kmers) emit(k.c_str());     // :554-591
                if (is_pe && it.pe2_valid) {                                               // :594-623
                    record_kmers(it.r2, pe2_kmers);
                    for (const std::string &k : pe2_kmers) emit(k.c_str());
                }
            }
        }
        t_emit += secs(tC, now());
    }
    if (rc == 0 && (s2_reader_damaged(r1) || (r2 && s2_reader_damaged(r2)))) {
        // corrupt DEFLATE data / CRC mismatch: the reference never returns from such a file (include/strainer2_b200.h)
        snprintf(msg, sizeof msg, "damaged gzip data in %s (gzread error)\n", s2_reader_damaged(r1) ? pe1 : pe2);
        job.err = msg;
        rc = EXIT_FAILURE;
    }
    if (rc == 0) {
        char foot[4][512];
        snprintf(foot[0], sizeof foot[0], "#%s\ttotal_kmer_evaluated\t%lld\n", pe1, (long long)evaluated);          // :633-636
        snprintf(foot[1], sizeof foot[1], "#%s\ttotal_reads_evaluated\t%lld\n", pe1, (long long)reads);
        snprintf(foot[2], sizeof foot[2], "#%s\ttotal_genome_kmers\t%lld\n", pe1, (long long)d.genome_kmers);
        snprintf(foot[3], sizeof foot[3], "#%s\ttotal_genome_informative_kmers\t%lld\n", pe1, (long long)d.genome_informative);
        for (auto &f : foot) job.out += f;
    }
    {
        std::lock_guard<std::mutex> g(d.stat_mu);
        d.t_read += t_read; d.t_gpu += t_gpu; d.t_emit += t_emit; d.n_bases += n_bases;
        d.files_host += r2 && r2 != r1 ? 2 : 1;
    }
    if (r2 && r2 != r1) s2_reader_close(r2);
    s2_reader_close(r1);
    return rc;
}

extern "C" int s2_strain_detect_main(int argc, char **argv)
{
    char *a_file = nullptr, *r_file = nullptr, *b_file = nullptr, *b_file2 = nullptr, *B_file = nullptr;
    char *seq_file_type = nullptr, *background_file = nullptr, *kmer_outfile = nullptr;
    int is_paired_end = NOT_PAIRED_END;
    int c;
    optind = 1;
    while ((c = getopt(argc, argv, "g:r:a:A:b:c:B:S:M:o:t:Hhuspn")) != EOF)     // src/strain_detect.c:84
        switch (c) {
        case 'a': a_file = optarg; break;
        case 'A': break;
        case 'b': b_file = optarg; break;
        case 'c': b_file2 = optarg; break;
        case 'B': B_file = optarg; break;
        case 'r': r_file = optarg; break;
        case 'g': background_file = optarg; break;
        case 'o': kmer_outfile = optarg; break;
        case 'n': is_paired_end = NOT_PAIRED_END; break;
        case 't': seq_file_type = optarg; break;
        case 'u': detect_usage(); break;
        case 'h': detect_usage(); break;
        default: detect_usage(); break;
        }
    if (!a_file || !kmer_outfile || !r_file) { detect_usage(); return 1; }                 // :104-111
    if (!b_file && !B_file) { detect_usage(); return 1; }
    if (seq_file_type) {
        is_paired_end = get_file_type(seq_file_type);
        if (is_paired_end == UNKNOWN_FILE_TYPE) {
            printf("unknown filetype specification. allowed are SE, PE, PEI\n\n");
            detect_usage();
            return 1;
        }
    }
    if (b_file && is_paired_end == IS_PAIRED_END && !b_file2) {
        printf("commandline PE mapping requires two files (-b [file1] and -c [file2])\n\n");
        detect_usage();
        return 1;
    }
    if (b_file && B_file) {
        printf("cannot have -B flag and -b flag\nEither have a file with metagenomics files to be detect the strain in or specify one "
               "metagenomic file to detect the strain in\n");
        detect_usage();
        return 1;
    }
    Detect d;
    d.batch_bytes = s2_env_u64("S2_DETECT_BATCH_MB", 32) << 20;
    d.gpu_ingest = s2_env_int("S2_GPU_INGEST", 1) != 0;
    const int n_threads = s2_default_reader_threads();
    d.ctx = background_file ? s2_init(s2_env_int("S2_DEVICE", 0), s2_env_u64("S2_BATCH_MB", 16) << 20, n_threads + 2)
                            : s2_init(s2_env_int("S2_DEVICE", 0), 8u << 20, 2);
    if (!d.ctx) { fprintf(stderr, "%s\n", s2_last_error()); return EXIT_FAILURE; }
    // the context's ingest pipelines (350 MB of device memory and 48 MB of pinned staging each: tens of milliseconds)
    // come into being beside the table build and the labelling instead of in front of the first metagenome
    struct Warmer { std::thread t; ~Warmer() { if (t.joinable()) t.join(); } } warmer;
    // (a detect call keeps its pipeline for a whole file - per-read results come back at its end - so three files in flight)
    setenv("S2_INGEST_PIPES", "3", 0);
    const bool gz_inputs = d.gpu_ingest && (B_file ? s2_list_starts_with_plain_gz(B_file) : false);
    if (d.gpu_ingest) warmer.t = std::thread([&d, n_threads, gz_inputs]() {
        const int pipes = std::min(n_threads, std::max(1, s2_env_int("S2_INGEST_PIPES", 3)));
        if (gz_inputs) s2_ingest_warm_gz(d.ctx, pipes, 96ull << 20); else s2_ingest_warm(d.ctx, pipes);
    });

    // GEN_hash_sequences_set_count_vec(r_file, 31, h, NON_INFORMATIVE, 0, 0, 6)             :139
    std::vector<uint8_t> flat;
    if (s2_load_flat(r_file, flat) != 0) {
        fprintf(stderr, "could not read file %s GEN_hash_sequences_set_count_vec()\n", r_file);
        return EXIT_FAILURE;
    }
    d.table = s2_table_build(d.ctx, flat.data(), flat.size(), 6, 0.0, 0);
    if (!d.table) { fprintf(stderr, "%s\n", s2_last_error()); return EXIT_FAILURE; }
    d.exotic = s2_exotic_build(flat.data(), flat.size(), 6);
    d.genome_kmers = (unsigned)(s2_table_n_keys(d.table) + s2_exotic_n_keys(d.exotic));
    unsigned n_found = 0;
    const int lrc = label_informative(d, a_file, &n_found);                                // :140
    if (lrc == -1) return EXIT_FAILURE;
    if (lrc) { fprintf(stderr, "%s\n", s2_last_error()); return EXIT_FAILURE; }
    if (background_file) {                                                                 // :142-143
        const int brc = background_filter(d, background_file, n_found, n_threads);
        if (brc) return brc;
    }

    // quantify_hits_all_files                                                             :263-384
    d.gzout = s2_gz_writer_open(kmer_outfile, s2_env_int("S2_GZ_THREADS", 0));                 // gzopen(outfile, "wb9") :299
    if (!d.gzout) {
        fprintf(stderr, "could not open *gzout file outfile %s in quantify_hits_all_files()\n", kmer_outfile);
        return EXIT_FAILURE;
    }
    // quantify_hits_all_files: one job per usable batch line, in order.  The stdout chatter about unusable
    // lines depends only on the line text, so it is printed now, in line order.
    std::vector<Job> jobs;
    if (B_file) {
        FILE *fp = fopen(B_file, "r");
        if (!fp) {
            fprintf(stderr, "could not read file file_of_filenames %s in quantify_hits_all_files()\n", B_file);
            return EXIT_FAILURE;
        }
        char *line = nullptr; size_t cap = 0;
        while (getline(&line, &cap, fp) != -1) {
            char *pos = strchr(line, '\n');
            if (pos) *pos = '\0';
            char *token = strtok(line, "\t");
            const int pe = token ? get_file_type(token) : UNKNOWN_FILE_TYPE;
            if (pe == UNKNOWN_FILE_TYPE) { printf("unknown file type skipping line (%s)\n", token ? token : "(null)"); continue; }
            char *file1 = strtok(nullptr, "\t");
            if (!file1) { printf("ERROR: no first file specified for %s\n", line); continue; }
            Job j; j.pe = pe; j.f1 = file1;
            if (pe == IS_PAIRED_END) {
                char *file2 = strtok(nullptr, "\t");
                if (!file2) { printf("ERROR: no second file specified for PE: %s\n", line); continue; }
                j.f2 = file2; j.has_f2 = true;
            }
            jobs.push_back(std::move(j));
        }
        free(line);
        fclose(fp);
    } else {
        Job j; j.pe = is_paired_end; j.f1 = b_file;
        if (b_file2) { j.f2 = b_file2; j.has_f2 = true; }
        jobs.push_back(std::move(j));
    }
    fflush(stdout);

    // S2_GPUS=n: the path shards by batch line - every GPU holds a replica of the labelled table, worker
    // thread w feeds GPU w % n, the finished blocks are still written in batch order; no collective
    const int n_workers = (int)std::max<size_t>(1, std::min<size_t>(jobs.size(), (size_t)s2_default_reader_threads()));
    std::vector<s2_ctx *> ctxs(1, d.ctx);
    std::vector<s2_table *> tables(1, d.table);
    {
        const int dev0 = s2_env_int("S2_DEVICE", 0);
        int n_gpus = std::min(std::min(s2_env_int("S2_GPUS", 1), s2_device_count() - dev0), n_workers);
        std::vector<uint64_t> labelled(d.informative.begin(), d.informative.end());
        for (int g = 1; g < n_gpus; ++g) {
            s2_ctx *gc = s2_init(dev0 + g, 8u << 20, 2);
            s2_table *t = gc ? s2_table_build(gc, flat.data(), flat.size(), 6, 0.0, 0) : nullptr;
            std::vector<uint8_t> found(labelled.size() + 1);
            if (!t || s2_table_flag(t, labelled.data(), labelled.size(), found.data())) {
                fprintf(stderr, "%s\n", s2_last_error());
                return EXIT_FAILURE;
            }
            ctxs.push_back(gc); tables.push_back(t);
        }
    }
    std::vector<uint8_t>().swap(flat);

    // worker threads read + scan files concurrently; this thread writes the finished blocks in order
    int rc = 0;
    const auto t_phase = std::chrono::steady_clock::now();
    double phase_s = 0;
    {
        std::mutex mu; std::condition_variable cv;
        std::atomic<size_t> next(0);
        std::atomic<bool> stop(false);
        auto worker = [&](int w) {
            s2_ctx *ctx = ctxs[(size_t)w % ctxs.size()];
            s2_table *table = tables[(size_t)w % tables.size()];
            for (;;) {
                const size_t i = next.fetch_add(1);
                if (i >= jobs.size() || stop.load()) break;
                jobs[i].rc = quantify_hits(d, jobs[i], ctx, table);
                { std::lock_guard<std::mutex> g(mu); jobs[i].done = true; }
                cv.notify_all();
            }
            s2_ingest_thread_cleanup();
        };
        std::vector<std::thread> pool;
        for (int w = 0; w < n_workers; ++w) pool.emplace_back(worker, w);
        for (size_t i = 0; i < jobs.size(); ++i) {
            { std::unique_lock<std::mutex> g(mu); cv.wait(g, [&]() { return jobs[i].done; }); }
            d.write_out(jobs[i].out);
            std::string().swap(jobs[i].out);
            if (jobs[i].rc) {                       // the reference exit()s here: later lines are never processed
                if (s2_ingest_engine_failed()) fprintf(stderr, "%s\n", s2_last_error());      // the cause; other lines fail with bare CUDA errors
                else fputs(jobs[i].err.c_str(), stderr);
                rc = jobs[i].rc;
                stop.store(true);
                break;
            }
        }
        for (auto &t : pool) t.join();
        phase_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_phase).count();
    }
    // on a fatal error the reference exit()s with the gz stream unfinished; we close it either way
    s2_gz_writer_close(d.gzout);
    fflush(stdout);
    if (s2_env_int("S2_STATS", 0)) {
        double kms = 0; uint64_t kl = 0;
        for (s2_ctx *gc : ctxs) { double k1 = 0; uint64_t l1 = 0; s2_kernel_time(gc, &k1, &l1, 0); kms += k1; kl += l1; }
        fprintf(stderr, "[s2 detect] gpus=%zu keys=%u informative=%u bytes=%llu read=%.3fs gpu_call=%.3fs emit=%.3fs kernel_ms=%.3f launches=%llu files_gpu_ingest=%llu files_host_reader=%llu "
                        "detect_phase=%.3fs detect_Gbases_per_s=%.3f\n",
                ctxs.size(), d.genome_kmers, d.genome_informative, (unsigned long long)d.n_bases, d.t_read, d.t_gpu, d.t_emit, kms, (unsigned long long)kl,
                (unsigned long long)d.files_gpu, (unsigned long long)d.files_host, phase_s, phase_s > 0 ? d.n_bases / phase_s / 1e9 : 0.0);
        // (read / gpu_call / emit are SUMS over the worker threads; detect_phase is the wall time of the whole batch list)
    }
    s2_exotic_free(d.exotic);
    for (size_t g = 0; g < ctxs.size(); ++g) { s2_table_free(tables[g]); s2_shutdown(ctxs[g]); }
    return rc;
}
