// s2_cli_filter.cpp - drop-in for scripts/kmer_scrub_filter.py: same options, same stdout / stderr
// bytes, same exit codes; the selection itself runs on the GPU (s2_filter.cu).
//
// Reference: /root/reference/scripts/kmer_scrub_filter.py (Python 3, dicts keyed by the k-mer string).  What the dicts
// do is restated with dense ids: every distinct key of any input file gets an id (ACGT 31-mers through a flat
// open-addressing table on their 62-bit code, anything else - the IUPAC rows of SURVEY D6, odd lengths - through a
// string map), pangenome / metagenome / drug membership and sums live in arrays over ids, and the strain dict of a
// file is the list of its distinct ids in first-occurrence order with the last reference count seen (:153-198).
// Floating point appears in two places and is reproduced bit for bit: the stopping rule of the joint scrub
// (`1-((num_scrubbed+1)/all_kmers) > min_fraction`, :128 - evaluated here in IEEE doubles exactly as written, it is
// monotone so the number of rows to remove is found by bisection) and the value every row is ranked by (two IEEE
// divisions and a max, done by the kernel).  Python's str(float) is reproduced by py_repr().
#include "../../include/strainer2_b200.h"
#include "s2_internal.h"

#include <sys/stat.h>
#include <zlib.h>

#include <algorithm>
#include <charconv>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

// ---- str(float) of Python 3: shortest digits that round-trip, fixed notation for 1e-4 <= |x| < 1e16 -------------
std::string py_repr(double x)
{
    if (x != x) return "nan";
    if (x == 1.0 / 0.0) return "inf";
    if (x == -1.0 / 0.0) return "-inf";
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, x, std::chars_format::scientific);      // d[.ddd]e[+-]XX, shortest
    std::string s(buf, r.ptr);
    std::string sign;
    if (!s.empty() && s[0] == '-') { sign = "-"; s.erase(0, 1); }
    const size_t e = s.find('e');
    std::string digits = s.substr(0, e);
    const int exp10 = atoi(s.c_str() + e + 1);
    digits.erase(std::remove(digits.begin(), digits.end(), '.'), digits.end());
    std::string out;
    if (exp10 < -4 || exp10 >= 16) {                                   // repr switches to exponent notation here
        out = digits.substr(0, 1);
        if (digits.size() > 1) out += "." + digits.substr(1);
        char eb[16];
        snprintf(eb, sizeof eb, "e%c%02d", exp10 < 0 ? '-' : '+', abs(exp10));
        out += eb;
    } else if (exp10 < 0) {
        out = "0." + std::string((size_t)(-exp10 - 1), '0') + digits;
    } else {
        if ((int)digits.size() <= exp10 + 1) out = digits + std::string((size_t)(exp10 + 1 - (int)digits.size()), '0') + ".0";
        else out = digits.substr(0, (size_t)exp10 + 1) + "." + digits.substr((size_t)exp10 + 1);
    }
    return sign + out;
}

// ---- keys -> dense ids ---------------------------------------------------------------------------------------------
struct KeyIndex {
    struct Slot { uint64_t key; uint32_t id, pad; };       // 62-bit code + 1 (0 = empty); one cache line touch per probe
    std::vector<Slot> slots;
    uint64_t mask = 0, used = 0;
    std::unordered_map<std::string, uint32_t> odd;
    uint32_t next_id = 0;

    static uint64_t slot_of(uint64_t code, uint64_t mask) { return ((code * 0x9E3779B97F4A7C15ull) >> 20) & mask; }
    void grow(uint64_t want = 0)
    {
        uint64_t cap = slots.empty() ? (1u << 16) : slots.size() * 2;
        while (cap < want) cap *= 2;
        std::vector<Slot> k(cap, Slot{ 0, 0, 0 });
        for (const Slot &x : slots)
            if (x.key) {
                uint64_t h = slot_of(x.key, cap - 1);
                while (k[h].key) h = (h + 1) & (cap - 1);
                k[h] = x;
            }
        slots.swap(k); mask = cap - 1;
    }
    void reserve(uint64_t n_keys)                   // room for n_keys plain keys without another rehash
    {
        if (slots.size() < (n_keys + 64) * 2) grow((n_keys + 64) * 2);
    }
    // 62-bit code + 1 of an ACGT 31-mer, 0 for anything else (branch-free: the letters are as good as random)
    static uint64_t encode(const char *s, size_t len)
    {
        static const struct Lut { uint8_t v[256]; Lut() { memset(v, 4, sizeof v); v['A'] = 0; v['C'] = 1; v['G'] = 2; v['T'] = 3; } } lut;
        if (len != 31) return 0;
        uint64_t c = 0; unsigned bad = 0;
        for (size_t i = 0; i < 31; ++i) { const unsigned x = lut.v[(uint8_t)s[i]]; bad |= x; c = (c << 2) | (x & 3u); }
        return (bad & 4u) ? 0 : c + 1;
    }
    void prefetch(uint64_t code) const { if (code) __builtin_prefetch(&slots[slot_of(code, mask)]); }
    // id of the key (code = encode(s, len)), new ids are handed out in order of first appearance; *is_new tells
    uint32_t get(uint64_t code, const char *s, size_t len, bool *is_new)
    {
        if (code) {
            if ((used + 1) * 2 > slots.size()) grow();
            uint64_t h = slot_of(code, mask);
            while (slots[h].key && slots[h].key != code) h = (h + 1) & mask;
            if (slots[h].key) { *is_new = false; return slots[h].id; }
            slots[h].key = code; slots[h].id = next_id; ++used;
            *is_new = true;
            return next_id++;
        }
        auto it = odd.find(std::string(s, len));
        if (it != odd.end()) { *is_new = false; return it->second; }
        odd.emplace(std::string(s, len), next_id);
        *is_new = true;
        return next_id++;
    }
};

struct Options {
    std::string file, list;
    bool has_file = false, has_list = false, independent = false;
    double min_fraction = 0.04;
};

const char *USAGE = "usage: kmer_scrub_filter.py [-h] [--scrub_count_file SCRUB_COUNT_FILE]\n"
                    "                            [--scrub_count_list SCRUB_COUNT_LIST]\n"
                    "                            [--min_fraction MIN_FRACTION] [--independent]\n";

// the script's argparse parser (:14-27): -s/-l/-m/-i, long names (unique prefixes accepted), --name=value, -mVALUE
int parse_args(int argc, char **argv, Options &o)
{
    static const char *longs[] = { "--scrub_count_file", "--scrub_count_list", "--min_fraction", "--independent", "--help" };
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i], val;
        bool has_val = false;
        int which = -1;
        if (a.size() >= 2 && a[0] == '-' && a[1] != '-') {
            const char *shorts = "slmih";
            const char *p = strchr(shorts, a[1]);
            if (!p) { fprintf(stderr, "%skmer_scrub_filter.py: error: unrecognized arguments: %s\n", USAGE, a.c_str()); return 2; }
            which = (int)(p - shorts);
            if (a.size() > 2) { val = a.substr(a[2] == '=' ? 3 : 2); has_val = true; }
        } else if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
            const size_t eq = a.find('=');
            const std::string name = a.substr(0, eq);
            if (eq != std::string::npos) { val = a.substr(eq + 1); has_val = true; }
            int n_match = 0;
            for (int k = 0; k < 5; ++k)
                if (name == longs[k]) { which = k; n_match = 1; break; }
                else if (strncmp(longs[k], name.c_str(), name.size()) == 0) { which = k; ++n_match; }
            if (n_match != 1) { fprintf(stderr, "%skmer_scrub_filter.py: error: unrecognized arguments: %s\n", USAGE, a.c_str()); return 2; }
        } else {
            fprintf(stderr, "%skmer_scrub_filter.py: error: unrecognized arguments: %s\n", USAGE, a.c_str());
            return 2;
        }
        if (which == 4) { fputs(USAGE, stdout); return -1; }                       // -h: usage, exit 0
        if (which == 3) { o.independent = true; continue; }
        if (!has_val) {
            if (i + 1 >= argc) { fprintf(stderr, "%skmer_scrub_filter.py: error: argument %s: expected one argument\n", USAGE, a.c_str()); return 2; }
            val = argv[++i];
        }
        if (which == 0) { o.file = val; o.has_file = !val.empty(); }
        else if (which == 1) { o.list = val; o.has_list = !val.empty(); }
        else {
            char *end = nullptr;
            o.min_fraction = strtod(val.c_str(), &end);
            if (val.empty() || *end) { fprintf(stderr, "%skmer_scrub_filter.py: error: argument --min_fraction/-m: invalid float value: '%s'\n", USAGE, val.c_str()); return 2; }
        }
    }
    return 0;
}

// the (gzip) file arrives in pieces from a thread of its own: zlib inflates the next 8 MB while the previous ones are parsed
struct GzPieces {
    gzFile f = nullptr;
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<std::string> q;
    bool done = false, ok = true, quit = false;
    bool open(const std::string &path)
    {
        f = gzopen(path.c_str(), "rb");
        if (!f) return false;
        gzbuffer(f, 1 << 20);
        th = std::thread([this]() {
            for (;;) {
                std::string buf;
                buf.resize(8u << 20);
                const int got = gzread(f, &buf[0], (unsigned)buf.size());
                std::unique_lock<std::mutex> g(mu);
                if (got <= 0) { ok = got == 0; done = true; cv.notify_all(); return; }
                buf.resize((size_t)got);
                cv.wait(g, [this]() { return q.size() < 3 || quit; });
                if (quit) { done = true; cv.notify_all(); return; }
                q.push_back(std::move(buf));
                cv.notify_all();
            }
        });
        return true;
    }
    bool next(std::string &out)                        // false at the end of the file
    {
        std::unique_lock<std::mutex> g(mu);
        cv.wait(g, [this]() { return !q.empty() || done; });
        if (q.empty()) return false;
        out = std::move(q.front());
        q.pop_front();
        cv.notify_all();
        return true;
    }
    ~GzPieces()
    {
        if (th.joinable()) { { std::lock_guard<std::mutex> g(mu); quit = true; } cv.notify_all(); th.join(); }
        if (f) gzclose(f);
    }
};

bool parse_int(const char *s, const char *e, long long *out)
{
    while (s < e && (*s == ' ')) ++s;
    while (e > s && (e[-1] == ' ' || e[-1] == '\r')) --e;
    if (s == e) return false;
    bool neg = false;
    if (*s == '-' || *s == '+') { neg = *s == '-'; ++s; }
    if (s == e) return false;
    long long v = 0;
    for (; s < e; ++s) { if (*s < '0' || *s > '9') return false; v = v * 10 + (*s - '0'); }
    *out = neg ? -v : v;
    return true;
}

struct Traceback { std::string last_line; };

}  // namespace

// str(float) as Python 3 prints it (the script's stderr / stdout carry such numbers); out must hold 32 bytes
extern "C" void s2_py_float_repr(double x, char *out)
{
    const std::string s = py_repr(x);
    memcpy(out, s.c_str(), s.size() + 1);
}

// scripts/kmer_scrub_filter.py:146-229
extern "C" int s2_kmer_scrub_filter_main(int argc, char **argv)
{
    Options opt;
    const int prc = parse_args(argc, argv, opt);
    if (prc) return prc < 0 ? 0 : prc;
    std::string out_head;          // stdout so far (flushed before anything that ends the run)
    auto fail_traceback = [&](const std::string &last) {
        fputs(out_head.c_str(), stdout); fflush(stdout);
        fprintf(stderr, "Traceback (most recent call last):\n  (kmer_scrub_filter, B200 build)\n%s\n", last.c_str());
        return 1;
    };
    if (opt.min_fraction < 0.0 || opt.min_fraction > 1.0)                           // :147-148 raises while building its message
        return fail_traceback("TypeError: can only concatenate str (not \"float\") to str");
    if (!opt.has_file && !opt.has_list) fputs("error: one of scrub_count_file or scrub_count_list must be provided.", stderr);
    if (opt.has_file && opt.has_list) fputs("error: can provide only one of either scrub_count_file or scrub_count_list.", stderr);
    std::vector<std::string> files;
    if (opt.has_file) files.push_back(opt.file);
    else if (opt.has_list) {
        FILE *lf = fopen(opt.list.c_str(), "r");
        if (!lf) return fail_traceback("FileNotFoundError: [Errno 2] No such file or directory: '" + opt.list + "'");
        char *line = nullptr; size_t cap = 0; ssize_t len;
        while ((len = getline(&line, &cap, lf)) != -1) {
            while (len > 0 && (line[len - 1] == '\n' || line[len - 1] == '\r' || line[len - 1] == ' ' || line[len - 1] == '\t')) --len;   // rstrip()
            files.emplace_back(line, (size_t)len);
        }
        free(line);
        fclose(lf);
    }

    const bool stats = s2_env_int("S2_STATS", 0) != 0;
    auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_start = now();
    // the CUDA context comes up (hundreds of milliseconds) while the table is inflated and parsed
    struct EarlyCtx {
        s2_ctx *ctx = nullptr; std::string error; std::thread th; bool started = false;
        void start() { started = true; th = std::thread([this]() { ctx = s2_init(s2_env_int("S2_DEVICE", 0), 1 << 20, 1); if (!ctx) error = s2_last_error(); }); }
        s2_ctx *get() { if (th.joinable()) th.join(); return ctx; }
        ~EarlyCtx() { if (th.joinable()) th.join(); if (ctx) s2_shutdown(ctx); }
    } early;
    if (!files.empty()) early.start();
    double t_inflate = 0, t_parse = 0, t_gpu = 0;
    KeyIndex index;
    std::vector<uint64_t> pan_sum, meta_sum;            // over ids; membership flags beside them
    std::vector<uint8_t> in_pan, in_meta, in_drug;
    std::string name_pool;                              // id -> key: bytes [name_off[id], name_off[id + 1])
    std::vector<uint64_t> name_off(1, 0);
    bool drug_filter = false;
    // the strain dict of the current and of the previous file: distinct ids in first-occurrence order + reference count by id
    std::vector<uint32_t> strain_ids, prev_ids;
    std::vector<long long> ref_of, prev_ref_of;
    std::vector<uint32_t> stamp;                        // id -> 1 + index of the file whose strain dict holds it
    uint64_t all_kmers = 0;
    std::string text;
    for (size_t fi = 0; fi < files.size(); ++fi) {
        if (fi > 1) { prev_ids = strain_ids; prev_ref_of = ref_of; }              // :163 (sic): only from the third file on
        strain_ids.clear();
        all_kmers = 0;
        GzPieces pieces;
        if (!pieces.open(files[fi])) return fail_traceback("FileNotFoundError: [Errno 2] No such file or directory: '" + files[fi] + "'");
        {                                                                         // sizes are known roughly: few reallocations / rehashes while parsing
            struct stat sb;
            const uint64_t est = (stat(files[fi].c_str(), &sb) == 0 ? (uint64_t)sb.st_size * 3 : (64u << 20)) / 36 + 1024;     // count tables compress about 3 : 1
            index.reserve(index.used + est);
            for (auto *v : { &pan_sum, &meta_sum }) v->reserve(v->size() + est);
            for (auto *v : { &in_pan, &in_meta, &in_drug }) v->reserve(v->size() + est);
            ref_of.reserve(ref_of.size() + est); stamp.reserve(stamp.size() + est); name_off.reserve(name_off.size() + est);
            name_pool.reserve(name_pool.size() + est * 31);
            strain_ids.reserve(est);
        }
        // complete lines [p, end) -> the dicts; returns the last line of a Python traceback, or nothing
        auto parse_lines = [&](const char *p, const char *end) -> std::string {
        const bool has_cr = memchr(p, '\r', (size_t)(end - p)) != nullptr;             // universal newlines only cost something when there is a CR
        struct Line { const char *b, *e; uint64_t code; };
        Line block[64];
        while (p < end) {
            // the next 64 data lines; their keys' hash slots are prefetched before the lines are parsed
            int n_lines = 0;
            while (p < end && n_lines < 64) {
                const char *nl;
                if (!has_cr) { nl = (const char *)memchr(p, '\n', (size_t)(end - p)); if (!nl) nl = end; }
                else { nl = p; while (nl < end && *nl != '\n' && *nl != '\r') ++nl; }
                const char *next = nl < end ? nl + ((*nl == '\r' && nl + 1 < end && nl[1] == '\n') ? 2 : 1) : end;
                if (!(nl > p && *p == '#')) {
                    const char *tab = (const char *)memchr(p, '\t', (size_t)(nl - p));
                    const uint64_t code = tab ? KeyIndex::encode(p, (size_t)(tab - p)) : 0;
                    index.prefetch(code);
                    block[n_lines++] = { p, nl, code };
                }
                p = next;
            }
            for (int li = 0; li < n_lines; ++li) {
                const char *lb = block[li].b, *line_end = block[li].e;
                // split on tabs
                const char *f[6]; int nf = 0;
                const char *q = lb;
                f[nf++] = q;
                for (; q < line_end; ++q) if (*q == '\t') { if (nf < 6) f[nf] = q + 1; ++nf; }
                const int n_fields = nf;
                if (n_fields < 4) return std::string("IndexError: list index out of range");
                auto field_end = [&](int k) { return k + 1 < n_fields && k + 1 < 6 ? f[k + 1] - 1 : line_end; };
                long long c1, c2, c3, c4 = 0;
                if (!parse_int(f[1], field_end(1), &c1) || !parse_int(f[2], field_end(2), &c2) || !parse_int(f[3], field_end(3), &c3))
                    return std::string("ValueError: invalid literal for int() with base 10");
                bool is_new;
                const size_t klen = (size_t)(field_end(0) - f[0]);
                const uint32_t id = index.get(block[li].code, f[0], klen, &is_new);
                if (is_new) {
                    name_pool.append(f[0], klen); name_off.push_back(name_pool.size());
                    pan_sum.push_back(0); meta_sum.push_back(0); in_pan.push_back(0); in_meta.push_back(0); in_drug.push_back(0);
                    ref_of.push_back(0); stamp.push_back(0);
                }
                ++all_kmers;
                if (stamp[id] != fi + 1) { stamp[id] = (uint32_t)fi + 1; strain_ids.push_back(id); }
                ref_of[id] = c1;
                if (c2 > 0) { pan_sum[id] += (uint64_t)c2; in_pan[id] = 1; }
                if (c3 > 0) { meta_sum[id] += (uint64_t)c3; in_meta[id] = 1; }
                if (n_fields == 5) {
                    drug_filter = true;
                    if (!parse_int(f[4], field_end(4), &c4)) return std::string("ValueError: invalid literal for int() with base 10");
                    if (c4 > 0) in_drug[id] = 1;                                 // :194 adds content[3]: only membership is used
                }
            }
        }
            return std::string();
        };
        // a piece is parsed up to its last line break; the rest waits for the next piece (a CR at the very end may be half of a CR LF)
        std::string piece;
        text.clear();
        for (;;) {
            const double t_a = now();
            const bool more = pieces.next(piece);
            const double t_b = now();
            t_inflate += t_b - t_a;                                               // time spent WAITING for the inflating thread
            if (more) text += piece;
            size_t cut = text.size();
            if (more) {
                while (cut > 0 && text[cut - 1] != '\n' && !(text[cut - 1] == '\r' && cut < text.size())) --cut;
            }
            const std::string err = parse_lines(text.data(), text.data() + cut);
            if (!err.empty()) return fail_traceback(err);
            text.erase(0, cut);
            t_parse += now() - t_b;
            if (!more) break;
        }
        if (!pieces.ok) return fail_traceback("EOFError: Compressed file ended before the end-of-stream marker was reached");
        const double t_b = now();
        t_parse += now() - t_b;
        if (fi > 1) {                                                             // dict equality with the previous file's strain dict
            bool same = prev_ids.size() == strain_ids.size();
            if (same) {
                std::vector<uint8_t> in_prev(name_off.size(), 0);
                for (uint32_t id : prev_ids) in_prev[id] = 1;
                for (uint32_t id : strain_ids) if (!in_prev[id] || (id < prev_ref_of.size() ? prev_ref_of[id] : 0) != ref_of[id]) { same = false; break; }
            }
            if (!same) { fputs("error: input files do not have identical hash and strain hash values.\n", stderr); return 1; }
        }
    }
    text.clear(); text.shrink_to_fit();

    const uint64_t n_ids = name_off.size() - 1, n_strain = strain_ids.size();
    uint64_t n_pan = 0, n_meta = 0, n_drug = 0;
    for (uint64_t i = 0; i < n_ids; ++i) { n_pan += in_pan[i]; n_meta += in_meta[i]; n_drug += in_drug[i]; }
    char hb[256];
    snprintf(hb, sizeof hb, "#total kmers in strain:%llu,%llu pangenome: %llu metagenome: %llu\n", (unsigned long long)all_kmers,
             (unsigned long long)n_strain, (unsigned long long)n_pan, (unsigned long long)n_meta);
    out_head += hb;

    // the table in strain-dict order
    std::vector<uint64_t> pan(n_strain), meta(n_strain);
    std::vector<uint8_t> alive(n_strain, 1), keep(n_strain, 0);
    for (uint64_t j = 0; j < n_strain; ++j) { pan[j] = pan_sum[strain_ids[j]]; meta[j] = meta_sum[strain_ids[j]]; }
    uint64_t n_alive = n_strain, drug_scrubbed = 0;
    if (drug_filter) {                                                            // :207-216
        snprintf(hb, sizeof hb, "#total kmers cross drug:%llu\n", (unsigned long long)n_drug);
        out_head += hb;
        for (uint64_t j = 0; j < n_strain; ++j) if (in_drug[strain_ids[j]]) { alive[j] = 0; --n_alive; }
        if (all_kmers == 0) return fail_traceback("ZeroDivisionError: float division by zero");
        const double remaining = (double)n_alive / (double)all_kmers;
        drug_scrubbed = all_kmers - n_alive;
        out_head += "#fraction kmers remaining drug post scrub:" + py_repr(remaining) + "\n";
        snprintf(hb, sizeof hb, "#drug_scrubbed kmers:%lld\n", (long long)all_kmers - (long long)n_alive);
        out_head += hb;
        if (remaining < opt.min_fraction * 2)
            return fail_traceback("Exception: ERROR: too few kmers remain after drug scrub. Are your drug strains too similar?");
    }

    s2_ctx *ctx = nullptr;
    const double t_gpu0 = now();
    auto need_gpu = [&]() -> bool {
        if (!ctx) { if (!early.started) early.start(); ctx = early.get(); }
        if (!ctx) fprintf(stderr, "kmer_scrub_filter: %s\n", early.error.c_str());
        return ctx != nullptr;
    };
    int rc = 0;
    if (opt.independent) {                                                        // :72-84 with :31-58 twice
        std::vector<uint64_t> vals;
        for (int which = 0; which < 2 && rc == 0; ++which) {
            const std::vector<uint64_t> &sum = which == 0 ? pan_sum : meta_sum;
            const std::vector<uint8_t> &in = which == 0 ? in_pan : in_meta;
            vals.clear();
            for (uint64_t i = 0; i < n_ids; ++i) if (in[i]) vals.push_back(sum[i]);       // the dict's values
            if (all_kmers == 0) { rc = fail_traceback("ZeroDivisionError: float division by zero"); break; }
            std::vector<uint64_t> hist(65537, 0);
            if (!need_gpu() || s2_scrub_histogram(ctx, vals.data(), vals.size(), hist.data())) { rc = 1; break; }
            // hits(t) = entries > t, from the top of the histogram down
            std::vector<uint64_t> above(65537, 0);                                // above[t] = entries with value > t, t < 65536
            uint64_t acc = hist[65536];
            for (int v = 65535; v >= 0; --v) { above[v] = acc; acc += hist[v]; }
            const double total = (double)all_kmers;
            long long t = -1;
            double kept = -1.0;
            uint64_t hits = 0;
            while (kept < opt.min_fraction) {
                ++t;
                if (t < 65536) hits = above[t];
                else if (s2_scrub_count_above(ctx, vals.data(), vals.size(), (uint64_t)t, &hits)) { rc = 1; break; }
                kept = 1 - ((double)hits / total);
                fprintf(stderr, "kept %s with threshold %lld\n", py_repr(kept).c_str(), t);
            }
            if (rc) break;
            fprintf(stderr, "threshold was %lld left with %llu out of %s that will be scrubbed\n", t, (unsigned long long)hits, py_repr(total).c_str());
            for (uint64_t j = 0; j < n_strain; ++j) {
                const uint32_t id = strain_ids[j];
                if (in[id] && (long long)sum[id] > t) alive[j] = 0;
            }
        }
        if (rc == 0) keep = alive;
    } else {                                                                      // :88-143
        uint64_t psum = 0, msum = 0;
        for (uint64_t i = 0; i < n_ids; ++i) { psum += pan_sum[i]; msum += meta_sum[i]; }
        // rows to remove: the longest prefix of the ranking for which the script's test holds (monotone in num_scrubbed)
        const double m = opt.min_fraction, base = (double)drug_scrubbed;
        auto test = [&](uint64_t j) { return (1 - (((base + (double)j) + 1) / (double)all_kmers)) > m; };
        uint64_t lo = 0, hi = n_alive;                                           // first j in [0, n_alive] with !test(j); n_alive if none
        while (lo < hi) { const uint64_t mid = lo + (hi - lo) / 2; if (test(mid)) lo = mid + 1; else hi = mid; }
        const uint64_t n_scrub = lo;
        if (n_strain) {
            if (!need_gpu() || s2_scrub_joint(ctx, pan.data(), meta.data(), alive.data(), n_strain, psum, msum, n_scrub, keep.data())) rc = 1;
        }
    }
    t_gpu = now() - t_gpu0;
    if (stats) fprintf(stderr, "[s2 filter] rows=%llu inflate=%.3fs parse=%.3fs wait for context + select=%.3fs total so far=%.3fs\n",
                       (unsigned long long)all_kmers, t_inflate, t_parse, t_gpu, now() - t_start);
    if (rc) { if (rc == 1 && ctx && s2_last_error()[0]) fprintf(stderr, "kmer_scrub_filter: %s\n", s2_last_error()); return rc; }

    uint64_t n_keep = 0;
    for (uint64_t j = 0; j < n_strain; ++j) n_keep += keep[j] ? 1 : 0;
    snprintf(hb, sizeof hb, "#post scrub kmers %llu out of %llu\n", (unsigned long long)n_keep, (unsigned long long)all_kmers);
    out_head += hb;
    fputs(out_head.c_str(), stdout);
    std::string body;
    body.reserve((size_t)n_keep * 33);
    for (uint64_t j = 0; j < n_strain; ++j)
        if (keep[j]) { const uint32_t id = strain_ids[j]; body.append(name_pool, name_off[id], name_off[id + 1] - name_off[id]); body += '\n'; }
    fwrite(body.data(), 1, body.size(), stdout);
    fflush(stdout);
    return 0;
}
