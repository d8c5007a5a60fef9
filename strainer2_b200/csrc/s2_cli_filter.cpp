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

#include <zlib.h>

#include <algorithm>
#include <charconv>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
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
    std::vector<uint64_t> slots_key;       // 62-bit code + 1 (0 = empty)
    std::vector<uint32_t> slots_id;
    uint64_t mask = 0, used = 0;
    std::unordered_map<std::string, uint32_t> odd;
    uint32_t next_id = 0;

    void grow()
    {
        const uint64_t cap = slots_key.empty() ? (1u << 16) : slots_key.size() * 2;
        std::vector<uint64_t> k(cap, 0); std::vector<uint32_t> v(cap, 0);
        for (size_t i = 0; i < slots_key.size(); ++i)
            if (slots_key[i]) {
                uint64_t h = (slots_key[i] * 0x9E3779B97F4A7C15ull) >> 17 & (cap - 1);
                while (k[h]) h = (h + 1) & (cap - 1);
                k[h] = slots_key[i]; v[h] = slots_id[i];
            }
        slots_key.swap(k); slots_id.swap(v); mask = cap - 1;
    }
    static bool encode(const char *s, size_t len, uint64_t *code)
    {
        if (len != 31) return false;
        uint64_t c = 0;
        for (size_t i = 0; i < 31; ++i) {
            unsigned x;
            switch (s[i]) { case 'A': x = 0; break; case 'C': x = 1; break; case 'G': x = 2; break; case 'T': x = 3; break; default: return false; }
            c = (c << 2) | x;
        }
        *code = c + 1;
        return true;
    }
    // id of the key, new ids are handed out in order of first appearance; *is_new tells
    uint32_t get(const char *s, size_t len, bool *is_new)
    {
        uint64_t code;
        if (encode(s, len, &code)) {
            if ((used + 1) * 2 > slots_key.size()) grow();
            uint64_t h = (code * 0x9E3779B97F4A7C15ull) >> 17 & mask;
            while (slots_key[h] && slots_key[h] != code) h = (h + 1) & mask;
            if (slots_key[h]) { *is_new = false; return slots_id[h]; }
            slots_key[h] = code; slots_id[h] = next_id; ++used;
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

// whole (gzip) file -> text; universal newlines like Python's text mode
bool read_gz(const std::string &path, std::string &text)
{
    gzFile f = gzopen(path.c_str(), "rb");
    if (!f) return false;
    gzbuffer(f, 1 << 20);
    text.clear();
    std::vector<char> buf(4 << 20);
    int got;
    while ((got = gzread(f, buf.data(), (unsigned)buf.size())) > 0) text.append(buf.data(), (size_t)got);
    const bool ok = got == 0;
    gzclose(f);
    return ok;
}

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

    KeyIndex index;
    std::vector<uint64_t> pan_sum, meta_sum;            // over ids; membership flags beside them
    std::vector<uint8_t> in_pan, in_meta, in_drug;
    std::vector<std::string> names;                     // id -> key
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
        if (!read_gz(files[fi], text)) return fail_traceback("FileNotFoundError: [Errno 2] No such file or directory: '" + files[fi] + "'");
        const char *p = text.data(), *end = p + text.size();
        while (p < end) {
            const char *nl = p;
            while (nl < end && *nl != '\n' && *nl != '\r') ++nl;
            const char *line_end = nl;
            const char *next = nl < end ? nl + ((*nl == '\r' && nl + 1 < end && nl[1] == '\n') ? 2 : 1) : end;
            if (line_end > p && *p == '#') { p = next; continue; }
            // split on tabs
            const char *f[6]; int nf = 0;
            const char *q = p;
            f[nf++] = q;
            for (; q < line_end; ++q) if (*q == '\t') { if (nf < 6) f[nf] = q + 1; ++nf; }
            const int n_fields = nf;
            if (n_fields < 4) return fail_traceback("IndexError: list index out of range");
            auto field_end = [&](int k) { return k + 1 < n_fields && k + 1 < 6 ? f[k + 1] - 1 : line_end; };
            long long c1, c2, c3, c4 = 0;
            if (!parse_int(f[1], field_end(1), &c1) || !parse_int(f[2], field_end(2), &c2) || !parse_int(f[3], field_end(3), &c3))
                return fail_traceback("ValueError: invalid literal for int() with base 10");
            bool is_new;
            const uint32_t id = index.get(f[0], (size_t)(field_end(0) - f[0]), &is_new);
            if (is_new) {
                names.emplace_back(f[0], (size_t)(field_end(0) - f[0]));
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
                if (!parse_int(f[4], field_end(4), &c4)) return fail_traceback("ValueError: invalid literal for int() with base 10");
                if (c4 > 0) in_drug[id] = 1;                                     // :194 adds content[3]: only membership is used
            }
            p = next;
        }
        if (fi > 1) {                                                             // dict equality with the previous file's strain dict
            bool same = prev_ids.size() == strain_ids.size();
            if (same) {
                std::vector<uint8_t> in_prev(names.size(), 0);
                for (uint32_t id : prev_ids) in_prev[id] = 1;
                for (uint32_t id : strain_ids) if (!in_prev[id] || (id < prev_ref_of.size() ? prev_ref_of[id] : 0) != ref_of[id]) { same = false; break; }
            }
            if (!same) { fputs("error: input files do not have identical hash and strain hash values.\n", stderr); return 1; }
        }
    }
    text.clear(); text.shrink_to_fit();

    const uint64_t n_ids = names.size(), n_strain = strain_ids.size();
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
    auto need_gpu = [&]() -> bool {
        if (!ctx) ctx = s2_init(s2_env_int("S2_DEVICE", 0), 1 << 20, 1);
        if (!ctx) fprintf(stderr, "kmer_scrub_filter: %s\n", s2_last_error());
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
    if (ctx) s2_shutdown(ctx);
    if (rc) { if (rc == 1 && s2_last_error()[0]) fprintf(stderr, "kmer_scrub_filter: %s\n", s2_last_error()); return rc; }

    uint64_t n_keep = 0;
    for (uint64_t j = 0; j < n_strain; ++j) n_keep += keep[j] ? 1 : 0;
    snprintf(hb, sizeof hb, "#post scrub kmers %llu out of %llu\n", (unsigned long long)n_keep, (unsigned long long)all_kmers);
    out_head += hb;
    fputs(out_head.c_str(), stdout);
    std::string body;
    body.reserve((size_t)n_keep * 33);
    for (uint64_t j = 0; j < n_strain; ++j)
        if (keep[j]) { body += names[strain_ids[j]]; body += '\n'; }
    fwrite(body.data(), 1, body.size(), stdout);
    fflush(stdout);
    return 0;
}
