// s2_exotic.cpp - host string path for windows the 2-bit kernels cannot represent.
//
// The reference treats only the byte 'N' specially (src/genome_compare.c:443-451).  Any other byte
// (IUPAC codes R Y K M S W B D H V, 'U', 'X', '-', '.', '\r', ...) stays in the window, is complemented
// through COMPLEMENT[] (src/BIO_sequence.c:203-213), and the window is hashed and printed AS A STRING
// (SURVEY D6).  The device table holds only ACGT windows, so:
//   * if the -r genome contains no such byte (every shipped fixture, every BASELINE config) this file
//     does nothing: windows with a foreign byte can never equal a key, and the kernels skip them;
//   * otherwise the (few) windows of the -r genome that contain a foreign byte and no 'N' become
//     string keys here, scanned files are checked for foreign bytes, and only windows that contain
//     one are looked up here.  ACGT windows are never handled on the host.
// Bytes >= 0x80 index the reference's table out of bounds (undefined there); we give them "no
// complement" (-1) like every other unmapped byte.
#include "s2_internal.h"

#include <algorithm>
#include <cstring>
#include <mutex>
#include <unordered_map>

static inline int exo_complement(int c)                       // src/BIO_sequence.c:203-213
{
    switch (c) {
    case '-': return '-'; case '.': return '.'; case '^': return '^';
    case 'A': return 'T'; case 'B': return 'V'; case 'C': return 'G'; case 'D': return 'H';
    case 'G': return 'C'; case 'H': return 'D'; case 'K': return '.'; case 'M': return 'K';
    case 'N': return 'N'; case 'R': return 'Y'; case 'S': return 'S'; case 'T': return 'A';
    case 'U': return 'A'; case 'V': return 'B'; case 'W': return 'W'; case 'X': return 'X';
    case 'Y': return 'R';
    case 'a': return 't'; case 'b': return 'v'; case 'c': return 'g'; case 'd': return 'h';
    case 'g': return 'c'; case 'h': return 'd'; case 'k': return 'm'; case 'm': return 'k';
    case 'n': return 'n'; case 'r': return 'y'; case 's': return 's'; case 't': return 'a';
    case 'u': return 'a'; case 'v': return 'b'; case 'w': return 'w'; case 'x': return 'x';
    case 'y': return 'r';
    default: return -1;
    }
}

static inline char exo_upper(char c) { return (c >= 'a' && c <= 'z') ? (char)(c - 32) : c; }   // toupper, C locale
static inline bool exo_plain(unsigned char c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == 'N'; }

// orient_string / rc_strcmp (src/genome_compare.c:1100-1141): signed-char compare of window[i] with
// complement(window[k-1-i]); forward wins ties
static void exo_orient(const char *w, char *out)
{
    int cmp = 0;
    for (int i = 0; i < 31 && cmp == 0; ++i) {
        const char rc = (char)exo_complement((unsigned char)w[30 - i]);
        if (w[i] > rc) cmp = 1; else if (rc > w[i]) cmp = -1;
    }
    if (cmp >= 0) memcpy(out, w, 31);
    else for (int i = 0; i < 31; ++i) out[30 - i] = (char)exo_complement((unsigned char)w[i]);
    out[31] = '\0';
}

// pass-1 orientation of strain_detect (src/strain_detect.c:457-474): unsigned byte compare (strcmp) of
// the window with the matching stretch of the complemented-then-reversed read; the COPY wins ties
static void exo_orient_pass1(const char *w, char *out)
{
    char rc[32];
    for (int i = 0; i < 31; ++i) rc[30 - i] = (char)exo_complement((unsigned char)w[i]);
    memcpy(out, memcmp(w, rc, 31) > 0 ? w : rc, 31);
    out[31] = '\0';
}

static uint32_t exo_djb2(const char *s)                        // src/BIO_hash.c:208-216
{
    uint32_t h = 5381u;
    for (; *s; ++s) h = h * 33u + (uint32_t)(int32_t)(signed char)*s;
    return h;
}

struct ExoEntry { uint64_t first_pos; uint32_t counts[8]; bool informative; };

struct s2_exotic {
    std::unordered_map<std::string, ExoEntry> map;
    std::mutex mu;
    int n_cols = 4;
};

// windows [i, i+31) of the upper-cased record that contain a foreign byte and no 'N'
template <class F>
static void for_exotic_windows(const char *up, uint64_t len, F &&f)
{
    if (len < 31) return;
    // prefix counts keep this linear; records reaching here are rare (they contain a foreign byte)
    std::vector<uint32_t> nf(len + 1, 0), nn(len + 1, 0);
    for (uint64_t i = 0; i < len; ++i) {
        nf[i + 1] = nf[i] + (exo_plain((unsigned char)up[i]) ? 0u : 1u);
        nn[i + 1] = nn[i] + (up[i] == 'N' ? 1u : 0u);
    }
    if (nf[len] == 0) return;
    for (uint64_t i = 0; i + 31 <= len; ++i)
        if (nf[i + 31] != nf[i] && nn[i + 31] == nn[i]) f(i);
}

static bool has_foreign(const char *s, uint64_t len)
{
    for (uint64_t i = 0; i < len; ++i)
        if (!exo_plain((unsigned char)exo_upper(s[i]))) return true;
    return false;
}

s2_exotic *s2_exotic_build(const uint8_t *flat, uint64_t n, int n_cols)
{
    // cheap pre-check over the whole stream ('\n' is the record separator of the flat format)
    bool any = false;
    for (uint64_t i = 0; i < n && !any; ++i) any = flat[i] != '\n' && !exo_plain((unsigned char)exo_upper((char)flat[i]));
    if (!any) return nullptr;
    s2_exotic *ex = new s2_exotic();
    ex->n_cols = n_cols;
    std::string up;
    char key[32];
    uint64_t start = 0;
    while (start < n) {
        const uint8_t *nl = (const uint8_t *)memchr(flat + start, '\n', n - start);
        const uint64_t len = nl ? (uint64_t)(nl - (flat + start)) : n - start;
        up.assign((const char *)flat + start, len);
        for (auto &c : up) c = exo_upper(c);
        for_exotic_windows(up.data(), len, [&](uint64_t i) {
            exo_orient(up.data() + i, key);
            auto it = ex->map.find(key);
            if (it == ex->map.end()) {
                ExoEntry e; memset(&e, 0, sizeof e);
                e.first_pos = start + i; e.counts[0] = 1;          // default_count 1 (src/genome_compare.c:1011-1013)
                ex->map.emplace(key, e);
            } else {
                it->second.counts[0] += 1;                          // += increment (:1016)
            }
        });
        start += len + 1;
    }
    if (ex->map.empty()) { delete ex; return nullptr; }
    return ex;
}

void s2_exotic_free(s2_exotic *ex) { delete ex; }
uint64_t s2_exotic_n_keys(const s2_exotic *ex) { return ex ? ex->map.size() : 0; }

// GEN_calculate_kmer_count semantics (src/genome_compare.c:203-229) for the foreign-byte windows of one record
void s2_exotic_count_record(s2_exotic *ex, const char *seq, uint64_t len, int col)
{
    if (!ex || len < 31 || !has_foreign(seq, len)) return;
    std::string up(seq, len);
    for (auto &c : up) c = exo_upper(c);
    char key[32];
    for_exotic_windows(up.data(), len, [&](uint64_t i) {
        exo_orient(up.data() + i, key);
        std::lock_guard<std::mutex> g(ex->mu);
        auto it = ex->map.find(key);
        if (it != ex->map.end()) it->second.counts[col] += 1;
    });
}

void s2_exotic_rows(const s2_exotic *ex, std::vector<S2ExoRow> &rows)
{
    rows.clear();
    if (!ex) return;
    for (auto &kv : ex->map) {
        S2ExoRow r;
        r.key = kv.first; r.first_pos = kv.second.first_pos; r.djb2 = exo_djb2(kv.first.c_str());
        memcpy(r.counts, kv.second.counts, sizeof r.counts);
        rows.push_back(r);
    }
    std::sort(rows.begin(), rows.end(), [](const S2ExoRow &a, const S2ExoRow &b) { return a.first_pos < b.first_pos; });
}

// hash_scrubbed_kmers (src/strain_detect.c:687-717) for a 31-character line that is not plain upper-case
// ACGT: orient the RAW line (no upper-casing there) and look it up among the string keys
bool s2_exotic_flag(s2_exotic *ex, const char *line31)
{
    if (!ex) return false;
    char key[32];
    exo_orient(line31, key);
    auto it = ex->map.find(key);
    if (it == ex->map.end()) return false;
    it->second.informative = true;
    return true;
}

bool s2_exotic_is_informative(const s2_exotic *ex, const char *key)
{
    if (!ex) return false;
    auto it = ex->map.find(key);
    return it != ex->map.end() && it->second.informative;
}

void s2_exotic_set_informative(s2_exotic *ex, const char *key, bool v)
{
    if (!ex) return;
    auto it = ex->map.find(key);
    if (it != ex->map.end()) it->second.informative = v;
}

uint64_t s2_exotic_n_informative(const s2_exotic *ex)
{
    uint64_t n = 0;
    if (ex) for (auto &kv : ex->map) n += kv.second.informative ? 1 : 0;
    return n;
}

// pass 1 (src/strain_detect.c:465-491) over the foreign-byte windows of one upper-cased read
void s2_exotic_pass1(s2_exotic *ex, const char *seq, uint64_t len, int *hits, int *inf)
{
    if (!ex || len < 31 || !has_foreign(seq, len)) return;
    std::string up(seq, len);
    for (auto &c : up) c = exo_upper(c);
    char key[32];
    for_exotic_windows(up.data(), len, [&](uint64_t i) {
        exo_orient_pass1(up.data() + i, key);
        auto it = ex->map.find(key);
        if (it != ex->map.end()) { ++*hits; if (it->second.informative) ++*inf; }
    });
}

// pass 2 (src/strain_detect.c:554-591): informative foreign-byte windows in order, with their spelling
void s2_exotic_pass2(s2_exotic *ex, const char *seq, uint64_t len, std::vector<std::pair<uint64_t, std::string>> &out)
{
    if (!ex || len < 31 || !has_foreign(seq, len)) return;
    std::string up(seq, len);
    for (auto &c : up) c = exo_upper(c);
    char key[32];
    for_exotic_windows(up.data(), len, [&](uint64_t i) {
        exo_orient(up.data() + i, key);
        auto it = ex->map.find(key);
        if (it != ex->map.end() && it->second.informative) out.emplace_back(i, key);
    });
}
