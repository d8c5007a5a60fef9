// s2_host.cpp - host half of the C ABI: codecs, the BIO_hash row-order replay, the count-table
// formatter and the FASTA/FASTQ reader.  None of this computes k-mer hits; that is device work.
#include "../../include/strainer2_b200.h"
#include "s2_internal.h"
#include "s2_kmer.cuh"

#include <zlib.h>

#include <algorithm>
#include <cctype>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <thread>

int s2_env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

uint64_t s2_env_u64(const char *name, uint64_t dflt)
{
    const char *v = getenv(name);
    return (v && *v) ? strtoull(v, nullptr, 10) : dflt;
}

// ------------------------------------------------------------------------------------------------
// codecs
// ------------------------------------------------------------------------------------------------
// bit-compatible with encode_DNA_2_bit (/root/reference/src/up2bit.c:53-72): (c & 6) >> 1, MSB first
extern "C" uint64_t s2_encode_2bit(const char *dna, int len)
{
    uint64_t v = 0;
    for (int i = 0; i < len; ++i) v = (v << 2) + (uint64_t)(((unsigned char)dna[i] & 0x6u) >> 1);
    return v;
}

// decode_DNA_2_bit (src/up2bit.c:75-98), TWO_BIT_TO_DNA = A C T G (src/up2bit.c:14)
extern "C" void s2_decode_2bit(uint64_t v, int len, char *out)
{
    static const char letters[4] = { 'A', 'C', 'T', 'G' };
    if (len > 32) len = 32;
    if (len < 0) len = 0;
    if (len) v <<= (32 - len) * 2;
    for (int i = 0; i < len; ++i) { out[i] = letters[v >> 62]; v <<= 2; }
    out[len] = '\0';
}

extern "C" int s2_kmer_from_ascii(const char *s, uint64_t *out)
{
    uint64_t fwd = 0;
    for (int i = 0; i < S2_K; ++i) {
        const unsigned char c = (unsigned char)s[i] & 0xDFu;
        uint64_t code;
        switch (c) { case 'A': code = 0; break; case 'C': code = 1; break; case 'G': code = 2; break; case 'T': code = 3; break;
                     default: return -1; }
        fwd = (fwd << 2) | code;
    }
    *out = s2_canonical(fwd, s2_revcomp31(fwd));
    return 0;
}

extern "C" void s2_kmer_to_ascii(uint64_t k, char *out)
{
    for (int i = 0; i < S2_K; ++i) out[i] = s2_letter((uint32_t)(k >> (2 * (S2_K - 1 - i))) & 3u);
    out[S2_K] = '\0';
}

// ------------------------------------------------------------------------------------------------
// BIO_hash row order replay
// ------------------------------------------------------------------------------------------------
// The reference prints rows in ascending slot order of its string table (src/BIO_hash.c:174-188).  A
// key's slot depends only on djb2(key), the insertion order and the table's growth history:
//   insert : first empty slot from djb2 % M, linear probing           (src/BIO_hash.c:129-136)
//   grow   : after placing, if (N++ >= M/2): M *= 2 and every occupied old slot is re-inserted
//            in ascending old-slot order                               (src/BIO_hash.c:138, :39-61)
// so replaying those rules on 4-byte integers reproduces the order exactly.
extern "C" int s2_roworder_emulate(const uint32_t *djb2, uint64_t n, uint32_t initial_capacity,
                                   uint32_t *order_out, uint32_t *final_capacity)
{
    uint64_t M = initial_capacity ? initial_capacity : 8000000u;      // src/genome_compare.h:20
    if (M < 10) M = 10;                                               // src/BIO_hash.c:20-21
    if (n >= 0x7FFFFFFFull) { s2_set_error("too many keys for the row-order replay"); return -1; }
    std::vector<uint32_t> slots(M, 0u);                               // 0 = empty, else insertion index + 1
    uint64_t N = 0;
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t s = djb2[i] % M;
        while (slots[s]) s = (s + 1 == M) ? 0 : s + 1;
        slots[s] = (uint32_t)(i + 1);
        if (N++ >= M / 2) {
            if (M * 2 > 0xFFFFFFFFull) { s2_set_error("row-order replay: capacity overflow"); return -1; }
            std::vector<uint32_t> bigger(M * 2, 0u);
            const uint64_t M2 = M * 2;
            N = 0;
            for (uint64_t o = 0; o < M; ++o) {
                const uint32_t id = slots[o];
                if (!id) continue;
                uint64_t s2 = djb2[id - 1] % M2;
                while (bigger[s2]) s2 = (s2 + 1 == M2) ? 0 : s2 + 1;
                bigger[s2] = id;
                ++N;
            }
            slots.swap(bigger);
            M = M2;
        }
    }
    uint64_t k = 0;
    for (uint64_t s = 0; s < M; ++s) if (slots[s]) order_out[k++] = slots[s] - 1;
    if (final_capacity) *final_capacity = (uint32_t)M;
    return k == n ? 0 : -1;
}

// ------------------------------------------------------------------------------------------------
// count table formatter  (print_hash_counts, src/kmer_scrub_count.c:134-156)
// ------------------------------------------------------------------------------------------------
static inline char *put_int(char *p, uint32_t u)
{
    // the reference prints its unsigned counters with %d: values above 2^31-1 come out negative
    int32_t v = (int32_t)u;
    uint32_t a;
    if (v < 0) { *p++ = '-'; a = (uint32_t)(-(int64_t)v); } else a = (uint32_t)v;
    char tmp[12]; int n = 0;
    do { tmp[n++] = (char)('0' + a % 10); a /= 10; } while (a);
    while (n) *p++ = tmp[--n];
    return p;
}

extern "C" int s2_format_count_table(FILE *out, const uint64_t *keys, const uint32_t *order, uint64_t n,
                                     const uint32_t *const *cols, int n_print_cols, int n_threads)
{
    static const char header[] = "#kmer\treference_count\tpangenome_count\tmetagenome_count\tdrug_count\n";
    if (fwrite(header, 1, sizeof header - 1, out) != sizeof header - 1) { s2_set_error("write failed"); return -1; }
    if (n_threads < 1) n_threads = 1;
    const uint64_t chunk = 1u << 18;                         // rows per work item
    const size_t row_max = S2_K + 4 * 12 + 2;
    std::vector<std::vector<char>> bufs(n_threads);
    std::vector<size_t> used(n_threads);
    for (auto &b : bufs) b.resize(chunk * row_max);
    for (uint64_t base = 0; base < n; base += chunk * n_threads) {
        auto work = [&](int tid) {
            const uint64_t lo = base + (uint64_t)tid * chunk, hi = std::min(n, lo + chunk);
            char *p = bufs[tid].data();
            for (uint64_t r = lo; r < hi; ++r) {
                const uint32_t id = order ? order[r] : (uint32_t)r;
                const uint64_t k = keys[id];
                for (int i = 0; i < S2_K; ++i) *p++ = s2_letter((uint32_t)(k >> (2 * (S2_K - 1 - i))) & 3u);
                for (int c = 0; c < n_print_cols; ++c) { *p++ = '\t'; p = put_int(p, cols[c][id]); }
                *p++ = '\n';
            }
            used[tid] = lo < hi ? (size_t)(p - bufs[tid].data()) : 0;
        };
        std::vector<std::thread> th;
        for (int t = 1; t < n_threads; ++t) th.emplace_back(work, t);
        work(0);
        for (auto &t : th) t.join();
        for (int t = 0; t < n_threads; ++t)
            if (used[t] && fwrite(bufs[t].data(), 1, used[t], out) != used[t]) { s2_set_error("write failed"); return -1; }
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// FASTA / FASTQ reader
// ------------------------------------------------------------------------------------------------
// Our own block parser; what it must reproduce is the *observable* behaviour of the parser the
// reference vendors (src/kseq.h:171-211), because record boundaries decide which windows exist:
//  * a record starts at the next '>' or '@' (anywhere, when the previous record was FASTQ);
//  * the name ends at the first whitespace, the rest of the header line is ignored;
//  * sequence lines are concatenated; a line whose first byte is '>', '+' or '@' ends the sequence;
//    '\n'-only lines are skipped; ONE trailing '\r' is dropped per line when the sequence so far is
//    longer than one byte;
//  * after '+': the rest of that line is skipped, quality lines are consumed until their total
//    length reaches the sequence length; shorter / longer totals or EOF give -2;
//  * after a FASTA record that ended at EOF the next call resets the length to 0 before it
//    notices EOF; after a FASTQ record the stale length survives (strain_detect's PE2 logic can
//    observe this, src/strain_detect.c:496-504).
struct s2_reader {
    gzFile f = nullptr;
    std::vector<unsigned char> buf;
    size_t begin = 0, end = 0;
    bool eof = false;
    bool damaged = false;          // gzread reported an error (Z_DATA_ERROR: corrupt DEFLATE data or a CRC mismatch)
    int last_char = 0;
    std::vector<char> seq, qual;
    size_t seq_len = 0, qual_len = 0;

    int fill()
    {
        if (eof) return 0;
        begin = 0;
        int got = gzread(f, buf.data(), (unsigned)buf.size());
        // 0 = end of file, also for a file cut short (zlib reports that as Z_BUF_ERROR and returns what it has, so the
        // reference ends there too).  < 0 = damaged data: the reference's kseq never returns from that (kseq.h:72,99
        // take only 0 for the end, so it re-reads the error for ever); here the stream ends and the callers fail loudly
        if (got < 0) damaged = true;
        if (got <= 0) { end = 0; eof = true; return 0; }
        end = (size_t)got;
        return 1;
    }
    inline int getc()
    {
        if (begin >= end && !fill()) return -1;
        return buf[begin++];
    }
    inline bool at_eof() { return begin >= end && !fill(); }
    inline void push(char c)
    {
        if (seq_len + 2 > seq.size()) seq.resize(std::max<size_t>(256, seq.size() * 2));
        seq[seq_len++] = c;
    }
    // rest of the current line, appended to `dst` (or discarded when dst == nullptr); swallows the
    // '\n'; then drops ONE trailing '\r' if the string is longer than one byte.  Returns false when
    // called at end of stream (nothing consumed, nothing stripped).
    bool rest_of_line(std::vector<char> *dst, size_t *len)
    {
        if (at_eof()) return false;
        for (;;) {
            if (begin >= end && !fill()) break;
            unsigned char *s = buf.data() + begin;
            unsigned char *nl = (unsigned char *)memchr(s, '\n', end - begin);
            const size_t take = nl ? (size_t)(nl - s) : end - begin;
            if (take && dst) {
                if (*len + take + 2 > dst->size()) dst->resize(std::max(dst->size() * 2, *len + take + 2));
                memcpy(dst->data() + *len, s, take);
                *len += take;
            }
            begin += take + (nl ? 1 : 0);
            if (nl) break;
        }
        if (dst && *len > 1 && (*dst)[*len - 1] == '\r') --*len;
        return true;
    }
};

extern "C" s2_reader *s2_reader_open(const char *path)
{
    gzFile f = gzopen(path, "r");
    if (!f) { s2_set_error("could not read file %s", path); return nullptr; }
    gzbuffer(f, 1u << 18);
    s2_reader *r = new s2_reader();
    r->f = f;
    r->buf.resize(1u << 18);
    r->seq.resize(256);
    r->qual.resize(256);
    return r;
}

extern "C" void s2_reader_close(s2_reader *r)
{
    if (!r) return;
    gzclose(r->f);
    delete r;
}

extern "C" uint64_t s2_reader_len(const s2_reader *r) { return r->seq_len; }

extern "C" int s2_reader_damaged(const s2_reader *r) { return r && r->damaged ? 1 : 0; }

extern "C" int64_t s2_reader_next(s2_reader *r, const char **seq_out)
{
    int c;
    if (seq_out) *seq_out = r->seq.data();
    if (r->last_char == 0) {
        while ((c = r->getc()) != -1 && c != '>' && c != '@') { }
        if (c == -1) return -1;
        r->last_char = c;
    }
    r->seq_len = 0;
    // name: up to the first whitespace; then the rest of the header line unless that was the newline
    if (r->at_eof()) return -1;
    int delim = 0;
    while ((c = r->getc()) != -1) if (isspace(c)) { delim = c; break; }
    if (delim != '\n') r->rest_of_line(nullptr, nullptr);
    // sequence lines
    while ((c = r->getc()) != -1 && c != '>' && c != '+' && c != '@') {
        if (c == '\n') continue;
        r->push((char)c);
        r->rest_of_line(&r->seq, &r->seq_len);
    }
    if (c == '>' || c == '@') r->last_char = c;
    if (r->seq_len + 2 > r->seq.size()) r->seq.resize(r->seq_len + 2);
    r->seq[r->seq_len] = '\0';
    if (seq_out) *seq_out = r->seq.data();
    if (c != '+') return (int64_t)r->seq_len;
    while ((c = r->getc()) != -1 && c != '\n') { }
    if (c == -1) return -2;
    r->qual_len = 0;
    while (r->rest_of_line(&r->qual, &r->qual_len) && r->qual_len < r->seq_len) { }
    r->last_char = 0;
    if (r->qual_len != r->seq_len) return -2;
    return (int64_t)r->seq_len;
}
