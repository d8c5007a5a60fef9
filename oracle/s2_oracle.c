/* s2_oracle.c - TEST INFRASTRUCTURE ONLY.  See s2_oracle.h for scope and parity status (PINNED).
 *
 * A deliberately simple, string-based, single-threaded restatement of what the reference computes on
 * the k-mer scan path.  Every function cites the reference lines it follows.  Nothing in here is
 * tuned: it is the checker, never the thing measured or shipped.
 */
#define _GNU_SOURCE
#include "s2_oracle.h"
#include <ctype.h>
#include <errno.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <zlib.h>

#define K S2O_K

/* ======================================================================================== */
/* primitives                                                                               */
/* ======================================================================================== */

/* src/BIO_hash.c:208-216: h = h*33 + (signed char)c in 32-bit wraparound, start 5381 */
uint32_t s2o_djb2(const char *s)
{
    uint32_t h = 5381u;
    for (; *s; ++s) h = (h << 5) + h + (uint32_t)(int32_t)(signed char)*s;
    return h;
}

/* src/BIO_sequence.c:203-213, restated as a switch (verified against the real array by
 * tests/test_oracle_golden.py::test_complement_table_matches_reference). -1 = "no complement". */
int s2o_complement(int c)
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

/* src/genome_compare.c:443-451 */
int s2o_contains_N(const char *s)
{
    for (; *s; ++s) if (*s == 'N') return 1;
    return 0;
}

/* src/genome_compare.c:1122-1141: compare w[i] with complement(w[k-1-i]) as (signed) char */
int s2o_rc_strcmp(const char *w, int k)
{
    for (int i = 0; i < k; ++i) {
        char rc = (char)s2o_complement((unsigned char)w[k - 1 - i]);
        if (w[i] > rc) return 1;
        if (rc > w[i]) return -1;
    }
    return 0;
}

/* src/genome_compare.c:1100-1120: forward wins ties; otherwise build the reverse complement */
const char *s2o_orient(const char *w, char *scratch, int k)
{
    if (s2o_rc_strcmp(w, k) >= 0) return w;
    scratch[k] = '\0';
    for (int i = k - 1, j = 0; i >= 0; --i, ++j) scratch[i] = (char)s2o_complement((unsigned char)w[j]);
    return scratch;
}

/* src/up2bit.c:53-72: MSB-first, code (c & 6) >> 1 => A0 C1 T2 G3 */
uint64_t s2o_encode_2bit(const char *dna, int len)
{
    uint64_t v = 0;
    for (int i = 0; i < len; ++i) v = (v << 2) + (uint64_t)((dna[i] & 0x6) >> 1);
    return v;
}

/* src/up2bit.c:75-98 */
void s2o_decode_2bit(uint64_t v, int len, char *out)
{
    static const char map[4] = { 'A', 'C', 'T', 'G' };   /* src/up2bit.c:14 */
    v <<= (32 - len) * 2;
    for (int i = 0; i < len; ++i) { out[i] = map[v >> 62]; v <<= 2; }
    out[len] = '\0';
}

/* src/BIO_sequence.c:228-234 (libc toupper, C locale, up to the first NUL) */
static void upper_inplace(char *s)
{
    for (; *s; ++s) *s = (char)toupper((unsigned char)*s);
}

/* ======================================================================================== */
/* reader: byte-at-a-time restatement of kseq_read() (src/kseq.h:171-211)                   */
/* ======================================================================================== */
typedef struct { size_t l, m; char *s; } ostr;

struct s2o_reader {
    gzFile f;
    unsigned char buf[16384];   /* src/kseq.h:221 */
    int begin, end, is_eof;
    int last_char;
    ostr seq, qual;
};

static void ostr_push(ostr *s, int c)
{
    if (s->l + 2 > s->m) { s->m = s->m ? s->m * 2 : 256; s->s = (char *)realloc(s->s, s->m); }
    s->s[s->l++] = (char)c;
}

/* ks_getc, src/kseq.h:66-76 */
static int rd_getc(s2o_reader *r)
{
    if (r->is_eof && r->begin >= r->end) return -1;
    if (r->begin >= r->end) {
        r->begin = 0;
        r->end = gzread(r->f, r->buf, sizeof r->buf);
        /* gzread < 0 (damaged DEFLATE data, CRC mismatch): the reference tests only `== 0` (src/kseq.h:72,:99) and
         * re-reads the error for ever - measured, it never returns from such a file.  No behaviour to restate: the
         * oracle ends the stream; the product ends the run with an error (DESIGN.md 4.4 "Damaged data"). */
        if (r->end <= 0) { r->end = 0; r->is_eof = 1; return -1; }
    }
    return r->buf[r->begin++];
}

/* true iff no byte is left in the stream (the "!gotany && ks_eof" exit of ks_getuntil2, :140) */
static int rd_at_eof(s2o_reader *r)
{
    int c = rd_getc(r);
    if (c < 0) return 1;
    r->begin--;
    return 0;
}

/* ks_getuntil2 with KS_SEP_LINE (src/kseq.h:93-146): append the rest of the current line to str
 * (str may be NULL = discard), swallow the '\n', strip ONE trailing '\r' if the string is longer
 * than 1.  Returns -1 when called with the stream already exhausted, else the string length. */
static long rd_line(s2o_reader *r, ostr *str)
{
    int c;
    if (rd_at_eof(r)) return -1;
    while ((c = rd_getc(r)) != -1 && c != '\n') if (str) ostr_push(str, c);
    if (str) {
        if (str->l > 1 && str->s[str->l - 1] == '\r') --str->l;
        if (!str->s) ostr_push(str, 0), str->l = 0;
        str->s[str->l] = '\0';
        return (long)str->l;
    }
    return 0;
}

/* ks_getuntil with KS_SEP_SPACE (record name).  *dret = delimiter hit, 0 if EOF ended it. */
static long rd_name(s2o_reader *r, int *dret)
{
    int c;
    *dret = 0;
    if (rd_at_eof(r)) return -1;
    while ((c = rd_getc(r)) != -1) if (isspace(c)) { *dret = c; break; }
    return 0;
}

s2o_reader *s2o_reader_open(const char *path)
{
    gzFile f = gzopen(path, "r");
    if (!f) return NULL;
    s2o_reader *r = (s2o_reader *)calloc(1, sizeof *r);
    r->f = f;
    return r;
}

int s2o_reader_next(s2o_reader *r)
{
    int c;
    if (r->last_char == 0) {                                    /* :176-180 */
        while ((c = rd_getc(r)) != -1 && c != '>' && c != '@') { }
        if (c == -1) return -1;
        r->last_char = c;
    }
    r->seq.l = r->qual.l = 0;                                   /* :181 */
    if (rd_name(r, &c) < 0) return -1;                          /* :182 */
    if (c != '\n') rd_line(r, NULL);                            /* :183 comment */
    if (!r->seq.s) { r->seq.m = 256; r->seq.s = (char *)malloc(r->seq.m); }
    while ((c = rd_getc(r)) != -1 && c != '>' && c != '+' && c != '@') {   /* :188-192 */
        if (c == '\n') continue;
        ostr_push(&r->seq, c);
        rd_line(r, &r->seq);
    }
    if (c == '>' || c == '@') r->last_char = c;                 /* :193 */
    if (r->seq.l + 2 > r->seq.m) { r->seq.m = r->seq.l + 2; r->seq.s = (char *)realloc(r->seq.s, r->seq.m); }
    r->seq.s[r->seq.l] = '\0';                                  /* :199 */
    if (c != '+') return (int)r->seq.l;                         /* :200 FASTA */
    while ((c = rd_getc(r)) != -1 && c != '\n') { }             /* :205 rest of '+' line */
    if (c == -1) return -2;                                     /* :206 */
    while (rd_line(r, &r->qual) >= 0 && r->qual.l < r->seq.l) { }   /* :207 */
    r->last_char = 0;                                           /* :208 */
    if (r->seq.l != r->qual.l) return -2;                       /* :209 */
    return (int)r->seq.l;
}

char *s2o_reader_seq(s2o_reader *r) { return r->seq.s; }
size_t s2o_reader_len(s2o_reader *r) { return r->seq.l; }

void s2o_reader_close(s2o_reader *r)
{
    if (!r) return;
    gzclose(r->f);
    free(r->seq.s); free(r->qual.s);
    free(r);
}

/* ======================================================================================== */
/* table = BIO_hash restated (src/BIO_hash.c)                                               */
/* ======================================================================================== */
typedef struct { char key[K + 1]; unsigned vec[6]; } oentry;

struct s2o_table {
    unsigned M, N;          /* capacity, live keys (src/BIO_hash.h:47-51) */
    uint32_t *slot;         /* 0 = empty, else entry index + 1 */
    oentry *ent; size_t n_ent, m_ent;
    int vec_size;
};

s2o_table *s2o_table_new(unsigned cap, int vec_size)
{
    s2o_table *t = (s2o_table *)calloc(1, sizeof *t);
    if (!cap) cap = 1000; else if (cap < 10) cap = 10;          /* src/BIO_hash.c:18-21 */
    t->M = cap; t->vec_size = vec_size;
    t->slot = (uint32_t *)calloc(cap, sizeof *t->slot);
    return t;
}

void s2o_table_free(s2o_table *t) { if (t) { free(t->slot); free(t->ent); free(t); } }
unsigned s2o_table_size(const s2o_table *t) { return t->N; }
unsigned s2o_table_capacity(const s2o_table *t) { return t->M; }

unsigned *s2o_table_search(const s2o_table *t, const char *key)      /* src/BIO_hash.c:161-172 */
{
    unsigned i = s2o_djb2(key) % t->M;
    while (t->slot[i]) {
        oentry *e = &t->ent[t->slot[i] - 1];
        if (strcmp(key, e->key) == 0) return e->vec;
        i = (i + 1) % t->M;
    }
    return NULL;
}

static void place(s2o_table *t, uint32_t id)
{
    unsigned i = s2o_djb2(t->ent[id - 1].key) % t->M;
    while (t->slot[i]) i = (i + 1) % t->M;
    t->slot[i] = id;
}

static void expand(s2o_table *t)                                      /* src/BIO_hash.c:39-61 */
{
    unsigned oldM = t->M;
    uint32_t *old = t->slot;
    t->M += t->M;
    t->N = 0;
    t->slot = (uint32_t *)calloc(t->M, sizeof *t->slot);
    for (unsigned i = 0; i < oldM; ++i)
        if (old[i]) { place(t, old[i]); t->N++; }   /* re-insertion in ascending old-slot order */
    free(old);
}

void s2o_table_add(s2o_table *t, const char *key, const unsigned *vec)   /* src/BIO_hash.c:129-139 */
{
    if (t->n_ent == t->m_ent) {
        t->m_ent = t->m_ent ? t->m_ent * 2 : 1u << 20;
        t->ent = (oentry *)realloc(t->ent, t->m_ent * sizeof *t->ent);
    }
    oentry *e = &t->ent[t->n_ent++];
    memset(e, 0, sizeof *e);
    strncpy(e->key, key, K);
    if (vec) memcpy(e->vec, vec, sizeof(unsigned) * (size_t)t->vec_size);
    place(t, (uint32_t)t->n_ent);
    if (t->N++ >= t->M / 2) expand(t);       /* doubling fires while inserting key number M/2+1 */
}

const char *s2o_table_key_at(const s2o_table *t, unsigned i, unsigned **vec)   /* :174-188 */
{
    /* i-th occupied slot in ascending slot order; O(M) walk cached by a static cursor would be
     * faster, but callers below iterate slots directly.  Kept for tests on small tables. */
    unsigned seen = 0;
    for (unsigned s = 0; s < t->M; ++s)
        if (t->slot[s]) {
            if (seen == i) { if (vec) *vec = t->ent[t->slot[s] - 1].vec; return t->ent[t->slot[s] - 1].key; }
            ++seen;
        }
    return NULL;
}

/* ======================================================================================== */
/* the path                                                                                 */
/* ======================================================================================== */

/* src/genome_compare.c:967-1030.  Contigs shorter than k-1 make the reference's unsigned loop bound
 * underflow (undefined, aborts in practice - SURVEY D9); the oracle defines them as "no windows". */
int s2o_build(s2o_table *t, const char *ref_file, int default_count, int increment, int vec_idx)
{
    s2o_reader *r = s2o_reader_open(ref_file);
    char scratch[K + 1], win[K + 1];
    unsigned vec[6];
    if (!r) return -1;
    win[K] = '\0';
    while (s2o_reader_next(r) >= 0) {
        char *s = s2o_reader_seq(r);
        size_t l = s2o_reader_len(r);
        upper_inplace(s);                                                /* :996 */
        for (size_t i = 0; i + K <= l; ++i) {                            /* :1000 */
            memcpy(win, s + i, K);
            const char *o = s2o_orient(win, scratch, K);                 /* :1005 */
            if (!s2o_contains_N(o)) {                                    /* :1007 */
                unsigned *c = s2o_table_search(t, o);
                if (!c) {
                    memset(vec, 0, sizeof vec);
                    vec[vec_idx] = (unsigned)default_count;              /* :1011-1013 */
                    s2o_table_add(t, o, vec);
                } else {
                    c[vec_idx] += (unsigned)increment;                   /* :1016 */
                }
            }
        }
    }
    s2o_reader_close(r);
    return 0;
}

/* src/genome_compare.c:179-236 */
int s2o_count_file(s2o_table *t, const char *file, unsigned col, uint64_t *n_bases, uint64_t *n_windows)
{
    s2o_reader *r = s2o_reader_open(file);
    char scratch[K + 1], win[K + 1];
    if (!r) return -1;
    win[K] = '\0';
    while (s2o_reader_next(r) >= 0) {                                    /* :203 */
        char *s = s2o_reader_seq(r);
        size_t l = s2o_reader_len(r);
        if (n_bases) *n_bases += l;
        if (l < K) continue;                                             /* :204 */
        upper_inplace(s);                                                /* :208 */
        int has_N = s2o_contains_N(s);                                   /* :210 */
        for (size_t i = 0; i + K <= l; ++i) {                            /* :213 */
            memcpy(win, s + i, K);
            const char *o = s2o_orient(win, scratch, K);                 /* :217 */
            if (!has_N || !s2o_contains_N(o)) {                          /* :219 */
                unsigned *c = s2o_table_search(t, o);                    /* :220 */
                if (c) c[col] += 1;                                      /* :222 */
            }
            if (n_windows) ++*n_windows;
        }
    }
    s2o_reader_close(r);
    return 0;
}

/* src/genome_compare.c:115-146 (skip_file != NULL) and :149-177 */
int s2o_count_list(s2o_table *t, const char *list_file, const char *skip_file, unsigned col,
                   FILE *progress, FILE *err)
{
    FILE *fp = fopen(list_file, "r");
    char *line = NULL, *pos;
    size_t cap = 0;
    if (!fp) {
        fprintf(err, "could not read file %s in GEN_all_kmer_counts()\n", list_file);   /* :125,:159 */
        return EXIT_FAILURE;
    }
    while (getline(&line, &cap, fp) != -1) {
        if ((pos = strchr(line, '\n')) != NULL) *pos = '\0';
        if (progress) {
            time_t now = time(NULL);
            fprintf(progress, "%s\t%s", line, asctime(localtime(&now)));                 /* :167-170 */
        }
        if (skip_file && strcmp(skip_file, line) == 0) {
            fprintf(err, "skipping %s (identical match)\n", line);                       /* :141 */
            continue;
        }
        if (s2o_count_file(t, line, col, NULL, NULL) != 0) {
            fprintf(err, "could not read file %s in GEN_calculate_kmer_count()\n", line); /* :196 */
            fclose(fp); free(line);
            return EXIT_FAILURE;
        }
    }
    fclose(fp);
    free(line);
    return 0;
}

/* src/kmer_scrub_count.c:134-156 (header always 5 columns; %d of unsigned) */
void s2o_print_counts(const s2o_table *t, int with_C, FILE *out)
{
    fprintf(out, "#kmer\treference_count\tpangenome_count\tmetagenome_count\tdrug_count\n");
    for (unsigned s = 0; s < t->M; ++s) {
        if (!t->slot[s]) continue;
        const oentry *e = &t->ent[t->slot[s] - 1];
        if (with_C) fprintf(out, "%s\t%d\t%d\t%d\t%d\n", e->key, e->vec[0], e->vec[1], e->vec[2], e->vec[3]);
        else        fprintf(out, "%s\t%d\t%d\t%d\n", e->key, e->vec[0], e->vec[1], e->vec[2]);
    }
}

/* src/kmer_scrub_count.c:72-123 */
int s2o_kmer_scrub_count(const char *r_file, const char *A_file, const char *B_file,
                         const char *C_file, const char *p_file, FILE *out, FILE *err)
{
    FILE *progress = NULL;
    int rc;
    if (!r_file || !A_file || !B_file) return 1;
    if (p_file) {
        progress = fopen(p_file, "w");
        if (!progress) { fprintf(err, "could not open progress file %s\n", p_file); return EXIT_FAILURE; }
        fprintf(progress, "adding kmer counts for:\n");
    }
    s2o_table *t = s2o_table_new(S2O_INITIAL_CAPACITY, 4);
    if (s2o_build(t, r_file, 1, 1, 0) != 0) {
        fprintf(err, "could not read file %s GEN_hash_sequences_set_count_vec()\n", r_file);
        return EXIT_FAILURE;
    }
    if ((rc = s2o_count_list(t, A_file, NULL, 1, progress, err)) != 0) return rc;
    if ((rc = s2o_count_list(t, B_file, NULL, 2, progress, err)) != 0) return rc;
    if (C_file && (rc = s2o_count_list(t, C_file, r_file, 3, progress, err)) != 0) return rc;
    s2o_print_counts(t, C_file != NULL, out);
    s2o_table_free(t);
    if (progress) fclose(progress);
    return 0;
}

/* src/strain_detect.c:668-726.  Lines are NOT upper-cased; gzgets with a 100-byte buffer. */
int s2o_flag_informative(s2o_table *t, const char *a_file, FILE *msg, unsigned *n_flagged)
{
    gzFile fp = gzopen(a_file, "r");
    char line[100], scratch[101], *pos;
    unsigned n = 0;
    if (!fp) return -1;
    while (gzgets(fp, line, 100)) {
        if (line[0] == '#') continue;
        if ((pos = strchr(line, '\n')) != NULL) *pos = '\0';
        if (strlen(line) == K) {
            const char *o = s2o_orient(line, scratch, K);
            unsigned *c = s2o_table_search(t, o);
            if (c) { c[0] = 2; ++n; }                                            /* :702-707 */
            else fprintf(msg, "error could not find informative kmer %s in the total kmer list\n", line);
        } else {
            fprintf(msg, "error string length in the scrubbed kmer file (%s) must be the same size as the kmer "
                         "length (scrubbed kmer, scrubbed kmer len, seed len): %s, %d, %d\n",
                    a_file, line, (int)strlen(line), K);
        }
    }
    gzclose(fp);
    if (n_flagged) *n_flagged = n;
    return 0;
}

/* pass 1 over one read (src/strain_detect.c:457-491 / :508-539): whole-read reverse complement,
 * per-window unsigned-byte compare (strcmp), reverse complement wins ties. */
static void pass1(const s2o_table *t, const char *s, size_t l, char **rcbuf, size_t *rccap,
                  int *hits, int *inf, unsigned long long *evaluated)
{
    if (*rccap < l + 1) { *rccap = l + 1; *rcbuf = (char *)realloc(*rcbuf, *rccap); }
    char *rc = *rcbuf, win[K + 1];
    for (size_t j = 0; j < l; ++j) rc[j] = (char)s2o_complement((unsigned char)s[l - 1 - j]);
    rc[l] = '\0';
    int has_N = s2o_contains_N(s);
    win[K] = '\0';
    for (size_t i = 0; i + K <= l; ++i) {
        const char *f = s + i, *r = rc + (l - K - i);
        memcpy(win, memcmp(f, r, K) > 0 ? f : r, K);
        if (!has_N || !s2o_contains_N(win)) {
            unsigned *c = s2o_table_search(t, win);
            if (c) { ++*hits; if (c[0] == 2) ++*inf; }
        }
        ++*evaluated;
    }
}

/* pass 2 over one read (src/strain_detect.c:554-591 / :594-623) */
static void pass2(const s2o_table *t, const char *s, size_t l, const char *pe1_file,
                  int h1, int i1, int h2, int i2, FILE *out)
{
    char scratch[K + 1], win[K + 1];
    win[K] = '\0';
    for (size_t i = 0; i + K <= l; ++i) {
        memcpy(win, s + i, K);
        const char *o = s2o_orient(win, scratch, K);
        if (!s2o_contains_N(o)) {
            unsigned *c = s2o_table_search(t, o);
            if (c && c[0] == 2) fprintf(out, "%s\t%d\t%d\t%d\t%d\t%s\n", pe1_file, h1, i1, h2, i2, o);
        }
    }
}

/* src/strain_detect.c:387-663, including the stale-state behaviour for reads shorter than k
 * (counters and the PE1 copy are only refreshed inside the length guards, :444-448 / :497-500). */
int s2o_quantify_hits(s2o_table *t, const char *pe1, const char *pe2, int is_pe,
                      unsigned genome_kmers, unsigned genome_informative, FILE *out, FILE *err)
{
    s2o_reader *r1 = s2o_reader_open(pe1), *r2 = NULL;
    if (!r1) {
        fprintf(err, "could not read file (read1) %s in quantify_hits_PE() (error: %s)\n", pe1, strerror(errno));
        return EXIT_FAILURE;
    }
    if (is_pe == 1) {
        r2 = s2o_reader_open(pe2);
        if (!r2) {
            fprintf(err, "could not read file (read2) is_PE %s in quantify_hits_PE() (error: %s)\n", pe2, "(null)");
            return EXIT_FAILURE;
        }
    } else if (is_pe == 2) r2 = r1;

    int h1 = 0, i1 = 0, h2 = 0, i2 = 0;
    char *copy = NULL, *rcbuf = NULL; size_t copycap = 0, rccap = 0, copylen = 0;
    unsigned long long evaluated = 0, reads = 0;

    while (s2o_reader_next(r1) >= 0) {
        if (s2o_reader_len(r1) >= K) {
            char *s = s2o_reader_seq(r1); size_t l = s2o_reader_len(r1);
            ++reads; h1 = 0; i1 = 0; copylen = l;
            upper_inplace(s);
            if (copycap < l + 1) { copycap = l + 1; copy = (char *)realloc(copy, copycap); }
            memcpy(copy, s, l + 1);
            pass1(t, s, l, &rcbuf, &rccap, &h1, &i1, &evaluated);
        }
        if (is_pe) {
            int l2 = s2o_reader_next(r2);
            if (s2o_reader_len(r2) >= K) {
                h2 = 0; i2 = 0;
                if (l2 < 0) {
                    fprintf(err, "reached end of PE2 (%s) before end of PE1 (%s), check that file names are correct\n",
                            pe2 ? pe2 : "(null)", pe1);
                    return EXIT_FAILURE;
                }
                upper_inplace(s2o_reader_seq(r2));
                pass1(t, s2o_reader_seq(r2), s2o_reader_len(r2), &rcbuf, &rccap, &h2, &i2, &evaluated);
            }
        }
        if (h1 + h2 >= 1 && i1 + i2 >= 1) {                                       /* :547 */
            if (copylen >= K) pass2(t, copy, copylen, pe1, h1, i1, h2, i2, out);
            if (is_pe && s2o_reader_len(r2) >= K)
                pass2(t, s2o_reader_seq(r2), s2o_reader_len(r2), pe1, h1, i1, h2, i2, out);
        }
    }
    fprintf(out, "#%s\ttotal_kmer_evaluated\t%lld\n", pe1, (long long)evaluated);              /* :633-636 */
    fprintf(out, "#%s\ttotal_reads_evaluated\t%lld\n", pe1, (long long)reads);
    fprintf(out, "#%s\ttotal_genome_kmers\t%lld\n", pe1, (long long)genome_kmers);
    fprintf(out, "#%s\ttotal_genome_informative_kmers\t%lld\n", pe1, (long long)genome_informative);
    free(copy); free(rcbuf);
    if (r2 && r2 != r1) s2o_reader_close(r2);
    s2o_reader_close(r1);
    return 0;
}

static int file_type(const char *s)                                     /* src/strain_detect.c:728-747 */
{
    if (!strcmp(s, "SE") || !strcmp(s, "se")) return 0;
    if (!strcmp(s, "PE") || !strcmp(s, "pe")) return 1;
    if (!strcmp(s, "PEI") || !strcmp(s, "pei") || !strcmp(s, "IPE") || !strcmp(s, "ipe")) return 2;
    return -1;
}

static int cmp_desc(const void *a, const void *b)                         /* src/strain_detect.c:255-261 */
{
    return (int)(*(const unsigned *)b - *(const unsigned *)a);
}

static int removed_at(unsigned thr, const unsigned *v, unsigned n)        /* kmer_removed, :242-252 */
{
    int c = 0;
    for (unsigned i = 0; i < n; ++i) if (v[i] >= thr) ++c;
    return c;
}

/* background_filter, src/strain_detect.c:160-240 (fraction_to_remove = 0.5, :82).  num_inform = number of
 * informative-list LINES that matched (duplicates count).  Returns 0 or the reference's exit status. */
int s2o_background_filter(s2o_table *t, const char *background_file, unsigned num_inform, FILE *msg, FILE *err)
{
    const double fraction = 0.5;
    unsigned keep = (unsigned)(int)(num_inform * fraction);
    unsigned *v = (unsigned *)calloc(num_inform ? num_inform : 1, sizeof *v);
    unsigned n = 0, thr = 1;
    int rc;
    fprintf(msg, "#removing %f proportion of %s kmers; informative %d keep at least %d\n", fraction, background_file,
            num_inform, keep);
    if ((rc = s2o_count_list(t, background_file, NULL, 5, NULL, err)) != 0) return rc;
    for (unsigned s = 0; s < t->M; ++s) {
        if (!t->slot[s]) continue;
        const oentry *e = &t->ent[t->slot[s] - 1];
        if (e->vec[0] != 2) continue;
        if (n >= num_inform) { fprintf(err, "Error: too many background kmers\n"); return 1; }
        v[n++] = e->vec[5];
    }
    qsort(v, num_inform, sizeof *v, cmp_desc);
    if (keep >= 1 && v[keep - 1] > thr) thr = v[keep - 1];
    while ((unsigned)removed_at(thr, v, num_inform) > keep) ++thr;
    unsigned demoted = 0;
    for (size_t i = 0; i < t->n_ent; ++i)
        if (t->ent[i].vec[0] == 2 && t->ent[i].vec[5] >= thr) { t->ent[i].vec[0] = 1; ++demoted; }
    fprintf(msg, "#final_threshold %d removes %d background kmers %d removed\n", thr, removed_at(thr, v, num_inform), demoted);
    free(v);
    return 0;
}

/* src/strain_detect.c:137-146 + :263-384 (output text uncompressed) */
int s2o_strain_detect(const char *r_file, const char *a_file, const char *B_file,
                      const char *b_file, const char *c_file, const char *type,
                      FILE *out, FILE *msg, FILE *err)
{
    return s2o_strain_detect_g(r_file, a_file, NULL, B_file, b_file, c_file, type, out, msg, err);
}

int s2o_strain_detect_g(const char *r_file, const char *a_file, const char *g_file, const char *B_file,
                        const char *b_file, const char *c_file, const char *type,
                        FILE *out, FILE *msg, FILE *err)
{
    s2o_table *t = s2o_table_new(S2O_INITIAL_CAPACITY, 6);
    unsigned n_lines = 0, n_inf = 0;
    int rc = 0;
    if (s2o_build(t, r_file, 1, 0, 0) != 0) {
        fprintf(err, "could not read file %s GEN_hash_sequences_set_count_vec()\n", r_file);
        return EXIT_FAILURE;
    }
    if (s2o_flag_informative(t, a_file, msg, &n_lines) != 0) {
        fprintf(err, "could not read file %s in hash_scrubbed_kmers()\n", a_file);
        return EXIT_FAILURE;
    }
    if (g_file && (rc = s2o_background_filter(t, g_file, n_lines, msg, err)) != 0) return rc;   /* :142-143 */
    for (size_t e = 0; e < t->n_ent; ++e) if (t->ent[e].vec[0] == 2) ++n_inf;    /* :285-290 */
    unsigned n_keys = s2o_table_size(t);

    if (B_file) {
        FILE *fp = fopen(B_file, "r");
        char *line = NULL, *pos; size_t cap = 0;
        if (!fp) {
            fprintf(err, "could not read file file_of_filenames %s in quantify_hits_all_files()\n", B_file);
            return EXIT_FAILURE;
        }
        while (rc == 0 && getline(&line, &cap, fp) != -1) {
            if ((pos = strchr(line, '\n')) != NULL) *pos = '\0';
            char *tok = strtok(line, "\t");
            int pe = tok ? file_type(tok) : -1;
            if (pe < 0) { fprintf(msg, "unknown file type skipping line (%s)\n", tok ? tok : "(null)"); continue; }
            char *f1 = strtok(NULL, "\t");
            if (!f1) { fprintf(msg, "ERROR: no first file specified for %s\n", line); continue; }
            if (pe == 1) {
                char *f2 = strtok(NULL, "\t");
                if (!f2) { fprintf(msg, "ERROR: no second file specified for PE: %s\n", line); continue; }
                rc = s2o_quantify_hits(t, f1, f2, 1, n_keys, n_inf, out, err);
            } else {
                rc = s2o_quantify_hits(t, f1, NULL, pe, n_keys, n_inf, out, err);
            }
        }
        fclose(fp); free(line);
    } else {
        int pe = type ? file_type(type) : 0;
        rc = s2o_quantify_hits(t, b_file, c_file, pe, n_keys, n_inf, out, err);
    }
    s2o_table_free(t);
    return rc;
}
