/* s2_oracle.h - TEST INFRASTRUCTURE ONLY (never linked, imported or executed by the product path).
 *
 * Plain-C CPU restatement of the strainer2 k-mer scan path, written from the behaviour of the
 * reference (citations are file:line under /root/reference/).  It exists to be the *checker* for the
 * CUDA path in tests/, in __graft_entry__.smoke() and as bench.py's cpu_baseline fallback.
 *
 * PARITY STATUS: PINNED.  tests/test_oracle_golden.py runs the unmodified reference
 * (oracle/_ref, built from /root/reference/src by oracle/Makefile) beside this restatement on
 * config #1 (test/example.sh) and on the synthetic edge-case corpus and requires byte-identical
 * output; the small reference-generated vectors are committed under tests/golden/ so that the same
 * check runs where /root/reference does not exist.
 */
#ifndef S2_ORACLE_H
#define S2_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S2O_K 31                       /* src/kmer_scrub_count.c:39, src/strain_detect.c:78 */
#define S2O_INITIAL_CAPACITY 8000000u  /* src/genome_compare.h:20 */

/* ---- primitives ------------------------------------------------------------------------- */
uint32_t s2o_djb2(const char *s);                              /* src/BIO_hash.c:208-216 (before % M) */
int      s2o_complement(int c);                                /* src/BIO_sequence.c:203-213           */
int      s2o_contains_N(const char *s);                        /* src/genome_compare.c:443-451         */
int      s2o_rc_strcmp(const char *w, int k);                  /* src/genome_compare.c:1122-1141       */
const char *s2o_orient(const char *w, char *scratch, int k);   /* src/genome_compare.c:1100-1120       */
uint64_t s2o_encode_2bit(const char *dna, int len);            /* src/up2bit.c:53-72                   */
void     s2o_decode_2bit(uint64_t v, int len, char *out);      /* src/up2bit.c:75-98                   */

/* ---- FASTA/FASTQ reader with the semantics of the reference's parser (src/kseq.h:171-211) -- */
typedef struct s2o_reader s2o_reader;
s2o_reader *s2o_reader_open(const char *path);                 /* NULL if gzopen fails */
/* returns seq length (>=0), -1 at EOF, -2 on truncated/mismatched quality.  *seq / *len expose the
 * reader's persistent sequence buffer exactly like kseq_t.seq (stale after a -1 return). */
int  s2o_reader_next(s2o_reader *r);
char *s2o_reader_seq(s2o_reader *r);
size_t s2o_reader_len(s2o_reader *r);
void s2o_reader_close(s2o_reader *r);

/* ---- string keyed table = restatement of BIO_hash (src/BIO_hash.c) ------------------------ */
typedef struct s2o_table s2o_table;
s2o_table *s2o_table_new(unsigned initial_capacity, int vec_size);
void       s2o_table_free(s2o_table *t);
unsigned   s2o_table_size(const s2o_table *t);                 /* number of keys (h->N)   */
unsigned   s2o_table_capacity(const s2o_table *t);             /* slots (h->M)            */
unsigned  *s2o_table_search(const s2o_table *t, const char *key);        /* BIO_searchHash */
void       s2o_table_add(s2o_table *t, const char *key, const unsigned *vec); /* BIO_addHashData */
/* slot-order walk (BIO_getHashKeys, src/BIO_hash.c:174-188): i in [0,size) */
const char *s2o_table_key_at(const s2o_table *t, unsigned i, unsigned **vec);

/* ---- the path itself ---------------------------------------------------------------------- */
/* GEN_hash_sequences_set_count_vec, src/genome_compare.c:967-1030. returns 0, or -1 if unreadable */
int s2o_build(s2o_table *t, const char *ref_file, int default_count, int increment, int vec_idx);
/* GEN_calculate_kmer_count, src/genome_compare.c:179-236. *n_windows (optional) += windows evaluated */
int s2o_count_file(s2o_table *t, const char *file, unsigned col, uint64_t *n_bases, uint64_t *n_windows);
/* GEN_all_kmer_counts / _skip_file, src/genome_compare.c:115-177 (skip_file may be NULL) */
int s2o_count_list(s2o_table *t, const char *list_file, const char *skip_file, unsigned col,
                   FILE *progress, FILE *err);
/* print_hash_counts, src/kmer_scrub_count.c:134-156 */
void s2o_print_counts(const s2o_table *t, int with_C, FILE *out);
/* whole kmer_scrub_count main (src/kmer_scrub_count.c:29-123) minus getopt */
int s2o_kmer_scrub_count(const char *r_file, const char *A_file, const char *B_file,
                         const char *C_file, const char *p_file, FILE *out, FILE *err);

/* hash_scrubbed_kmers, src/strain_detect.c:668-726 (diagnostics go to 'msg' = reference stdout) */
int s2o_flag_informative(s2o_table *t, const char *a_file, FILE *msg, unsigned *n_flagged);
/* quantify_hits_PE, src/strain_detect.c:387-663; text goes uncompressed to 'out'.
 * is_pe: 0 SE, 1 PE (two files), 2 interleaved.  returns 0 or the reference's exit status */
int s2o_quantify_hits(s2o_table *t, const char *pe1, const char *pe2, int is_pe,
                      unsigned genome_kmers, unsigned genome_informative, FILE *out, FILE *err);
/* whole strain_detect main for the -B batch form or the -b/-c/-t form; text output uncompressed */
int s2o_strain_detect(const char *r_file, const char *a_file, const char *B_file,
                      const char *b_file, const char *c_file, const char *type,
                      FILE *out, FILE *msg, FILE *err);

/* same with the -g background filter (src/strain_detect.c:142-143, :160-240); g_file may be NULL */
int s2o_strain_detect_g(const char *r_file, const char *a_file, const char *g_file, const char *B_file,
                        const char *b_file, const char *c_file, const char *type,
                        FILE *out, FILE *msg, FILE *err);
int s2o_background_filter(s2o_table *t, const char *background_file, unsigned num_inform, FILE *msg, FILE *err);

#ifdef __cplusplus
}
#endif
#endif
