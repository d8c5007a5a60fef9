/* strainer2_b200.h - C ABI of the B200-native k-mer scan path of strainer2.
 *
 * The reference (jeremiahfaith/strainer2) has no plugin / FFI interface: its three mains call the C
 * functions of src/genome_compare.h directly and those hand out raw pointers into BIO_hash slots, so
 * they cannot front a device-resident table.  This header is the thin replacement boundary: every
 * entry point names the reference function (file:line under /root/reference/) whose work it takes
 * over.  Plain pointers and sizes only; no CUDA or torch types appear in a signature.  Device
 * pointers are passed as const void* / void* with an explicit on_device flag.
 *
 * Error convention.  The reference prints to stderr and exit(EXIT_FAILURE)s from inside the library
 * (e.g. src/genome_compare.c:124-127).  A shared library must not kill its host process, so every
 * call returns 0 / a valid handle on success and -1 / NULL on failure with the text available from
 * s2_last_error(); the drop-in executables (kmer_scrub_count, strain_detect) turn that into the
 * reference's "message on stderr + EXIT_FAILURE".  There is NO CPU fallback: without a usable
 * sm_100 device s2_init() fails.
 *
 * Threading.  A context belongs to one GPU.  s2_batch_acquire / s2_batch_submit_* / s2_sync are
 * thread safe (reader threads fill pinned batches concurrently); everything else on a context or a
 * table must be called from one thread at a time.
 */
#ifndef STRAINER2_B200_H
#define STRAINER2_B200_H
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S2_K 31                 /* seed length, hard-coded in src/kmer_scrub_count.c:39, src/strain_detect.c:78 */
#define S2_ABI_VERSION 1

typedef struct s2_ctx s2_ctx;         /* one GPU: streams, pinned batch ring, scratch            */
typedef struct s2_table s2_table;     /* device-resident strain table = the BIO_hash replacement */

typedef struct s2_scan_stats {
    uint64_t hits;            /* windows found in the table (count[vec_column] += 1 executions)  */
    uint64_t valid_windows;   /* windows probed, i.e. not skipped by the N rule                  */
} s2_scan_stats;

/* ---------------------------------------------------------------- context ------------------- */
int         s2_abi_version(void);
const char *s2_last_error(void);                 /* text of the calling thread's last failure    */
int         s2_device_count(void);               /* -1 if the CUDA runtime is unusable           */
/* batch_bytes: capacity of each pinned/device batch buffer (0 = 64 MiB); n_lanes: batches in flight
 * (0 = 4).  Fails (NULL) when `device` is not an sm_100 GPU. */
s2_ctx     *s2_init(int device, uint64_t batch_bytes, int n_lanes);
void        s2_shutdown(s2_ctx *ctx);
int         s2_ctx_device(const s2_ctx *ctx);
int         s2_ctx_sm_count(const s2_ctx *ctx);

/* ---------------------------------------------------------------- strain table -------------- */
/* Replaces BIO_initHash(8,000,000) + GEN_hash_sequences_set_count_vec()
 * (src/kmer_scrub_count.c:87-89, src/genome_compare.c:967-1030).
 * `bases` is the reference genome as ONE flat byte stream: the records' sequence bytes exactly as the
 * parser yields them (any case), consecutive records separated by one byte that is not in ACGTacgt
 * ('\n' by convention).  Every 31-byte window made only of ACGTacgt is canonicalised and inserted;
 * column 0 of a new key starts at 1 and grows by 1 per repeat (default_count = increment = 1; the
 * strain_detect caller ignores column 0).  n_cols: 4 (kmer_scrub_count) or 6 (strain_detect).
 * load_factor: keys / slots, 0 = default 0.5.  on_device: `bases` is a device pointer. */
s2_table   *s2_table_build(s2_ctx *ctx, const void *bases, uint64_t n_bytes, int n_cols,
                           double load_factor, int on_device);
void        s2_table_free(s2_table *t);
uint64_t    s2_table_n_keys(const s2_table *t);          /* BIO_getHashSize()                     */
uint64_t    s2_table_n_slots(const s2_table *t);
uint64_t    s2_table_hbm_bytes(const s2_table *t);       /* fingerprints + keys + counters        */
uint64_t    s2_table_probe_bytes(const s2_table *t);     /* the fingerprint array only            */
/* keys[n_keys] (62-bit, A0 C1 G2 T3, first base highest), djb2[n_keys] = hashU() of each key's ASCII
 * spelling before "% M" (src/BIO_hash.c:208-216) and first_pos[n_keys] = byte offset in the build
 * stream of the window that inserted the key, all in FIRST-OCCURRENCE order = the order the reference
 * inserted them, which is all s2_roworder_emulate() needs.  Any of the three may be NULL. */
int         s2_table_export(s2_table *t, uint64_t *keys, uint32_t *djb2, uint32_t *first_pos);
/* counter column `col` in first-occurrence order (host buffers of n_keys uint32) */
int         s2_table_counts_fetch(s2_table *t, int col, uint32_t *host_out);
int         s2_table_counts_store(s2_table *t, int col, const uint32_t *host_in);
int         s2_table_counts_clear(s2_table *t, int col);
/* same, device to device, for the one collective of the path: the caller all-reduces (sum, uint32)
 * the dense n_keys vector across GPUs (NCCL) between gather and scatter.  First-occurrence order is
 * identical on every replica, whatever slot each replica's build happened to give a key. */
int         s2_table_counts_gather_dev(s2_table *t, int col, void *dev_out);
int         s2_table_counts_scatter_dev(s2_table *t, int col, const void *dev_in);
/* The collective itself for a single process driving several GPUs (what the executables do with
 * S2_GPUS > 1): tabs[i] are replicas built from the same bytes on n different GPUs; after the call
 * every replica's column `col` holds the sum over all replicas.  One ncclAllReduce(sum, uint32) over
 * NVLink on the dense first-occurrence-order vectors; libnccl.so.2 is dlopen()ed on first use. */
int         s2_tables_allreduce(s2_table **tabs, int n, int col);
/* hash_scrubbed_kmers() labelling, src/strain_detect.c:687-717: mark canonical 62-bit k-mers as
 * INFORMATIVE.  found[i] (may be NULL) = 1 if kmers[i] is a key of the table. */
int         s2_table_flag(s2_table *t, const uint64_t *kmers, uint64_t n, uint8_t *found);
/* counter column `col` for arbitrary canonical k-mers (0 for keys the table does not hold): how the
 * multi-strain batch reads one strain's rows out of the union table */
int         s2_table_counts_by_key(s2_table *t, int col, const uint64_t *kmers, uint64_t n, uint32_t *host_out);
/* background_filter() demotion, src/strain_detect.c:218-228: back to NON_INFORMATIVE */
int         s2_table_unflag(s2_table *t, const uint64_t *kmers, uint64_t n);
int         s2_table_lookup(s2_table *t, const uint64_t *kmers, uint64_t n, uint32_t *slot_out);

/* ---------------------------------------------------------------- count scan ---------------- */
/* The per-window loop of GEN_calculate_kmer_count() (src/genome_compare.c:203-229) over one batch:
 * flat byte stream as for s2_table_build (records shorter than 31 contribute no window by
 * construction).  Adds 1 to counter column `col` for every window found.  Synchronous.
 * If on_device, `bases` must be 16-byte aligned and readable up to the next multiple of 16. */
int         s2_scan_count(s2_ctx *ctx, s2_table *t, const void *bases, uint64_t n_bytes, int col,
                          int on_device, s2_scan_stats *stats);
/* enqueue-only form for device-resident batches: launches on the context's first stream and
 * returns immediately; s2_sync() waits and returns the accumulated stats. */
int         s2_scan_count_enqueue(s2_ctx *ctx, s2_table *t, const void *dev_bases, uint64_t n_bytes, int col);
/* pinned host memory for batches handed to s2_scan_count(on_device = 0) at full PCIe rate */
void       *s2_pinned_alloc(uint64_t n_bytes);
void        s2_pinned_free(void *p);
/* pipelined form: pinned double(+)-buffered batches, H2D copy and kernel overlapped on a stream per
 * lane.  acquire blocks until a lane is free and returns its pinned host buffer. */
uint8_t    *s2_batch_acquire(s2_ctx *ctx, uint64_t *capacity);
int         s2_batch_submit_count(s2_ctx *ctx, s2_table *t, uint8_t *batch, uint64_t n_bytes, int col);
int         s2_batch_release(s2_ctx *ctx, uint8_t *batch);               /* give back unused      */
/* wait for every batch in flight; returns totals accumulated since the previous s2_sync */
int         s2_sync(s2_ctx *ctx, s2_scan_stats *totals);

/* GPU-side ingest: GEN_calculate_kmer_count() for one FILE without the host inflating or
 * parsing it.  BGZF-compressed (bgzip) strict 4-line FASTQ / strict FASTA (and the same uncompressed, unless
 * S2_GPU_INGEST_PLAIN=0) is inflated by the Blackwell hardware decompression engine and split into records by kernels; every
 * chunk is proven regular on the device before its scan starts, so an irregular file is never counted.
 * Returns 0 = done, 1 = not handled (nothing was counted; use the reader + s2_batch_submit_count), -1 = error.
 * Thread safe: a context owns a small pool of ingest pipelines (S2_INGEST_PIPES, default 3) that all calling threads
 * share - a call takes a free pipeline for as long as it enqueues work; s2_shutdown() releases them.  s2_ingest_warm()
 * creates n of them ahead of the first call (device buffers + pinned staging: tens of milliseconds each).
 * Knobs: S2_INGEST_CHUNK_MB (compressed bytes per chunk, 16), S2_INGEST_TEXT_MB (text per chunk, 64). */
int         s2_ingest_count_file(s2_ctx *ctx, s2_table *t, const char *path, int col, uint64_t *bases, uint64_t *lookups);
/* the same for a file IMAGE in host memory (the bytes of a BGZF or plain FASTA / FASTQ file; pinned memory from
 * s2_pinned_alloc gives the full PCIe rate): only the compressed bytes cross PCIe.  The image must stay valid
 * until the call returns. */
int         s2_ingest_count_mem(s2_ctx *ctx, s2_table *t, const void *image, uint64_t n_bytes, int col, uint64_t *bases, uint64_t *lookups);
/* many files per call (the list loop of GEN_all_kmer_counts, src/genome_compare.c:149-177, minus the progress lines):
 * files that fit one chunk travel in groups - their texts back to back, one launch sequence per group - and the
 * verdicts are collected once at the end.  rc_each[i] = 0 done / 1 not handled (nothing of file i was counted);
 * bases / lookups = totals of the handled files.  Returns 0, or -1 on error. */
int         s2_ingest_count_mem_batch(s2_ctx *ctx, s2_table *t, const void *const *images, const uint64_t *n_bytes, int n, int col,
                                      int *rc_each, uint64_t *bases, uint64_t *lookups);
int         s2_ingest_count_files(s2_ctx *ctx, s2_table *t, const char *const *paths, int n, int col,
                                  int *rc_each, uint64_t *bases, uint64_t *lookups);
/* asynchronous form of the two calls above: submit returns as soon as every file that fits a chunk is enqueued (file
 * images must stay valid until the wait), wait blocks until the job's files are counted or handed back, fills rc_each /
 * bases / lookups as above and releases the job.  Submitting the next job before waiting for the previous one hides the
 * fill and drain of the copy -> inflate -> kernels pipeline.  A job is waited for on the thread that submitted it, jobs
 * of one thread in the order they were submitted.  submit returns NULL on error. */
typedef struct s2_ingest_job s2_ingest_job;
s2_ingest_job *s2_ingest_submit_mem_batch(s2_ctx *ctx, s2_table *t, const void *const *images, const uint64_t *n_bytes, int n, int col);
s2_ingest_job *s2_ingest_submit_files(s2_ctx *ctx, s2_table *t, const char *const *paths, int n, int col);
int         s2_ingest_wait(s2_ingest_job *job, int *rc_each, uint64_t *bases, uint64_t *lookups);
/* the detect form: pass 1 of quantify_hits_PE for every read of one file.  len / hits / inf are per record in
 * file order (ALL records, also those shorter than 31, which the pairing loop needs); inf_* list the informative
 * windows sorted by (record, offset) with their canonical k-mer.  The arrays live in one pinned host buffer owned by the library
 * until s2_ingest_detect_free gives it back (do not free() them).  FASTQ (four lines
 * per record) and FASTA reads (two lines per record: test/target_metagenomes.txt), uncompressed, BGZF or ordinary .gz. */
typedef struct s2_ingest_detect_result {
    uint64_t n_records, n_inf, bases;
    uint32_t *len, *hits, *inf;
    uint32_t *inf_rec, *inf_off;
    uint64_t *inf_kmer;
    /* 1: two-line FASTA reads, 0: FASTQ.  The pairing loop needs it: after the last record of a FASTQ file the parser's
     * sequence length keeps its last value (src/kseq.h:174-177 returns before the reset), after a FASTA file it is 0
     * (last_char stays '>', so the reset at :179 runs first) - src/strain_detect.c:497 tests that length */
    uint32_t fasta, pad_;
} s2_ingest_detect_result;
int         s2_ingest_detect_file(s2_ctx *ctx, s2_table *t, const char *path, s2_ingest_detect_result *out);
void        s2_ingest_detect_free(s2_ingest_detect_result *r);
void        s2_ingest_thread_cleanup(void);               /* no-op since the pipelines belong to the context */
int         s2_ingest_warm(s2_ctx *ctx, int n_pipes);
/* the same for the gunzip stage of those pipelines (ordinary .gz inputs): its symbol area is 1 MB per 32 KB of compressed
 * bytes, and cudaMalloc maps about 10 GB/s - worth doing beside the table build when the first listed file is a .gz */
int         s2_ingest_warm_gz(s2_ctx *ctx, int n_pipes, uint64_t batch_comp_bytes);
void        s2_ingest_reset(s2_ctx *ctx);                 /* releases the context's pipelines (no job may be in flight); the next call makes new ones */
/* 1 after the hardware decompression engine met a DEFLATE block it cannot decode (a damaged BGZF member): the engine
 * reports that as a sticky launch failure (measured, profiles/r1s_hw_decompression_error_probe.txt), the CUDA context
 * of the process is lost and every later call fails.  The executables exit with an error that says so - as they do
 * when zlib finds the damage on the host path (s2_reader_damaged); the reference hangs on such a file. */
int         s2_ingest_engine_failed(void);

/* ---------------------------------------------------------------- gzip output --------------- */
/* The kmer_hits stream of strain_detect (gzopen(outfile, "wb9") + gzprintf, src/strain_detect.c:299,567,608).
 * threads == 0: one zlib stream at level 9 - the file is byte-identical to the reference's.  threads > 0 (SURVEY 8f
 * rank 2): 256 KB blocks deflated independently at level 9 on that many threads (each primed with the 32 KB before
 * it) and concatenated into one gzip member, as pigz does: the same text for any gunzip, different compressed bytes. */
typedef struct s2_gz_writer s2_gz_writer;
s2_gz_writer *s2_gz_writer_open(const char *path, int threads);
int         s2_gz_writer_write(s2_gz_writer *w, const void *data, uint64_t n);
int         s2_gz_writer_close(s2_gz_writer *w);

/* ---------------------------------------------------------------- scrub filter -------------- */
/* The selection steps of scripts/kmer_scrub_filter.py on table columns.
 * joint scrub (:88-143): rows are ranked by max(pan / pan_sum, meta / meta_sum) (IEEE doubles, as the script computes
 * them), descending, ties in table order; the top n_scrub ALIVE rows go.  keep_out[i] = 1 for the survivors (dead
 * rows - those the drug scrub removed, alive[i] = 0 - are never kept).  0 <= n_scrub <= number of alive rows. */
int         s2_scrub_joint(s2_ctx *ctx, const uint64_t *pan, const uint64_t *meta, const uint8_t *alive, uint64_t n,
                           uint64_t pan_sum, uint64_t meta_sum, uint64_t n_scrub, uint8_t *keep_out);
/* independent scrub (:31-58): hist[v] = entries equal to v for v < 65536, hist[65536] = entries >= 65536;
 * count_above = entries > t (for thresholds beyond the histogram) */
int         s2_scrub_histogram(s2_ctx *ctx, const uint64_t *vals, uint64_t n, uint64_t *hist65537);
int         s2_scrub_count_above(s2_ctx *ctx, const uint64_t *vals, uint64_t n, uint64_t t, uint64_t *count);
/* str(float) of Python 3 (shortest round-trip digits, exponent notation below 1e-4 and from 1e16): the script prints
 * such numbers on stdout / stderr.  out must hold 32 bytes. */
void        s2_py_float_repr(double x, char *out);

/* ---------------------------------------------------------------- detect scan --------------- */
/* Pass 1 of quantify_hits_PE() (src/strain_detect.c:465-491, :514-539) for every record of a batch,
 * plus the positions pass 2 (:554-623) will print.  rec_off[n_rec+1] are the ascending byte offsets of
 * the records inside `bases` (rec_off[n_rec] = n_bytes; the separator byte belongs to the record
 * before it).  Outputs (host): read_hits[n_rec], read_inf[n_rec]; inf_pos[<= inf_cap] = byte offsets
 * of windows whose k-mer is INFORMATIVE, ascending; *n_inf = how many there were (may exceed inf_cap:
 * call again with a larger buffer). */
int         s2_scan_detect(s2_ctx *ctx, s2_table *t, const void *bases, uint64_t n_bytes,
                           const uint64_t *rec_off, uint32_t n_rec, uint32_t *read_hits,
                           uint32_t *read_inf, uint64_t *inf_pos, uint64_t inf_cap, uint64_t *n_inf,
                           int on_device, s2_scan_stats *stats);

/* ---------------------------------------------------------------- timing -------------------- */
/* CUDA-event time (ms) and launch count of the scan kernels since the last reset, measured on the
 * streams they were launched on. */
int         s2_kernel_time(s2_ctx *ctx, double *ms, uint64_t *launches, int reset);
/* four user events on the stream the enqueue-only scans run on: bracket a timed region on the device */
int         s2_event_record(s2_ctx *ctx, int which);
int         s2_event_elapsed_ms(s2_ctx *ctx, int from, int to, double *ms);

/* ---------------------------------------------------------------- tuning -------------------- */
/* The scan kernel is compiled in several shapes (loads in flight per lane, CTAs per SM, software
 * pipelining).  v < 0 returns the number of shapes; otherwise selects shape v for subsequent scans
 * (process-wide).  The default is the shape DESIGN.md names; S2_SCAN_VARIANT overrides it. */
int         s2_tune_scan_variant(s2_ctx *ctx, int v);
const char *s2_tune_scan_variant_name(int v);

/* ---------------------------------------------------------------- codecs -------------------- */
/* bit-compatible with encode_DNA_2_bit / decode_DNA_2_bit (src/up2bit.c:53-98): A0 C1 T2 G3 */
uint64_t    s2_encode_2bit(const char *dna, int len);
void        s2_decode_2bit(uint64_t v, int len, char *out);
/* device 2-bit pack kernel (order-preserving code A0 C1 G2 T3 + validity), results to host:
 * words[ceil(n/16)], masks[ceil(n/16)] */
int         s2_pack_2bit(s2_ctx *ctx, const void *bases, uint64_t n_bytes, int on_device,
                         uint32_t *words, uint16_t *masks);
/* canonical 62-bit k-mer of 31 ASCII bases (orient_string, src/genome_compare.c:1100-1120);
 * returns 0 and sets *out, or -1 if a byte is not in ACGTacgt */
int         s2_kmer_from_ascii(const char *s, uint64_t *out);
void        s2_kmer_to_ascii(uint64_t kmer, char *out32);      /* 31 letters + NUL */

/* ---------------------------------------------------------------- host helpers -------------- */
/* Replays BIO_hash's slot placement (src/BIO_hash.c:129-139 insert + doubling at N++ >= M/2,
 * :39-61 re-insertion in old-slot order, :174-188 slot-order walk) on the djb2 values of the keys in
 * insertion order; order_out[i] = insertion index of the i-th emitted row.  initial_capacity = 0
 * means the reference's DEFAULT_GENOME_HASH_SIZE (8,000,000, src/genome_compare.h:20). */
int         s2_roworder_emulate(const uint32_t *djb2, uint64_t n, uint32_t initial_capacity,
                                uint32_t *order_out, uint32_t *final_capacity);
/* print_hash_counts() (src/kmer_scrub_count.c:134-156): header + one row per key in `order`.
 * cols[c] are first-occurrence-order counter columns; n_print_cols = 3 or 4 (-C given). */
int         s2_format_count_table(FILE *out, const uint64_t *keys, const uint32_t *order, uint64_t n,
                                  const uint32_t *const *cols, int n_print_cols, int n_threads);

/* the same table formatted ON THE DEVICE straight from a strain table (keys, counters and row text never
 * exist on the host except as the finished bytes): header + one row per key in `order`. */
int         s2_table_format(s2_table *t, const uint32_t *order, int n_print_cols, FILE *out);

/* FASTA/FASTQ (plain or gzip) reader with the record semantics of the parser the reference vendors
 * (src/kseq.h:171-211).  s2_reader_next returns the sequence length, -1 at end of file, -2 on a
 * truncated / mismatched quality string; the sequence stays valid until the next call. */
typedef struct s2_reader s2_reader;
s2_reader  *s2_reader_open(const char *path);
int64_t     s2_reader_next(s2_reader *r, const char **seq);
uint64_t    s2_reader_len(const s2_reader *r);      /* kseq's seq.l, stale after EOF like the original */
/* 1 once gzread has reported damaged data (corrupt DEFLATE payload, CRC mismatch).  The reference never returns from
 * such a file: kseq (src/kseq.h:72,:99) takes only a 0 from gzread for the end and re-reads the error for ever
 * (measured: 100 % CPU until killed).  Here the stream ends at the damage and the executables exit with an error. */
int         s2_reader_damaged(const s2_reader *r);
void        s2_reader_close(s2_reader *r);

/* whole programs, argv-compatible with the reference executables */
int         s2_kmer_scrub_count_main(int argc, char **argv);   /* src/kmer_scrub_count.c:29-123 */
int         s2_kmer_scrub_filter_main(int argc, char **argv);  /* scripts/kmer_scrub_filter.py:146-229 */
int         s2_strain_detect_main(int argc, char **argv);      /* src/strain_detect.c:61-158    */
/* many strains against the same -A/-B/-C lists in ONE pass over the inputs (union table); every output
 * table is byte-identical to kmer_scrub_count run on that strain alone.  No reference counterpart:
 * README.md:47 runs one process per strain. */
int         s2_kmer_scrub_count_batch_main(int argc, char **argv);

#ifdef __cplusplus
}
#endif
#endif
