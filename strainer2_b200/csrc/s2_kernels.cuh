// s2_kernels.cuh - launch interfaces of the sm_100a kernels (implemented in s2_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// Device view of one strain table.  Layout in HBM (DESIGN.md "data layout"):
//   fp     : n_buckets x 16 x uint16  - 32-byte fingerprint buckets, the only array a miss touches
//   keys   : n_slots  x uint64        - canonical 62-bit k-mer (+ informative flag in bit 63), EMPTY = ~0
//   counts : n_cols x n_slots x uint32 (column major) - what BIO_hash hangs off each key as unsigned[4|6]
struct S2TableView {
    uint16_t *fp;
    uint64_t *keys;
    uint32_t *counts;
    uint32_t n_buckets;
    uint64_t n_slots;
    int n_cols;
};

// a batch that was produced on the device (GPU ingest): its length lives there, together with a veto (the ingest
// kernels found the text irregular: count nothing) and the increment (1, or 0xFFFFFFFF = -1 when a file that turned
// out irregular after some of its chunks were counted is replayed to take its contribution back out)
struct S2DevBatch {
    unsigned long long n_bytes;
    unsigned int skip;
    unsigned int inc;
};

struct S2DetectOut {
    const uint64_t *rec_off;     // n_rec + 1 ascending byte offsets of the records inside the batch
    uint32_t n_rec;
    const uint32_t *n_rec_dev;   // when set, the record count is read from device memory (the batch is an S2DevBatch)
    uint32_t *read_hits;         // per record: table hits                 (src/strain_detect.c:481)
    uint32_t *read_inf;          // per record: informative hits           (src/strain_detect.c:482-483)
    uint64_t *inf_pos;           // batch byte offsets of informative windows (unordered)
    unsigned long long *inf_count;
    uint64_t inf_cap;
};

enum { S2_MODE_COUNT = 0, S2_MODE_DETECT = 1 };

// scan stats written by the scan kernels: [0] table hits, [1] valid (non-N) windows probed
void s2_launch_scan_count(const uint8_t *bases, uint64_t n_bytes, const S2TableView &t, int col,
                          unsigned long long *stats, int grid_blocks, cudaStream_t stream);
void s2_launch_scan_detect(const uint8_t *bases, uint64_t n_bytes, const S2TableView &t,
                           const S2DetectOut &out, unsigned long long *stats, int grid_blocks,
                           cudaStream_t stream);
void s2_launch_scan_count_dev(const uint8_t *bases, const S2DevBatch *dev, const S2TableView &t, int col,
                              unsigned long long *stats, int grid_blocks, cudaStream_t stream);
void s2_launch_scan_detect_dev(const uint8_t *bases, const S2DevBatch *dev, const S2TableView &t, const S2DetectOut &out,
                               unsigned long long *stats, int grid_blocks, cudaStream_t stream);
int  s2_scan_blocks_per_sm(int mode);
// two-phase (radix partition, then per-partition probe) count scan for tables larger than L2
#define S2_NPART_LOG2 7
#define S2_NPART (1 << S2_NPART_LOG2)
void s2_launch_scan_count_partitioned(const uint8_t *bases, uint64_t n_bytes, const S2TableView &t, int col,
                                      unsigned long long *stats, uint64_t *part_pool, uint64_t region_cap,
                                      unsigned long long *cursor, uint32_t *overflow, int n_sm, int grid_blocks,
                                      cudaStream_t stream, const S2DevBatch *dev = nullptr);
// kernel shape selection (sweep tool / S2_SCAN_VARIANT); see s2_kernels.cu
int  s2_scan_variant_count(void);
const char *s2_scan_variant_name(int v);
int  s2_scan_variant_get(void);
int  s2_scan_variant_set(int v);

// table build (one-off, not the hot path)
void s2_launch_build_insert(const uint8_t *bases, uint64_t n_bytes, const S2TableView &t,
                            uint32_t *first_pos, uint32_t *slot_of_pos, cudaStream_t stream);
// number of distinct keys is returned through *d_n_keys (device); rank_slot must hold >= n windows
void s2_launch_build_rank(uint64_t n_bytes, const uint32_t *first_pos, const uint32_t *slot_of_pos,
                          uint32_t *block_sums, uint32_t n_blocks, uint32_t *rank_slot, uint32_t *rank_pos,
                          unsigned long long *d_n_keys, cudaStream_t stream);
void s2_launch_export(const S2TableView &t, const uint32_t *rank_slot, uint64_t n_keys,
                      uint64_t *keys_out, uint32_t *djb2_out, cudaStream_t stream);
void s2_launch_gather_counts(const S2TableView &t, int col, const uint32_t *rank_slot, uint64_t n_keys,
                             uint32_t *out, cudaStream_t stream);
void s2_launch_scatter_counts(const S2TableView &t, int col, const uint32_t *rank_slot, uint64_t n_keys,
                              const uint32_t *in, cudaStream_t stream);
// all-reduce over peer memory for a process that drives several GPUs: dense vectors of all replicas -> sum -> own column
#define S2_MAX_PEERS 16
struct S2PeerVecs { const uint32_t *v[S2_MAX_PEERS]; int n; };
void s2_launch_peer_sum_scatter(const S2TableView &t, int col, const uint32_t *rank_slot, uint64_t n_keys, const S2PeerVecs &pv,
                                cudaStream_t stream);
void s2_launch_flag(const S2TableView &t, const uint64_t *kmers, uint64_t n, uint8_t *found, int set,
                    cudaStream_t stream);
void s2_launch_lookup(const S2TableView &t, const uint64_t *kmers, uint64_t n, uint32_t *slot_out,
                      cudaStream_t stream);
void s2_launch_counts_by_key(const S2TableView &t, int col, const uint64_t *kmers, uint64_t n, uint32_t *out, cudaStream_t stream);
void s2_launch_pack(const uint8_t *bases, uint64_t n_bytes, uint32_t *words, uint16_t *masks,
                    cudaStream_t stream);
void s2_launch_fill_u32(uint32_t *p, uint64_t n, uint32_t v, cudaStream_t stream);

// device-side count-table formatting: phase 0 = row lengths + offsets (+ total bytes), phase 1 = write the text
void s2_launch_format(const uint64_t *keys, const uint32_t *order, uint64_t n, const uint32_t *const *cols, int n_cols,
                      unsigned long long *block_sums, unsigned long long *d_total, char *out, int phase, cudaStream_t stream);
