// s2_kernels.cu - hand-written sm_100a kernels for the strainer2 k-mer scan path.
//
// What they replace in the reference (paths under /root/reference/):
//   scan kernels   : the per-window loop of GEN_calculate_kmer_count  src/genome_compare.c:213-229
//                    and pass 1 of quantify_hits_PE                    src/strain_detect.c:465-491,514-539
//   build kernels  : GEN_hash_sequences_set_count_vec                  src/genome_compare.c:967-1030
//   probe          : BIO_searchHash + hashU                            src/BIO_hash.c:161-172,208-216
//
// This path is hashing + random access over bytes: HBM/L2-bound integer work, so no tensor cores.
// Design: one warp owns a 512-base tile of the flat ASCII stream; every lane packs
// 16 bases with one 128-bit load, neighbours' packed words arrive by warp shuffle (the 30-base halo),
// forward and reverse-complement k-mers are funnel-shift extractions from three packed words, the
// canonical k-mer is an integer max, and the probe is ONE 256-bit load of a 32-byte fingerprint
// bucket tested with 8 HSET2 (half2 ==) instructions.  Only fingerprint matches (true hits and
// ~2^-15 false ones) touch the key array and the counters, in a compacted slow path.
#include "s2_kernels.cuh"
#include "s2_kmer.cuh"
#include <cuda_fp16.h>
#include <cstdlib>

#define S2_THREADS 256
#define S2_NONE 0xFFFFFFFFu

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ld_bucket256(const uint16_t *fp, uint32_t bucket, uint32_t (&x)[8])
{
    const uint16_t *p = fp + (uint64_t)bucket * S2_BUCKET_SLOTS;
    // one 32-byte sector, read-only path, do not allocate in L1 (the table never fits, the input does)
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_last.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]), "=r"(x[4]), "=r"(x[5]), "=r"(x[6]), "=r"(x[7])
                 : "l"(p));
}

// the same without an L2 policy (tables that are swept once per batch, slice by slice: nothing should stay behind)
__device__ __forceinline__ void ld_bucket256_plain(const uint16_t *fp, uint32_t bucket, uint32_t (&x)[8])
{
    const uint16_t *p = fp + (uint64_t)bucket * S2_BUCKET_SLOTS;
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]), "=r"(x[4]), "=r"(x[5]), "=r"(x[6]), "=r"(x[7])
                 : "l"(p));
}

// 16 fingerprints against one: 8 HSET2.  Bit i of the result = low half of word i matched (slot 2i),
// bit 16+i = high half of word i matched (slot 2i+1); 0 = no match.
__device__ __forceinline__ uint32_t fp_match_bits(const uint32_t (&x)[8], uint32_t fp2)
{
    const __half2 f = *reinterpret_cast<const __half2 *>(&fp2);
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) m |= __heq2_mask(*reinterpret_cast<const __half2 *>(&x[i]), f) & (0x00010001u << i);
    return m;
}

// 16 bases of the stream -> packed word + validity mask; bytes at or beyond n_bytes are invalid.
// The buffer must be 16-byte aligned and readable up to the next multiple of 16 after n_bytes.
__device__ __forceinline__ void load_chunk(const uint8_t *__restrict__ bases, uint64_t n_bytes, uint64_t chunk,
                                           uint32_t &w, uint32_t &m)
{
    const uint64_t s = chunk * 16;
    w = 0; m = 0;
    if (s < n_bytes) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(bases) + chunk);
        s2_pack16(v.x, v.y, v.z, v.w, &w, &m);
        const uint64_t rem = n_bytes - s;
        if (rem < 16) m &= (0xFFFFu << (16 - (unsigned)rem)) & 0xFFFFu;
    }
}

// One warp tile = 32 chunks = 512 bases.  After this call lane L holds the packed words / masks of
// chunks L, L+1, L+2 of the tile (48 bases = its 16 window starts + the 30-base halo).  The two chunks
// past the tile are loaded by lanes 0 and 1 and handed to lanes 30/31 by shuffle.
__device__ __forceinline__ void load_tile(const uint8_t *__restrict__ bases, uint64_t n_bytes, uint64_t tile,
                                          int lane, uint32_t &w0, uint32_t &w1, uint32_t &w2,
                                          uint32_t &m0, uint32_t &m1, uint32_t &m2)
{
    const uint64_t chunk0 = tile * 32;
    load_chunk(bases, n_bytes, chunk0 + lane, w0, m0);
    uint32_t wx = 0, mx = 0;
    if (lane < 2) load_chunk(bases, n_bytes, chunk0 + 32 + lane, wx, mx);
    const uint32_t a1 = __shfl_sync(0xFFFFFFFFu, w0, (lane + 1) & 31), b1 = __shfl_sync(0xFFFFFFFFu, wx, (lane + 1) & 31);
    const uint32_t a2 = __shfl_sync(0xFFFFFFFFu, w0, (lane + 2) & 31), b2 = __shfl_sync(0xFFFFFFFFu, wx, (lane + 2) & 31);
    const uint32_t c1 = __shfl_sync(0xFFFFFFFFu, m0, (lane + 1) & 31), d1 = __shfl_sync(0xFFFFFFFFu, mx, (lane + 1) & 31);
    const uint32_t c2 = __shfl_sync(0xFFFFFFFFu, m0, (lane + 2) & 31), d2 = __shfl_sync(0xFFFFFFFFu, mx, (lane + 2) & 31);
    w1 = lane < 31 ? a1 : b1;  m1 = lane < 31 ? c1 : d1;
    w2 = lane < 30 ? a2 : b2;  m2 = lane < 30 ? c2 : d2;
}

// canonical k-mer of window j (0..15) of the lane's 48 bases; r0:r1:r2 is the reverse complement
// of w0:w1:w2, so the window's reverse complement starts at base 17-j of it.
__device__ __forceinline__ uint64_t window_canon(uint32_t w0, uint32_t w1, uint32_t w2,
                                                 uint32_t r0, uint32_t r1, uint32_t r2, unsigned j)
{
    return s2_canonical(s2_extract31(w0, w1, w2, j), s2_extract31(r0, r1, r2, 17u - j));
}

// exact probe (slow path, flagging, tests): fingerprint bucket -> key compare -> slot
__device__ __forceinline__ bool probe_exact(const S2TableView &t, uint64_t canon, uint32_t &slot_out, uint64_t &key_out)
{
    const s2_hash_t hh = s2_hash(canon);
    uint32_t b = s2_bucket_of(hh.h, t.n_buckets);
    for (;;) {
        uint32_t x[8];
        ld_bucket256(t.fp, b, x);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if ((x[i] & 0xFFFFu) == hh.fp) {
                const uint32_t slot = b * S2_BUCKET_SLOTS + 2 * i;
                const uint64_t key = t.keys[slot];
                if ((key & S2_KMER_MASK) == canon) { slot_out = slot; key_out = key; return true; }
            }
            if ((x[i] >> 16) == hh.fp) {
                const uint32_t slot = b * S2_BUCKET_SLOTS + 2 * i + 1;
                const uint64_t key = t.keys[slot];
                if ((key & S2_KMER_MASK) == canon) { slot_out = slot; key_out = key; return true; }
            }
        }
        if ((x[7] >> 16) == 0) return false;          // slots fill in order: last slot empty => bucket not full
        b = (b + 1 == t.n_buckets) ? 0 : b + 1;
    }
}

// ------------------------------------------------------------------------------------------------
// scan + probe + count   (the hot kernel)
// ------------------------------------------------------------------------------------------------
// Template knobs (the shipped configuration is chosen in s2_scan_config(); the others exist so that
// the sweep tool can measure them on the same binary):
//   G     windows whose bucket loads are issued together (independent 256-bit loads in flight per lane)
//   MINB  __launch_bounds__ minimum CTAs per SM (register budget 65536 / (256 * MINB))
//   PIPE  software pipelining: bucket loads of group g+1 and the next tile's bases are issued before
//         group g is consumed
// Slow path: windows whose fingerprint matched (or whose bucket is full) are pushed onto a per-warp
// shared-memory queue and resolved 32 at a time with every lane busy, instead of diverging per lane.
#define S2_WARPS (S2_THREADS / 32)
#define S2_QCAP 64

// raw 16 input bytes with an L2 evict-first policy: the stream is read once, the table must stay
__device__ __forceinline__ uint4 ld_stream16(const uint8_t *__restrict__ bases, uint64_t n_bytes, uint64_t chunk, uint64_t policy)
{
    uint4 v = make_uint4(0, 0, 0, 0);
    if (chunk * 16 < n_bytes)
        asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
            : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
            : "l"(reinterpret_cast<const uint4 *>(bases) + chunk), "l"(policy));
    return v;
}

__device__ __forceinline__ void pack_chunk(const uint4 &v, uint64_t n_bytes, uint64_t chunk, uint32_t &w, uint32_t &m)
{
    const uint64_t s = chunk * 16;
    s2_pack16(v.x, v.y, v.z, v.w, &w, &m);
    if (s >= n_bytes) m = 0;
    else if (n_bytes - s < 16) m &= (0xFFFFu << (16 - (unsigned)(n_bytes - s))) & 0xFFFFu;
}

template <int G>
__device__ __forceinline__ void issue_group(const S2TableView &t, uint32_t w0, uint32_t w1, uint32_t w2,
                                            uint32_t r0, uint32_t r1, uint32_t r2, uint32_t vmask, uint32_t dummy_bucket,
                                            int g, uint32_t (&x)[G][8], uint32_t (&fp2)[G])
{
#pragma unroll
    for (int u = 0; u < G; ++u) {
        const unsigned j = G * g + u;
        const s2_hash_t hh = s2_hash(window_canon(w0, w1, w2, r0, r1, r2, j));
        fp2[u] = hh.fp * 0x00010001u;
        // windows broken by N / a record boundary probe the warp's dummy bucket: all such lanes of the
        // warp ask for the same sector (one extra L1 wavefront), different warps ask different L2
        // slices (a single shared dummy sector becomes an L2 hot spot: 20 % of the windows of 150-base
        // reads are broken), and the result is masked by vmask.  (A predicated load makes ptxas keep
        // all previous bucket registers alive and spill.)
        ld_bucket256(t.fp, ((vmask >> j) & 1u) ? s2_bucket_of(hh.h, t.n_buckets) : dummy_bucket, x[u]);
    }
}

template <int G>
__device__ __forceinline__ void consume_group(uint32_t vmask, int g, const uint32_t (&x)[G][8],
                                              const uint32_t (&fp2)[G], uint32_t &cand, uint32_t &cp_lo, uint32_t &cp_hi)
{
#pragma unroll
    for (int u = 0; u < G; ++u) {
        const unsigned j = G * g + u;
        const uint32_t m = fp_match_bits(x[u], fp2[u]);
        const bool full = x[u][7] > 0xFFFFu;
        // branch-free: which of the 16 slots matched (any one; 4 bits per window) so that the slow path
        // can go straight to the key.  A full bucket without a match records garbage and falls back.
        const uint32_t b = 31u - (uint32_t)__clz(m);
        const uint32_t slot4 = ((b & 15u) << 1) | ((b >> 4) & 1u);
        if (((vmask >> j) & 1u) && (m != 0 || full)) {
            cand |= 1u << j;
            if (j < 8) cp_lo |= slot4 << (4 * j); else cp_hi |= slot4 << (4 * (j - 8));
        }
    }
}

// resolve `count` (<= 32) queued windows from the top of the warp's queue: the recorded slot is checked
// against the key array first (one 8-byte load); anything else (false fingerprint match, second match
// in the bucket, overflowed bucket) goes through the exact probe.
template <int MODE>
__device__ __forceinline__ void drain_queue(const S2TableView &t, const uint64_t *q_canon, const uint32_t *q_slot,
                                            const uint64_t *q_pos, uint32_t &qlen, uint32_t count, int lane,
                                            uint32_t *__restrict__ counts_col, const S2DetectOut &dout, uint32_t &n_hits, uint32_t inc)
{
    const uint32_t base = qlen - count;
    if ((uint32_t)lane < count) {
        const uint64_t canon = q_canon[base + lane];
        uint32_t slot = q_slot[base + lane];
        const uint64_t pos = MODE == S2_MODE_DETECT ? q_pos[base + lane] : 0;
        uint64_t key = t.keys[slot];
        bool hit = (key & S2_KMER_MASK) == canon && key != S2_EMPTY_KEY;
        if (!hit) hit = probe_exact(t, canon, slot, key);
        if (hit) {
            ++n_hits;
            if (MODE == S2_MODE_COUNT) {
                atomicAdd(&counts_col[slot], inc);                    // count[vec_column] += 1
            } else {
                uint32_t lo = 0, hi = dout.n_rec;                     // record r: rec_off[r] <= pos < rec_off[r+1]
                while (hi - lo > 1) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (dout.rec_off[mid] <= pos) lo = mid; else hi = mid;
                }
                atomicAdd(&dout.read_hits[lo], 1u);
                if (key & S2_INFORMATIVE_BIT) {
                    atomicAdd(&dout.read_inf[lo], 1u);
                    const unsigned long long idx = atomicAdd(dout.inf_count, 1ull);
                    if (idx < dout.inf_cap) dout.inf_pos[idx] = pos;
                }
            }
        }
    }
    qlen = base;
    __syncwarp();
}

template <int MODE, int G, int MINB, bool PIPE>
__global__ void __launch_bounds__(S2_THREADS, MINB)
s2_scan_kernel(const uint8_t *__restrict__ bases, uint64_t n_bytes, S2TableView t,
               uint32_t *__restrict__ counts_col, S2DetectOut dout, unsigned long long *__restrict__ stats,
               const uint32_t *__restrict__ run_if, const S2DevBatch *__restrict__ dev)
{
    constexpr int NG = 16 / G;
    if (run_if && *run_if == 0) return;            // fallback launch after a partition overflow: normally a no-op
    uint32_t inc = 1u;
    if (dev) {                                     // batch produced on the device (GPU ingest): length, veto and sign live there
        if (dev->skip) return;
        n_bytes = dev->n_bytes;
        inc = dev->inc;
    }
    if (MODE == S2_MODE_DETECT && dout.n_rec_dev) dout.n_rec = *dout.n_rec_dev;
    __shared__ uint64_t q_canon_s[S2_WARPS][S2_QCAP];
    __shared__ uint32_t q_slot_s[S2_WARPS][S2_QCAP];
    __shared__ uint64_t q_pos_s[MODE == S2_MODE_DETECT ? S2_WARPS : 1][MODE == S2_MODE_DETECT ? S2_QCAP : 1];

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint64_t *q_canon = q_canon_s[wid];
    uint32_t *q_slot = q_slot_s[wid];
    uint64_t *q_pos = q_pos_s[MODE == S2_MODE_DETECT ? wid : 0];
    const uint64_t gwarp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t n_tiles = (n_bytes + 511) / 512;
    uint32_t n_hits = 0, n_valid = 0, qlen = 0;
    const uint32_t dummy_bucket = s2_bucket_of((uint32_t)gwarp * 0x9E3779B1u, t.n_buckets);
    uint64_t policy;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));

    uint64_t tile = gwarp;
    uint4 raw = make_uint4(0, 0, 0, 0), rawx = make_uint4(0, 0, 0, 0);
    if (tile < n_tiles) {
        raw = ld_stream16(bases, n_bytes, tile * 32 + lane, policy);
        if (lane < 2) rawx = ld_stream16(bases, n_bytes, tile * 32 + 32 + lane, policy);
    }
    for (; tile < n_tiles; tile += n_warps) {
        // ---- this tile's 48 packed bases per lane; (PIPE) the next tile's bytes start moving now ------
        uint32_t w0, w1, w2, m0, m1, m2;
        {
            uint32_t wx = 0, mx = 0;
            pack_chunk(raw, n_bytes, tile * 32 + lane, w0, m0);
            if (lane < 2) pack_chunk(rawx, n_bytes, tile * 32 + 32 + lane, wx, mx);
            if (PIPE) {
                const uint64_t nt = tile + n_warps;
                raw = make_uint4(0, 0, 0, 0); rawx = raw;
                if (nt < n_tiles) {
                    raw = ld_stream16(bases, n_bytes, nt * 32 + lane, policy);
                    if (lane < 2) rawx = ld_stream16(bases, n_bytes, nt * 32 + 32 + lane, policy);
                }
            }
            const uint32_t a1 = __shfl_sync(0xFFFFFFFFu, w0, (lane + 1) & 31), b1 = __shfl_sync(0xFFFFFFFFu, wx, (lane + 1) & 31);
            const uint32_t a2 = __shfl_sync(0xFFFFFFFFu, w0, (lane + 2) & 31), b2 = __shfl_sync(0xFFFFFFFFu, wx, (lane + 2) & 31);
            const uint32_t c1 = __shfl_sync(0xFFFFFFFFu, m0, (lane + 1) & 31), d1 = __shfl_sync(0xFFFFFFFFu, mx, (lane + 1) & 31);
            const uint32_t c2 = __shfl_sync(0xFFFFFFFFu, m0, (lane + 2) & 31), d2 = __shfl_sync(0xFFFFFFFFu, mx, (lane + 2) & 31);
            w1 = lane < 31 ? a1 : b1;  m1 = lane < 31 ? c1 : d1;
            w2 = lane < 30 ? a2 : b2;  m2 = lane < 30 ? c2 : d2;
        }
        uint32_t r0 = s2_rc16(w2), r1 = s2_rc16(w1), r2 = s2_rc16(w0);

        // ---- fast path: 16 windows per lane, G bucket loads in flight (2G when pipelined) -------------
        // The whole tile is one basic block; without a fence the compiler hoists the hashing of all 16
        // windows above the first load and spills.  An empty asm that "rewrites" the packed words pins
        // each group's arithmetic between its neighbours (no instruction is emitted).
#define S2_GROUP_FENCE() asm volatile("" : "+r"(w0), "+r"(w1), "+r"(w2), "+r"(r0), "+r"(r1), "+r"(r2), "+r"(vmask))
        uint32_t vmask = 0, cand = 0;
        uint32_t cp_lo = 0, cp_hi = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) vmask |= (s2_window_valid(m0, m1, m2, j) ? 1u : 0u) << j;
        {
            uint32_t xa[G][8], xb[PIPE ? G : 1][8], fa[G], fb[PIPE ? G : 1];
            if (PIPE) {
                issue_group<G>(t, w0, w1, w2, r0, r1, r2, vmask, dummy_bucket, 0, xa, fa);
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    if (g & 1) {
                        if (g + 1 < NG) issue_group<G>(t, w0, w1, w2, r0, r1, r2, vmask, dummy_bucket, g + 1, xa, fa);
                        consume_group<G>(vmask, g, reinterpret_cast<uint32_t (&)[G][8]>(xb), reinterpret_cast<uint32_t (&)[G]>(fb), cand, cp_lo, cp_hi);
                        S2_GROUP_FENCE();
                    } else {
                        if (g + 1 < NG) issue_group<G>(t, w0, w1, w2, r0, r1, r2, vmask, dummy_bucket, g + 1, reinterpret_cast<uint32_t (&)[G][8]>(xb), reinterpret_cast<uint32_t (&)[G]>(fb));
                        consume_group<G>(vmask, g, xa, fa, cand, cp_lo, cp_hi);
                        S2_GROUP_FENCE();
                    }
                }
            } else {
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    issue_group<G>(t, w0, w1, w2, r0, r1, r2, vmask, dummy_bucket, g, xa, fa);
                    consume_group<G>(vmask, g, xa, fa, cand, cp_lo, cp_hi);
                    S2_GROUP_FENCE();
                }
            }
        }
        n_valid += __popc(vmask);

        // ---- slow path: queue the candidates of the whole warp, resolve them 32 at a time -------------
        while (__any_sync(0xFFFFFFFFu, cand != 0)) {
            const uint32_t has = cand != 0;
            const uint32_t bal = __ballot_sync(0xFFFFFFFFu, has);
            if (has) {
                const unsigned j = __ffs(cand) - 1;
                cand &= cand - 1;
                const uint32_t at = qlen + __popc(bal & ((1u << lane) - 1u));
                const uint64_t canon = window_canon(w0, w1, w2, r0, r1, r2, j);
                q_canon[at] = canon;
                q_slot[at] = s2_bucket_of(s2_hash(canon).h, t.n_buckets) * S2_BUCKET_SLOTS + (((j < 8 ? cp_lo >> (4 * j) : cp_hi >> (4 * (j - 8)))) & 15u);
                if (MODE == S2_MODE_DETECT) q_pos[at] = tile * 512 + (uint64_t)lane * 16 + j;
            }
            qlen += __popc(bal);
            __syncwarp();
            if (qlen >= 32) drain_queue<MODE>(t, q_canon, q_slot, q_pos, qlen, 32, lane, counts_col, dout, n_hits, inc);
        }
        if (!PIPE) {
            const uint64_t nt = tile + n_warps;
            raw = make_uint4(0, 0, 0, 0); rawx = raw;
            if (nt < n_tiles) {
                raw = ld_stream16(bases, n_bytes, nt * 32 + lane, policy);
                if (lane < 2) rawx = ld_stream16(bases, n_bytes, nt * 32 + 32 + lane, policy);
            }
        }
    }
    if (qlen) drain_queue<MODE>(t, q_canon, q_slot, q_pos, qlen, qlen, lane, counts_col, dout, n_hits, inc);

    n_hits = __reduce_add_sync(0xFFFFFFFFu, n_hits);
    n_valid = __reduce_add_sync(0xFFFFFFFFu, n_valid);
    if (lane == 0 && stats) {
        const long long sign = (int)inc;                                              // +1, or -1 for a take-back replay
        if (n_hits) atomicAdd(&stats[0], (unsigned long long)(sign * (long long)n_hits));
        if (n_valid && !run_if) atomicAdd(&stats[1], (unsigned long long)(sign * (long long)n_valid));   // fallback run: phase A counted them
    }
}

// ---- variants -------------------------------------------------------------------------------------
typedef void (*s2_scan_fn)(const uint8_t *, uint64_t, S2TableView, uint32_t *, S2DetectOut, unsigned long long *, const uint32_t *,
                           const S2DevBatch *);
struct S2ScanVariant { const char *name; s2_scan_fn count_fn, detect_fn; };

#define S2_VARIANT(G, MINB, PIPE) \
    { "G" #G "_B" #MINB "_P" #PIPE, s2_scan_kernel<S2_MODE_COUNT, G, MINB, PIPE>, s2_scan_kernel<S2_MODE_DETECT, G, MINB, PIPE> }

static const S2ScanVariant g_variants[] = {
    S2_VARIANT(4, 2, false),   // 0: round-1a shape
    S2_VARIANT(4, 3, false),   // 1
    S2_VARIANT(2, 3, false),   // 2
    S2_VARIANT(2, 4, false),   // 3
    S2_VARIANT(4, 2, true),    // 4
    S2_VARIANT(2, 3, true),    // 5
    S2_VARIANT(2, 4, true),    // 6
    S2_VARIANT(8, 2, false),   // 7
    S2_VARIANT(4, 3, true),    // 8
    S2_VARIANT(1, 4, true),    // 9
};
static int g_variant = 9;   // G1_B4_Ptrue: best on the config-2 workload in the round-1 sweeps (profiles/r1d_scan_sweep.txt)

int s2_scan_variant_count(void) { return (int)(sizeof g_variants / sizeof g_variants[0]); }
const char *s2_scan_variant_name(int v) { return (v >= 0 && v < s2_scan_variant_count()) ? g_variants[v].name : "?"; }
int s2_scan_variant_get(void) { return g_variant; }
int s2_scan_variant_set(int v)
{
    if (v < 0 || v >= s2_scan_variant_count()) return -1;
    g_variant = v;
    return 0;
}

int s2_scan_blocks_per_sm(int mode)
{
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(
        &n, mode == S2_MODE_COUNT ? g_variants[g_variant].count_fn : g_variants[g_variant].detect_fn, S2_THREADS, 0);
    return n > 0 ? n : 1;
}

void s2_launch_scan_count(const uint8_t *bases, uint64_t n_bytes, const S2TableView &t, int col,
                          unsigned long long *stats, int grid_blocks, cudaStream_t stream)
{
    if (n_bytes == 0) return;
    S2DetectOut none = {};
    g_variants[g_variant].count_fn<<<grid_blocks, S2_THREADS, 0, stream>>>(
        bases, n_bytes, t, t.counts + (uint64_t)col * t.n_slots, none, stats, nullptr, nullptr);
}

// same, but the batch was produced by the GPU ingest kernels: length, veto and increment are read from device memory
void s2_launch_scan_count_dev(const uint8_t *bases, const S2DevBatch *dev, const S2TableView &t, int col,
                              unsigned long long *stats, int grid_blocks, cudaStream_t stream)
{
    S2DetectOut none = {};
    g_variants[g_variant].count_fn<<<grid_blocks, S2_THREADS, 0, stream>>>(
        bases, 0, t, t.counts + (uint64_t)col * t.n_slots, none, stats, nullptr, dev);
}

void s2_launch_scan_detect(const uint8_t *bases, uint64_t n_bytes, const S2TableView &t,
                           const S2DetectOut &out, unsigned long long *stats, int grid_blocks,
                           cudaStream_t stream)
{
    if (n_bytes == 0) return;
    g_variants[g_variant].detect_fn<<<grid_blocks, S2_THREADS, 0, stream>>>(bases, n_bytes, t, nullptr, out, stats, nullptr, nullptr);
}


// detect scan of a batch that was produced on the device (GPU ingest): lengths are in device memory
void s2_launch_scan_detect_dev(const uint8_t *bases, const S2DevBatch *dev, const S2TableView &t, const S2DetectOut &out,
                               unsigned long long *stats, int grid_blocks, cudaStream_t stream)
{
    g_variants[g_variant].detect_fn<<<grid_blocks, S2_THREADS, 0, stream>>>(bases, 0, t, nullptr, out, stats, nullptr, dev);
}

// ------------------------------------------------------------------------------------------------
// two-phase count scan for tables whose fingerprints do not fit L2 (multi-strain union tables)
// ------------------------------------------------------------------------------------------------
// Probing a 1.3 GB fingerprint array at random is DRAM row/latency bound (36 G lookups/s measured).
// Phase A radix-partitions the canonical k-mers of a batch by the top 7 bits of their hash into 128
// streams (8 bytes written + 8 read per lookup, fully coalesced); phase B probes one partition at a
// time, whose 1/128 slice of the table (10 MB for 64 strains) stays L2 resident.  (Round 1 used 32
// partitions: ncu showed 84 % of the probes missing L2 - profiles/r2a_two_phase_ncu.txt - because the grid
// straddles two partitions while it prefetches a third, and three 40 MB slices plus the entry streams
// do not fit the part of the 126 MB L2 that data shared by both dies can use.)

#define S2_PSTAGE 64           /* staged entries per (CTA, partition) and round: 2x the even share of 4096 windows (64 KB: 3 CTAs per SM) */

struct S2PartView {
    uint64_t *pool;            // S2_NPART regions of region_cap entries
    uint64_t region_cap;
    unsigned long long *cursor;    // [S2_NPART] entries written per partition
    uint32_t *overflow;        // set when a region would overflow: phase B is skipped and the direct kernel runs
};

// Phase A.  A CTA works in rounds of 8 tiles (4096 windows): every valid window's canonical k-mer goes to
// the shared-memory stage of its partition (one shared atomic for the slot), then the CTA reserves room
// in each partition's global region with ONE global atomic per partition and round and copies the stage
// out in coalesced runs (about 1 KB each).  Skewed rounds that overflow a stage append straight to global.
__global__ void __launch_bounds__(S2_THREADS, 3)
s2_partition_kernel(const uint8_t *__restrict__ bases, uint64_t n_bytes, S2PartView pv, unsigned long long *__restrict__ stats,
                    const S2DevBatch *__restrict__ dev)
{
    long long sign = 1;
    if (dev) {                                                    // batch produced on the device (GPU ingest)
        if (dev->skip) return;
        n_bytes = dev->n_bytes;
        sign = (int)dev->inc;
    }
    extern __shared__ uint64_t stage[];                           // [S2_NPART][S2_PSTAGE]
    __shared__ uint32_t cnt[S2_NPART];
    __shared__ unsigned long long gbase[S2_NPART];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint64_t n_tiles = (n_bytes + 511) / 512;
    const uint64_t n_rounds = (n_tiles + S2_WARPS - 1) / S2_WARPS;
    uint32_t n_valid = 0;
    if (threadIdx.x < S2_NPART) cnt[threadIdx.x] = 0;
    __syncthreads();
    for (uint64_t round = blockIdx.x; round < n_rounds; round += gridDim.x) {
        const uint64_t tile = round * S2_WARPS + wid;
        if (tile < n_tiles) {
            uint32_t w0, w1, w2, m0, m1, m2;
            load_tile(bases, n_bytes, tile, lane, w0, w1, w2, m0, m1, m2);
            const uint32_t r0 = s2_rc16(w2), r1 = s2_rc16(w1), r2 = s2_rc16(w0);
#pragma unroll 4
            for (unsigned j = 0; j < 16; ++j) {
                if (s2_window_valid(m0, m1, m2, j)) {
                    const uint64_t canon = window_canon(w0, w1, w2, r0, r1, r2, j);
                    const uint32_t p = s2_hash(canon).h >> (32 - S2_NPART_LOG2);
                    const uint32_t at = atomicAdd(&cnt[p], 1u);
                    ++n_valid;
                    if (at < S2_PSTAGE) {
                        stage[p * S2_PSTAGE + at] = canon;
                    } else {                                      // rare: this round is skewed towards one partition
                        const unsigned long long g = atomicAdd(&pv.cursor[p], 1ull);
                        if (g < pv.region_cap) pv.pool[(uint64_t)p * pv.region_cap + g] = canon;
                        else atomicOr(pv.overflow, 1u);
                    }
                }
            }
        }
        __syncthreads();
        if (threadIdx.x < S2_NPART) {                             // one global reservation per partition and round
            const uint32_t c = min(cnt[threadIdx.x], (uint32_t)S2_PSTAGE);
            gbase[threadIdx.x] = c ? atomicAdd(&pv.cursor[threadIdx.x], (unsigned long long)c) : 0ull;
        }
        __syncthreads();
        for (int p = wid; p < S2_NPART; p += S2_WARPS) {          // each warp copies out its share of the partitions
            const uint32_t c = min(cnt[p], (uint32_t)S2_PSTAGE);
            const unsigned long long g = gbase[p];
            if (g + c > pv.region_cap) { if (lane == 0 && c) atomicOr(pv.overflow, 1u); continue; }
            uint64_t *dst = pv.pool + (uint64_t)p * pv.region_cap + g;
            for (uint32_t i = lane; i < c; i += 32) dst[i] = stage[p * S2_PSTAGE + i];
        }
        __syncthreads();
        if (threadIdx.x < S2_NPART) cnt[threadIdx.x] = 0;
        __syncthreads();
    }
    n_valid = __reduce_add_sync(0xFFFFFFFFu, n_valid);
    if (lane == 0 && stats && n_valid) atomicAdd(&stats[1], (unsigned long long)(sign * (long long)n_valid));
}

// phase B, all partitions in ONE launch: CTAs draw work items (4096 entries of one partition) from a global
// counter in partition order, so at any moment the whole grid works on one or two neighbouring partitions
// and their slices of the table stay L2 resident without a barrier between partitions.  While it probes
// partition p an item also pulls its share of slice p+1 into L2 with sequential loads.
#define S2_PITEM 4096
__global__ void __launch_bounds__(S2_THREADS, 4)
s2_probe_all_kernel(S2PartView pv, S2TableView t, uint32_t *__restrict__ counts_col,
                    unsigned long long *__restrict__ stats, unsigned long long *__restrict__ work_counter, const S2DevBatch *__restrict__ dev, uint32_t flags)
{
    if (*pv.overflow) return;
    uint32_t inc = 1u;
    if (dev) { if (dev->skip) return; inc = dev->inc; }
    __shared__ unsigned long long pre[S2_NPART + 1];
    __shared__ unsigned long long item_s;
    if (threadIdx.x == 0) {
        unsigned long long acc = 0;
        for (int p = 0; p < S2_NPART; ++p) { pre[p] = acc; acc += (pv.cursor[p] + S2_PITEM - 1) / S2_PITEM; }
        pre[S2_NPART] = acc;
    }
    __syncthreads();
    const unsigned long long total = pre[S2_NPART];
    uint32_t n_hits = 0, sink = 0;
    for (;;) {
        if (threadIdx.x == 0) item_s = atomicAdd(work_counter, 1ull);
        __syncthreads();
        const unsigned long long item = item_s;
        __syncthreads();
        if (item >= total) break;
        int part = 0;                                             // largest part with pre[part] <= item < pre[part + 1]
#pragma unroll
        for (int step = S2_NPART / 2; step > 0; step >>= 1) if (pre[part + step] <= item) part += step;
        const uint64_t n = pv.cursor[part];
        const uint64_t off = (item - pre[part]) * S2_PITEM;
        const uint64_t *__restrict__ src = pv.pool + (uint64_t)part * pv.region_cap;
        // this item's share of the next partition's slice -> L2
        if (part + 1 < S2_NPART && !(flags & 1u)) {
            const uint64_t items_here = pre[part + 1] - pre[part];
            const uint64_t lo = ((uint64_t)(part + 1) * t.n_buckets + S2_NPART - 1) / S2_NPART;
            const uint64_t hi = ((uint64_t)(part + 2) * t.n_buckets + S2_NPART - 1) / S2_NPART;
            if (flags & 8u) {                                     // (round 1's form: loads whose results a thread waits for)
                const uint64_t share = (hi - lo + items_here - 1) / items_here;
                const uint64_t b0 = lo + (item - pre[part]) * share;
                for (uint64_t b = b0 + threadIdx.x; b < b0 + share && b < hi; b += S2_THREADS) {
                    uint32_t x[8];
                    if (flags & 4u) ld_bucket256_plain(t.fp, (uint32_t)b, x); else ld_bucket256(t.fp, (uint32_t)b, x);
                    sink |= x[0] ^ x[7];
                }
            } else if (threadIdx.x == 0) {
                // one bulk prefetch per item, nobody waits for it; the first half of a partition's items bring in the whole next
                // slice, so that it is there when the grid moves on
                const uint64_t half = (items_here + 1) / 2;
                const uint64_t k = item - pre[part];
                if (k < half) {
                    const uint64_t share = (hi - lo + half - 1) / half;
                    const uint64_t b0 = lo + k * share, b1 = min(b0 + share, hi);
                    if (b1 > b0) {
                        const uint16_t *a = t.fp + b0 * S2_BUCKET_SLOTS;
                        const uint32_t bytes = (uint32_t)((b1 - b0) * S2_BUCKET_SLOTS * sizeof(uint16_t));
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(a), "r"(bytes) : "memory");
                    }
                }
            }
        }
#pragma unroll 1
        for (int round = 0; round < S2_PITEM / (4 * S2_THREADS); ++round) {
            uint64_t canon[4]; uint32_t x[4][8], fp2[4], bucket[4]; bool have[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint64_t i = off + (uint64_t)(round * 4 + u) * S2_THREADS + threadIdx.x;
                have[u] = i < n;
                canon[u] = have[u] ? __ldcs(src + i) : 0;
                const s2_hash_t hh = s2_hash(canon[u]);
                fp2[u] = hh.fp * 0x00010001u;
                bucket[u] = s2_bucket_of(hh.h, t.n_buckets);
                if (flags & 2u) ld_bucket256_plain(t.fp, bucket[u], x[u]); else ld_bucket256(t.fp, bucket[u], x[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t m = fp_match_bits(x[u], fp2[u]);
                const bool full = x[u][7] > 0xFFFFu;
                if (have[u] && (m != 0 || full)) {
                    const uint32_t b = 31u - (uint32_t)__clz(m);
                    uint32_t slot = bucket[u] * S2_BUCKET_SLOTS + ((((b & 15u) << 1) | ((b >> 4) & 1u)) & 15u);
                    uint64_t key = t.keys[slot];
                    bool hit = (key & S2_KMER_MASK) == canon[u] && key != S2_EMPTY_KEY;
                    if (!hit) hit = probe_exact(t, canon[u], slot, key);
                    if (hit) { atomicAdd(&counts_col[slot], inc); ++n_hits; }
                }
            }
        }
    }
    n_hits = __reduce_add_sync(0xFFFFFFFFu, n_hits);
    if ((threadIdx.x & 31) == 0 && n_hits && stats) atomicAdd(&stats[0], (unsigned long long)((long long)(int)inc * (long long)n_hits));
    if (sink == 0x12345679u) atomicOr(pv.overflow + 1, sink);     // never true; keeps the prefetch loads alive
}

// ---- second versions (S2_PART_A=2 / S2_PART_B=2; profiles/r2d_two_phase_128_partitions_ncu_metrics.csv showed phase A
// waiting for its tile loads and at its barriers, phase B waiting for probes that miss L2 and at two barriers per item) ----
// Phase A, v2: the next round's tile is on its way while this round's windows are staged (as in the scan kernel), and a
// round has three barriers instead of four (the counters are cleared by the threads that read them for the reservation).
template <int PSTAGE, int MINB>
__global__ void __launch_bounds__(S2_THREADS, MINB)
s2_partition_kernel_v2(const uint8_t *__restrict__ bases, uint64_t n_bytes, S2PartView pv, unsigned long long *__restrict__ stats,
                       const S2DevBatch *__restrict__ dev)
{
    long long sign = 1;
    if (dev) {
        if (dev->skip) return;
        n_bytes = dev->n_bytes;
        sign = (int)dev->inc;
    }
    extern __shared__ uint64_t stage[];                           // [S2_NPART][PSTAGE]
    __shared__ uint32_t cnt[S2_NPART], cnt_c[S2_NPART];
    __shared__ unsigned long long gbase[S2_NPART];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint64_t n_tiles = (n_bytes + 511) / 512;
    const uint64_t n_rounds = (n_tiles + S2_WARPS - 1) / S2_WARPS;
    uint32_t n_valid = 0;
    uint64_t policy;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    if (threadIdx.x < S2_NPART) cnt[threadIdx.x] = 0;
    uint64_t round = blockIdx.x;
    uint4 raw = make_uint4(0, 0, 0, 0), rawx = make_uint4(0, 0, 0, 0);
    if (round < n_rounds) {
        const uint64_t tile = round * S2_WARPS + wid;
        if (tile < n_tiles) {
            raw = ld_stream16(bases, n_bytes, tile * 32 + lane, policy);
            if (lane < 2) rawx = ld_stream16(bases, n_bytes, tile * 32 + 32 + lane, policy);
        }
    }
    __syncthreads();
    for (; round < n_rounds; round += gridDim.x) {
        const uint64_t tile = round * S2_WARPS + wid;
        uint32_t w0 = 0, w1 = 0, w2 = 0, m0 = 0, m1 = 0, m2 = 0;
        {
            uint32_t wx = 0, mx = 0;
            if (tile < n_tiles) {
                pack_chunk(raw, n_bytes, tile * 32 + lane, w0, m0);
                if (lane < 2) pack_chunk(rawx, n_bytes, tile * 32 + 32 + lane, wx, mx);
            }
            const uint64_t nt = (round + gridDim.x) * S2_WARPS + wid;
            raw = make_uint4(0, 0, 0, 0); rawx = raw;
            if (round + gridDim.x < n_rounds && nt < n_tiles) {
                raw = ld_stream16(bases, n_bytes, nt * 32 + lane, policy);
                if (lane < 2) rawx = ld_stream16(bases, n_bytes, nt * 32 + 32 + lane, policy);
            }
            const uint32_t a1 = __shfl_sync(0xFFFFFFFFu, w0, (lane + 1) & 31), b1 = __shfl_sync(0xFFFFFFFFu, wx, (lane + 1) & 31);
            const uint32_t a2 = __shfl_sync(0xFFFFFFFFu, w0, (lane + 2) & 31), b2 = __shfl_sync(0xFFFFFFFFu, wx, (lane + 2) & 31);
            const uint32_t c1 = __shfl_sync(0xFFFFFFFFu, m0, (lane + 1) & 31), d1 = __shfl_sync(0xFFFFFFFFu, mx, (lane + 1) & 31);
            const uint32_t c2 = __shfl_sync(0xFFFFFFFFu, m0, (lane + 2) & 31), d2 = __shfl_sync(0xFFFFFFFFu, mx, (lane + 2) & 31);
            w1 = lane < 31 ? a1 : b1;  m1 = lane < 31 ? c1 : d1;
            w2 = lane < 30 ? a2 : b2;  m2 = lane < 30 ? c2 : d2;
        }
        if (tile < n_tiles) {
            const uint32_t r0 = s2_rc16(w2), r1 = s2_rc16(w1), r2 = s2_rc16(w0);
#pragma unroll 4
            for (unsigned j = 0; j < 16; ++j) {
                if (s2_window_valid(m0, m1, m2, j)) {
                    const uint64_t canon = window_canon(w0, w1, w2, r0, r1, r2, j);
                    const uint32_t p = s2_hash(canon).h >> (32 - S2_NPART_LOG2);
                    const uint32_t at = atomicAdd(&cnt[p], 1u);
                    ++n_valid;
                    if (at < PSTAGE) {
                        stage[p * PSTAGE + at] = canon;
                    } else {
                        const unsigned long long g = atomicAdd(&pv.cursor[p], 1ull);
                        if (g < pv.region_cap) pv.pool[(uint64_t)p * pv.region_cap + g] = canon;
                        else atomicOr(pv.overflow, 1u);
                    }
                }
            }
        }
        __syncthreads();
        if (threadIdx.x < S2_NPART) {                             // one global reservation per partition and round
            const uint32_t c = min(cnt[threadIdx.x], (uint32_t)PSTAGE);
            cnt[threadIdx.x] = 0;
            cnt_c[threadIdx.x] = c;
            gbase[threadIdx.x] = c ? atomicAdd(&pv.cursor[threadIdx.x], (unsigned long long)c) : 0ull;
        }
        __syncthreads();
        for (int p = wid; p < S2_NPART; p += S2_WARPS) {
            const uint32_t c = cnt_c[p];
            const unsigned long long g = gbase[p];
            if (g + c > pv.region_cap) { if (lane == 0 && c) atomicOr(pv.overflow, 1u); continue; }
            uint64_t *dst = pv.pool + (uint64_t)p * pv.region_cap + g;
            for (uint32_t i = lane; i < c; i += 32) __stcs(dst + i, stage[p * PSTAGE + i]);
        }
        __syncthreads();
    }
    n_valid = __reduce_add_sync(0xFFFFFFFFu, n_valid);
    if (lane == 0 && stats && n_valid) atomicAdd(&stats[1], (unsigned long long)(sign * (long long)n_valid));
}

// Phase B, v2: WARPS draw the work items (no barrier inside the loop), and the share of the next partition's slice that an
// item pulls into L2 is one bulk prefetch instruction instead of loads whose results somebody has to wait for.
__global__ void __launch_bounds__(S2_THREADS, 4)
s2_probe_all_kernel_v2(S2PartView pv, S2TableView t, uint32_t *__restrict__ counts_col,
                       unsigned long long *__restrict__ stats, unsigned long long *__restrict__ work_counter, const S2DevBatch *__restrict__ dev, uint32_t flags)
{
    if (*pv.overflow) return;
    uint32_t inc = 1u;
    if (dev) { if (dev->skip) return; inc = dev->inc; }
    __shared__ unsigned long long pre[S2_NPART + 1];
    if (threadIdx.x == 0) {
        unsigned long long acc = 0;
        for (int p = 0; p < S2_NPART; ++p) { pre[p] = acc; acc += (pv.cursor[p] + S2_PITEM - 1) / S2_PITEM; }
        pre[S2_NPART] = acc;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const unsigned long long total = pre[S2_NPART];
    uint32_t n_hits = 0;
    for (;;) {
        unsigned long long item = 0;
        if (lane == 0) item = atomicAdd(work_counter, 1ull);
        item = __shfl_sync(0xFFFFFFFFu, item, 0);
        if (item >= total) break;
        int part = 0;
#pragma unroll
        for (int step = S2_NPART / 2; step > 0; step >>= 1) if (pre[part + step] <= item) part += step;
        const uint64_t n = pv.cursor[part];
        const uint64_t off = (item - pre[part]) * S2_PITEM;
        const uint64_t *__restrict__ src = pv.pool + (uint64_t)part * pv.region_cap;
        if (part + 1 < S2_NPART && !(flags & 1u) && lane == 0) {
            const uint64_t items_here = pre[part + 1] - pre[part];
            const uint64_t lo = ((uint64_t)(part + 1) * t.n_buckets + S2_NPART - 1) / S2_NPART;
            const uint64_t hi = ((uint64_t)(part + 2) * t.n_buckets + S2_NPART - 1) / S2_NPART;
            const uint64_t share = (hi - lo + items_here - 1) / items_here;
            const uint64_t b0 = lo + (item - pre[part]) * share;
            const uint64_t b1 = min(b0 + share, hi);
            if (b1 > b0) {
                const uint16_t *a = t.fp + b0 * S2_BUCKET_SLOTS;                // buckets are 32 bytes: the range is 16-byte aligned
                const uint32_t bytes = (uint32_t)((b1 - b0) * S2_BUCKET_SLOTS * sizeof(uint16_t));
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(a), "r"(bytes) : "memory");
            }
        }
#pragma unroll 1
        for (int round = 0; round < S2_PITEM / (4 * 32); ++round) {
            uint64_t canon[4]; uint32_t x[4][8], fp2[4], bucket[4]; bool have[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint64_t i = off + (uint64_t)(round * 4 + u) * 32 + lane;
                have[u] = i < n;
                canon[u] = have[u] ? __ldcs(src + i) : 0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const s2_hash_t hh = s2_hash(canon[u]);
                fp2[u] = hh.fp * 0x00010001u;
                bucket[u] = s2_bucket_of(hh.h, t.n_buckets);
                if (flags & 2u) ld_bucket256_plain(t.fp, bucket[u], x[u]); else ld_bucket256(t.fp, bucket[u], x[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t m = fp_match_bits(x[u], fp2[u]);
                const bool full = x[u][7] > 0xFFFFu;
                if (have[u] && (m != 0 || full)) {
                    const uint32_t b = 31u - (uint32_t)__clz(m);
                    uint32_t slot = bucket[u] * S2_BUCKET_SLOTS + ((((b & 15u) << 1) | ((b >> 4) & 1u)) & 15u);
                    uint64_t key = t.keys[slot];
                    bool hit = (key & S2_KMER_MASK) == canon[u] && key != S2_EMPTY_KEY;
                    if (!hit) hit = probe_exact(t, canon[u], slot, key);
                    if (hit) { atomicAdd(&counts_col[slot], inc); ++n_hits; }
                }
            }
            if (off + (uint64_t)(round + 1) * 128 >= n) break;
        }
    }
    n_hits = __reduce_add_sync(0xFFFFFFFFu, n_hits);
    if (lane == 0 && n_hits && stats) atomicAdd(&stats[0], (unsigned long long)((long long)(int)inc * (long long)n_hits));
}

// pull one partition's slice of the fingerprint array into L2 with sequential 256-bit loads (evict-last):
// the probes that follow then find it there instead of fetching it at random, sector by sector
__global__ void __launch_bounds__(S2_THREADS)
s2_prefetch_slice_kernel(const uint16_t *__restrict__ fp, uint32_t bucket_lo, uint32_t bucket_hi, uint32_t *__restrict__ sink)
{
    uint32_t acc = 0;
    for (uint64_t b = bucket_lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < bucket_hi; b += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t x[8];
        ld_bucket256(fp, (uint32_t)b, x);
        acc |= x[0] ^ x[7];
    }
    if (acc == 0x12345679u) *sink = acc;              // never true in practice; keeps the loads alive
}

size_t s2_partition_smem_bytes(void) { return (size_t)S2_NPART * S2_PSTAGE * sizeof(uint64_t); }

// whole two-phase scan on one stream.  part_pool holds S2_NPART * region_cap entries; cursor[S2_NPART]
// and overflow[1] are zeroed here.  If a region overflows (pathological low-complexity input) phase B
// does nothing and the direct scan kernel takes over, so the counters are exact either way.
void s2_launch_scan_count_partitioned(const uint8_t *bases, uint64_t n_bytes, const S2TableView &t, int col,
                                      unsigned long long *stats, uint64_t *part_pool, uint64_t region_cap,
                                      unsigned long long *cursor, uint32_t *overflow, int n_sm, int grid_blocks,
                                      cudaStream_t stream, const S2DevBatch *dev)
{
    if (n_bytes == 0 && !dev) return;
    unsigned long long *work_counter = cursor + S2_NPART;        // cursor[] has one spare slot for the item counter
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(s2_partition_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s2_partition_smem_bytes());
        attr_set = true;
    }
    cudaMemsetAsync(cursor, 0, S2_NPART * sizeof(unsigned long long), stream);
    cudaMemsetAsync(overflow, 0, sizeof(uint32_t), stream);
    S2PartView pv = { part_pool, region_cap, cursor, overflow };
    // Measured (profiles/r2j_part_probe.txt, 64-strain table): phase A with 52-entry stages and four CTAs per SM is the
    // fastest form at every batch size; phase B with warp-drawn items wins once a partition holds a thousand items or more
    // (batches of 640 Mbases: 93.8 against 84.4 G lookups/s) and loses badly below that (the grid then spans a dozen
    // partitions and their slices fall out of L2: 54 against 80 at 200 Mbases).
    static const int env_a = getenv("S2_PART_A") ? atoi(getenv("S2_PART_A")) : 0, env_b = getenv("S2_PART_B") ? atoi(getenv("S2_PART_B")) : 0;
    const int ver_a = env_a ? env_a : 3;
    const int ver_b = env_b ? env_b : (!dev && n_bytes >= (512ull << 20) ? 2 : 1);
    if (ver_a == 2) {
        cudaFuncSetAttribute(s2_partition_kernel_v2<S2_PSTAGE, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s2_partition_smem_bytes());
        s2_partition_kernel_v2<S2_PSTAGE, 3><<<n_sm * 3, S2_THREADS, s2_partition_smem_bytes(), stream>>>(bases, n_bytes, pv, stats, dev);
    } else if (ver_a == 3) {             // 52 staged entries per partition and round (1.6 x the even share): four CTAs per SM
        const size_t smem = (size_t)S2_NPART * 52 * sizeof(uint64_t);
        cudaFuncSetAttribute(s2_partition_kernel_v2<52, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        s2_partition_kernel_v2<52, 4><<<n_sm * 4, S2_THREADS, smem, stream>>>(bases, n_bytes, pv, stats, dev);
    } else
    s2_partition_kernel<<<n_sm * 3, S2_THREADS, s2_partition_smem_bytes(), stream>>>(bases, n_bytes, pv, stats, dev);
    uint32_t *counts_col = t.counts + (uint64_t)col * t.n_slots;
    // buckets whose hash has top bits == p: [ceil(p * nb / 32), ceil((p+1) * nb / 32)); slice 0 is prefetched
    // by its own small kernel, every later slice by the items of the partition before it
    const uint32_t hi0 = (uint32_t)(((uint64_t)t.n_buckets + S2_NPART - 1) / S2_NPART);
    s2_prefetch_slice_kernel<<<n_sm * 4, S2_THREADS, 0, stream>>>(t.fp, 0, hi0, overflow + 1);
    cudaMemsetAsync(work_counter, 0, sizeof(unsigned long long), stream);
    static const uint32_t probe_flags = getenv("S2_PROBE_FLAGS") ? (uint32_t)atoi(getenv("S2_PROBE_FLAGS")) : 0u;
    if (ver_b == 2) s2_probe_all_kernel_v2<<<n_sm * 4, S2_THREADS, 0, stream>>>(pv, t, counts_col, stats, work_counter, dev, probe_flags);
    else
    s2_probe_all_kernel<<<n_sm * 4, S2_THREADS, 0, stream>>>(pv, t, counts_col, stats, work_counter, dev, probe_flags);
    S2DetectOut none = {};
    g_variants[g_variant].count_fn<<<grid_blocks, S2_THREADS, 0, stream>>>(bases, n_bytes, t, counts_col, none, stats, overflow, dev);
}

// ------------------------------------------------------------------------------------------------
// table build: insert-if-absent with first-occurrence position and reference count (column 0)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t table_insert(const S2TableView &t, uint64_t canon)
{
    const s2_hash_t hh = s2_hash(canon);
    uint32_t b = s2_bucket_of(hh.h, t.n_buckets);
    for (;;) {
        for (int i = 0; i < S2_BUCKET_SLOTS; ++i) {
            const uint32_t slot = b * S2_BUCKET_SLOTS + i;
            unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(&t.keys[slot]);
            if (cur == S2_EMPTY_KEY) {
                cur = atomicCAS(reinterpret_cast<unsigned long long *>(&t.keys[slot]), S2_EMPTY_KEY, canon);
                if (cur == S2_EMPTY_KEY) {                       // we own the slot: publish its fingerprint
                    t.fp[slot] = (uint16_t)hh.fp;
                    return slot;
                }
            }
            if ((cur & S2_KMER_MASK) == canon) return slot;      // somebody (maybe just now) inserted it
        }
        b = (b + 1 == t.n_buckets) ? 0 : b + 1;                  // bucket full: linear probing by bucket
    }
}

__global__ void __launch_bounds__(S2_THREADS)
s2_build_insert_kernel(const uint8_t *__restrict__ bases, uint64_t n_bytes, S2TableView t,
                       uint32_t *__restrict__ first_pos, uint32_t *__restrict__ slot_of_pos)
{
    const int lane = threadIdx.x & 31;
    const uint64_t gwarp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t n_tiles = (n_bytes + 511) / 512;
    for (uint64_t tile = gwarp; tile < n_tiles; tile += n_warps) {
        uint32_t w0, w1, w2, m0, m1, m2;
        load_tile(bases, n_bytes, tile, lane, w0, w1, w2, m0, m1, m2);
        const uint32_t r0 = s2_rc16(w2), r1 = s2_rc16(w1), r2 = s2_rc16(w0);
        for (unsigned j = 0; j < 16; ++j) {
            const uint64_t pos = tile * 512 + (uint64_t)lane * 16 + j;
            if (pos >= n_bytes) break;
            uint32_t slot = S2_NONE;
            if (s2_window_valid(m0, m1, m2, j)) {
                slot = table_insert(t, window_canon(w0, w1, w2, r0, r1, r2, j));
                atomicMin(&first_pos[slot], (uint32_t)pos);      // insertion order = first occurrence in file order
                atomicAdd(&t.counts[slot], 1u);                  // column 0: default 1, +1 per repeat
            }
            slot_of_pos[pos] = slot;
        }
    }
}

void s2_launch_build_insert(const uint8_t *bases, uint64_t n_bytes, const S2TableView &t,
                            uint32_t *first_pos, uint32_t *slot_of_pos, cudaStream_t stream)
{
    if (n_bytes == 0) return;
    const uint64_t n_tiles = (n_bytes + 511) / 512;
    const uint64_t blocks = (n_tiles + (S2_THREADS / 32) - 1) / (S2_THREADS / 32);
    s2_build_insert_kernel<<<(unsigned)(blocks < 148ull * 16 ? blocks : 148ull * 16), S2_THREADS, 0, stream>>>(
        bases, n_bytes, t, first_pos, slot_of_pos);
}

// ---- first-occurrence ranking: flag[p] = (p is the first position of its key); rank = exclusive scan
#define S2_RANK_PER_THREAD 4
#define S2_RANK_PER_BLOCK (S2_THREADS * S2_RANK_PER_THREAD)

__device__ __forceinline__ uint32_t rank_flag(uint64_t p, uint64_t n, const uint32_t *first_pos,
                                              const uint32_t *slot_of_pos, uint32_t &slot)
{
    slot = S2_NONE;
    if (p >= n) return 0;
    slot = slot_of_pos[p];
    return (slot != S2_NONE && first_pos[slot] == (uint32_t)p) ? 1u : 0u;
}

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t &total)
{
    __shared__ uint32_t warp_sums[S2_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += n; }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < S2_THREADS / 32; ++i) { const uint32_t s = warp_sums[i]; if (i < wid) base += s; tot += s; }
    __syncthreads();
    total = tot;
    return base + inc - v;
}

__global__ void __launch_bounds__(S2_THREADS)
s2_rank_count_kernel(uint64_t n, const uint32_t *__restrict__ first_pos, const uint32_t *__restrict__ slot_of_pos,
                     uint32_t *__restrict__ block_sums)
{
    const uint64_t base = (uint64_t)blockIdx.x * S2_RANK_PER_BLOCK + (uint64_t)threadIdx.x * S2_RANK_PER_THREAD;
    uint32_t c = 0, slot;
#pragma unroll
    for (int i = 0; i < S2_RANK_PER_THREAD; ++i) c += rank_flag(base + i, n, first_pos, slot_of_pos, slot);
    uint32_t total;
    block_exclusive_scan(c, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(S2_THREADS)
s2_rank_scan_kernel(uint32_t *__restrict__ block_sums, uint32_t n_blocks, unsigned long long *__restrict__ d_n_keys)
{
    // single CTA: exclusive scan of the per-block counts in place, carrying across 256-element strips
    uint32_t carry = 0;
    for (uint32_t base = 0; base < n_blocks; base += S2_THREADS) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n_blocks ? block_sums[i] : 0;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan(v, total);
        if (i < n_blocks) block_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) *d_n_keys = carry;
}

__global__ void __launch_bounds__(S2_THREADS)
s2_rank_write_kernel(uint64_t n, const uint32_t *__restrict__ first_pos, const uint32_t *__restrict__ slot_of_pos,
                     const uint32_t *__restrict__ block_offsets, uint32_t *__restrict__ rank_slot,
                     uint32_t *__restrict__ rank_pos)
{
    const uint64_t base = (uint64_t)blockIdx.x * S2_RANK_PER_BLOCK + (uint64_t)threadIdx.x * S2_RANK_PER_THREAD;
    uint32_t f[S2_RANK_PER_THREAD], s[S2_RANK_PER_THREAD], c = 0;
#pragma unroll
    for (int i = 0; i < S2_RANK_PER_THREAD; ++i) { f[i] = rank_flag(base + i, n, first_pos, slot_of_pos, s[i]); c += f[i]; }
    uint32_t total;
    uint32_t r = block_offsets[blockIdx.x] + block_exclusive_scan(c, total);
#pragma unroll
    for (int i = 0; i < S2_RANK_PER_THREAD; ++i)
        if (f[i]) { rank_slot[r] = s[i]; rank_pos[r] = (uint32_t)(base + i); ++r; }
}

void s2_launch_build_rank(uint64_t n_bytes, const uint32_t *first_pos, const uint32_t *slot_of_pos,
                          uint32_t *block_sums, uint32_t n_blocks, uint32_t *rank_slot, uint32_t *rank_pos,
                          unsigned long long *d_n_keys, cudaStream_t stream)
{
    if (n_blocks == 0) { cudaMemsetAsync(d_n_keys, 0, sizeof(unsigned long long), stream); return; }
    s2_rank_count_kernel<<<n_blocks, S2_THREADS, 0, stream>>>(n_bytes, first_pos, slot_of_pos, block_sums);
    s2_rank_scan_kernel<<<1, S2_THREADS, 0, stream>>>(block_sums, n_blocks, d_n_keys);
    s2_rank_write_kernel<<<n_blocks, S2_THREADS, 0, stream>>>(n_bytes, first_pos, slot_of_pos, block_sums, rank_slot, rank_pos);
}


// ------------------------------------------------------------------------------------------------
// count-table formatting on the device  (print_hash_counts, src/kmer_scrub_count.c:134-156)
// ------------------------------------------------------------------------------------------------
// One thread per row: "<31 letters>\t%d\t%d\t%d[\t%d]\n" with the counters printed as signed int like
// the reference's %d.  Three launches: row lengths per 1024-row block, scan of the block sums, write.
#define S2_FMT_MAXROW 96

__device__ __forceinline__ int fmt_int(char *p, uint32_t u)
{
    int32_t v = (int32_t)u;
    uint32_t a = v < 0 ? (uint32_t)(-(int64_t)v) : (uint32_t)v;
    char tmp[12]; int n = 0, w = 0;
    do { tmp[n++] = (char)('0' + a % 10); a /= 10; } while (a);
    if (v < 0) p[w++] = '-';
    while (n) p[w++] = tmp[--n];
    return w;
}

__device__ __forceinline__ int fmt_row(char *row, uint64_t key, const uint32_t *const *cols, int n_cols, uint32_t id)
{
    int w = 0;
#pragma unroll
    for (int i = 0; i < S2_K; ++i) row[w++] = s2_letter((uint32_t)(key >> (2 * (S2_K - 1 - i))) & 3u);
    for (int c = 0; c < n_cols; ++c) { row[w++] = '\t'; w += fmt_int(row + w, cols[c][id]); }
    row[w++] = '\n';
    return w;
}

struct S2FmtCols { const uint32_t *col[4]; int n; };

__global__ void __launch_bounds__(S2_THREADS)
s2_fmt_len_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ order, uint64_t n, S2FmtCols fc,
                  unsigned long long *__restrict__ block_sums)
{
    const uint64_t base = (uint64_t)blockIdx.x * S2_RANK_PER_BLOCK + (uint64_t)threadIdx.x * S2_RANK_PER_THREAD;
    uint32_t len = 0;
    char row[S2_FMT_MAXROW];
    for (int i = 0; i < S2_RANK_PER_THREAD; ++i)
        if (base + i < n) { const uint32_t id = order[base + i]; len += fmt_row(row, keys[id], fc.col, fc.n, id); }
    uint32_t total;
    block_exclusive_scan(len, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(S2_THREADS)
s2_fmt_scan_kernel(unsigned long long *__restrict__ block_sums, uint32_t n_blocks, unsigned long long *__restrict__ total_out)
{
    // single CTA, sequential over strips of 256 blocks (a few thousand blocks at most)
    __shared__ unsigned long long carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t b0 = 0; b0 < n_blocks; b0 += S2_THREADS) {
        const uint32_t i = b0 + threadIdx.x;
        const uint32_t v = i < n_blocks ? (uint32_t)block_sums[i] : 0;     // a block is < 100 KB of text
        uint32_t total;
        const uint32_t ex = block_exclusive_scan(v, total);
        const unsigned long long carry = carry_s;
        if (i < n_blocks) block_sums[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry_s;
}

__global__ void __launch_bounds__(S2_THREADS)
s2_fmt_write_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ order, uint64_t n, S2FmtCols fc,
                    const unsigned long long *__restrict__ block_offsets, char *__restrict__ out)
{
    const uint64_t base = (uint64_t)blockIdx.x * S2_RANK_PER_BLOCK + (uint64_t)threadIdx.x * S2_RANK_PER_THREAD;
    char rows[S2_RANK_PER_THREAD][S2_FMT_MAXROW];
    int lens[S2_RANK_PER_THREAD];
    uint32_t len = 0;
    for (int i = 0; i < S2_RANK_PER_THREAD; ++i) {
        lens[i] = 0;
        if (base + i < n) { const uint32_t id = order[base + i]; lens[i] = fmt_row(rows[i], keys[id], fc.col, fc.n, id); }
        len += lens[i];
    }
    uint32_t total;
    unsigned long long at = block_offsets[blockIdx.x] + block_exclusive_scan(len, total);
    for (int i = 0; i < S2_RANK_PER_THREAD; ++i) {
        for (int b = 0; b < lens[i]; ++b) out[at + b] = rows[i][b];
        at += lens[i];
    }
}

void s2_launch_format(const uint64_t *keys, const uint32_t *order, uint64_t n, const uint32_t *const *cols, int n_cols,
                      unsigned long long *block_sums, unsigned long long *d_total, char *out, int phase, cudaStream_t stream)
{
    if (n == 0) { if (phase == 0) cudaMemsetAsync(d_total, 0, sizeof(unsigned long long), stream); return; }
    S2FmtCols fc; fc.n = n_cols;
    for (int c = 0; c < 4; ++c) fc.col[c] = c < n_cols ? cols[c] : nullptr;
    const uint32_t n_blocks = (uint32_t)((n + S2_RANK_PER_BLOCK - 1) / S2_RANK_PER_BLOCK);
    if (phase == 0) {
        s2_fmt_len_kernel<<<n_blocks, S2_THREADS, 0, stream>>>(keys, order, n, fc, block_sums);
        s2_fmt_scan_kernel<<<1, S2_THREADS, 0, stream>>>(block_sums, n_blocks, d_total);
    } else {
        s2_fmt_write_kernel<<<n_blocks, S2_THREADS, 0, stream>>>(keys, order, n, fc, block_sums, out);
    }
}

// ------------------------------------------------------------------------------------------------
// export / gather / scatter / flag / lookup / pack / fill
// ------------------------------------------------------------------------------------------------
__global__ void s2_export_kernel(S2TableView t, const uint32_t *__restrict__ rank_slot, uint64_t n_keys,
                                 uint64_t *__restrict__ keys_out, uint32_t *__restrict__ djb2_out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_keys) return;
    const uint64_t k = t.keys[rank_slot[i]] & S2_KMER_MASK;
    keys_out[i] = k;
    djb2_out[i] = s2_djb2_of_kmer(k);          // hashU of the key's ASCII spelling, before "% M"
}

__global__ void s2_gather_kernel(const uint32_t *__restrict__ col, const uint32_t *__restrict__ rank_slot,
                                 uint64_t n_keys, uint32_t *__restrict__ out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_keys) out[i] = col[rank_slot[i]];
}

__global__ void s2_scatter_kernel(uint32_t *__restrict__ col, const uint32_t *__restrict__ rank_slot,
                                  uint64_t n_keys, const uint32_t *__restrict__ in)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_keys) col[rank_slot[i]] = in[i];
}

__global__ void s2_flag_kernel(S2TableView t, const uint64_t *__restrict__ kmers, uint64_t n, uint8_t *__restrict__ found, int set)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t slot; uint64_t key;
    const bool hit = probe_exact(t, kmers[i] & S2_KMER_MASK, slot, key);
    if (hit) {
        if (set) atomicOr(reinterpret_cast<unsigned long long *>(&t.keys[slot]), S2_INFORMATIVE_BIT);
        else atomicAnd(reinterpret_cast<unsigned long long *>(&t.keys[slot]), ~S2_INFORMATIVE_BIT);
    }
    if (found) found[i] = hit ? 1 : 0;
}

__global__ void s2_lookup_kernel(S2TableView t, const uint64_t *__restrict__ kmers, uint64_t n, uint32_t *__restrict__ slot_out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t slot; uint64_t key;
    slot_out[i] = probe_exact(t, kmers[i] & S2_KMER_MASK, slot, key) ? slot : S2_NONE;
}

// counter of column `col` for each given canonical k-mer (0 when the key is absent): the multi-strain batch
// reads a strain's rows out of the union table with this
__global__ void s2_counts_by_key_kernel(S2TableView t, const uint32_t *__restrict__ col, const uint64_t *__restrict__ kmers,
                                        uint64_t n, uint32_t *__restrict__ out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t slot; uint64_t key;
    out[i] = probe_exact(t, kmers[i] & S2_KMER_MASK, slot, key) ? col[slot] : 0u;
}

// standalone 2-bit pack: one 128-bit coalesced load per thread -> 32-bit word + 16-bit validity mask
__global__ void s2_pack_kernel(const uint8_t *__restrict__ bases, uint64_t n_bytes,
                               uint32_t *__restrict__ words, uint16_t *__restrict__ masks)
{
    const uint64_t chunk = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (chunk * 16 >= n_bytes) return;
    uint32_t w, m;
    load_chunk(bases, n_bytes, chunk, w, m);
    words[chunk] = w;
    masks[chunk] = (uint16_t)m;
}

__global__ void s2_fill_u32_kernel(uint32_t *p, uint64_t n, uint32_t v)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = v;
}

static inline unsigned blocks_for(uint64_t n, unsigned threads) { return (unsigned)((n + threads - 1) / threads); }

void s2_launch_export(const S2TableView &t, const uint32_t *rank_slot, uint64_t n_keys,
                      uint64_t *keys_out, uint32_t *djb2_out, cudaStream_t stream)
{
    if (n_keys) s2_export_kernel<<<blocks_for(n_keys, 256), 256, 0, stream>>>(t, rank_slot, n_keys, keys_out, djb2_out);
}

void s2_launch_gather_counts(const S2TableView &t, int col, const uint32_t *rank_slot, uint64_t n_keys,
                             uint32_t *out, cudaStream_t stream)
{
    if (n_keys) s2_gather_kernel<<<blocks_for(n_keys, 256), 256, 0, stream>>>(t.counts + (uint64_t)col * t.n_slots, rank_slot, n_keys, out);
}

void s2_launch_scatter_counts(const S2TableView &t, int col, const uint32_t *rank_slot, uint64_t n_keys,
                              const uint32_t *in, cudaStream_t stream)
{
    if (n_keys) s2_scatter_kernel<<<blocks_for(n_keys, 256), 256, 0, stream>>>(t.counts + (uint64_t)col * t.n_slots, rank_slot, n_keys, in);
}

// The all-reduce of a counter column when ONE process drives all the GPUs (the executables): every GPU has gathered its
// column into a dense first-occurrence-order vector (identical order on all replicas); this kernel, run on every GPU,
// reads ALL replicas' vectors straight out of peer memory over NVLink (coalesced 128-bit loads), adds them with uint32
// wrap-around and scatters the sums into its own column - collective and scatter in one pass, no communicator, no staging.
__global__ void __launch_bounds__(256)
s2_peer_sum_scatter_kernel(S2PeerVecs pv, uint32_t *__restrict__ counts_col, const uint32_t *__restrict__ rank_slot, uint64_t n_keys)
{
    const uint64_t n4 = n_keys / 4;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * blockDim.x) {
        uint4 acc = make_uint4(0, 0, 0, 0);
        for (int p = 0; p < pv.n; ++p) {
            const uint4 v = __ldcv(reinterpret_cast<const uint4 *>(pv.v[p]) + i);       // peer memory: never from a stale cache line
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        const uint4 s = *(reinterpret_cast<const uint4 *>(rank_slot) + i);
        counts_col[s.x] = acc.x; counts_col[s.y] = acc.y; counts_col[s.z] = acc.z; counts_col[s.w] = acc.w;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n_keys & 3)) {
        const uint64_t r = n4 * 4 + threadIdx.x;
        uint32_t acc = 0;
        for (int p = 0; p < pv.n; ++p) acc += __ldcv(pv.v[p] + r);
        counts_col[rank_slot[r]] = acc;
    }
}

void s2_launch_peer_sum_scatter(const S2TableView &t, int col, const uint32_t *rank_slot, uint64_t n_keys, const S2PeerVecs &pv, cudaStream_t stream)
{
    if (n_keys) s2_peer_sum_scatter_kernel<<<148 * 8, 256, 0, stream>>>(pv, t.counts + (uint64_t)col * t.n_slots, rank_slot, n_keys);
}

void s2_launch_flag(const S2TableView &t, const uint64_t *kmers, uint64_t n, uint8_t *found, int set, cudaStream_t stream)
{
    if (n) s2_flag_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(t, kmers, n, found, set);
}

void s2_launch_lookup(const S2TableView &t, const uint64_t *kmers, uint64_t n, uint32_t *slot_out, cudaStream_t stream)
{
    if (n) s2_lookup_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(t, kmers, n, slot_out);
}

void s2_launch_counts_by_key(const S2TableView &t, int col, const uint64_t *kmers, uint64_t n, uint32_t *out, cudaStream_t stream)
{
    if (n) s2_counts_by_key_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(t, t.counts + (uint64_t)col * t.n_slots, kmers, n, out);
}

void s2_launch_pack(const uint8_t *bases, uint64_t n_bytes, uint32_t *words, uint16_t *masks, cudaStream_t stream)
{
    const uint64_t n_chunks = (n_bytes + 15) / 16;
    if (n_chunks) s2_pack_kernel<<<blocks_for(n_chunks, 256), 256, 0, stream>>>(bases, n_bytes, words, masks);
}

void s2_launch_fill_u32(uint32_t *p, uint64_t n, uint32_t v, cudaStream_t stream)
{
    if (n) s2_fill_u32_kernel<<<148 * 8, 256, 0, stream>>>(p, n, v);
}
