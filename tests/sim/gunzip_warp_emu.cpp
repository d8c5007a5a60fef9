// TEST-ONLY: the DEVICE control flow of strainer2_b200/csrc/s2_gunzip.cuh on the CPU - 32 threads are the 32 lanes of a warp,
// every one runs gz_subchunk(..., lane, 32) on the warp's shared tables, symbol region and queue, and the places where the
// lanes of a warp synchronise (GZ_SYNC = __syncwarp, GZ_FENCE = "the lanes run in step", GZ_BALLOT, GZ_BCAST0) are barriers.
// What the one-lane host harness cannot see runs here: the block finder's queue of survivors, the lane's share of a match
// copy (lanes behind the end of a match repeating its last lane, the deferred store, the guard slot), table builds and
// header parsing spread over 32 lanes.  The text must be zlib's.  Usage: gunzip_warp_emu [1] (tests/test_host.py builds and runs it).
#include <pthread.h>
#include <atomic>
#include <cstdint>

struct EmuWarp {
    pthread_barrier_t bar;
    std::atomic<uint32_t> mask[2];
    uint32_t bcast;
};
static thread_local EmuWarp *tl_warp;
static thread_local int tl_lane;
static thread_local unsigned tl_ballots;
static inline void emu_sync() { pthread_barrier_wait(&tl_warp->bar); }
static inline uint32_t emu_ballot(bool p)
{
    EmuWarp *w = tl_warp;
    std::atomic<uint32_t> &m = w->mask[tl_ballots++ & 1u];
    if (p) m.fetch_or(1u << tl_lane);
    emu_sync();
    const uint32_t v = m.load();
    emu_sync();
    if (tl_lane == 0) m.store(0);            // (its next use is two ballots away: behind the other word's barriers)
    return v;
}
template <class T> static inline T emu_bcast0(T v)
{
    EmuWarp *w = tl_warp;
    if (tl_lane == 0) w->bcast = (uint32_t)v;
    emu_sync();
    const T r = (T)w->bcast;
    emu_sync();
    return r;
}
#define GZ_EMULATE_WARP 1
#define GZ_SYNC() emu_sync()
#define GZ_FENCE() emu_sync()
#define GZ_BALLOT(p) emu_ballot(p)
#define GZ_BCAST0(v) emu_bcast0(v)
#include "../../strainer2_b200/csrc/s2_gunzip.cuh"
#include "../../strainer2_b200/csrc/s2_inflate.cuh"      // s2_gzip_header_len

#include <zlib.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>

struct LaneArgs {
    EmuWarp *warp; int lane;
    const uint32_t *words; uint64_t n_words, known_start, cut, next_cut; uint16_t *out; uint32_t cap; GzTables *t; const uint8_t *kraft9; GzSubResult *res;
};
static void *lane_main(void *p)
{
    LaneArgs *a = (LaneArgs *)p;
    tl_warp = a->warp; tl_lane = a->lane; tl_ballots = 0;
    gz_subchunk(a->words, a->n_words, a->known_start, a->cut, a->next_cut, 8ull << 20, a->out, a->cap, *a->t, a->kraft9, a->res, a->lane, 32);
    return nullptr;
}

// one .gz -> text through warp-emulated sub-chunks + the host's chain / translate; 0 = equal to `want`
static int run(const std::vector<uint8_t> &gz, const std::vector<uint8_t> &want, uint32_t sub_bytes, long *n_marker_syms)
{
    const uint64_t hl = s2_gzip_header_len(gz.data(), gz.size());
    if (!hl) return -100;
    const uint64_t n = gz.size(), n_words = (n + 3) / 4;
    std::vector<uint32_t> words(n_words + 4, 0u);
    memcpy(words.data(), gz.data(), n);
    const uint64_t n_sub = (n + sub_bytes - 1) / sub_bytes;
    const uint32_t sub_cap = (uint32_t)want.size() + (1u << 16) + GZ_WINDOW;
    std::vector<uint16_t> sym_alloc((size_t)n_sub * sub_cap + GZ_WINDOW);
    for (uint64_t i = 0; i < n_sub; ++i) gz_marker_prefix(sym_alloc.data() + i * sub_cap, 0, 1);
    uint16_t *const sym = sym_alloc.data() + GZ_WINDOW;
    std::vector<GzSubResult> res(n_sub);
    static GzTables tables;
    static uint8_t kraft9[512];
    gz_kraft9_fill(kraft9, 0, 1);
    for (uint64_t i = 0; i < n_sub; ++i) {
        EmuWarp warp;
        pthread_barrier_init(&warp.bar, nullptr, 32);
        warp.mask[0] = 0; warp.mask[1] = 0; warp.bcast = 0;
        LaneArgs args[32];
        pthread_t th[32];
        for (int l = 0; l < 32; ++l) {
            args[l] = { &warp, l, words.data(), n_words, i == 0 ? hl * 8 : ~0ull, i * sub_bytes * 8ull, (i + 1) * sub_bytes * 8ull, sym + i * sub_cap,
                        sub_cap - GZ_WINDOW - 1u, &tables, kraft9, &res[i] };
            if (pthread_create(&th[l], nullptr, lane_main, &args[l])) return -200;
        }
        for (int l = 0; l < 32; ++l) pthread_join(th[l], nullptr);
        pthread_barrier_destroy(&warp.bar);
    }
    // chain + translate as the one-lane harness does them
    uint64_t cur = hl * 8, total = 0;
    bool done = false;
    std::vector<uint8_t> win((size_t)(n_sub + 1) * GZ_WINDOW, 0), text(want.size() + 1);
    for (uint64_t i = 0; i < n_sub && !done; ++i) {
        uint8_t *prev = win.data() + i * GZ_WINDOW, *next = win.data() + (i + 1) * GZ_WINDOW;
        if (res[i].start_bit != cur) return -101;
        if (res[i].status < 0) return res[i].status;
        if (total + res[i].n_out > want.size()) return -103;
        for (uint32_t k = 0; k < res[i].n_out; ++k) *n_marker_syms += sym[i * sub_cap + k] >= 256;
        gz_translate(prev, sym + i * sub_cap, res[i].n_out, text.data() + total, 0, 1);
        gz_next_window(prev, sym + i * sub_cap, res[i].n_out, next, 0, 1);
        total += res[i].n_out;
        cur = res[i].end_bit;
        done = res[i].status == GZ_FINAL;
    }
    if (!done) return -102;
    if (total != want.size() || memcmp(text.data(), want.data(), total)) return -104;
    return 0;
}

int main(int argc, char **argv)
{
    const bool full = argc > 1 && atoi(argv[1]) > 0;                      // 1: larger texts, every combination (three minutes)
    unsigned long long s = 0x2545F4914F6CDD1Dull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    long runs = 0, markers = 0;
    for (int kind = 0; kind < 4; ++kind) {
        std::vector<uint8_t> t;
        const size_t n = (size_t)(full ? 150000 : 45000) + rnd() % 20000;
        if (kind == 0) while (t.size() < n) { for (int k = 0; k < 80; ++k) t.push_back("ACGT"[rnd() & 3]); t.push_back('\n'); }                      // FASTA
        else if (kind == 1) while (t.size() < n) { t.push_back('@'); for (int k = 0; k < 100; ++k) t.push_back("ACGT"[rnd() & 3]); t.push_back('\n'); t.push_back('+'); t.push_back('\n');
                                                   for (int k = 0; k < 100; ++k) t.push_back(rnd() % 5 ? 'I' : (uint8_t)('#' + rnd() % 40)); t.push_back('\n'); }   // FASTQ
        else if (kind == 2) { std::vector<uint8_t> unit; for (int k = 0; k < 37; ++k) unit.push_back("ACGT"[rnd() & 3]);                               // long, self-overlapping matches
                              while (t.size() < n) { t.insert(t.end(), unit.begin(), unit.end()); if (rnd() % 50 == 0) t.push_back('N'); } }
        else while (t.size() < n) t.push_back((uint8_t)(rnd() % 3 ? "ACGT\n"[rnd() % 5] : rnd()));                                                    // literals with long codes
        for (int lvl : { 1, 6, 9 }) {
            if (kind && lvl == 9) continue;
            if (!full && kind && lvl != 6) continue;
            z_stream z; memset(&z, 0, sizeof z);
            deflateInit2(&z, lvl, Z_DEFLATED, 31, 8, Z_DEFAULT_STRATEGY);
            std::vector<uint8_t> c(deflateBound(&z, t.size()) + 64);
            z.next_in = t.data(); z.avail_in = (uInt)t.size(); z.next_out = c.data(); z.avail_out = (uInt)c.size();
            deflate(&z, Z_FINISH); c.resize(z.total_out); deflateEnd(&z);
            for (uint32_t sub : { 8192u, 65536u }) {
                if (sub == 65536u && (kind == 2 || lvl == 1 || (!full && kind))) continue;
                const int rc = run(c, t, sub, &markers);
                if (rc) { printf("kind %d level %d sub-chunk %u: rc %d\n", kind, lvl, sub, rc); return 1; }
                ++runs;
            }
        }
    }
    printf("warp emulation done: %ld streams decoded by 32 lanes each are zlib's text (%ld marker symbols resolved)\n", runs, markers);
    return 0;
}
