// s2_pgz.cpp - parallel gzip writer for the text outputs (second half).
//
// The reference writes kmer_hits through ONE zlib stream at level 9 (gzopen(outfile, "wb9"),
// /root/reference/src/strain_detect.c:299; gzprintf per hit line :567,:608) - about 20-30 MB/s, which becomes the
// tail of a run once the scan takes milliseconds and the strain is abundant.  With threads == 0 this writer is that
// same single stream (gzwrite on a "wb9" gzFile: the .gz stays byte-identical to the reference's).  With threads > 0 it
// compresses 256 KB blocks independently (raw DEFLATE at level 9, each primed with the 32 KB before it as dictionary,
// ended on a byte boundary with a sync flush) on a pool of threads and concatenates them into ONE gzip member with the
// CRC-32 / length of the whole text - the layout pigz writes.  Any gunzip yields the identical text; only the
// compressed bytes differ from the single-stream file.
#include "../../include/strainer2_b200.h"
#include "s2_internal.h"

#include <zlib.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#define PGZ_BLOCK (256u << 10)
#define PGZ_DICT 32768u

struct s2_gz_writer {
    int threads = 0;
    gzFile gz = nullptr;              // threads == 0
    FILE *fp = nullptr;               // threads > 0
    std::string pending;              // text not yet compressed (blocks are cut from its front)
    std::string dict;                 // last 32 KB of the text already compressed
    uint32_t crc = 0;
    uint64_t total = 0;
    bool failed = false;
};

static bool pgz_deflate_block(const char *data, size_t n, const char *dict, size_t dict_len, std::string &out)
{
    z_stream z; memset(&z, 0, sizeof z);
    if (deflateInit2(&z, 9, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
    if (dict_len) deflateSetDictionary(&z, (const Bytef *)dict, (uInt)dict_len);
    out.resize(deflateBound(&z, (uLong)n) + 16);
    z.next_in = (Bytef *)data; z.avail_in = (uInt)n;
    z.next_out = (Bytef *)&out[0]; z.avail_out = (uInt)out.size();
    const int rc = deflate(&z, Z_SYNC_FLUSH);                 // ends on a byte boundary, stream not finished
    const bool ok = rc == Z_OK && z.avail_in == 0;
    out.resize(out.size() - z.avail_out);
    deflateEnd(&z);
    return ok;
}

// compress every complete block of `pending` (all of it when `all`), in parallel, and write the pieces in order
static void pgz_drain(s2_gz_writer *w, bool all)
{
    const size_t n_blocks = all ? (w->pending.size() + PGZ_BLOCK - 1) / PGZ_BLOCK : w->pending.size() / PGZ_BLOCK;
    if (!n_blocks) return;
    const size_t n_bytes = all ? w->pending.size() : n_blocks * PGZ_BLOCK;
    const std::string text = w->dict + w->pending.substr(0, n_bytes);          // dictionary bytes in front of the blocks
    const size_t d0 = w->dict.size();
    std::vector<std::string> pieces(n_blocks);
    std::vector<char> ok(n_blocks, 0);
    auto work = [&](size_t first, size_t step) {
        for (size_t b = first; b < n_blocks; b += step) {
            const size_t off = d0 + b * PGZ_BLOCK;
            const size_t len = std::min<size_t>(PGZ_BLOCK, text.size() - off);
            const size_t dl = std::min<size_t>(PGZ_DICT, off);
            ok[b] = pgz_deflate_block(text.data() + off, len, text.data() + off - dl, dl, pieces[b]) ? 1 : 0;
        }
    };
    const size_t n_thr = std::min<size_t>((size_t)w->threads, n_blocks);
    std::vector<std::thread> pool;
    for (size_t t = 1; t < n_thr; ++t) pool.emplace_back(work, t, n_thr);
    work(0, n_thr);
    for (auto &t : pool) t.join();
    for (size_t b = 0; b < n_blocks; ++b) {
        if (!ok[b] || fwrite(pieces[b].data(), 1, pieces[b].size(), w->fp) != pieces[b].size()) w->failed = true;
    }
    w->crc = (uint32_t)crc32(w->crc, (const Bytef *)text.data() + d0, (uInt)n_bytes);
    w->total += n_bytes;
    w->dict = text.size() > PGZ_DICT ? text.substr(text.size() - PGZ_DICT) : text;
    w->pending.erase(0, n_bytes);
}

extern "C" s2_gz_writer *s2_gz_writer_open(const char *path, int threads)
{
    s2_gz_writer *w = new s2_gz_writer();
    w->threads = threads < 0 ? 0 : threads;
    if (w->threads == 0) {
        w->gz = gzopen(path, "wb9");
        if (!w->gz) { delete w; return nullptr; }
        return w;
    }
    w->fp = fopen(path, "wb");
    if (!w->fp) { delete w; return nullptr; }
    static const unsigned char head[10] = { 0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 2, 3 };      // no name, no time, XFL = best, OS = unix
    fwrite(head, 1, sizeof head, w->fp);
    w->crc = (uint32_t)crc32(0L, Z_NULL, 0);
    return w;
}

extern "C" int s2_gz_writer_write(s2_gz_writer *w, const void *data, uint64_t n)
{
    if (!w) return -1;
    if (w->threads == 0) {
        const char *p = (const char *)data;
        while (n) {
            const unsigned step = (unsigned)std::min<uint64_t>(n, 1u << 30);
            if (gzwrite(w->gz, p, step) != (int)step) { w->failed = true; return -1; }
            p += step; n -= step;
        }
        return 0;
    }
    w->pending.append((const char *)data, (size_t)n);
    if (w->pending.size() >= (size_t)w->threads * PGZ_BLOCK * 4) pgz_drain(w, false);      // enough for every thread to have work
    return w->failed ? -1 : 0;
}

extern "C" int s2_gz_writer_close(s2_gz_writer *w)
{
    if (!w) return -1;
    int rc = 0;
    if (w->threads == 0) {
        rc = gzclose(w->gz) == Z_OK && !w->failed ? 0 : -1;
    } else {
        pgz_drain(w, true);
        static const unsigned char fin[2] = { 0x03, 0x00 };                             // final empty fixed-Huffman block
        unsigned char tail[8];
        const uint32_t isize = (uint32_t)(w->total & 0xFFFFFFFFu);
        for (int i = 0; i < 4; ++i) { tail[i] = (unsigned char)(w->crc >> (8 * i)); tail[4 + i] = (unsigned char)(isize >> (8 * i)); }
        if (fwrite(fin, 1, 2, w->fp) != 2 || fwrite(tail, 1, 8, w->fp) != 8) w->failed = true;
        if (fclose(w->fp) != 0) w->failed = true;
        rc = w->failed ? -1 : 0;
    }
    delete w;
    return rc;
}
