"""Damaged compressed input through the executables (GPU).  Kept in a file of its own that sorts AFTER every other
test module: the damaged BGZF member makes the hardware decompression engine fault inside the child process, and while
that is confined to the child's CUDA context in every run so far, this way the test process holds no context of its
own at that moment (the other modules' fixtures are closed) and nothing runs after it."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def s2():
    import strainer2_b200 as s2
    return s2


def test_executables_fail_loudly_on_damaged_gzip_data(s2, tmp_path, capsys):
    """Corrupt DEFLATE data inside an intact container.  The reference never returns from such a file (kseq re-reads
    the gzread error for ever, src/kseq.h:72,:99), so there is no output to match; what must not happen is a table
    that silently misses part of a file.  GPU ingest path: the hardware engine either reports a wrong length (verdict
    irregular -> host reader -> zlib finds the damage) or takes the CUDA context down (sticky launch failure); host
    path (ordinary .gz): zlib finds it.  Every way ends in exit code 1, a message that names the cause, no table."""
    import gzip
    from strainer2_b200 import synth
    tmp = str(tmp_path)
    rng = synth.rng_for(7, 6)
    strain = synth.genome(rng, 300_000, 4, n_runs=2)
    clean = [np.where(c == ord("N"), ord("A"), c).astype(np.uint8) for c in strain]
    reads = synth.sample_reads(rng, clean + synth.genome(rng, 600_000, 2), 30_000, 150, sub_rate=0.005, n_rate=1e-4)
    synth.write_fasta(os.path.join(tmp, "strain.fa"), strain)
    text = synth.fastq_bytes(reads)
    good = synth.bgzf_bytes(text)
    members, off = [], 0
    while off < len(good):
        bsize = int.from_bytes(good[off + 16:off + 18], "little") + 1
        members.append((off, bsize))
        off += bsize
    o, b = members[len(members) // 2]
    junk = bytes((((i * 2654435761) & 0xFFFFFFFF) >> 13) & 0xFF for i in range(b - 26))
    open(os.path.join(tmp, "bad_bgzf.fastq.gz"), "wb").write(good[:o + 18] + junk + good[o + b - 8:])
    z = bytearray(gzip.compress(text, 6))
    z[-6] ^= 1                                                    # ordinary gzip, wrong CRC-32
    open(os.path.join(tmp, "bad_gzip.fastq.gz"), "wb").write(bytes(z))
    open(os.path.join(tmp, "good.fastq.gz"), "wb").write(good)
    open(os.path.join(tmp, "A.txt"), "w").write("")
    with open(os.path.join(tmp, "inf.txt"), "wb") as f:
        c0 = bytes(strain[0]).replace(b"N", b"A")
        for i in range(0, len(c0) - 31, 500):
            f.write(c0[i:i + 31] + b"\n")
    for bad in ("bad_bgzf.fastq.gz", "bad_gzip.fastq.gz"):
        open(os.path.join(tmp, "B.txt"), "w").write("good.fastq.gz\n%s\ngood.fastq.gz\n" % bad)
        for env in ({}, {"S2_GPU_INGEST": "0"}):
            p = s2.run_kmer_scrub_count(["-r", "strain.fa", "-A", "A.txt", "-B", "B.txt"], cwd=tmp, env=env, timeout=120)
            with capsys.disabled():
                print("\n[%s %s] kmer_scrub_count rc=%d: %s" % (bad, env, p.returncode, p.stderr.decode(errors="replace").strip()[-300:]))
            assert p.returncode == 1 and p.stdout == b"", (bad, env)
            assert b"damaged" in p.stderr, (bad, env)
            q = s2.run_strain_detect(["-r", "strain.fa", "-a", "inf.txt", "-b", bad, "-t", "SE", "-o", "hits.gz"], cwd=tmp, env=env, timeout=120)
            with capsys.disabled():
                print("[%s %s] strain_detect rc=%d: %s" % (bad, env, q.returncode, q.stderr.decode(errors="replace").strip()[-300:]))
            assert q.returncode == 1 and b"damaged" in q.stderr, (bad, env)
    # the same lists without the damaged file still work afterwards (a fresh process has a fresh context)
    open(os.path.join(tmp, "B.txt"), "w").write("good.fastq.gz\n")
    assert s2.run_kmer_scrub_count(["-r", "strain.fa", "-A", "A.txt", "-B", "B.txt"], cwd=tmp).returncode == 0


@pytest.mark.parametrize("sub_kb", [8, 64, 256])
def test_gpu_gunzip_of_ordinary_gz_files_equals_host_reader(s2, tmp_path, monkeypatch, sub_kb):
    """ordinary single-member .gz files (FASTA genomes at several compression levels, FASTQ files of several MB) decoded by
    the chunk-parallel gunzip (s2_gunzip.cu: block finder, speculative decode, chain, translate, CRC-32) inside the ingest
    pipeline: counters equal the host reader's (zlib); files the decoder cannot vouch for (a second member, a wrong ISIZE, a
    wrong CRC-32, a flipped bit, a truncated stream) are handed back untouched"""
    import gzip
    from strainer2_b200 import synth
    monkeypatch.setenv("S2_GZ_SUB_KB", str(sub_kb))
    tmp = str(tmp_path)
    rng = synth.rng_for(7, 11)
    strain = synth.genome(rng, 300_000, 4, n_runs=2)
    clean = [np.where(c == ord("N"), ord("A"), c).astype(np.uint8) for c in strain]
    synth.write_fasta(os.path.join(tmp, "strain.fa"), strain)
    ctx = s2.Context(0, batch_bytes=8 << 20, n_lanes=2)
    t = s2.StrainTable(ctx, s2.load_flat(os.path.join(tmp, "strain.fa")), n_cols=4)
    paths = []
    for i in range(24):                                           # relatives and strangers, 0.3 - 0.6 Mb each
        g = [c.copy() for c in clean] if i % 3 == 0 else synth.genome(rng, 300_000 + 10_000 * i, 3)
        p = os.path.join(tmp, "g%d.fa.gz" % i)
        open(p, "wb").write(gzip.compress(synth.fasta_bytes(g, 80), (1, 6, 9)[i % 3]))
        paths.append(p)
    open(os.path.join(tmp, "empty.fa.gz"), "wb").write(gzip.compress(b"", 6))
    fqs = []
    for k, n_reads in enumerate((20_000, 150_000)):                # 6 MB and 47 MB of FASTQ text
        reads = synth.sample_reads(rng, clean + synth.genome(rng, 600_000, 2), n_reads, 150, sub_rate=0.005, n_rate=1e-4)
        fq = os.path.join(tmp, "m%d.fastq.gz" % k)
        open(fq, "wb").write(gzip.compress(synth.fastq_bytes(reads), 6 if k else 1))
        fqs.append((fq, reads.size))
    want_hits = 0
    for p in paths:
        want_hits += ctx.scan_count(t, s2.load_flat(p), 1).hits
    rc, bases, lookups = ctx.ingest_count_files(t, paths, 2)
    st = ctx.sync()
    assert rc == [0] * len(paths)
    assert st.hits == want_hits > 1000 and np.array_equal(t.counts(1), t.counts(2))
    for fq, n_bases in fqs:
        t.clear_counts(1); t.clear_counts(2)
        want = ctx.scan_count(t, s2.load_flat(fq), 1)
        rc, bases, lookups = ctx.ingest_count_files(t, [fq], 2)
        st = ctx.sync()
        assert rc == [0] and bases == n_bases and st.hits == want.hits and np.array_equal(t.counts(1), t.counts(2))
    # the same as file images in host memory, genomes and reads in one call
    t.clear_counts(1); t.clear_counts(2)
    want_hits = sum(ctx.scan_count(t, s2.load_flat(p), 1).hits for p in paths[:6] + [fqs[0][0]])
    images = [np.fromfile(p, dtype=np.uint8) for p in paths[:6] + [fqs[0][0], os.path.join(tmp, "empty.fa.gz")]]
    rc, bases, lookups = ctx.ingest_count_mem_batch(t, [im.ctypes.data for im in images], [im.size for im in images], 2)
    st = ctx.sync()
    assert list(rc)[:7] == [0] * 7 and st.hits == want_hits and np.array_equal(t.counts(1), t.counts(2))
    # not vouched for: nothing counted, rc 1
    t.clear_counts(2)
    z = gzip.compress(synth.fasta_bytes(clean, 80), 6)
    mid = len(z) // 2
    bad = {"two_members.fa.gz": z + z, "wrong_isize.fa.gz": z[:-4] + b"\x01\x00\x00\x00", "wrong_crc.fa.gz": z[:-8] + bytes([z[-8] ^ 1]) + z[-7:],
           "flipped.fa.gz": z[:mid] + bytes([z[mid] ^ 0x10]) + z[mid + 1:], "cut.fa.gz": z[:mid] + z[-8:]}
    for name, data in bad.items():
        open(os.path.join(tmp, name), "wb").write(data)
    rc, _, _ = ctx.ingest_count_files(t, [os.path.join(tmp, n) for n in bad] + paths[:2], 2)
    st = ctx.sync()
    assert rc[:len(bad)] == [1] * len(bad) and rc[len(bad):] == [0, 0], rc
    want_two = sum(ctx.scan_count(t, s2.load_flat(p), 3).hits for p in paths[:2])
    assert st.hits == want_two and np.array_equal(t.counts(2), t.counts(3))
    t.free()
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("batch_mb,text_mb", [(4, 8), (2, 64)])
def test_gpu_gunzip_streams_a_big_gz_file_in_pieces(s2, tmp_path, monkeypatch, batch_mb, text_mb):
    """one ordinary .gz file larger than a gz batch (config #3's shape: one long DEFLATE stream of FASTQ) goes through the gz
    stage in PIECES - the block finder starts every later piece, the chain and the 32 KB window carry over, the stream's
    size and CRC-32 are checked over all pieces - and each piece's text through the ring in slices; FASTA alike.  A stream
    whose CRC-32 is wrong at the very end is taken back out (increment -1) and handed to the host reader"""
    import gzip
    from strainer2_b200 import synth
    monkeypatch.setenv("S2_GZ_BATCH_MB", str(batch_mb))
    monkeypatch.setenv("S2_INGEST_TEXT_MB", str(text_mb))
    tmp = str(tmp_path)
    rng = synth.rng_for(17, 3)
    strain = synth.genome(rng, 300_000, 4, n_runs=2)
    clean = [np.where(c == ord("N"), ord("A"), c).astype(np.uint8) for c in strain]
    synth.write_fasta(os.path.join(tmp, "strain.fa"), strain)
    ctx = s2.Context(0, batch_bytes=8 << 20, n_lanes=2)
    t = s2.StrainTable(ctx, s2.load_flat(os.path.join(tmp, "strain.fa")), n_cols=4)
    reads = synth.sample_reads(rng, clean + synth.genome(rng, 600_000, 2), 150_000, 150, sub_rate=0.005, n_rate=1e-4)
    fq = os.path.join(tmp, "big.fastq.gz")
    zq = gzip.compress(synth.fastq_bytes(reads), 6)                     # 47 MB of text, about 10 MB of .gz: several pieces
    open(fq, "wb").write(zq)
    fa = os.path.join(tmp, "big.fa.gz")
    genomes = [c for i in range(6) for c in (clean if i % 2 else synth.genome(rng, 2_000_000, 3))]
    open(fa, "wb").write(gzip.compress(synth.fasta_bytes(genomes, 80), 6))
    assert os.path.getsize(fq) > 2 * (batch_mb << 20) * 3 // 4
    for path, n_bases in ((fq, reads.size), (fa, sum(c.size for c in genomes))):
        t.clear_counts(1); t.clear_counts(2)
        want = ctx.scan_count(t, s2.load_flat(path), 1)
        rc, bases, lookups = ctx.ingest_count_files(t, [path], 2)
        st = ctx.sync()
        assert rc == [0] and bases == n_bases, (path, rc, bases)
        assert st.hits == want.hits > 1000 and np.array_equal(t.counts(1), t.counts(2))
        # as an image in host memory
        t.clear_counts(2)
        image = np.fromfile(path, dtype=np.uint8)
        rc, bases, lookups = ctx.ingest_count_mem_batch(t, [image.ctypes.data], [image.size], 2)
        st = ctx.sync()
        assert list(rc) == [0] and st.hits == want.hits and np.array_equal(t.counts(1), t.counts(2))
    # damaged at the very end / in the middle / a second member behind it: nothing stays counted
    t.clear_counts(2)
    mid = len(zq) // 2
    bad = {"wrong_crc.fastq.gz": zq[:-8] + bytes([zq[-8] ^ 1]) + zq[-7:], "wrong_isize.fastq.gz": zq[:-4] + bytes([zq[-4] ^ 1]) + zq[-3:],
           "flipped.fastq.gz": zq[:mid] + bytes([zq[mid] ^ 0x10]) + zq[mid + 1:], "two_members.fastq.gz": zq + zq, "cut.fastq.gz": zq[:mid]}
    for name, data in bad.items():
        open(os.path.join(tmp, name), "wb").write(data)
        rc, _, _ = ctx.ingest_count_files(t, [os.path.join(tmp, name)], 2)
        st = ctx.sync()
        assert rc == [1], (name, rc)
        assert not t.counts(2).any(), name
    t.free()
    ctx.close()


@pytest.mark.gpu
def test_gpu_gunzip_symbol_region_that_overflows_hands_the_file_back(s2, tmp_path):
    """a .gz whose sub-chunks inflate to more symbols than a region holds (a short unit repeated: 250 : 1) - literals and the
    deferred copies of the decoder store into the region's guard slot from then on, the sub-chunk reports the overflow, the
    file is NOT handled (rc 1, nothing counted: the host reader takes it) - and the files decoded beside it in the same
    launch, whose regions lie behind the overflowing ones, are counted exactly as the host reader counts them"""
    import gzip
    from strainer2_b200 import synth
    tmp = str(tmp_path)
    rng = synth.rng_for(23, 5)
    strain = synth.genome(rng, 300_000, 4)
    synth.write_fasta(os.path.join(tmp, "strain.fa"), strain)
    ctx = s2.Context(0, batch_bytes=8 << 20, n_lanes=2)
    t = s2.StrainTable(ctx, s2.load_flat(os.path.join(tmp, "strain.fa")), n_cols=4)
    unit = bytes(strain[0][:1000])
    rep = os.path.join(tmp, "repeat.fa.gz")
    open(rep, "wb").write(gzip.compress(b">r\n" + b"\n".join(unit * 50 for _ in range(400)) + b"\n", 6))      # 20 MB of text in 80 KB
    good = []
    for i in range(3):
        p = os.path.join(tmp, "g%d.fa.gz" % i)
        open(p, "wb").write(gzip.compress(synth.fasta_bytes([c.copy() for c in strain] if i == 1 else synth.genome(rng, 400_000, 3), 80), 6))
        good.append(p)
    want = sum(ctx.scan_count(t, s2.load_flat(p), 1).hits for p in good)
    rc, bases, lookups = ctx.ingest_count_files(t, [good[0], rep, good[1], rep, good[2]], 2)
    st = ctx.sync()
    assert list(rc) == [0, 1, 0, 1, 0], rc
    assert st.hits == want > 1000 and np.array_equal(t.counts(1), t.counts(2))
    # through the executable the file is simply read by the host instead: same table as with the GPU ingest switched off
    open(os.path.join(tmp, "A.txt"), "w").write("repeat.fa.gz\ng1.fa.gz\n")
    open(os.path.join(tmp, "B.txt"), "w").write("")
    a = s2.run_kmer_scrub_count(["-r", "strain.fa", "-A", "A.txt", "-B", "B.txt"], cwd=tmp)
    b = s2.run_kmer_scrub_count(["-r", "strain.fa", "-A", "A.txt", "-B", "B.txt"], cwd=tmp, env={"S2_GPU_INGEST": "0"})
    assert a.returncode == 0 and b.returncode == 0 and a.stdout == b.stdout and len(a.stdout) > 1000
    t.free()
    ctx.close()
