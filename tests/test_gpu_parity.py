"""GPU parity tests: the CUDA path, called through the C ABI (ctypes) and through the drop-in
executables, against the CPU oracle and the reference-generated golden vectors.  Bit-exact: this path
is integer / byte work, there is no tolerance anywhere."""
import os
import subprocess

import numpy as np
import pytest

import oracle_util as ou

pytestmark = pytest.mark.gpu

CODE = {65: 0, 67: 1, 71: 2, 84: 3}


@pytest.fixture(scope="module")
def s2():
    import strainer2_b200 as s2
    return s2


@pytest.fixture(scope="module")
def ctx(s2):
    c = s2.Context(0, batch_bytes=8 << 20, n_lanes=3)
    yield c
    c.close()


def table_bytes(s2, table, n_print, tmp_path, name="t.tsv"):
    keys, djb2 = table.export()
    order, _ = s2.roworder_emulate(djb2)
    out = os.path.join(str(tmp_path), name)
    s2.format_count_table(out, keys, order, [table.counts(c) for c in range(n_print)])
    return open(out, "rb").read()


# ---------------------------------------------------------------------------------------------
def test_pack_kernel_matches_definition(s2, ctx):
    rng = np.random.default_rng(1)
    for n in (0, 1, 15, 16, 17, 511, 512, 513, 100003):
        b = rng.choice(np.frombuffer(b"ACGTacgtNn\n>RY", dtype=np.uint8), size=n)
        if n > 40:
            b[7:30] = rng.integers(0, 256, size=23, dtype=np.uint8)
        words, masks = ctx.pack_2bit(b)
        assert words.size == (n + 15) // 16
        pad = np.concatenate([b, np.full((-n) % 16, ord("\n"), np.uint8)]).reshape(-1, 16)
        up = np.where((pad >= 97) & (pad <= 122), pad - 32, pad)
        ok = np.isin(up, [65, 67, 71, 84])
        code = np.zeros_like(up, dtype=np.uint32)
        for a, c in CODE.items():
            code[up == a] = c
        want_m = (ok.astype(np.uint32) << (15 - np.arange(16, dtype=np.uint32))).sum(axis=1)
        assert np.array_equal(masks.astype(np.uint32), want_m)
        want_w = (np.where(ok, code, 0) << (30 - 2 * np.arange(16, dtype=np.uint32))).sum(axis=1).astype(np.uint32)
        keep = np.repeat(ok.astype(np.uint32) * 3, 1, axis=1)
        sel = (keep << (30 - 2 * np.arange(16, dtype=np.uint32))).sum(axis=1).astype(np.uint32)
        assert np.array_equal(words & sel, want_w)


def test_table_build_matches_oracle_on_golden_genome(s2, ctx, golden_dir, tmp_path):
    d = os.path.join(golden_dir, "count_edge")
    ref = os.path.join(d, "ref.fa.gz")
    t = s2.StrainTable(ctx, s2.load_flat(ref), n_cols=4)
    o = ou.OracleTable(4)
    o.build(ref)
    assert t.n_keys == o.size
    got = table_bytes(s2, t, 3, tmp_path)
    assert got == o.table_text(False, str(tmp_path / "o.tsv"))
    # export: keys really are in first-occurrence order and djb2 is hashU of the spelling
    keys, djb2 = t.export()
    L = ou.lib()
    for k, h in list(zip(keys, djb2))[:200]:
        assert L.s2o_djb2(s2.kmer_to_ascii(int(k))) == int(h)
    assert np.unique(keys).size == keys.size
    t.free(); o.free()


@pytest.mark.parametrize("name,args", [
    ("ABC", ["-r", "ref.fa.gz", "-A", "listA.txt", "-B", "listB.txt", "-C", "listC.txt"]),
    ("AB", ["-r", "ref.fa.gz", "-A", "listA.txt", "-B", "listB.txt"]),
    ("A_only", ["-r", "ref.fa.gz", "-A", "listA.txt", "-B", "listB_empty.txt"]),
])
@pytest.mark.parametrize("threads", ["1", "4"])
def test_kmer_scrub_count_executable_matches_reference_bytes(s2, golden_dir, tmp_path, name, args, threads):
    d = os.path.join(golden_dir, "count_edge")
    prog = str(tmp_path / "progress")
    p = s2.run_kmer_scrub_count(args + ["-p", prog], cwd=d, env={"S2_THREADS": threads, "S2_BATCH_MB": "1"})
    assert p.returncode == 0, p.stderr
    assert p.stdout == open(os.path.join(d, f"expected_{name}.tsv"), "rb").read()
    assert p.stderr == open(os.path.join(d, f"expected_{name}.stderr"), "rb").read()
    exp = os.path.join(d, f"expected_{name}.progress")
    if os.path.exists(exp):
        assert ou.mask_progress(open(prog).read()) == open(exp).read()


def test_kmer_scrub_count_executable_error_paths(s2, golden_dir):
    d = os.path.join(golden_dir, "count_edge")
    p = s2.run_kmer_scrub_count(["-r", "ref.fa.gz", "-A", "listA.txt"], cwd=d)
    assert p.returncode == 1 and p.stdout == b""
    assert p.stderr == open(os.path.join(d, "expected_usage.stderr"), "rb").read()
    p = s2.run_kmer_scrub_count(["-r", "ref.fa.gz", "-A", "listA_missing.txt", "-B", "listB_empty.txt"], cwd=d,
                                env={"S2_THREADS": "1"})
    assert p.returncode == 1 and p.stdout == b""
    assert p.stderr == open(os.path.join(d, "expected_missing.stderr"), "rb").read()
    p = s2.run_kmer_scrub_count(["-r", "nope.fa", "-A", "listA.txt", "-B", "listB.txt"], cwd=d)
    assert p.returncode == 1 and p.stdout == b""
    assert p.stderr == b"could not read file nope.fa GEN_hash_sequences_set_count_vec()\n"


def _write_inputs(s2, tmp, strain_bp, n_rel, n_rand, n_reads, seed=0):
    from strainer2_b200 import synth
    rng = synth.rng_for(2, seed)
    strain = synth.genome(rng, strain_bp, 8, n_runs=6)
    synth.write_fasta(os.path.join(tmp, "strain.fa.gz"), strain)
    A, B = [], []
    for i in range(n_rel):
        g = [synth.mutate(c, 0.005 + 0.01 * i, rng) for c in strain]
        p = os.path.join(tmp, f"rel{i}.fa.gz")
        synth.write_fasta(p, g)
        A.append(p)
    for i in range(n_rand):
        p = os.path.join(tmp, f"rand{i}.fa")
        synth.write_fasta(p, synth.genome(rng, strain_bp, 5))
        A.append(p)
    clean = [np.where(c == ord("N"), ord("A"), c).astype(np.uint8) for c in strain]
    for i in range(2):
        other = synth.genome(rng, strain_bp, 4)
        reads = synth.sample_reads(rng, clean + other * 3, n_reads, 150, sub_rate=0.005, n_rate=1e-4)
        p = os.path.join(tmp, f"meta{i}.fastq.gz")
        synth.write_reads_fastq(p, reads)
        B.append(p)
    open(os.path.join(tmp, "A.txt"), "w").write("".join(a + "\n" for a in A))
    open(os.path.join(tmp, "B.txt"), "w").write("".join(b + "\n" for b in B))
    open(os.path.join(tmp, "C.txt"), "w").write(os.path.join(tmp, "strain.fa.gz") + "\n" + A[0] + "\n")
    return ["-r", os.path.join(tmp, "strain.fa.gz"), "-A", os.path.join(tmp, "A.txt"), "-B", os.path.join(tmp, "B.txt"),
            "-C", os.path.join(tmp, "C.txt")]


def test_synthetic_count_run_matches_oracle(s2, tmp_path):
    """down-scaled configs #2/#3: 400 kb strain, relatives + random genomes, FASTQ.gz metagenomes"""
    args = _write_inputs(s2, str(tmp_path), 400_000, 3, 2, 30_000)
    o = ou.oracle_cli(["count"] + args)
    p = s2.run_kmer_scrub_count(args, env={"S2_BATCH_MB": "2"})
    assert o.returncode == 0 and p.returncode == 0, p.stderr
    assert p.stdout == o.stdout
    assert p.stderr == o.stderr
    k, v = ou.parse_table(p.stdout)
    assert v[:, 1].sum() > 100000 and v[:, 2].sum() > 1000 and v[:, 3].sum() > 0       # the case is not vacuous


def test_five_megabase_strain_crosses_the_table_doubling(s2, tmp_path):
    """5 Mb strain => > 4,000,000 keys => the reference table doubles once (8M -> 16M): row order
    must still be identical (BASELINE config #2 strain shape, 2 genomes instead of 2,000)."""
    args = _write_inputs(s2, str(tmp_path), 5_000_000, 1, 1, 20_000, seed=1)
    o = ou.oracle_cli(["count"] + args)
    p = s2.run_kmer_scrub_count(args)
    assert o.returncode == 0 and p.returncode == 0, p.stderr
    assert p.stdout.count(b"\n") > 4_000_001
    assert p.stdout == o.stdout


def test_scan_api_host_and_device_inputs_agree_and_are_additive(s2, ctx):
    import torch
    from strainer2_b200 import synth
    rng = synth.rng_for(3, 7)
    strain = synth.genome(rng, 200_000, 4, n_runs=3)
    flat = synth.contigs_to_flat(strain)
    t = s2.StrainTable(ctx, flat, n_cols=4)
    clean = [np.where(c == ord("N"), ord("C"), c).astype(np.uint8) for c in strain]
    reads = synth.sample_reads(rng, clean + synth.genome(rng, 200_000, 2), 40_000, 150, sub_rate=0.01, n_rate=1e-3)
    batch = synth.reads_to_flat(reads)
    st_host = ctx.scan_count(t, batch, 1)
    dev = torch.from_numpy(batch).cuda()
    st_dev = ctx.scan_count(t, dev, 2)
    c1, c2 = t.counts(1), t.counts(2)
    assert st_host.hits == st_dev.hits == int(c1.sum()) == int(c2.sum()) > 0
    assert st_host.valid_windows == st_dev.valid_windows
    assert np.array_equal(c1, c2)
    # additivity: scanning two halves (split on a record boundary) into one column = the whole
    half = (reads.shape[0] // 2) * 151
    ctx.scan_count(t, batch[:half], 3)
    ctx.scan_count(t, batch[half:], 3)
    assert np.array_equal(t.counts(3), c1)
    # every ragged length: the tail masking must not invent or lose windows
    t.clear_counts(3)
    total = 0
    for n in (0, 1, 30, 31, 32, 47, 48, 511, 512, 513, 1000, 4097):
        total += ctx.scan_count(t, dev[:n], 3).valid_windows
        want = sum(1 for i in range(max(0, n - 30)) if all(c in b"ACGTacgt" for c in batch[i:i + 31].tobytes()))
        assert ctx.scan_count(t, batch[:n].copy(), 3).valid_windows == want, n
    t.free()


def test_scan_counts_equal_oracle_dictionary(s2, ctx, tmp_path):
    """API-level: counter columns, key by key, against a dictionary built with the oracle's orient()"""
    from strainer2_b200 import synth
    rng = synth.rng_for(3, 11)
    strain = synth.genome(rng, 30_000, 3, n_runs=2)
    synth.write_fasta(str(tmp_path / "s.fa"), strain, wrap=70)
    reads = synth.sample_reads(rng, [np.where(c == ord("N"), ord("G"), c).astype(np.uint8) for c in strain], 3000, 100,
                               sub_rate=0.02, n_rate=2e-3)
    synth.write_reads_fastq(str(tmp_path / "m.fastq"), reads)
    o = ou.OracleTable(4)
    o.build(str(tmp_path / "s.fa"))
    o.count_file(str(tmp_path / "m.fastq"), 2)
    want = o.table_text(False, str(tmp_path / "o.tsv"))
    t = s2.StrainTable(ctx, s2.load_flat(str(tmp_path / "s.fa")), n_cols=4)
    ctx.scan_count(t, s2.load_flat(str(tmp_path / "m.fastq")), 2)
    assert table_bytes(s2, t, 3, tmp_path) == want
    t.free(); o.free()


def test_detect_scan_matches_python_restatement(s2, ctx):
    from strainer2_b200 import synth
    rng = synth.rng_for(4, 3)
    strain = synth.genome(rng, 20_000, 2)
    flat = synth.contigs_to_flat(strain)
    t = s2.StrainTable(ctx, flat, n_cols=6)
    keys, _ = t.export()
    inf_keys = keys[::50]
    found = t.flag(np.concatenate([inf_keys, np.array([12345], dtype=np.uint64)]))
    assert found[:-1].all() and not found[-1]
    reads = synth.sample_reads(rng, strain + synth.genome(rng, 20_000, 1), 1500, 120, sub_rate=0.01, n_rate=2e-3)
    recs = [r.tobytes() for r in reads] + [b"ACGT", b"", b"N" * 50]
    batch, off = s2.flatten_records(recs)
    hits, inf, pos, st = ctx.scan_detect(t, batch, off)
    keyset = {s2.kmer_to_ascii(int(k)) for k in keys}
    infset = {s2.kmer_to_ascii(int(k)) for k in inf_keys}
    want_hits, want_inf, want_pos = [], [], []
    for r, s in enumerate(recs):
        h = i_ = 0
        for j in range(len(s) - 30):
            w = s[j:j + 31].upper()
            if b"N" in w:
                continue
            k = ou.orient(w)
            if k in keyset:
                h += 1
                if k in infset:
                    i_ += 1
                    want_pos.append(int(off[r]) + j)
        want_hits.append(h); want_inf.append(i_)
    assert hits.tolist() == want_hits
    assert inf.tolist() == want_inf
    assert pos.tolist() == want_pos
    assert st.hits == sum(want_hits)
    t.free()


# ---------------------------------------------------------------------------------------------
# strain_detect executable against the reference-generated golden vectors
# ---------------------------------------------------------------------------------------------
DETECT_RUNS = {
    "bg_batch": ["-r", "ref.fa", "-a", "informative.txt.gz", "-g", "background.txt", "-B", "batch.txt"],
    "bg_single": ["-r", "ref.fa", "-a", "informative_plain.txt", "-g", "background.txt", "-b", "s1_R1.fastq.gz", "-c",
                  "s1_R2.fastq.gz", "-t", "PE"],
    "batch": ["-r", "ref.fa", "-a", "informative.txt.gz", "-B", "batch.txt"],
    "single_pe": ["-r", "ref.fa", "-a", "informative_plain.txt", "-b", "s1_R1.fastq.gz", "-c", "s1_R2.fastq.gz", "-t", "PE"],
    "single_se_default": ["-r", "ref.fa", "-a", "informative.txt.gz", "-b", "s3_single.fa.gz"],
    "single_pei": ["-r", "ref.fa", "-a", "informative.txt.gz", "-b", "s2_interleaved.fa", "-t", "PEI"],
}


@pytest.mark.parametrize("name", sorted(DETECT_RUNS))
@pytest.mark.parametrize("batch_mb", ["32", "0"])
def test_strain_detect_executable_matches_reference_bytes(s2, golden_dir, tmp_path, name, batch_mb):
    """batch_mb=0 forces one read pair per GPU batch: the stale-state replay must not depend on batching"""
    import hashlib
    d = os.path.join(golden_dir, "detect_edge")
    out = str(tmp_path / "hits.gz")
    p = s2.run_strain_detect(DETECT_RUNS[name] + ["-o", out], cwd=d, env={"S2_DETECT_BATCH_MB": batch_mb})
    assert p.returncode == 0, p.stderr
    assert ou.gunzip(out) == ou.gunzip(os.path.join(d, f"expected_{name}.hits.txt.gz"))
    assert p.stdout == open(os.path.join(d, f"expected_{name}.stdout"), "rb").read()
    assert p.stderr == open(os.path.join(d, f"expected_{name}.stderr"), "rb").read()
    # even the compressed file is identical (same zlib, level 9, same byte stream)
    md5_file = os.path.join(d, f"expected_{name}.hits.gz.md5")
    if os.path.exists(md5_file):
        assert hashlib.md5(open(out, "rb").read()).hexdigest() == open(md5_file).read().strip()


def test_strain_detect_executable_error_paths(s2, golden_dir, tmp_path):
    d = os.path.join(golden_dir, "detect_edge")
    out = str(tmp_path / "o.gz")
    p = s2.run_strain_detect(["-r", "ref.fa", "-a", "informative.txt.gz", "-b", "s6_R1.fastq", "-c", "s6_R2.fastq", "-t", "PE", "-o", out], cwd=d)
    assert p.returncode == 1
    assert p.stderr == open(os.path.join(d, "expected_pe2_short.stderr"), "rb").read()
    for name, args in {"usage_missing": ["-r", "ref.fa"],
                       "usage_bad_type": ["-r", "ref.fa", "-a", "informative.txt.gz", "-b", "s3_single.fa.gz", "-t", "QQ", "-o", out],
                       "usage_pe_needs_c": ["-r", "ref.fa", "-a", "informative.txt.gz", "-b", "s3_single.fa.gz", "-t", "PE", "-o", out]}.items():
        p = s2.run_strain_detect(args, cwd=d)
        assert p.returncode == 1
        assert p.stdout == open(os.path.join(d, f"expected_{name}.stdout"), "rb").read()
        assert p.stderr == open(os.path.join(d, f"expected_{name}.stderr"), "rb").read()


def test_strain_detect_synthetic_matches_oracle(s2, tmp_path):
    """down-scaled config #4: 300 kb strain, 1 % informative k-mers, PE + SE + interleaved metagenomes"""
    from strainer2_b200 import synth
    tmp = str(tmp_path)
    rng = synth.rng_for(4, 0)
    strain = synth.genome(rng, 300_000, 6, n_runs=3)
    synth.write_fasta(os.path.join(tmp, "strain.fa"), strain)
    clean = [np.where(c == ord("N"), ord("T"), c).astype(np.uint8) for c in strain]
    other = synth.genome(rng, 300_000, 3)
    # informative = every 100th window of contig 0 (plain text, as kmer_scrub_filter.py prints them)
    c0 = clean[0].tobytes()
    with open(os.path.join(tmp, "inf.txt"), "wb") as f:
        f.write(b"#kmer\n")
        for i in range(0, len(c0) - 31, 100):
            f.write(c0[i:i + 31] + b"\n")
    r1 = synth.sample_reads(rng, clean + other * 4, 20_000, 150, sub_rate=0.004, n_rate=2e-4)
    r2 = synth.sample_reads(rng, clean + other * 4, 20_000, 150, sub_rate=0.004, n_rate=2e-4)
    synth.write_reads_fastq(os.path.join(tmp, "a_R1.fastq.gz"), r1)
    synth.write_reads_fastq(os.path.join(tmp, "a_R2.fastq.gz"), r2)
    synth.write_fasta(os.path.join(tmp, "b_se.fa"), [r for r in r1[:5000]], wrap=0)
    inter = np.empty((10_000, 150), dtype=np.uint8)
    inter[0::2] = r1[:5000]; inter[1::2] = r2[:5000]
    synth.write_reads_fastq(os.path.join(tmp, "c_inter.fastq"), inter)
    open(os.path.join(tmp, "batch.txt"), "w").write("PE\ta_R1.fastq.gz\ta_R2.fastq.gz\nSE\tb_se.fa\nPEI\tc_inter.fastq\n")
    args = ["-r", "strain.fa", "-a", "inf.txt", "-B", "batch.txt"]
    o = ou.oracle_cli(["detect"] + args + ["-m", os.path.join(tmp, "msg")], cwd=tmp)
    p = s2.run_strain_detect(args + ["-o", os.path.join(tmp, "hits.gz")], cwd=tmp, env={"S2_DETECT_BATCH_MB": "1"})
    assert o.returncode == 0 and p.returncode == 0, p.stderr
    got = ou.gunzip(os.path.join(tmp, "hits.gz"))
    assert got == o.stdout
    assert got.count(b"\n") > 1000
    assert p.stdout == open(os.path.join(tmp, "msg"), "rb").read()


def test_example_pipeline_count_filter_detect_chained(s2, tmp_path):
    """the three steps of test/example.sh chained, every step fed by the previous step's own output:
    kmer_scrub_count | gzip --best  ->  kmer_scrub_filter -m  ->  strain_detect -B; against the same chain of the oracle
    (C restatement for steps 1 and 3, the restated script for step 2).  Inputs mix BGZF, ordinary gzip and plain files."""
    import gzip
    from oracle import scrub_filter_oracle as fo
    from strainer2_b200 import synth
    tmp = str(tmp_path)
    rng = synth.rng_for(1, 7)
    strain = synth.genome(rng, 400_000, 5, n_runs=3)
    synth.write_fasta(os.path.join(tmp, "strain.fna.gz"), strain)
    clean = [np.where(c == ord("N"), ord("G"), c).astype(np.uint8) for c in strain]
    others = synth.genome(rng, 800_000, 4)
    # step-1 inputs: genomes (relatives of the strain among them) and metagenomes
    A, B = [], []
    for i in range(6):
        g = [synth.mutate(c, 0.02 * (i + 1), rng) for c in strain[: 1 + i % 3]] + synth.genome(rng, 200_000, 2) if i < 4 else synth.genome(rng, 400_000, 3)
        name = "g%d.fa" % i + ("" if i % 3 == 0 else ".gz")
        if i % 3 == 1:
            synth.write_bgzf(os.path.join(tmp, name), synth.fasta_bytes(g, 70))
        else:
            synth.write_fasta(os.path.join(tmp, name), g)
        A.append(name)
    for i in range(3):
        reads = synth.sample_reads(rng, clean[: 1 + i] + others, 30_000, 150, sub_rate=0.01, n_rate=1e-4)
        name = "scrub_m%d.fastq.gz" % i
        if i == 1:
            synth.write_reads_fastq(os.path.join(tmp, name), reads)
        else:
            synth.write_bgzf(os.path.join(tmp, name), synth.fastq_bytes(reads))
        B.append(name)
    open(os.path.join(tmp, "genomes_to_scrub.txt"), "w").write("".join(a + "\n" for a in A))
    open(os.path.join(tmp, "metagenomes_to_scrub.txt"), "w").write("".join(b + "\n" for b in B))
    # step-3 inputs: target metagenomes that contain the strain
    r1 = synth.sample_reads(rng, clean + others, 40_000, 150, sub_rate=0.003, n_rate=1e-4)
    r2 = synth.sample_reads(rng, clean + others, 40_000, 150, sub_rate=0.003, n_rate=1e-4)
    synth.write_bgzf(os.path.join(tmp, "t_R1.fastq.gz"), synth.fastq_bytes(r1))
    synth.write_bgzf(os.path.join(tmp, "t_R2.fastq.gz"), synth.fastq_bytes(r2))
    synth.write_reads_fastq(os.path.join(tmp, "t_se.fastq.gz"), r1[:10_000])
    open(os.path.join(tmp, "target_metagenomes.txt"), "w").write("PE\tt_R1.fastq.gz\tt_R2.fastq.gz\nSE\tt_se.fastq.gz\n")

    step1 = ["-r", "strain.fna.gz", "-A", "genomes_to_scrub.txt", "-B", "metagenomes_to_scrub.txt"]
    # ---- the oracle's chain
    o1 = ou.oracle_cli(["count"] + step1, cwd=tmp)
    assert o1.returncode == 0
    with gzip.GzipFile(os.path.join(tmp, "o.scrub_kmer_counts.gz"), "wb", compresslevel=9) as f:
        f.write(o1.stdout)
    rc, o2, _ = fo.run(["-s", "o.scrub_kmer_counts.gz", "-m", "0.02"], cwd=tmp)
    assert rc == 0
    with gzip.GzipFile(os.path.join(tmp, "o.scrubbed_kmers.gz"), "wb", compresslevel=9) as f:
        f.write(o2)
    o3 = ou.oracle_cli(["detect", "-r", "strain.fna.gz", "-a", "o.scrubbed_kmers.gz", "-B", "target_metagenomes.txt", "-m", os.path.join(tmp, "o.msg")], cwd=tmp)
    assert o3.returncode == 0
    # ---- the drop-ins' chain
    p1 = s2.run_kmer_scrub_count(step1 + ["-p", "strain.progress"], cwd=tmp)
    assert p1.returncode == 0, p1.stderr
    assert p1.stdout == o1.stdout
    with gzip.GzipFile(os.path.join(tmp, "p.scrub_kmer_counts.gz"), "wb", compresslevel=9) as f:
        f.write(p1.stdout)
    p2 = s2.run_kmer_scrub_filter(["-s", "p.scrub_kmer_counts.gz", "-m", "0.02"], cwd=tmp)
    assert p2.returncode == 0, p2.stderr
    assert p2.stdout == o2 and p2.stdout.count(b"\n") > 5000
    with gzip.GzipFile(os.path.join(tmp, "p.scrubbed_kmers.gz"), "wb", compresslevel=9) as f:
        f.write(p2.stdout)
    p3 = s2.run_strain_detect(["-r", "strain.fna.gz", "-a", "p.scrubbed_kmers.gz", "-B", "target_metagenomes.txt", "-o", "p.kmer_hits.gz"], cwd=tmp)
    assert p3.returncode == 0, p3.stderr
    hits = ou.gunzip(os.path.join(tmp, "p.kmer_hits.gz"))
    assert hits == o3.stdout and hits.count(b"\n") > 2000
    assert p3.stdout == open(os.path.join(tmp, "o.msg"), "rb").read()


def test_config1_reference_example_sh_has_the_reference_digests(s2, tmp_path):
    """BASELINE config #1 = the reference's own test/example.sh (steps 1-3: test/example.sh:4,11,18) through the three
    drop-in executables on the reference's own inputs (single-member .gz FASTA), against the digests of the unmodified
    reference.  The inputs are a copy of /root/reference/test made by oracle/Makefile under oracle/_ref/test
    (git-ignored data that travels to the GPU box like the compiled reference)."""
    import gzip
    import hashlib
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "test")
    strain = "strains/Bacteroides_ovatus_1001283st1_B8_1001283B150210_160208"
    if not os.path.exists(os.path.join(d, strain + ".fna.gz")):
        pytest.skip("oracle/_ref/test (copy of the reference's example inputs, `make -C oracle`) is not present")
    tmp = str(tmp_path)
    md5 = lambda b: hashlib.md5(b).hexdigest()
    p1 = s2.run_kmer_scrub_count(["-r", strain + ".fna.gz", "-A", "genomes_to_scrub.txt", "-B", "metagenomes_to_scrub.txt",
                                  "-p", os.path.join(tmp, "progress")], cwd=d)
    assert p1.returncode == 0, p1.stderr
    assert p1.stdout.count(b"\n") == 6_698_541
    assert md5(p1.stdout) == "75989a9bc31ef0b6f53a5112a60920bd"
    prog = open(os.path.join(tmp, "progress"), "rb").read().split(b"\n")
    assert prog[0] == b"adding kmer counts for:" and [l.split(b"\t")[0] for l in prog[1:3]] == [
        b"strains/Bacteroides_ovatus_1001302st1_D4_1001302B_160321.fna.gz", b"metagenomes/1001099B_150804_B6_s09_tiny_PE1.fasta.gz"]
    with gzip.GzipFile(os.path.join(tmp, "counts.gz"), "wb", compresslevel=6) as f:
        f.write(p1.stdout)
    p2 = s2.run_kmer_scrub_filter(["-s", os.path.join(tmp, "counts.gz"), "-m", "0.01"], cwd=d)
    assert p2.returncode == 0, p2.stderr
    lines = p2.stdout.split(b"\n")
    assert sum(1 for l in lines if l and not l.startswith(b"#")) == 66_986
    assert md5(p2.stdout) == "fe981fa571be70e602875ac3463ecdac"          # unmodified scripts/kmer_scrub_filter.py, recorded in the dev container
    with gzip.GzipFile(os.path.join(tmp, "scrubbed.gz"), "wb", compresslevel=9) as f:
        f.write(p2.stdout)
    p3 = s2.run_strain_detect(["-r", strain + ".fna.gz", "-a", os.path.join(tmp, "scrubbed.gz"), "-B", "target_metagenomes.txt",
                               "-o", os.path.join(tmp, "kmer_hits.gz")], cwd=d)
    assert p3.returncode == 0, p3.stderr
    raw = open(os.path.join(tmp, "kmer_hits.gz"), "rb").read()
    text = gzip.decompress(raw)
    assert text.count(b"\n") == 1_130
    assert md5(text) == "e1799e705d4f693240573da32540efcc"
    assert md5(raw) == "997c3e1b8c1272a736168909c6be359b"                # same zlib, level 9, same byte stream
    # -C = {the strain itself, the other strain}: the self entry is skipped with the reference's note, the drug column
    # then equals the pangenome column (same file)
    open(os.path.join(tmp, "listC.txt"), "w").write(strain + ".fna.gz\nstrains/Bacteroides_ovatus_1001302st1_D4_1001302B_160321.fna.gz\n")
    p4 = s2.run_kmer_scrub_count(["-r", strain + ".fna.gz", "-A", "genomes_to_scrub.txt", "-B", "metagenomes_to_scrub.txt",
                                  "-C", os.path.join(tmp, "listC.txt")], cwd=d)
    assert p4.returncode == 0, p4.stderr
    assert p4.stderr == ("skipping %s.fna.gz (identical match)\n" % strain).encode()
    import io
    import pandas as pd
    v = pd.read_csv(io.BytesIO(p4.stdout), sep="\t", header=0, usecols=[1, 2, 3, 4]).to_numpy()
    assert v.shape[0] == 6_698_540
    assert v[:, 0].sum() == 6_721_161 and v[:, 1].sum() == 1_376_122 and v[:, 2].sum() == 50_506 and v[:, 3].sum() == 1_376_122


def test_iupac_bytes_are_hashed_as_strings_like_the_reference(s2, golden_dir, tmp_path):
    """windows with bytes outside ACGTN go through the host string path; output bytes equal the
    reference's (rows containing R/Y/K/M/-/E ..., merged into the replayed row order)"""
    d = os.path.join(golden_dir, "iupac")
    p = s2.run_kmer_scrub_count(["-r", "ref.fa", "-A", "listA.txt", "-B", "listB.txt", "-C", "listC.txt"], cwd=d)
    assert p.returncode == 0, p.stderr
    assert p.stdout == open(os.path.join(d, "expected_count.tsv"), "rb").read()
    assert p.stderr == open(os.path.join(d, "expected_count.stderr"), "rb").read()
    out = str(tmp_path / "hits.gz")
    p = s2.run_strain_detect(["-r", "ref.fa", "-a", "informative.txt", "-B", "batch.txt", "-o", out], cwd=d)
    assert p.returncode == 0, p.stderr
    assert ou.gunzip(out) == ou.gunzip(os.path.join(d, "expected_detect.hits.txt.gz"))
    assert p.stdout == open(os.path.join(d, "expected_detect.stdout"), "rb").read()


def test_executable_shards_files_over_two_gpus_with_one_allreduce(s2, golden_dir, tmp_path):
    """S2_GPUS=2: replicas on two GPUs, files dealt to both, counters summed by ncclAllReduce; bytes unchanged"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    d = os.path.join(golden_dir, "count_edge")
    p = s2.run_kmer_scrub_count(["-r", "ref.fa.gz", "-A", "listA.txt", "-B", "listB.txt", "-C", "listC.txt"], cwd=d,
                                env={"S2_GPUS": "2", "S2_THREADS": "4", "S2_BATCH_MB": "1"})
    assert p.returncode == 0, p.stderr
    assert p.stdout == open(os.path.join(d, "expected_ABC.tsv"), "rb").read()
    assert p.stderr == open(os.path.join(d, "expected_ABC.stderr"), "rb").read()
    args = _write_inputs(s2, str(tmp_path), 400_000, 3, 2, 30_000)
    one = s2.run_kmer_scrub_count(args, env={"S2_BATCH_MB": "2"})
    two = s2.run_kmer_scrub_count(args, env={"S2_BATCH_MB": "2", "S2_GPUS": "2"})
    assert one.returncode == 0 and two.returncode == 0, two.stderr
    assert one.stdout == two.stdout


def test_multi_strain_batch_equals_one_reference_run_per_strain(s2, tmp_path):
    """BASELINE config #5, down-scaled: 5 strains (pairs share 10 % of their sequence), shared -A/-B lists, -C = all
    strain genomes (self skipped per strain) + one outsider listed twice; ONE pass over the inputs.  Every
    strain's table must be byte-identical to the oracle run on that strain alone."""
    from strainer2_b200 import synth
    tmp = str(tmp_path)
    rng = synth.rng_for(5, 1)
    shared = synth.random_bases(rng, 20_000)
    strains = []
    for i in range(5):
        g = synth.genome(rng, 200_000, 4, n_runs=2)
        if i % 2 == 0 or i == 1:
            g[0][1000:21_000] = synth.mutate(shared, 0.002 * i, rng)
        p = os.path.join(tmp, f"strain{i}.fa" + (".gz" if i % 2 else ""))
        synth.write_fasta(p, g)
        strains.append((p, g))
    A = []
    for i in range(3):
        p = os.path.join(tmp, f"a{i}.fa")
        synth.write_fasta(p, [synth.mutate(c, 0.01, rng) for c in strains[i][1]] if i < 2 else synth.genome(rng, 200_000, 3))
        A.append(p)
    clean = [np.where(c == ord("N"), ord("A"), c).astype(np.uint8) for _, g in strains for c in g]
    B = []
    for i in range(2):
        p = os.path.join(tmp, f"m{i}.fastq.gz")
        rd = synth.sample_reads(rng, clean + synth.genome(rng, 400_000, 2), 20_000, 150, sub_rate=0.004, n_rate=1e-4)
        if i == 0:
            synth.write_reads_fastq(p, rd)                               # ordinary gzip: host inflate
        else:
            synth.write_bgzf(p, synth.fastq_bytes(rd))                  # BGZF: GPU ingest into the union table
        B.append(p)
    open(os.path.join(tmp, "R.txt"), "w").write("".join(p + "\n" for p, _ in strains))
    open(os.path.join(tmp, "A.txt"), "w").write("".join(p + "\n" for p in A))
    open(os.path.join(tmp, "B.txt"), "w").write("".join(p + "\n" for p in B))
    open(os.path.join(tmp, "C.txt"), "w").write("".join(p + "\n" for p, _ in strains) + A[0] + "\n" + strains[3][0] + "\n")
    args = ["-R", os.path.join(tmp, "R.txt"), "-A", os.path.join(tmp, "A.txt"), "-B", os.path.join(tmp, "B.txt"), "-C", os.path.join(tmp, "C.txt")]
    want = {}
    for path, _ in strains:
        o = ou.oracle_cli(["count", "-r", path, "-A", os.path.join(tmp, "A.txt"), "-B", os.path.join(tmp, "B.txt"), "-C", os.path.join(tmp, "C.txt")])
        assert o.returncode == 0
        want[path] = o.stdout
    # second run: the union table is forced through the two-phase (partitioned) scan, host batches and GPU ingest alike
    for k_run, env in enumerate(({"S2_BATCH_MB": "2"}, {"S2_BATCH_MB": "2", "S2_PARTITION_MIN_MB": "0", "S2_PARTITION_MIN_BATCH_KB": "0"})):
        out = os.path.join(tmp, "out%d" % k_run)
        p = s2.run_kmer_scrub_count_batch(args + ["-O", out], env=env)
        assert p.returncode == 0, p.stderr
        drug_sums = []
        for path, _ in strains:
            got = open(os.path.join(out, os.path.basename(path) + ".scrub_kmer_counts"), "rb").read()
            assert got == want[path], (path, env)
            k, v = ou.parse_table(got)
            assert v[:, 2].sum() > 0
            drug_sums.append(int(v[:, 3].sum()))
        assert sum(1 for d in drug_sums if d > 0) >= 3          # the strains that share sequence see each other in -C


def test_two_phase_partitioned_scan_equals_direct_scan(s2, tmp_path, monkeypatch):
    """tables whose fingerprints exceed L2 are scanned by radix-partition + per-partition probe; forced on
    here for a small table: counters must equal the direct kernel's, including when a partition overflows
    (poly-A input: every window lands in one partition) and the direct kernel takes over."""
    import torch
    from strainer2_b200 import synth
    rng = synth.rng_for(6, 0)
    strain = synth.genome(rng, 300_000, 4, n_runs=3)
    strain[1][5000:5200] = ord("A")                                  # poly-A keys in the table
    flat = synth.contigs_to_flat(strain)
    clean = [np.where(c == ord("N"), ord("C"), c).astype(np.uint8) for c in strain]
    reads = synth.sample_reads(rng, clean + synth.genome(rng, 600_000, 2), 60_000, 150, sub_rate=0.01, n_rate=1e-3)
    batch = synth.reads_to_flat(reads)
    polyA = np.full(3_000_000, ord("A"), dtype=np.uint8)
    polyA[::1000] = ord("\n")
    with s2.Context(0, batch_bytes=4 << 20, n_lanes=3) as ctx:
        direct = s2.StrainTable(ctx, flat, n_cols=4)
        monkeypatch.setenv("S2_PARTITION_MIN_MB", "0")
        monkeypatch.setenv("S2_PARTITION_MIN_BATCH_KB", "0")
        part = s2.StrainTable(ctx, flat, n_cols=4)
        for col, data in ((1, batch), (2, polyA)):
            a = ctx.scan_count(direct, torch.from_numpy(data).cuda(), col)
            b = ctx.scan_count(part, torch.from_numpy(data).cuda(), col)
            assert (a.hits, a.valid_windows) == (b.hits, b.valid_windows) and a.hits > 0
            assert np.array_equal(direct.counts(col), part.counts(col))
            if col == 1:
                a_hits_reads = a.hits
        # host batches through the lanes use the same two-phase path
        c = ctx.scan_count(part, batch, 3)
        assert c.hits == int(direct.counts(1).sum())
        assert np.array_equal(part.counts(3), direct.counts(1))
        # and so does the GPU ingest (batch length, veto and increment come from device memory): a BGZF image of the
        # reads, streamed in several chunks; then an irregular one whose counted chunks are taken back out
        monkeypatch.setenv("S2_INGEST_CHUNK_MB", "1")
        monkeypatch.setenv("S2_INGEST_TEXT_MB", "4")
        ctx.ingest_reset()
        ctx.sync()
        part.clear_counts(3)
        text = synth.fastq_bytes(reads)
        rc, n_bases, _ = ctx.ingest_count_mem(part, np.frombuffer(synth.bgzf_bytes(text), dtype=np.uint8), 3)
        st = ctx.sync()
        assert rc == 0 and n_bases == reads.size and st.hits == a_hits_reads
        assert np.array_equal(part.counts(3), direct.counts(1))
        bad = text + b"@x\nACGT\n+\nII\n"
        assert ctx.ingest_count_mem(part, np.frombuffer(synth.bgzf_bytes(bad), dtype=np.uint8), 3)[0] == 1
        assert ctx.sync().hits == 0 and np.array_equal(part.counts(3), direct.counts(1))
        ctx.ingest_reset()
        direct.free(); part.free()


def test_device_formatter_equals_host_formatter(s2, ctx, golden_dir, tmp_path):
    """s2_table_format (rows formatted by a kernel) against s2_format_count_table (host) and the golden bytes,
    including counters above 2^31 that the reference prints negative (%d of unsigned)"""
    d = os.path.join(golden_dir, "count_edge")
    t = s2.StrainTable(ctx, s2.load_flat(os.path.join(d, "ref.fa.gz")), n_cols=4)
    keys, djb2 = t.export()
    order, _ = s2.roworder_emulate(djb2)
    big = np.arange(t.n_keys, dtype=np.uint64) * 1_000_003 % (2 ** 32)
    big[:5] = [0, 2 ** 31 - 1, 2 ** 31, 2 ** 32 - 1, 10]
    t.set_counts(2, big.astype(np.uint32))
    t.set_counts(3, (big[::-1] // 7).astype(np.uint32))
    for n_print in (3, 4):
        host, dev = str(tmp_path / f"h{n_print}.tsv"), str(tmp_path / f"d{n_print}.tsv")
        s2.format_count_table(host, keys, order, [t.counts(c) for c in range(n_print)])
        t.format_to(dev, order, n_print)
        assert open(dev, "rb").read() == open(host, "rb").read()
    p = s2.run_kmer_scrub_count(["-r", "ref.fa.gz", "-A", "listA.txt", "-B", "listB.txt", "-C", "listC.txt"], cwd=d,
                                env={"S2_HOST_FORMAT": "1"})
    assert p.returncode == 0 and p.stdout == open(os.path.join(d, "expected_ABC.tsv"), "rb").read()
    t.free()


# ---------------------------------------------------------------------------------------------
# GPU-side ingest: hardware DEFLATE (BGZF) + record splitting kernels
# ---------------------------------------------------------------------------------------------
def _ingest_fixture(s2, tmp, n_reads=60_000, seed=0):
    from strainer2_b200 import synth
    rng = synth.rng_for(7, seed)
    strain = synth.genome(rng, 300_000, 4, n_runs=2)
    clean = [np.where(c == ord("N"), ord("A"), c).astype(np.uint8) for c in strain]
    reads = synth.sample_reads(rng, clean + synth.genome(rng, 600_000, 2), n_reads, 150, sub_rate=0.005, n_rate=1e-4)
    synth.write_fasta(os.path.join(tmp, "strain.fa"), strain)
    return strain, reads


def _ingest_chunks(ctx, monkeypatch, chunk_mb):
    """small chunks stream a file through the three-slot ring in many pieces; the defaults take it as one chunk"""
    if chunk_mb:
        monkeypatch.setenv("S2_INGEST_CHUNK_MB", str(chunk_mb[0]))
        monkeypatch.setenv("S2_INGEST_TEXT_MB", str(chunk_mb[1]))
    ctx.ingest_reset()


@pytest.mark.parametrize("chunk_mb", [(2, 8), None], ids=["streamed", "one_chunk"])
def test_gpu_ingest_bgzf_and_plain_fastq_equal_host_reader(s2, ctx, tmp_path, monkeypatch, chunk_mb):
    from strainer2_b200 import synth
    monkeypatch.setenv("S2_GPU_INGEST_PLAIN", "1")
    _ingest_chunks(ctx, monkeypatch, chunk_mb)
    tmp = str(tmp_path)
    strain, reads = _ingest_fixture(s2, tmp, 340_000)                  # ~100 MB of text: several ingest chunks
    data = synth.fastq_bytes(reads)
    # a few irregular-but-legal records at known places: short reads, an empty read, N runs
    data += b"@short\nACGT\n+\nIIII\n@empty\n\n+\n\n@n\n" + b"N" * 80 + b"\n+\n" + b"#" * 80 + b"\n"
    synth.write_bgzf(os.path.join(tmp, "m.fastq.gz"), data)
    open(os.path.join(tmp, "m.fastq"), "wb").write(data[:-1])          # plain, last line without '\n'
    t = s2.StrainTable(ctx, s2.load_flat(os.path.join(tmp, "strain.fa")), n_cols=4)
    want = ctx.scan_count(t, s2.load_flat(os.path.join(tmp, "m.fastq.gz")), 1)
    n_bases = len(reads) * 150 + 4 + 80
    for col, name in ((2, "m.fastq.gz"), (3, "m.fastq")):
        rc, bases, lookups = ctx.ingest_count_file(t, os.path.join(tmp, name), col)
        st = ctx.sync()
        assert rc == 0
        assert bases == n_bases and lookups == len(reads) * 120 + 50
        assert st.hits == want.hits and st.valid_windows == want.valid_windows
        assert np.array_equal(t.counts(col), t.counts(1))
    # the same files as images in host memory (s2_ingest_count_mem): only the compressed bytes cross PCIe
    for name in ("m.fastq.gz", "m.fastq"):
        t.clear_counts(2)
        image = np.fromfile(os.path.join(tmp, name), dtype=np.uint8)
        rc, bases, lookups = ctx.ingest_count_mem(t, image, 2)
        st = ctx.sync()
        assert rc == 0 and bases == n_bases and lookups == len(reads) * 120 + 50
        assert st.hits == want.hits and np.array_equal(t.counts(2), t.counts(1))
    assert ctx.ingest_count_mem(t, np.frombuffer(b"junk\n" + data[:5000], dtype=np.uint8), 3)[0] == 1
    t.free()
    ctx.ingest_reset()


def test_gpu_ingest_streamed_bgzf_chunk_may_end_inside_a_member_header(s2, ctx, tmp_path, monkeypatch):
    """a streamed BGZF file whose 1 MiB chunk ends 1..17 bytes into the next member's header (empty members are used as
    padding to put a header exactly there): the walk must take that as an incomplete member and start the next chunk at
    it, not hand the file back as "not BGZF" """
    import struct
    from strainer2_b200 import synth
    _ingest_chunks(ctx, monkeypatch, (1, 8))
    tmp = str(tmp_path)
    strain, reads = _ingest_fixture(s2, tmp, 60_000)
    data = synth.fastq_bytes(reads)
    t = s2.StrainTable(ctx, s2.load_flat(os.path.join(tmp, "strain.fa")), n_cols=4)
    eof = synth.bgzf_bytes(b"")
    assert len(eof) == 28
    tested = 0
    for k in (1, 5, 11, 17):                                            # bytes of the straddling header inside chunk 0
        members, size, pos = [], 0, 0
        while pos < len(data):
            m = synth.bgzf_bytes(data[pos:pos + 65280])[:-28]
            if size + len(m) > (1 << 20) - k and size <= (1 << 20) - k:
                gap = (1 << 20) - k - size                              # pad with empty members up to 2^20 - k
                if gap % 28:
                    # shrink the previous member's text so that the gap becomes a multiple of 28 (try a few cuts)
                    prev_pos = pos - 65280
                    for cut in range(1, 4000):
                        pm = synth.bgzf_bytes(data[prev_pos:pos - cut])[:-28]
                        if ((1 << 20) - k - (size - len(members[-1]) + len(pm))) % 28 == 0:
                            size += len(pm) - len(members[-1]); members[-1] = pm; pos -= cut
                            break
                    else:
                        break
                    gap = (1 << 20) - k - size
                    m = synth.bgzf_bytes(data[pos:pos + 65280])[:-28]
                members.extend([eof] * (gap // 28)); size += gap
            members.append(m); size += len(m); pos += 65280
        image = b"".join(members) + eof
        import gzip
        assert gzip.decompress(image) == data
        # a member header really starts k bytes before the 1 MiB mark
        off, starts = 0, set()
        while off < len(image):
            starts.add(off); off += struct.unpack_from("<H", image, off + 16)[0] + 1
        if (1 << 20) - k not in starts:
            continue
        tested += 1
        path = os.path.join(tmp, "straddle%d.fastq.gz" % k)
        open(path, "wb").write(image)
        t.clear_counts(1); t.clear_counts(2)
        want = ctx.scan_count(t, s2.load_flat(path), 1)
        rc, bases, lookups = ctx.ingest_count_file(t, path, 2)
        st = ctx.sync()
        assert rc == 0, "the file was handed back to the host reader"
        assert bases == len(reads) * 150 and st.hits == want.hits
        assert np.array_equal(t.counts(2), t.counts(1))
    assert tested >= 2
    t.free()
    ctx.ingest_reset()


def test_gpu_ingest_pipeline_follows_its_context(s2, ctx, tmp_path):
    """the calling thread's ingest pipeline belongs to one context: closing that context drops it, and a context
    made afterwards (possibly at the same address) gets a pipeline of its own with the same results"""
    from strainer2_b200 import synth
    tmp = str(tmp_path)
    strain, reads = _ingest_fixture(s2, tmp, 20_000, seed=5)
    synth.write_bgzf(os.path.join(tmp, "m.fastq.gz"), synth.fastq_bytes(reads))
    flat = s2.load_flat(os.path.join(tmp, "strain.fa"))
    ctx.ingest_reset()
    seen = []
    for _ in range(3):
        c2 = s2.Context(0, batch_bytes=8 << 20, n_lanes=2)
        t = s2.StrainTable(c2, flat, n_cols=3)
        want = c2.scan_count(t, synth.reads_to_flat(reads), 1)
        rc, bases, _ = c2.ingest_count_file(t, os.path.join(tmp, "m.fastq.gz"), 2)
        st = c2.sync()
        assert rc == 0 and bases == reads.size and st.hits == want.hits > 1000
        assert np.array_equal(t.counts(1), t.counts(2))
        seen.append(t.counts(2).copy())
        t.free()
        c2.close()                                   # takes this thread's pipeline with it
    assert all(np.array_equal(seen[0], x) for x in seen[1:])
    # the module's context still works afterwards (its pipeline is rebuilt on demand)
    t = s2.StrainTable(ctx, flat, n_cols=3)
    assert ctx.ingest_count_file(t, os.path.join(tmp, "m.fastq.gz"), 2)[0] == 0
    ctx.sync()
    assert np.array_equal(t.counts(2), seen[0])
    t.free()
    ctx.ingest_reset()


@pytest.mark.parametrize("chunk_mb", [(1, 4), None], ids=["streamed", "one_chunk"])
def test_gpu_ingest_hands_irregular_files_back_untouched(s2, ctx, tmp_path, monkeypatch, chunk_mb):
    """irregular text is never counted: in one chunk the verdict precedes the scan; in a streamed file the chunks
    that were counted before the irregular one are taken back out (replay with increment -1)"""
    from strainer2_b200 import synth
    _ingest_chunks(ctx, monkeypatch, chunk_mb)
    tmp = str(tmp_path)
    strain, reads = _ingest_fixture(s2, tmp, 60_000 if chunk_mb else 3000, seed=1)       # streamed: ~20 MB of text, 5+ chunks
    good = synth.fastq_bytes(reads)
    r0 = reads[0].tobytes()
    cases = {
        "crlf": good.replace(b"\n", b"\r\n"),
        "multiline": good + b"@ml\n" + r0[:75] + b"\n" + r0[75:] + b"\n+\n" + b"I" * 75 + b"\n" + b"I" * 75 + b"\n",
        "truncated": good + b"@t\n" + r0 + b"\n+\n",
        "qual_len": good + b"@q\n" + r0 + b"\n+\n" + b"I" * 149 + b"\n",
        "fasta_inside": good + b">fa\n" + r0 + b"\n",
        "junk_first": b"junk\n" + good,
        "middle": good[:len(good) // 2] + b"@x\nACGT\n+\nII\n" + good[len(good) // 2:],
    }
    t = s2.StrainTable(ctx, s2.load_flat(os.path.join(tmp, "strain.fa")), n_cols=4)
    want = ctx.scan_count(t, synth.reads_to_flat(reads), 2)
    assert want.hits > 1000
    ctx.sync()
    for name, data in cases.items():
        p = os.path.join(tmp, name + ".fastq.gz")
        synth.write_bgzf(p, data)
        rc, _, _ = ctx.ingest_count_file(t, p, 1)
        st = ctx.sync()
        assert rc == 1, name
        assert int(t.counts(1).sum()) == 0, name                     # nothing counted, or everything taken back
        assert st.hits == 0, name
    # a member whose CRC-32 does not match its text (the hardware engine checks none; zlib would end the run): the member
    # CRC pass vetoes the chunk, the file is handed back - streamed: after the earlier chunks were taken back out
    whole = bytearray(synth.bgzf_bytes(good))
    members, pos = [], 0
    while pos + 18 <= len(whole):
        bsize = (whole[pos + 16] | whole[pos + 17] << 8) + 1
        members.append((pos, bsize))
        pos += bsize
    m_pos, m_size = members[len(members) * 2 // 3]
    whole[m_pos + m_size - 8] ^= 0x40
    open(os.path.join(tmp, "bad_crc.fastq.gz"), "wb").write(bytes(whole))
    assert ctx.ingest_count_file(t, os.path.join(tmp, "bad_crc.fastq.gz"), 1)[0] == 1
    assert int(t.counts(1).sum()) == 0 and ctx.sync().hits == 0
    monkeypatch.setenv("S2_BGZF_CRC", "0")                                # (read when a pipeline is created)
    ctx.ingest_reset()
    assert ctx.ingest_count_file(t, os.path.join(tmp, "bad_crc.fastq.gz"), 1)[0] == 0          # round 1's behaviour: ISIZE alone
    assert ctx.sync().hits == want.hits
    t.clear_counts(1)
    monkeypatch.delenv("S2_BGZF_CRC")
    ctx.ingest_reset()
    # a BGZF file cut in the middle of a member (streamed: the earlier chunks are taken back)
    whole = synth.bgzf_bytes(good)
    open(os.path.join(tmp, "cut.fastq.gz"), "wb").write(whole[:len(whole) * 3 // 4])
    assert ctx.ingest_count_file(t, os.path.join(tmp, "cut.fastq.gz"), 1)[0] == 1
    assert int(t.counts(1).sum()) == 0 and ctx.sync().hits == 0
    # a plain single-member gzip is not BGZF: the hardware engine cannot take it - host reader when the software gunzip is
    # switched off (round 1's only behaviour), else the chunk-parallel gunzip of s2_gunzip.cu
    synth.write_reads_fastq(os.path.join(tmp, "plain.fastq.gz"), reads[:3000])
    monkeypatch.setenv("S2_GPU_GUNZIP", "0")
    assert ctx.ingest_count_file(t, os.path.join(tmp, "plain.fastq.gz"), 1)[0] == 1
    assert int(t.counts(1).sum()) == 0 and ctx.sync().hits == 0
    monkeypatch.delenv("S2_GPU_GUNZIP")
    t.clear_counts(3)
    want3 = ctx.scan_count(t, synth.reads_to_flat(reads[:3000]), 3)
    assert ctx.ingest_count_file(t, os.path.join(tmp, "plain.fastq.gz"), 1)[0] == 0
    assert ctx.sync().hits == want3.hits and np.array_equal(t.counts(1), t.counts(3))
    t.clear_counts(1)
    # and the regular file itself is counted
    synth.write_bgzf(os.path.join(tmp, "good.fastq.gz"), good)
    assert ctx.ingest_count_file(t, os.path.join(tmp, "good.fastq.gz"), 1)[0] == 0
    assert ctx.sync().hits == want.hits and np.array_equal(t.counts(1), t.counts(2))
    t.free()
    ctx.ingest_reset()


@pytest.mark.parametrize("chunk_mb", [(1, 4), None], ids=["streamed", "one_chunk"])
def test_gpu_ingest_long_lines_are_spread_over_blocks(s2, ctx, tmp_path, monkeypatch, chunk_mb):
    """unwrapped genomes and long reads: a line of hundreds of KB spans many 16 KB text blocks, each of which copies the
    part that lies in it; FASTA lines may be longer than a chunk (no bytes are carried between chunks, only what kind of
    line is open), FASTQ records must fit the 1 MB carry"""
    from strainer2_b200 import synth
    _ingest_chunks(ctx, monkeypatch, chunk_mb)
    tmp = str(tmp_path)
    rng = synth.rng_for(12, 0)
    strain = synth.genome(rng, 300_000, 3, n_runs=2)
    synth.write_fasta(os.path.join(tmp, "strain.fa"), strain)
    t = s2.StrainTable(ctx, s2.load_flat(os.path.join(tmp, "strain.fa")), n_cols=4)
    rel = [synth.mutate(c, 0.01, rng) for c in strain]
    # FASTA: one line per contig (100 kb each), a 900 kb line, short and empty-ish records in between, a wrapped record
    contigs = rel + [synth.random_bases(rng, 900_000)] + [synth.random_bases(rng, n) for n in (31, 30, 1, 16_384, 16_383, 16_385, 512, 513)] + rel[:1]
    text = b"".join(b">c%d\n%s\n" % (i, c.tobytes()) for i, c in enumerate(contigs)) + synth.fasta_bytes(rel[1:2], 80)
    open(os.path.join(tmp, "u.fa.gz"), "wb").write(synth.bgzf_bytes(text))
    want = ctx.scan_count(t, s2.load_flat(os.path.join(tmp, "u.fa.gz")), 1)
    assert want.hits > 300_000
    rc, bases, _ = ctx.ingest_count_file(t, os.path.join(tmp, "u.fa.gz"), 2)
    st = ctx.sync()
    assert rc == 0 and bases == sum(c.size for c in contigs) + rel[1].size
    assert st.hits == want.hits and st.valid_windows == want.valid_windows and np.array_equal(t.counts(2), t.counts(1))
    # lines longer than a chunk (streamed: 4 MB of text per chunk): a 9 Mb chromosome on one line that holds the strain
    # twice, a header of 5 MB, and the file from above behind them
    chrom = synth.random_bases(rng, 9_000_000)
    chrom[2_000_000:2_100_000] = rel[0][:100_000]
    chrom[8_388_600:8_388_600 + rel[1].size] = rel[1]
    big = b">chromosome\n" + chrom.tobytes() + b"\n>" + b"h" * 5_000_000 + b"\n" + text
    open(os.path.join(tmp, "big.fa.gz"), "wb").write(synth.bgzf_bytes_parallel(big, threads=8))
    t.clear_counts(1); t.clear_counts(2)
    want = ctx.scan_count(t, s2.load_flat(os.path.join(tmp, "big.fa.gz")), 1)
    rc, bases, _ = ctx.ingest_count_file(t, os.path.join(tmp, "big.fa.gz"), 2)
    st = ctx.sync()
    assert rc == 0 and bases == chrom.size + sum(c.size for c in contigs) + rel[1].size
    assert st.hits == want.hits and st.valid_windows == want.valid_windows and np.array_equal(t.counts(2), t.counts(1))
    # FASTQ with long reads (quality lines just as long) between short ones
    reads = []
    for i, n in enumerate((150, 40_000, 150, 31, 30, 100_000, 16_384, 20, 250_000, 150, 16_385, 400_000, 151)):      # records stay below the 1 MB carry limit
        src = np.concatenate([rel[i % 3], synth.random_bases(rng, max(0, n - rel[i % 3].size))])[:n]
        reads.append(src)
    fq = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, r.tobytes(), b"I" * r.size) for i, r in enumerate(reads)) * 2
    open(os.path.join(tmp, "long.fastq.gz"), "wb").write(synth.bgzf_bytes(fq))
    t.clear_counts(1); t.clear_counts(2)
    want = ctx.scan_count(t, s2.load_flat(os.path.join(tmp, "long.fastq.gz")), 1)
    assert want.hits > 100_000
    rc, bases, lookups = ctx.ingest_count_file(t, os.path.join(tmp, "long.fastq.gz"), 2)
    st = ctx.sync()
    assert rc == 0 and bases == 2 * sum(r.size for r in reads) and lookups == 2 * sum(r.size - 30 for r in reads if r.size >= 31)
    assert st.hits == want.hits and st.valid_windows == want.valid_windows and np.array_equal(t.counts(2), t.counts(1))
    t.free()
    ctx.ingest_reset()


def test_gpu_ingest_many_short_lines_per_block(s2, ctx, tmp_path):
    """more lines than threads in one 16 KB text block (the per-block loops run in several rounds): FASTA wrapped at 12
    columns and at 3 (4,096 lines per block), FASTQ reads of 31-40 bases"""
    from strainer2_b200 import synth
    tmp = str(tmp_path)
    ctx.ingest_reset()
    rng = synth.rng_for(13, 0)
    strain = synth.genome(rng, 200_000, 3, n_runs=2)
    synth.write_fasta(os.path.join(tmp, "strain.fa"), strain)
    t = s2.StrainTable(ctx, s2.load_flat(os.path.join(tmp, "strain.fa")), n_cols=4)
    rel = [synth.mutate(c, 0.01, rng) for c in strain] + synth.genome(rng, 300_000, 2)
    clean = [np.where(c == ord("N"), ord("A"), c).astype(np.uint8) for c in rel]
    reads = [synth.sample_reads(rng, clean, 30_000, n, sub_rate=0.005) for n in (31, 33, 40)]
    fq = b"".join(b"@%d\n%s\n+\n%s\n" % (i, r.tobytes(), b"I" * r.size) for rr in reads for i, r in enumerate(rr))
    for name, text, ok in (("w12.fa.gz", synth.fasta_bytes(rel, 12), True), ("w3.fa.gz", synth.fasta_bytes(rel[:1], 3), True), ("short.fq.gz", fq, True)):
        open(os.path.join(tmp, name), "wb").write(synth.bgzf_bytes(text))
        t.clear_counts(1); t.clear_counts(2)
        want = ctx.scan_count(t, s2.load_flat(os.path.join(tmp, name)), 1)
        rc, _, _ = ctx.ingest_count_file(t, os.path.join(tmp, name), 2)
        st = ctx.sync()
        if ok:
            assert rc == 0 and st.hits == want.hits > 10_000 and st.valid_windows == want.valid_windows, name
            assert np.array_equal(t.counts(2), t.counts(1)), name
        else:
            assert rc == 1 and st.hits == 0 and int(t.counts(2).sum()) == 0, name
    t.free()


def test_gpu_ingest_groups_of_small_files(s2, ctx, tmp_path, monkeypatch):
    """many files per call: small files share a chunk (texts back to back); an irregular member sends its group back
    to be run file by file; rc_each tells which files were not handled; counters = sum over the handled files"""
    from strainer2_b200 import synth
    monkeypatch.setenv("S2_INGEST_CHUNK_MB", "2")
    monkeypatch.setenv("S2_INGEST_TEXT_MB", "8")
    ctx.ingest_reset()
    tmp = str(tmp_path)
    rng = synth.rng_for(11, 0)
    strain = synth.genome(rng, 300_000, 4, n_runs=2)
    synth.write_fasta(os.path.join(tmp, "strain.fa"), strain)
    t = s2.StrainTable(ctx, s2.load_flat(os.path.join(tmp, "strain.fa")), n_cols=4)
    clean = [np.where(c == ord("N"), ord("A"), c).astype(np.uint8) for c in strain]
    files, handled = [], []
    # 30 FASTA genomes of 0.2 - 1.2 Mb (several groups of 8 MB text), relatives of the strain among them
    for i in range(30):
        contigs = [synth.mutate(c, 0.01 * (1 + i % 3), rng) for c in strain[: 1 + i % 4]] + [synth.random_bases(rng, 50_000 * (1 + i % 5))]
        text = synth.fasta_bytes(contigs, 60 + i)
        if i == 7:
            text = text[:-1]                                  # no final newline inside a group: irregular there, fine alone
        if i == 12:
            text = text.replace(b"\n", b"\r\n", 5)             # irregular wherever it is
        files.append(("g%d.fa.gz" % i, synth.bgzf_bytes(text)))
        handled.append(i != 12)
    # FASTQ files (a group of their own), one truncated, one with a record count that is not a multiple of 4 lines
    for i in range(6):
        reads = synth.sample_reads(rng, clean, 4000 + 500 * i, 100 + 10 * i, sub_rate=0.01, n_rate=1e-4)
        text = synth.fastq_bytes(reads)
        if i == 2:
            text = text[:-40]
        if i == 4:
            text += b"@half\nACGT\n"
        files.append(("r%d.fq.gz" % i, synth.bgzf_bytes(text)))
        handled.append(i not in (2, 4))
    files.append(("ordinary.fa.gz", __import__("gzip").compress(synth.fasta_bytes(strain))))      # single-member .gz: software gunzip stage
    handled.append(True)
    files.append(("big.fa.gz", synth.bgzf_bytes(synth.fasta_bytes([synth.mutate(np.concatenate(clean), 0.02, rng)] * 30, 80))))    # 9 MB of text: streamed
    handled.append(True)
    for name, data in files:
        open(os.path.join(tmp, name), "wb").write(data)
    # expected: the host reader on the handled files
    want_hits = 0
    for (name, _), h in zip(files, handled):
        if h:
            want_hits += ctx.scan_count(t, s2.load_flat(os.path.join(tmp, name)), 1).hits
    assert want_hits > 100_000
    images = [np.frombuffer(d, dtype=np.uint8) for _, d in files]
    rcs, bases, lookups = ctx.ingest_count_mem_batch(t, [a.ctypes.data for a in images], [a.size for a in images], 2)
    st = ctx.sync()
    assert [rc == 0 for rc in rcs] == handled
    assert st.hits == want_hits and np.array_equal(t.counts(2), t.counts(1))
    rcs, bases2, lookups2 = ctx.ingest_count_files(t, [os.path.join(tmp, n) for n, _ in files], 3)
    st = ctx.sync()
    assert [rc == 0 for rc in rcs] == handled and bases2 == bases
    assert st.hits == want_hits and np.array_equal(t.counts(3), t.counts(1))
    # asynchronous jobs: three in flight (the same files into two columns, and once more), waited for in order
    t.clear_counts(2); t.clear_counts(3)
    ptrs, sizes = [a.ctypes.data for a in images], [a.size for a in images]
    jobs = [ctx.ingest_submit_mem_batch(t, ptrs, sizes, 2), ctx.ingest_submit_mem_batch(t, ptrs, sizes, 3), ctx.ingest_submit_mem_batch(t, ptrs, sizes, 3)]
    for j in jobs:
        rcs, b3, _ = ctx.ingest_wait(j)
        assert [rc == 0 for rc in rcs] == handled and b3 == bases
    st = ctx.sync()
    assert st.hits == 3 * want_hits and np.array_equal(t.counts(2), t.counts(1)) and np.array_equal(t.counts(3), 2 * t.counts(1))
    t.free()
    ctx.ingest_reset()


def test_executable_with_bgzf_inputs_matches_oracle(s2, golden_dir, tmp_path):
    """-B lists mixing BGZF, plain and ordinary .gz FASTQ (GPU ingest for the first two, host reader for the rest,
    and for every irregular golden edge case re-packed as BGZF): stdout bytes equal the oracle's"""
    from strainer2_b200 import synth
    tmp = str(tmp_path)
    strain, reads = _ingest_fixture(s2, tmp, 40_000, seed=2)
    synth.write_bgzf(os.path.join(tmp, "a.fastq.gz"), synth.fastq_bytes(reads[:20_000]))
    open(os.path.join(tmp, "b.fastq"), "wb").write(synth.fastq_bytes(reads[20_000:30_000]))
    synth.write_reads_fastq(os.path.join(tmp, "c.fastq.gz"), reads[30_000:])
    edge = os.path.join(golden_dir, "count_edge")
    names = []
    for f in ("m1_reads.fastq.gz", "m3_truncated_quality.fastq", "m4_crlf.fastq", "m2_multiline_reads.fa"):
        raw = ou.gunzip(os.path.join(edge, f)) if f.endswith(".gz") else open(os.path.join(edge, f), "rb").read()
        synth.write_bgzf(os.path.join(tmp, "edge_" + f + ".bgz"), raw)
        names.append("edge_" + f + ".bgz")
    open(os.path.join(tmp, "A.txt"), "w").write("")
    open(os.path.join(tmp, "B.txt"), "w").write("\n".join(["a.fastq.gz", "b.fastq", "c.fastq.gz"] + names) + "\n")
    args = ["-r", "strain.fa", "-A", "A.txt", "-B", "B.txt"]
    o = ou.oracle_cli(["count"] + args, cwd=tmp)
    assert o.returncode == 0
    for env in ({}, {"S2_GPU_INGEST": "0"}, {"S2_THREADS": "1", "S2_GPU_INGEST_PLAIN": "0"}):
        p = s2.run_kmer_scrub_count(args, cwd=tmp, env=env)
        assert p.returncode == 0, p.stderr
        assert p.stdout == o.stdout, env


def test_strain_detect_gpu_ingest_matches_oracle_including_stale_state(s2, tmp_path):
    """strain_detect on BGZF / plain FASTQ (GPU ingest) with reads shorter than 31 sprinkled in (the reference's
    stale-state behaviour), SE + PE + PEI lines, a PE2 that ends early on a short read (silent) and one that
    ends early on a long read (error exit)"""
    from strainer2_b200 import synth
    tmp = str(tmp_path)
    rng = synth.rng_for(8, 0)
    strain = synth.genome(rng, 200_000, 4, n_runs=2)
    synth.write_fasta(os.path.join(tmp, "strain.fa"), strain)
    clean = [np.where(c == ord("N"), ord("G"), c).astype(np.uint8) for c in strain]
    c0 = clean[0].tobytes()
    with open(os.path.join(tmp, "inf.txt"), "wb") as f:
        for i in range(0, len(c0) - 31, 60):
            f.write(c0[i:i + 31] + b"\n")

    def fq(reads, shorten_every, seed):
        r = np.random.default_rng(seed)
        out = []
        for i, x in enumerate(reads):
            s = x.tobytes()
            if i % shorten_every == 3:
                s = s[:int(r.choice([0, 7, 30]))]
            out.append(b"@r%d\n%s\n+\n%s\n" % (i, s, b"I" * len(s)))
        return out

    r1 = fq(synth.sample_reads(rng, clean + synth.genome(rng, 200_000, 2), 6000, 120, sub_rate=0.004, n_rate=3e-4), 9, 1)
    r2 = fq(synth.sample_reads(rng, clean + synth.genome(rng, 200_000, 2), 6000, 120, sub_rate=0.004, n_rate=3e-4), 11, 2)
    synth.write_bgzf(os.path.join(tmp, "a_R1.fastq.gz"), b"".join(r1))
    synth.write_bgzf(os.path.join(tmp, "a_R2.fastq.gz"), b"".join(r2))
    open(os.path.join(tmp, "b_se.fastq"), "wb").write(b"".join(r1[:3000]))
    inter = [x for pair in zip(r1[:2000], r2[:2000]) for x in pair] + [r1[2000]]          # odd count: PE2 runs out
    synth.write_bgzf(os.path.join(tmp, "c_inter.fastq.gz"), b"".join(inter[:-1]))
    # PE2 shorter than PE1 and ending on a short (stale < 31) read: the reference silently keeps going
    synth.write_bgzf(os.path.join(tmp, "d_R2_short.fastq.gz"), b"".join(r2[:1500]) + b"@s\nACGT\n+\nIIII\n")
    open(os.path.join(tmp, "batch.txt"), "w").write(
        "PE\ta_R1.fastq.gz\ta_R2.fastq.gz\nSE\tb_se.fastq\nPEI\tc_inter.fastq.gz\nPE\ta_R1.fastq.gz\td_R2_short.fastq.gz\nse\ta_R2.fastq.gz\n")
    args = ["-r", "strain.fa", "-a", "inf.txt", "-B", "batch.txt"]
    o = ou.oracle_cli(["detect"] + args + ["-m", os.path.join(tmp, "msg")], cwd=tmp)
    assert o.returncode == 0, o.stderr
    assert o.stdout.count(b"\n") > 500
    # S2_GPUS: batch lines are sharded over table replicas (as many as there are GPUs; one here on a single-GPU box)
    for env in ({}, {"S2_GPU_INGEST": "0"}, {"S2_THREADS": "1", "S2_GPU_INGEST_PLAIN": "0"}, {"S2_GPUS": "4"}, {"S2_GPUS": "2", "S2_GPU_INGEST": "0"},
                {"S2_GZ_THREADS": "4"}):                                 # parallel gzip writer: same text, other compressed bytes
        out = os.path.join(tmp, "hits.gz")
        p = s2.run_strain_detect(args + ["-o", out], cwd=tmp, env=env)
        assert p.returncode == 0, p.stderr
        assert ou.gunzip(out) == o.stdout, env
        assert p.stdout == open(os.path.join(tmp, "msg"), "rb").read()
    # PE2 ends early on a long read: error exit with the reference's message
    synth.write_bgzf(os.path.join(tmp, "e_R2.fastq.gz"), b"".join(x for x in r2[:100] if len(x) > 200))
    o = ou.oracle_cli(["detect", "-r", "strain.fa", "-a", "inf.txt", "-b", "a_R1.fastq.gz", "-c", "e_R2.fastq.gz", "-t", "PE", "-m", os.path.join(tmp, "m2")], cwd=tmp)
    p = s2.run_strain_detect(["-r", "strain.fa", "-a", "inf.txt", "-b", "a_R1.fastq.gz", "-c", "e_R2.fastq.gz", "-t", "PE", "-o", os.path.join(tmp, "x.gz")], cwd=tmp)
    assert o.returncode == 1 and p.returncode == 1
    assert p.stderr == o.stderr


@pytest.mark.parametrize("small_pieces", [False, True], ids=["defaults", "streamed_in_pieces"])
def test_strain_detect_gpu_ingest_of_fasta_reads_and_ordinary_gz(s2, tmp_path, small_pieces):
    """strain_detect on what the reference's own test/target_metagenomes.txt holds - two-line FASTA reads in ordinary
    single-member .gz - plus BGZF / plain FASTA reads and ordinary-.gz FASTQ: all of them inflated and split on the GPU
    (chunk-parallel gunzip, record kernels with two lines per record), short reads and the stale-state rules included;
    a FASTA file with wrapped sequences is not two-line and goes to the host parser.  Output equals the oracle's"""
    import gzip
    from strainer2_b200 import synth
    tmp = str(tmp_path)
    rng = synth.rng_for(9, 1)
    strain = synth.genome(rng, 200_000, 4, n_runs=2)
    synth.write_fasta(os.path.join(tmp, "strain.fa"), strain)
    clean = [np.where(c == ord("N"), ord("G"), c).astype(np.uint8) for c in strain]
    c0 = clean[0].tobytes()
    with open(os.path.join(tmp, "inf.txt"), "wb") as f:
        for i in range(0, len(c0) - 31, 60):
            f.write(c0[i:i + 31] + b"\n")

    def recs(reads, shorten_every, seed, fasta):
        r = np.random.default_rng(seed)
        out = []
        for i, x in enumerate(reads):
            q = x.tobytes()
            if i % shorten_every == 3:
                q = q[:int(r.choice([0, 7, 30]))]
            out.append(b">r%d 1\n%s\n" % (i, q) if fasta else b"@r%d\n%s\n+\n%s\n" % (i, q, b"I" * len(q)))
        return out

    n = 40_000                                                            # 6 MB of FASTA text, 10 MB of FASTQ per file
    a1 = synth.sample_reads(rng, clean + synth.genome(rng, 200_000, 2), n, 150, sub_rate=0.004, n_rate=3e-4)
    a2 = synth.sample_reads(rng, clean + synth.genome(rng, 200_000, 2), n, 150, sub_rate=0.004, n_rate=3e-4)
    f1, f2 = recs(a1, 9, 1, True), recs(a2, 11, 2, True)
    q1, q2 = recs(a1, 7, 3, False), recs(a2, 13, 4, False)
    w = lambda name, data: open(os.path.join(tmp, name), "wb").write(data)
    w("a_PE1.fasta.gz", gzip.compress(b"".join(f1), 6))
    w("a_PE2.fasta.gz", gzip.compress(b"".join(f2), 6))
    w("b_se.fasta", b"".join(f1[:9000]))
    synth.write_bgzf(os.path.join(tmp, "c_inter.fasta.gz"), b"".join(x for pair in zip(f1[:5000], f2[:5000]) for x in pair))
    w("d_R1.fastq.gz", gzip.compress(b"".join(q1), 1))
    w("d_R2.fastq.gz", gzip.compress(b"".join(q2), 9))
    w("e_R2_short.fasta.gz", gzip.compress(b"".join(f2[:1500]) + b">s\nACGT\n", 6))       # PE2 runs out on a short read: silent
    wrapped = b"".join(b">w%d\n%s\n%s\n" % (i, x.tobytes()[:80], x.tobytes()[80:]) for i, x in enumerate(a1[:2000]))
    w("f_wrapped.fasta.gz", gzip.compress(wrapped, 6))
    open(os.path.join(tmp, "batch.txt"), "w").write(
        "PE\ta_PE1.fasta.gz\ta_PE2.fasta.gz\nSE\tb_se.fasta\nPEI\tc_inter.fasta.gz\nPE\td_R1.fastq.gz\td_R2.fastq.gz\n"
        "PE\ta_PE1.fasta.gz\te_R2_short.fasta.gz\nSE\tf_wrapped.fasta.gz\nse\ta_PE2.fasta.gz\n")
    args = ["-r", "strain.fa", "-a", "inf.txt", "-B", "batch.txt"]
    o = ou.oracle_cli(["detect"] + args + ["-m", os.path.join(tmp, "msg")], cwd=tmp)
    assert o.returncode == 0, o.stderr
    assert o.stdout.count(b"\n") > 2000
    small = {"S2_GZ_BATCH_MB": "1", "S2_INGEST_CHUNK_MB": "1", "S2_INGEST_TEXT_MB": "4"} if small_pieces else {}
    # (S2_PARTITION_MIN_MB=0: the table counts as "larger than L2" - the GPU ingest then probes it with the direct detect kernel)
    for env in ({"S2_STATS": "1"}, {"S2_THREADS": "1"}, {"S2_GPU_INGEST": "0"}, {"S2_GPU_GUNZIP": "0"}, {"S2_STATS": "1", "S2_PARTITION_MIN_MB": "0"}):
        env = dict(env, **small)
        out = os.path.join(tmp, "hits.gz")
        p = s2.run_strain_detect(args + ["-o", out], cwd=tmp, env=env)
        assert p.returncode == 0, p.stderr
        assert ou.gunzip(out) == o.stdout, env
        assert p.stdout == open(os.path.join(tmp, "msg"), "rb").read()
        if "S2_STATS" in env:                                             # 10 file reads, all but the wrapped one on the GPU
            assert b"files_gpu_ingest=9 files_host_reader=1 " in p.stderr, p.stderr


@pytest.mark.parametrize("chunk_mb", [(2, 8), None], ids=["streamed", "one_chunk"])
def test_gpu_ingest_fasta_genomes_equal_host_reader(s2, ctx, golden_dir, tmp_path, monkeypatch, chunk_mb):
    monkeypatch.setenv("S2_GPU_INGEST_PLAIN", "1")
    _ingest_chunks(ctx, monkeypatch, chunk_mb)
    """multi-line FASTA (BGZF and plain) through the GPU ingest: sequence lines joined on the device, headers become
    separators, chunk boundaries inside long contigs; irregular FASTA goes back to the host reader"""
    from strainer2_b200 import synth
    tmp = str(tmp_path)
    rng = synth.rng_for(9, 0)
    strain = synth.genome(rng, 400_000, 3, n_runs=3)
    synth.write_fasta(os.path.join(tmp, "strain.fa"), strain)
    # ~35 MB of FASTA text: one 24 Mb contig (crosses ingest chunks) + relatives of the strain + short contigs
    big = [synth.random_bases(rng, 24_000_000)] + [synth.mutate(c, 0.01, rng) for c in strain] * 20 + [synth.random_bases(rng, n) for n in (5, 30, 31, 200)]
    big[0][1_000_000:1_400_000] = strain[0][:400_000] if strain[0].size >= 400_000 else big[0][1_000_000:1_400_000]
    import io
    def fasta_text(recs, wrap):
        out = io.BytesIO()
        for i, r in enumerate(recs):
            b = r.tobytes()
            out.write(b">c%d some description\n" % i)
            for j in range(0, len(b), wrap):
                out.write(b[j:j + wrap] + b"\n")
            if i % 7 == 3:
                out.write(b"\n")                                # empty lines inside / between records are legal
        return out.getvalue()
    text = fasta_text(big, 70)
    synth.write_bgzf(os.path.join(tmp, "g.fa.gz"), text)
    open(os.path.join(tmp, "g.fa"), "wb").write(text[:-1])           # plain, unterminated last line
    t = s2.StrainTable(ctx, s2.load_flat(os.path.join(tmp, "strain.fa")), n_cols=4)
    want = ctx.scan_count(t, s2.load_flat(os.path.join(tmp, "g.fa.gz")), 1)
    assert want.hits > 100_000
    for col, name in ((2, "g.fa.gz"), (3, "g.fa")):
        rc, bases, lookups = ctx.ingest_count_file(t, os.path.join(tmp, name), col)
        st = ctx.sync()
        assert rc == 0
        assert bases == sum(r.size for r in big)
        assert st.hits == want.hits and st.valid_windows == want.valid_windows
        assert np.array_equal(t.counts(col), t.counts(1))
    t.clear_counts(2)
    for name, bad in {"junk": b"junk\n" + text[:5000], "crlf": text[:5000].replace(b"\n", b"\r\n"),
                      "fastq_inside": text[:5000] + b"\n@r\nACGT\n+\nIIII\n", "plus_line": text[:5000] + b"\n+\n"}.items():
        synth.write_bgzf(os.path.join(tmp, name + ".fa.gz"), bad)
        assert ctx.ingest_count_file(t, os.path.join(tmp, name + ".fa.gz"), 2)[0] == 1, name
        assert int(t.counts(2).sum()) == 0
    t.free()
    ctx.ingest_reset()
    if chunk_mb:
        return
    # the executable on the golden count case with every input re-packed as BGZF: bytes equal the reference's
    d = os.path.join(golden_dir, "count_edge")
    for lst in ("listA.txt", "listB.txt", "listC.txt"):
        names = [l.strip() for l in open(os.path.join(d, lst)) if l.strip()]
        new = []
        for n in names:
            raw = ou.gunzip(os.path.join(d, n)) if n.endswith(".gz") else open(os.path.join(d, n), "rb").read()
            synth.write_bgzf(os.path.join(tmp, n.replace("/", "_") + ".bgz"), raw)
            new.append(os.path.join(tmp, n.replace("/", "_") + ".bgz"))
        open(os.path.join(tmp, lst), "w").write("".join(x + "\n" for x in new))
    p = s2.run_kmer_scrub_count(["-r", os.path.join(d, "ref.fa.gz"), "-A", os.path.join(tmp, "listA.txt"), "-B", os.path.join(tmp, "listB.txt")], cwd=tmp)
    assert p.returncode == 0, p.stderr
    assert p.stdout == open(os.path.join(d, "expected_AB.tsv"), "rb").read()
