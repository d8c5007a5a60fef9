"""CPU-side tests (no GPU needed): the C ABI loads and exports every declared symbol, the host half of
the library (reader, row-order replay, formatter, codecs) agrees with the oracle and the golden
vectors, and the bit primitives shared with the kernels agree with string arithmetic."""
import ctypes as C
import os
import random
import re
import subprocess

import numpy as np
import pytest

import oracle_util as ou

ROOT = ou.ROOT


def test_abi_exports_every_declared_symbol():
    import strainer2_b200 as s2
    from strainer2_b200 import _lib
    header = open(os.path.join(ROOT, "include", "strainer2_b200.h")).read()
    declared = set(re.findall(r"\b(s2_[a-z0-9_]+)\s*\(", header))
    declared -= {"s2_roworder_emulate()"}
    assert len(declared) >= 40
    for name in sorted(declared):
        assert hasattr(s2.lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert s2.lib.s2_abi_version() == 1


def test_no_gpu_means_loud_failure_not_fallback():
    import strainer2_b200 as s2
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(s2.S2Error):
        s2.Context(0)


def test_up2bit_codec_known_answers():
    import strainer2_b200 as s2
    L = ou.lib()
    r = random.Random(11)
    for _ in range(300):
        n = r.randint(1, 32)
        s = "".join(r.choice("ACGTacgtNRY") for _ in range(n)).encode()
        v = s2.encode_2bit(s)
        assert v == L.s2o_encode_2bit(s, n)
        out = C.create_string_buffer(40)
        L.s2o_decode_2bit(v, n, out)
        assert s2.decode_2bit(v, n) == out.value
    assert s2.encode_2bit(b"ACTG") == 0b00011011
    assert s2.decode_2bit(0b00011011, 4) == b"ACTG"


def test_kmer_ascii_roundtrip_and_orientation():
    import strainer2_b200 as s2
    r = random.Random(3)
    assert s2.kmer_to_ascii(s2.kmer_from_ascii(b"A" * 31)) == b"T" * 31
    assert s2.kmer_to_ascii(s2.kmer_from_ascii(b"ATGCAAATGACGCTTGTATCAGCGGATTTCA")) == b"TGAAATCCGCTGATACAAGCGTCATTTGCAT"
    for _ in range(500):
        w = "".join(r.choice("ACGTacgt") for _ in range(31)).encode()
        assert s2.kmer_to_ascii(s2.kmer_from_ascii(w)) == ou.orient(w.upper())
    assert s2.kmer_from_ascii(b"ACGTN" + b"A" * 26) is None


def test_reader_matches_reference_parser_dumps(golden_dir):
    import strainer2_b200 as s2
    d = os.path.join(golden_dir, "count_edge")
    n = 0
    for f in sorted(os.listdir(os.path.join(d, "kseq"))):
        src = f[:-len(".dump.gz")]
        rd = s2.Reader(os.path.join(d, src))
        out = []
        while True:
            ret, seq = rd.next()
            if ret < 0:
                out.append(b"%d\t%d\t<END>\n" % (ret, len(seq)))
                break
            out.append(b"%d\t%d\t%s\n" % (ret, len(seq), seq))
        rd.close()
        assert b"".join(out) == ou.gunzip(os.path.join(d, "kseq", f)), src
        n += 1
    assert n >= 10


def test_reader_randomised_against_oracle_reader(tmp_path):
    """fuzz: random mixes of FASTA/FASTQ fragments, CR/LF, blank lines, stray '>' '@' '+'"""
    import strainer2_b200 as s2
    r = random.Random(99)
    pieces = [">h1 c\n", "@q1\n", "ACGT", "acgtn", "\n", "\r\n", "+\n", "+x\n", "IIII", ">", "@", " ", "\t", "N" * 40,
              "ACGTACGTACGTACGTACGTACGTACGTACGTACGT\n", "\r", "\n\n", "+", "GATTACA\r\n"]
    for t in range(120):
        text = "".join(r.choice(pieces) for _ in range(r.randint(0, 60)))
        p = tmp_path / f"f{t}.txt"
        p.write_bytes(text.encode())
        o = ou.oracle_cli(["kseq", str(p)]).stdout
        rd = s2.Reader(str(p))
        out = []
        while True:
            ret, seq = rd.next()
            if ret < 0:
                out.append(b"%d\t%d\t<END>\n" % (ret, len(seq)))
                break
            out.append(b"%d\t%d\t%s\n" % (ret, len(seq), seq))
        rd.close()
        assert b"".join(out) == o, text


def test_reader_flags_damaged_gzip_data(tmp_path):
    """Corrupt DEFLATE data makes gzread return -1.  The reference never returns from such a file (kseq takes only a
    0 from gzread for the end, src/kseq.h:72,:99, and re-reads the error for ever: measured, 100 % CPU until killed),
    so there is nothing to match: the reader ends the stream there and says so, load_flat raises, the executables exit
    with an error.  A file that is merely cut short is an ordinary end of file for zlib, the reference and the reader."""
    import gzip
    import strainer2_b200 as s2
    from strainer2_b200 import synth
    rng = synth.rng_for(9, 0)
    reads = synth.sample_reads(rng, synth.genome(rng, 100_000, 2), 6000, 100)
    text = synth.fastq_bytes(reads)
    good = synth.bgzf_bytes(text)
    members, off = [], 0
    while off < len(good):
        bsize = int.from_bytes(good[off + 16:off + 18], "little") + 1
        members.append((off, bsize))
        off += bsize
    assert len(members) > 4
    o, b = members[len(members) // 2]
    junk = bytes((((i * 2654435761) & 0xFFFFFFFF) >> 13) & 0xFF for i in range(b - 26))
    z = bytearray(gzip.compress(text, 6))
    for i in range(len(z) // 2, len(z) // 2 + 64):
        z[i] ^= 0x5A
    crc = bytearray(gzip.compress(text, 6))
    crc[-6] ^= 1                                                  # only the CRC-32 in the trailer is wrong
    cases = {"bgzf_junk_member.gz": (good[:o + 18] + junk + good[o + b - 8:], True),
             "gzip_flipped_bytes.gz": (bytes(z), True),
             "gzip_bad_crc.gz": (bytes(crc), True),
             "gzip_cut_short.gz": (gzip.compress(text, 6)[:-5000], False),
             "good.gz": (good, False)}
    first = reads[0].tobytes()
    for name, (data, damaged) in cases.items():
        p = tmp_path / name
        p.write_bytes(data)
        rd = s2.Reader(str(p))
        n = 0
        while True:
            ret, seq = rd.next()
            if ret == -1:
                break
            if ret == -2:                                          # garbage text in front of the damage may parse as a broken record
                continue
            assert n > 0 or seq == first, name
            n += 1
        assert rd.damaged == damaged, name                         # read to the end: the damage has been met
        rd.close()
        assert (n == len(reads)) == (name == "good.gz"), (name, n)
    with pytest.raises(s2.S2Error):
        s2.load_flat(str(tmp_path / "bgzf_junk_member.gz"))
    with pytest.raises(s2.S2Error):
        s2.load_flat(str(tmp_path / "gzip_bad_crc.gz"))
    assert len(s2.load_flat(str(tmp_path / "gzip_cut_short.gz"))) < len(s2.load_flat(str(tmp_path / "good.gz"))) == reads.size + len(reads)


# ---- the DEFLATE / gzip decoder written for host and device (groundwork, not on the product path) ----------
@pytest.fixture(scope="module")
def infl(tmp_path_factory):
    so = tmp_path_factory.mktemp("infl") / "libinfl.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++",
                           os.path.join(ROOT, "tests", "sim", "inflate_harness.cpp"), "-o", str(so)])
    L = C.CDLL(str(so))
    L.sim_gunzip.restype = C.c_int
    L.sim_inflate_raw.restype = C.c_int

    class Infl:
        @staticmethod
        def gunzip(data, cap):
            dst = (C.c_ubyte * (cap + 64))()
            C.memset(dst, 0xAB, cap + 64)
            out = C.c_ulonglong()
            rc = L.sim_gunzip(data, C.c_ulonglong(len(data)), dst, C.c_ulonglong(cap), C.byref(out))
            raw = bytes(dst)
            assert raw[cap:] == b"\xAB" * 64, "wrote past the destination"
            return rc, raw[:out.value]

        @staticmethod
        def raw(data, cap):
            dst = (C.c_ubyte * (cap + 64))()
            C.memset(dst, 0xAB, cap + 64)
            out, used = C.c_ulonglong(), C.c_ulonglong()
            rc = L.sim_inflate_raw(data, C.c_ulonglong(len(data)), dst, C.c_ulonglong(cap), C.byref(out), C.byref(used))
            raw = bytes(dst)
            assert raw[cap:] == b"\xAB" * 64, "wrote past the destination"
            return rc, raw[:out.value], used.value
    return Infl


def _inflate_texts():
    r = random.Random(1)
    fasta = b">c1 test\n" + b"\n".join(bytes(r.choice(b"ACGT") for _ in range(80)) for _ in range(1500)) + b"\n"
    fastq = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, bytes(r.choice(b"ACGT") for _ in range(150)), b"I" * 150) for i in range(700))
    return {"empty": b"", "one": b"A", "runs": b"A" * 100000 + b"CG" * 5000, "fasta": fasta, "fastq": fastq,
            "random": bytes(r.getrandbits(8) for _ in range(70000)), "prose": open(os.path.join(ROOT, "DESIGN.md"), "rb").read()}


def test_inflate_matches_zlib_on_every_block_type(infl):
    """stored / fixed / dynamic blocks, every zlib strategy and level, long overlapping matches, several members, a
    header with a file name, sync-flush pieces: the decoder yields zlib's bytes and knows where the stream ended"""
    import gzip
    import io
    import zlib
    texts = _inflate_texts()
    for name, t in texts.items():
        for lvl in (0, 1, 6, 9):
            rc, o = infl.gunzip(gzip.compress(t, lvl), len(t))
            assert rc == 0 and o == t, (name, lvl, rc)
        for strat in (zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED):
            co = zlib.compressobj(6, zlib.DEFLATED, -15, 8, strat)
            z = co.compress(t) + co.flush()
            rc, o, used = infl.raw(z + b"TRAILING", len(t))
            assert rc == 0 and o == t and used == len(z), (name, strat, rc, used, len(z))
        if len(t) > 1:                                            # a destination one byte too small
            rc, o = infl.gunzip(gzip.compress(t, 6), len(t) - 1)
            assert rc == -7 and o == t[:len(o)], (name, rc)
    buf = io.BytesIO()
    with gzip.GzipFile(filename="some_name.fa", mode="wb", fileobj=buf, compresslevel=6) as f:
        f.write(texts["fasta"])
    z = buf.getvalue() + gzip.compress(texts["fastq"], 9) + gzip.compress(b"") + b"\0\0\0\0"
    rc, o = infl.gunzip(z, len(texts["fasta"]) + len(texts["fastq"]))
    assert rc == 0 and o == texts["fasta"] + texts["fastq"]
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    fq = texts["fastq"]
    z = b"".join(co.compress(fq[i:i + 5000]) + co.flush(zlib.Z_SYNC_FLUSH) for i in range(0, len(fq), 5000)) + co.flush()
    rc, o, used = infl.raw(z, len(fq))
    assert rc == 0 and o == fq and used == len(z)
    assert infl.gunzip(b"not a gzip file at all, just text\n", 100)[0] == -8


def test_inflate_survives_damaged_input(infl):
    """every truncation is an error; random bit flips end in an error or in the bytes zlib also produces; nothing is
    ever written past the destination (the harness checks a canary behind it on every call)"""
    import gzip
    import zlib
    r = random.Random(2)
    t = _inflate_texts()["fastq"][:20000]
    z = gzip.compress(t, 6)
    for cut in range(0, len(z), 5):
        assert infl.gunzip(z[:cut], len(t))[0] < 0, cut
    for _ in range(1500):
        zz = bytearray(z)
        for _ in range(r.randint(1, 4)):
            zz[r.randrange(len(zz))] ^= 1 << r.randrange(8)
        rc, o = infl.gunzip(bytes(zz), len(t))
        if rc == 0:                                               # the CRC-32 is not checked here: compare with zlib's raw inflate
            d = zlib.decompressobj(-15).decompress(bytes(zz)[10:])
            assert d == o


def test_inflate_fuzz_under_sanitizers(tmp_path):
    """the decoder on damaged / truncated streams and short destinations, built with -fsanitize=address,undefined and run
    on exact-size heap buffers: every access stays inside its buffer and intact streams decode to their text"""
    exe = tmp_path / "inflate_fuzz"
    cc = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                         os.path.join(ROOT, "tests", "sim", "inflate_fuzz.cpp"), "-o", str(exe), "-lz"], capture_output=True)
    if cc.returncode != 0:
        pytest.skip("no sanitizer runtime for this g++: " + cc.stderr.decode()[-200:])
    p = subprocess.run([str(exe), "300"], capture_output=True, timeout=600)
    assert p.returncode == 0, (p.stdout + p.stderr).decode()[-2000:]
    assert b"fuzz done" in p.stdout


def test_gunzip_fuzz_under_sanitizers(tmp_path):
    """the chunk-parallel gunzip's sub-chunk decoder (block finder + symbol loop of s2_gunzip.cuh, host build of the device
    source) on intact, damaged and truncated streams, built with -fsanitize=address,undefined: every sub-chunk gets a symbol
    region of its own allocated to the slot (32,768 markers, cap symbols, the guard slot) and the compressed words are
    allocated to the word, regions are sometimes far too small - no access outside them, intact streams chain to their
    exact size"""
    exe = tmp_path / "gunzip_fuzz"
    cc = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                         os.path.join(ROOT, "tests", "sim", "gunzip_fuzz.cpp"), "-o", str(exe), "-lz"], capture_output=True)
    if cc.returncode != 0:
        pytest.skip("no sanitizer runtime for this g++: " + cc.stderr.decode()[-200:])
    p = subprocess.run([str(exe), "40"], capture_output=True, timeout=600)
    assert p.returncode == 0, (p.stdout + p.stderr).decode()[-2000:]
    assert b"fuzz done" in p.stdout


def test_gunzip_device_control_flow_on_32_emulated_lanes(tmp_path):
    """the DEVICE branches of s2_gunzip.cuh on the CPU: 32 threads are the lanes of a warp, barriers stand where the lanes
    synchronise (__syncwarp, the ballots and broadcasts, the places that rely on the lanes running in step).  The block
    finder's queue of survivors, a lane's share of a match copy (lanes behind its end, the deferred store, the guard slot),
    table builds and header parsing over 32 lanes: the text of every stream is zlib's (FASTA, FASTQ, self-overlapping
    matches, long codes; levels 1 / 6 / 9; sub-chunks of 8 and 64 KB)"""
    exe = tmp_path / "gunzip_warp_emu"
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-pthread", os.path.join(ROOT, "tests", "sim", "gunzip_warp_emu.cpp"), "-o", str(exe), "-lz"])
    p = subprocess.run([str(exe)], capture_output=True, timeout=900)
    assert p.returncode == 0, (p.stdout + p.stderr).decode()[-2000:]
    assert b"warp emulation done" in p.stdout


def _djb2_str(s: bytes) -> int:
    h = 5381
    for c in s:
        h = (h * 33 + c) & 0xFFFFFFFF
    return h


def test_roworder_replay_matches_oracle_table_small_capacity():
    """the doubling rule (N++ >= M/2) and re-insertion order, exercised with tiny capacities"""
    import strainer2_b200 as s2
    L = ou.lib()
    r = random.Random(8)
    for cap in (10, 16, 37, 100):
        keys = list({("".join(r.choice("ACGT") for _ in range(31))).encode() for _ in range(r.randint(1, 400))})
        r.shuffle(keys)
        t = L.s2o_table_new(cap, 4)
        L.s2o_table_add.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_uint)]
        L.s2o_table_key_at.restype = C.c_char_p
        L.s2o_table_key_at.argtypes = [C.c_void_p, C.c_uint, C.c_void_p]
        for k in keys:
            L.s2o_table_add(t, k, None)
        want = [L.s2o_table_key_at(t, i, None) for i in range(len(keys))]
        order, final_cap = s2.roworder_emulate([_djb2_str(k) for k in keys], cap)
        assert [keys[i] for i in order] == want
        assert final_cap == L.s2o_table_capacity(t)
        L.s2o_table_free(t)


def test_roworder_and_formatter_reproduce_golden_table(golden_dir, tmp_path):
    """rebuild expected_ABC.tsv from (keys in first-occurrence order, counts) with the product's
    row-order replay + formatter; the key list comes from the oracle, the bytes must equal the
    reference's."""
    import strainer2_b200 as s2
    d = os.path.join(golden_dir, "count_edge")
    want = open(os.path.join(d, "expected_ABC.tsv"), "rb").read()
    kmers, vals = ou.parse_table(want)
    # first-occurrence order = order of first appearance in the reference genome
    rd = s2.Reader(os.path.join(d, "ref.fa.gz"))
    seen, first = set(), []
    while True:
        ret, seq = rd.next()
        if ret < 0:
            break
        s = seq.upper()
        for i in range(len(s) - 30):
            w = s[i:i + 31]
            if b"N" in w:
                continue
            k = ou.orient(w)
            if k not in seen:
                seen.add(k)
                first.append(k)
    rd.close()
    assert len(first) == len(kmers)
    row_of = {k: i for i, k in enumerate(kmers)}
    keys = np.array([s2.kmer_from_ascii(k) for k in first], dtype=np.uint64)
    assert all(s2.kmer_to_ascii(int(k)) == f for k, f in zip(keys, first))
    djb2 = np.array([_djb2_str(k) for k in first], dtype=np.uint32)
    order, _ = s2.roworder_emulate(djb2)
    cols = [np.array([vals[row_of[k]][c] for k in first], dtype=np.uint32) for c in range(4)]
    out = tmp_path / "t.tsv"
    s2.format_count_table(str(out), keys, order, cols, n_threads=3)
    assert out.read_bytes() == want


def test_formatter_prints_counters_with_percent_d(tmp_path):
    import strainer2_b200 as s2
    keys = np.array([s2.kmer_from_ascii(b"T" * 31)], dtype=np.uint64)
    cols = [np.array([v], dtype=np.uint32) for v in (1, 0x80000000, 0xFFFFFFFF)]
    out = tmp_path / "t.tsv"
    s2.format_count_table(str(out), keys, np.array([0], np.uint32), cols)
    assert out.read_bytes().split(b"\n")[1] == b"T" * 31 + b"\t1\t-2147483648\t-1"


# ---- bit primitives shared with the kernels, compiled for the host ---------------------------------
@pytest.fixture(scope="module")
def sim(tmp_path_factory):
    so = tmp_path_factory.mktemp("sim") / "libsim.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++",
                           os.path.join(ROOT, "tests", "sim", "kmer_prims_harness.cpp"), "-o", str(so)])
    L = C.CDLL(str(so))
    L.sim_rc16.restype = C.c_uint32
    L.sim_rc16.argtypes = [C.c_uint32]
    L.sim_extract31.restype = C.c_uint64
    L.sim_extract31.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint]
    L.sim_revcomp31.restype = C.c_uint64
    L.sim_revcomp31.argtypes = [C.c_uint64]
    L.sim_djb2.restype = C.c_uint32
    L.sim_djb2.argtypes = [C.c_uint64]
    L.sim_window_canon.restype = C.c_uint64
    L.sim_window_canon.argtypes = [C.c_char_p, C.c_uint, C.POINTER(C.c_int)]
    L.sim_hash.argtypes = [C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.sim_bucket.restype = C.c_uint32
    L.sim_bucket.argtypes = [C.c_uint32, C.c_uint32]
    return L


CODE = {65: 0, 67: 1, 71: 2, 84: 3}


def test_pack16_all_byte_values(sim):
    r = random.Random(1)
    for t in range(400):
        b = bytes(r.randrange(256) for _ in range(16)) if t % 2 else bytes(r.choice(b"ACGTacgtNn\n>@RYKM-. ") for _ in range(16))
        w, m = C.c_uint32(), C.c_uint32()
        sim.sim_pack16(b, C.byref(w), C.byref(m))
        for i, c in enumerate(b):
            up = c - 32 if 97 <= c <= 122 else c
            ok = up in CODE
            assert ((m.value >> (15 - i)) & 1) == (1 if ok else 0), (b, i)
            if ok:
                assert ((w.value >> (30 - 2 * i)) & 3) == CODE[up]
    # every single byte value in every position class
    for c in range(256):
        b = bytes([c] * 16)
        w, m = C.c_uint32(), C.c_uint32()
        sim.sim_pack16(b, C.byref(w), C.byref(m))
        up = c - 32 if 97 <= c <= 122 else c
        assert m.value == (0xFFFF if up in CODE else 0), c


def test_window_canon_matches_reference_orientation(sim):
    import strainer2_b200 as s2
    r = random.Random(2)
    for _ in range(300):
        s = bytes(r.choice(b"ACGTacgt" if r.random() < 0.8 else b"ACGTN\n") for _ in range(48))
        for j in range(16):
            valid = C.c_int()
            k = sim.sim_window_canon(s, j, C.byref(valid))
            w = s[j:j + 31].upper()
            good = all(c in b"ACGT" for c in w)
            assert valid.value == (1 if good else 0)
            if good:
                assert s2.kmer_to_ascii(k) == ou.orient(w), (s, j)


def test_revcomp_and_djb2_of_packed_kmer(sim):
    import strainer2_b200 as s2
    L = ou.lib()
    r = random.Random(4)
    for _ in range(500):
        w = "".join(r.choice("ACGT") for _ in range(31)).encode()
        v = 0
        for c in w:
            v = (v << 2) | CODE[c]
        rc = w.translate(bytes.maketrans(b"ACGT", b"TGCA"))[::-1]
        vr = 0
        for c in rc:
            vr = (vr << 2) | CODE[c]
        assert sim.sim_revcomp31(v) == vr
        assert sim.sim_djb2(v) == L.s2o_djb2(w)
    assert sim.sim_djb2(s2.kmer_from_ascii(b"T" * 31)) == 3948423441


def test_hash_spreads_real_kmers(sim):
    """bucket occupancy of canonical k-mers from a random genome stays near Poisson and fingerprints
    are legal fp16 'normal' bit patterns (0x0400..0x7BFF), never 0"""
    import strainer2_b200 as s2
    rng = np.random.default_rng(5)
    n = 200000
    keys = rng.integers(0, 1 << 62, size=n, dtype=np.uint64)
    # canonical k-mers are skewed towards large values (max of forward / reverse complement)
    keys[:50000] = [max(int(k), sim.sim_revcomp31(int(k))) for k in keys[:50000]]
    nb = n // 8                                     # load 0.5 with 16-slot buckets
    cnt = np.zeros(nb, dtype=np.int64)
    h, fp = C.c_uint32(), C.c_uint32()
    fps = set()
    for k in keys[:50000]:
        sim.sim_hash(int(k), C.byref(h), C.byref(fp))
        assert 0x0400 <= fp.value <= 0x7BFF
        fps.add(fp.value)
        cnt[sim.sim_bucket(h.value, nb)] += 1
    assert len(fps) > 20000
    lam = 50000 / nb
    assert abs(cnt.mean() - lam) < 1e-9
    assert cnt.var() < 1.3 * lam        # Poisson variance = lam
    # overlapping windows of one sequence (consecutive k-mers share 30 bases) must spread as well
    seq = rng.integers(0, 4, size=60000, dtype=np.uint64)
    cnt2 = np.zeros(nb, dtype=np.int64)
    fwd = 0
    for i, b in enumerate(seq):
        fwd = ((fwd << 2) | int(b)) & ((1 << 62) - 1)
        if i >= 30:
            k = max(fwd, sim.sim_revcomp31(fwd))
            sim.sim_hash(k, C.byref(h), C.byref(fp))
            cnt2[sim.sim_bucket(h.value, nb)] += 1
    lam2 = cnt2.sum() / nb
    assert cnt2.var() < 1.3 * lam2 and cnt2.max() <= 12


def test_gz_writer_single_stream_and_parallel_blocks(tmp_path):
    """s2_gz_writer: threads=0 is gzopen("wb9") + gzwrite (the reference's stream, src/strain_detect.c:299); threads>0
    deflates 256 KB blocks in parallel into ONE gzip member - any gunzip must give back the identical text"""
    import ctypes as C
    import gzip
    import random
    import zlib
    from strainer2_b200 import lib
    r = random.Random(3)
    line = lambda: b"sample_%d.fastq.gz\t%d\t%d\t%d\t%d\t%s\n" % (r.randint(0, 9), r.randint(0, 120), r.randint(0, 9), r.randint(0, 120), r.randint(0, 9),
                                                                   bytes(r.choice(b"ACGT") for _ in range(31)))
    big = b"".join(line() for _ in range(40_000))                     # ~2.7 MB: several rounds of blocks
    cases = {"empty": b"", "one_byte": b"x", "block_edge": big[:262144], "block_edge_plus": big[:262145], "big": big}
    for name, text in cases.items():
        for threads in (0, 1, 3, 8):
            path = str(tmp_path / f"{name}_{threads}.gz").encode()
            w = lib.s2_gz_writer_open(path, threads)
            assert w
            pos, step = 0, 1
            while pos < len(text):                                    # ragged writes: a few bytes up to hundreds of KB
                n = min(len(text) - pos, step)
                buf = text[pos:pos + n]
                assert lib.s2_gz_writer_write(w, buf, n) == 0
                pos += n
                step = min(step * 3 + 1, 400_000)
            assert lib.s2_gz_writer_close(w) == 0
            raw = open(path, "rb").read()
            assert gzip.decompress(raw) == text, (name, threads)
            d = zlib.decompressobj(31)                                # exactly one gzip member, nothing after it
            assert d.decompress(raw) == text and d.eof and d.unused_data == b""
            if threads == 0:                                          # the single stream is what zlib's gzwrite gives at level 9
                ref = str(tmp_path / "ref.gz")
                with open(ref, "wb") as f:
                    pass
                g = zlib.compressobj(9, zlib.DEFLATED, 31)
                assert len(raw) == len(g.compress(text) + g.flush())


# ---- chunk-parallel gunzip (strainer2_b200/csrc/s2_gunzip.cuh), host build of the device source ----------------------
@pytest.fixture(scope="module")
def pgz(tmp_path_factory):
    so = tmp_path_factory.mktemp("pgz") / "libpgz.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++",
                           os.path.join(ROOT, "tests", "sim", "gunzip_harness.cpp"), "-o", str(so)])
    L = C.CDLL(str(so))
    L.sim_pgunzip.restype = C.c_int

    cap = 1 << 23
    dst = (C.c_ubyte * cap)()

    def run(z, sub_bytes, ratio=64):
        out, stats = C.c_ulonglong(), (C.c_ulonglong * 4)()
        rc = L.sim_pgunzip(z, C.c_ulonglong(len(z)), C.c_uint(sub_bytes), C.c_uint(ratio), dst, C.c_ulonglong(cap), C.byref(out), stats)
        return rc, C.string_at(dst, out.value) if rc == 0 else b"", list(stats)
    return run


def test_parallel_gunzip_matches_zlib(pgz):
    """one .gz stream cut into sub-chunks that are found, decoded with window markers, chained and translated
    independently: the text is zlib's, for every level, block type and sub-chunk size (cuts inside blocks, blocks larger
    than a sub-chunk, streams smaller than one)"""
    import gzip
    import zlib
    texts = _inflate_texts()
    r = random.Random(5)
    texts["fastq_big"] = b"".join(b"@read%d/1\n%s\n+\n%s\n" % (i, bytes(r.choice(b"ACGTN") for _ in range(150)),
                                                                  bytes(r.choice(b"FFFFF:,#") for _ in range(150))) for i in range(9000))
    n_multi = 0
    for name, t in texts.items():
        for lvl in (0, 1, 6, 9):
            z = gzip.compress(t, lvl)
            for sub in ((4096, 16384, 65536, 1 << 20) if lvl == 6 else (16384,)):
                rc, o, st = pgz(z, sub)
                if lvl == 0:                                      # stored blocks only: invisible to the block finder, so the first sub-chunk
                    assert (rc == 0 and o == t) or rc in (-5, -101), (name, sub, rc)       # decodes it all, or its symbol area overflows
                    continue
                assert rc == 0 and o == t, (name, lvl, sub, rc)
                n_multi += st[1] > 1
        for strat in (zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE):
            co = zlib.compressobj(6, zlib.DEFLATED, 31, 8, strat)
            z = co.compress(t) + co.flush()
            rc, o, st = pgz(z, 16384)
            # fixed-code and stored blocks are invisible to the block finder: such a stream either decodes (one sub-chunk
            # runs through them) or breaks the chain (-101: handed to the host reader) - never a wrong text
            assert (rc == 0 and o == t) or rc in (-5, -101), (name, strat, rc)
    assert n_multi >= 4                                           # streams really were decoded in several pieces


def test_parallel_gunzip_rejects_what_it_cannot_vouch_for(pgz):
    """truncations and random bit flips end in an error or in the text zlib also produces - the chain check, ISIZE and
    the bounds checks; the CRC-32 pass (device only) closes the rest"""
    import gzip
    import zlib
    r = random.Random(3)
    t = _inflate_texts()["fastq"]
    z = gzip.compress(t, 6)
    for cut in range(20, len(z), 211):
        assert pgz(z[:cut], 8192)[0] != 0, cut
    assert pgz(z + z, 8192)[0] == -104                            # a second member: not ours
    for _ in range(300):
        zz = bytearray(z)
        for _ in range(r.randint(1, 3)):
            zz[r.randrange(10, len(zz))] ^= 1 << r.randrange(8)
        rc, o, _ = pgz(bytes(zz), 8192)
        if rc == 0:
            try:
                assert o == zlib.decompress(bytes(zz), 47)
            except zlib.error:
                assert len(o) == len(t)                           # only the CRC-32 differs: the device pass compares it
