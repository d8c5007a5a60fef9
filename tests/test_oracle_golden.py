"""Pins the CPU oracle (oracle/s2_oracle.c): against the committed reference-generated golden vectors
(always), and against the compiled reference itself where oracle/_ref exists (dev container and any box
that received the prebuilt files)."""
import hashlib
import os
import subprocess

import pytest

import oracle_util as ou

COUNT_RUNS = {
    "ABC": ["-r", "ref.fa.gz", "-A", "listA.txt", "-B", "listB.txt", "-C", "listC.txt"],
    "AB": ["-r", "ref.fa.gz", "-A", "listA.txt", "-B", "listB.txt"],
    "A_only": ["-r", "ref.fa.gz", "-A", "listA.txt", "-B", "listB_empty.txt"],
}
DETECT_RUNS = {
    "bg_batch": ["-r", "ref.fa", "-a", "informative.txt.gz", "-g", "background.txt", "-B", "batch.txt"],
    "bg_single": ["-r", "ref.fa", "-a", "informative_plain.txt", "-g", "background.txt", "-b", "s1_R1.fastq.gz", "-c",
                  "s1_R2.fastq.gz", "-t", "PE"],
    "batch": ["-r", "ref.fa", "-a", "informative.txt.gz", "-B", "batch.txt"],
    "single_pe": ["-r", "ref.fa", "-a", "informative_plain.txt", "-b", "s1_R1.fastq.gz", "-c", "s1_R2.fastq.gz", "-t", "PE"],
    "single_se_default": ["-r", "ref.fa", "-a", "informative.txt.gz", "-b", "s3_single.fa.gz"],
    "single_pei": ["-r", "ref.fa", "-a", "informative.txt.gz", "-b", "s2_interleaved.fa", "-t", "PEI"],
}


@pytest.mark.parametrize("name", sorted(COUNT_RUNS))
def test_oracle_count_matches_golden(golden_dir, tmp_path, name):
    d = os.path.join(golden_dir, "count_edge")
    prog = str(tmp_path / "progress")
    p = ou.oracle_cli(["count"] + COUNT_RUNS[name] + ["-p", prog], cwd=d)
    assert p.returncode == 0, p.stderr
    assert p.stdout == open(os.path.join(d, f"expected_{name}.tsv"), "rb").read()
    assert p.stderr == open(os.path.join(d, f"expected_{name}.stderr"), "rb").read()
    exp = os.path.join(d, f"expected_{name}.progress")
    if os.path.exists(exp):
        assert ou.mask_progress(open(prog).read()) == open(exp).read()


@pytest.mark.parametrize("name", sorted(DETECT_RUNS))
def test_oracle_detect_matches_golden(golden_dir, tmp_path, name):
    d = os.path.join(golden_dir, "detect_edge")
    msg = str(tmp_path / "msg")
    p = ou.oracle_cli(["detect"] + DETECT_RUNS[name] + ["-m", msg], cwd=d)
    assert p.returncode == 0, p.stderr
    assert p.stdout == ou.gunzip(os.path.join(d, f"expected_{name}.hits.txt.gz"))
    assert open(msg, "rb").read() == open(os.path.join(d, f"expected_{name}.stdout"), "rb").read()


def test_oracle_detect_pe2_short_is_an_error(golden_dir, tmp_path):
    d = os.path.join(golden_dir, "detect_edge")
    p = ou.oracle_cli(["detect", "-r", "ref.fa", "-a", "informative.txt.gz", "-b", "s6_R1.fastq", "-c", "s6_R2.fastq",
                       "-t", "PE", "-m", str(tmp_path / "m")], cwd=d)
    assert p.returncode == 1
    assert p.stderr == open(os.path.join(d, "expected_pe2_short.stderr"), "rb").read()


def test_oracle_matches_golden_iupac_case(golden_dir, tmp_path):
    """IUPAC / foreign bytes are hashed as strings by the reference; so does the oracle"""
    d = os.path.join(golden_dir, "iupac")
    p = ou.oracle_cli(["count", "-r", "ref.fa", "-A", "listA.txt", "-B", "listB.txt", "-C", "listC.txt"], cwd=d)
    assert p.returncode == 0
    assert p.stdout == open(os.path.join(d, "expected_count.tsv"), "rb").read()
    assert p.stderr == open(os.path.join(d, "expected_count.stderr"), "rb").read()
    msg = str(tmp_path / "msg")
    p = ou.oracle_cli(["detect", "-r", "ref.fa", "-a", "informative.txt", "-B", "batch.txt", "-m", msg], cwd=d)
    assert p.returncode == 0
    assert p.stdout == ou.gunzip(os.path.join(d, "expected_detect.hits.txt.gz"))
    assert open(msg, "rb").read() == open(os.path.join(d, "expected_detect.stdout"), "rb").read()


def test_oracle_reader_matches_kseq_dumps(golden_dir):
    d = os.path.join(golden_dir, "count_edge")
    n = 0
    for f in sorted(os.listdir(os.path.join(d, "kseq"))):
        src = f[:-len(".dump.gz")]
        p = ou.oracle_cli(["kseq", src], cwd=d)
        assert p.stdout == ou.gunzip(os.path.join(d, "kseq", f)), src
        n += 1
    assert n >= 10


def test_known_answers():
    L = ou.lib()
    # SURVEY 8a row E / 8c
    assert L.s2o_djb2(b"T" * 31) == 3948423441
    assert L.s2o_djb2(b"TGAAATCCGCTGATACAAGCGTCATTTGCAT") % 8000000 == 1
    assert L.s2o_djb2(b"TGAAATCCGCTGATACAAGCGTCATTTGCAT") % 16000000 == 1
    assert ou.orient(b"A" * 31) == b"T" * 31
    assert ou.orient(b"ATGCAAATGACGCTTGTATCAGCGGATTTCA") == b"TGAAATCCGCTGATACAAGCGTCATTTGCAT"
    assert L.s2o_encode_2bit(b"ACTG", 4) == 0b00011011        # A0 C1 T2 G3, src/up2bit.c:14


# ---- against the compiled reference itself (dev container / prebuilt oracle/_ref) ------------------
def test_complement_table_matches_reference(ref_dir):
    import ctypes as C
    R = C.CDLL(os.path.join(ref_dir, "libref_prims.so"))
    table = (C.c_char * 255).in_dll(R, "COMPLEMENT")
    L = ou.lib()
    for c in range(255):
        ref = table[c][0]
        ref = ref - 256 if ref > 127 else ref
        assert L.s2o_complement(c) == ref, c


def test_up2bit_matches_reference(ref_dir):
    import ctypes as C
    import random
    R = C.CDLL(os.path.join(ref_dir, "libref_prims.so"))
    R.encode_DNA_2_bit.restype = C.c_uint64
    R.encode_DNA_2_bit.argtypes = [C.c_char_p, C.c_int]
    R.decode_DNA_2_bit.argtypes = [C.c_uint64, C.c_int, C.c_char_p]
    L = ou.lib()
    r = random.Random(5)
    for _ in range(200):
        n = r.randint(1, 32)
        s = "".join(r.choice("ACGTacgtN") for _ in range(n)).encode()
        v = R.encode_DNA_2_bit(s, n)
        assert v == L.s2o_encode_2bit(s, n)
        a, b = C.create_string_buffer(40), C.create_string_buffer(40)
        R.decode_DNA_2_bit(v, n, a)
        L.s2o_decode_2bit(v, n, b)
        assert a.value == b.value


@pytest.mark.slow
def test_oracle_matches_reference_on_config1(ref_dir, tmp_path):
    """config #1 = test/example.sh step 1 as shipped (6.7 M keys: exercises the 8M -> 16M doubling)."""
    t = "/root/reference/test"
    if not os.path.isdir(t):
        pytest.skip("/root/reference/test not present")
    args = ["-r", "strains/Bacteroides_ovatus_1001283st1_B8_1001283B150210_160208.fna.gz", "-A", "genomes_to_scrub.txt",
            "-B", "metagenomes_to_scrub.txt"]
    o = ou.oracle_cli(["count"] + args, cwd=t)
    assert o.returncode == 0
    assert hashlib.md5(o.stdout).hexdigest() == "75989a9bc31ef0b6f53a5112a60920bd"      # SURVEY 8c, re-measured with oracle/_ref
    assert o.stdout.count(b"\n") == 6698541
