#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference (oracle/_ref,
built from /root/reference/src by oracle/Makefile) on small, deliberately nasty inputs.

Run from the repo root in the dev container:   python tests/golden/make_golden.py
Inputs and expected outputs are committed so the same vectors check the oracle and the CUDA path on
machines where /root/reference does not exist.  Progress files are stored with the asctime() column
replaced by <T> (wall-clock).
"""
import gzip
import os
import random
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref")
COMP = {ord(a): ord(b) for a, b in zip("ACGTN", "TGCAN")}


def rnd(r, n):
    return "".join(r.choice("ACGT") for _ in range(n))


def revcomp(s):
    return s.translate(str.maketrans("ACGTNacgtn", "TGCANtgcan"))[::-1]


def mutate(r, s, rate):
    out = list(s)
    for i in range(len(out)):
        if r.random() < rate:
            out[i] = r.choice([c for c in "ACGT" if c != out[i].upper()])
    return "".join(out)


def fasta(records, wrap=60, nl="\n", blank_every=0):
    out = []
    for i, (name, s) in enumerate(records):
        out.append(">" + name + nl)
        if wrap:
            for j in range(0, len(s), wrap):
                out.append(s[j:j + wrap] + nl)
                if blank_every and (j // wrap) % blank_every == blank_every - 1:
                    out.append("\n")
        else:
            out.append(s + nl)
    return "".join(out)


def fastq(records, nl="\n", qual_char="I"):
    out = []
    for name, s in records:
        out.append("@" + name + nl + s + nl + "+" + nl + qual_char * len(s) + nl)
    return "".join(out)


def write(path, text, gz=False):
    data = text.encode("latin-1") if isinstance(text, str) else text
    if gz:
        with open(path, "wb") as raw:      # mtime=0 keeps the fixture bytes reproducible
            with gzip.GzipFile(fileobj=raw, mode="wb", mtime=0, filename="") as f:
                f.write(data)
    else:
        with open(path, "wb") as f:
            f.write(data)


def run(exe, args, cwd):
    p = subprocess.run([os.path.join(REF, exe)] + args, cwd=cwd, capture_output=True)
    return p.returncode, p.stdout, p.stderr


def mask_progress(path):
    lines = open(path).read().splitlines()
    return "\n".join([lines[0]] + [re.sub(r"\t.*$", "\t<T>", l) for l in lines[1:]]) + "\n"


def make_reference_genome(r):
    core = rnd(r, 2500)
    rep = core[300:420]                       # repeated region -> reference_count > 1
    contigs = [
        ("c1 first contig", core[:1200] + "NNNNNNNN" + core[1200:2000] + rep + "n" + core[2000:]),
        ("c2_lower", rnd(r, 900).lower()),
        ("c3_mixed", "".join(c.lower() if r.random() < 0.3 else c for c in rnd(r, 700))),
        ("c4_exact31", rnd(r, 31)),
        ("c5_len30", rnd(r, 30)),
        ("c6_N_mid", rnd(r, 22) + "N" + rnd(r, 22)),
        ("c7_revcomp_of_c1_part", revcomp(core[100:260])),
        ("c8_polyA", "A" * 40 + rnd(r, 35) + "T" * 40),
    ]
    return contigs


def reads_from(r, src, n, length, sub=0.0, n_rate=0.0, lower=0.0):
    out = []
    for i in range(n):
        st = r.randrange(0, len(src) - length + 1)
        s = src[st:st + length]
        if r.random() < 0.5:
            s = revcomp(s)
        s = mutate(r, s, sub) if sub else s
        if n_rate:
            s = "".join("N" if r.random() < n_rate else c for c in s)
        if lower and r.random() < lower:
            s = s.lower()
        out.append(("r%d/%d" % (i, 1), s))
    return out


def case_count(base):
    r = random.Random(20261018)
    d = os.path.join(base, "count_edge")
    os.makedirs(d, exist_ok=True)
    contigs = make_reference_genome(r)
    write(os.path.join(d, "ref.fa.gz"), fasta(contigs, wrap=60), gz=True)
    allseq = "".join(s for _, s in contigs if "N" not in s.upper()).upper()
    # -A genomes
    g1 = [(n, mutate(r, s.upper(), 0.02)) for n, s in contigs[:3]]
    write(os.path.join(d, "g1_relative_crlf.fa"), fasta(g1, wrap=70, nl="\r\n"))
    write(os.path.join(d, "g2_random.fa.gz"), fasta([("x1", rnd(r, 3000)), ("x2", rnd(r, 50))], wrap=80, blank_every=3), gz=True)
    write(os.path.join(d, "g3_identical_unwrapped.fa"), fasta([(n, s) for n, s in contigs], wrap=0))
    write(os.path.join(d, "g4_empty.fa"), "")
    write(os.path.join(d, "g5_header_only.fa"), ">only_a_header\n")
    write(os.path.join(d, "g6_gt_last_byte.fa"), ">a\n" + allseq[:80] + "\n>")
    write(os.path.join(d, "g7_no_trailing_newline.fa"), ">a desc\n" + allseq[50:200] + "\n" + allseq[200:260])
    write(os.path.join(d, "listA.txt"), "\n".join(["g1_relative_crlf.fa", "g2_random.fa.gz", "g3_identical_unwrapped.fa",
                                                   "g4_empty.fa", "g5_header_only.fa", "g6_gt_last_byte.fa",
                                                   "g7_no_trailing_newline.fa"]) + "\n")
    # -B metagenomes
    m1 = reads_from(r, allseq, 300, 150, sub=0.005, n_rate=0.002, lower=0.1) + [("rnd%d" % i, rnd(r, 150)) for i in range(100)]
    m1 += [("short1", allseq[10:40]), ("short2", allseq[10:30]), ("exact31", allseq[500:531]), ("allN", "N" * 60),
           ("lead_at", allseq[600:700])]
    r.shuffle(m1)
    fq = fastq(m1)
    # quality lines that begin with '@' or '>' must not confuse the parser
    fq += "@tricky\n" + allseq[700:760] + "\n+\n@" + "I" * 59 + "\n"
    fq += "@tricky2\n" + allseq[760:820] + "\n+tricky2 again\n>" + "#" * 59 + "\n"
    write(os.path.join(d, "m1_reads.fastq.gz"), fq, gz=True)
    m2 = reads_from(r, allseq, 120, 101, sub=0.01)
    write(os.path.join(d, "m2_multiline_reads.fa"), fasta(m2, wrap=50))
    m3 = fastq(reads_from(r, allseq, 20, 80)) + "@trunc\n" + allseq[0:70] + "\n+\n" + "I" * 30 + "\n@after_trunc\n" + allseq[100:170] + "\n+\n" + "I" * 70 + "\n"
    write(os.path.join(d, "m3_truncated_quality.fastq"), m3)
    m4 = fastq(reads_from(r, allseq, 10, 90), nl="\r\n")
    write(os.path.join(d, "m4_crlf.fastq"), m4)
    write(os.path.join(d, "listB.txt"), "\n".join(["m1_reads.fastq.gz", "m2_multiline_reads.fa", "m3_truncated_quality.fastq",
                                                   "m4_crlf.fastq"]) + "\n")
    write(os.path.join(d, "listB_empty.txt"), "")
    write(os.path.join(d, "listC.txt"), "ref.fa.gz\ng1_relative_crlf.fa\n")
    runs = {
        "ABC": ["-r", "ref.fa.gz", "-A", "listA.txt", "-B", "listB.txt", "-C", "listC.txt", "-p", "progress.tmp"],
        "AB": ["-r", "ref.fa.gz", "-A", "listA.txt", "-B", "listB.txt"],
        "A_only": ["-r", "ref.fa.gz", "-A", "listA.txt", "-B", "listB_empty.txt", "-p", "progress.tmp"],
    }
    for name, args in runs.items():
        rc, out, err = run("kmer_scrub_count", args, d)
        assert rc == 0, (name, rc, err)
        write(os.path.join(d, "expected_%s.tsv" % name), out)
        write(os.path.join(d, "expected_%s.stderr" % name), err)
        if "-p" in args:
            write(os.path.join(d, "expected_%s.progress" % name), mask_progress(os.path.join(d, "progress.tmp")))
            os.remove(os.path.join(d, "progress.tmp"))
    # error behaviours
    rc, out, err = run("kmer_scrub_count", ["-r", "ref.fa.gz", "-A", "listA.txt"], d)
    write(os.path.join(d, "expected_usage.stderr"), err)
    assert rc == 1 and out == b""
    write(os.path.join(d, "listA_missing.txt"), "g2_random.fa.gz\nno_such_file.fa\ng1_relative_crlf.fa\n")
    rc, out, err = run("kmer_scrub_count", ["-r", "ref.fa.gz", "-A", "listA_missing.txt", "-B", "listB_empty.txt"], d)
    assert rc == 1 and out == b"", (rc, out[:100])
    write(os.path.join(d, "expected_missing.stderr"), err)
    # kseq record dumps for every input file (pins our reader)
    dumps = os.path.join(d, "kseq")
    os.makedirs(dumps, exist_ok=True)
    for f in sorted(os.listdir(d)):
        if f.endswith((".fa", ".fa.gz", ".fastq", ".fastq.gz")):
            rc, out, err = run("kseq_dump", [f], d)
            write(os.path.join(dumps, f + ".dump.gz"), out, gz=True)
    return d


def case_detect(base):
    r = random.Random(77001)
    d = os.path.join(base, "detect_edge")
    os.makedirs(d, exist_ok=True)
    contigs = make_reference_genome(r)
    write(os.path.join(d, "ref.fa"), fasta(contigs, wrap=60))
    allseq = "".join(s for _, s in contigs[:3]).upper().replace("N", "")
    # informative k-mers: every 7th window of contig 2, in assorted spellings
    c2 = contigs[1][1].upper()
    inf = []
    for i in range(0, len(c2) - 31, 7):
        k = c2[i:i + 31]
        choice = (i // 7) % 4
        inf.append(k if choice == 0 else revcomp(k) if choice == 1 else k if choice == 2 else revcomp(k))
    lines = ["#comment line", "# another"] + inf
    lines.insert(5, "ACGT")                                   # wrong length -> stdout error
    lines.insert(9, rnd(r, 31))                               # not in the genome -> stdout error
    lines.insert(12, inf[3].lower())                          # lower case is NOT upper-cased by the reference
    lines.insert(15, rnd(r, 40))                              # too long
    write(os.path.join(d, "informative.txt.gz"), "\n".join(lines) + "\n", gz=True)
    write(os.path.join(d, "informative_plain.txt"), "\n".join(lines[:40]) + "\n")

    def pairs(n, length, with_short=True):
        p1, p2 = [], []
        for i in range(n):
            src = c2 if r.random() < 0.4 else allseq if r.random() < 0.5 else rnd(r, 600)
            st = r.randrange(0, len(src) - 2 * length)
            a = src[st:st + length]
            b = revcomp(src[st + length:st + 2 * length])
            if with_short and i % 9 == 4:
                a = a[:r.choice([5, 20, 30])]
            if with_short and i % 11 == 7:
                b = b[:r.choice([0, 12, 30])] if r.random() < 0.7 else b
            if i % 13 == 3:
                a = a[:40] + "N" + a[41:] if len(a) > 41 else a
            if i % 17 == 5:
                b = b.lower()
            p1.append(("p%d/1" % i, a))
            p2.append(("p%d/2" % i, b))
        return p1, p2

    p1, p2 = pairs(220, 100)
    write(os.path.join(d, "s1_R1.fastq.gz"), fastq(p1), gz=True)
    write(os.path.join(d, "s1_R2.fastq.gz"), fastq(p2), gz=True)
    q1, q2 = pairs(150, 75)
    inter = []
    for a, b in zip(q1, q2):
        inter += [a, b]
    write(os.path.join(d, "s2_interleaved.fa"), fasta(inter, wrap=0))
    se = [x for x in pairs(200, 90)[0]]
    write(os.path.join(d, "s3_single.fa.gz"), fasta(se, wrap=60), gz=True)
    # PE2 shorter than PE1, ending on a short stale read -> silently keeps going with stale state
    t1, t2 = pairs(40, 80, with_short=False)
    t2 = t2[:25] + [("stale_short", "ACGTACGTAC")]
    write(os.path.join(d, "s4_R1.fa"), fasta(t1, wrap=0))
    write(os.path.join(d, "s4_R2.fastq"), fastq(t2))
    # FASTA PE2 shorter than PE1: the parser resets seq.l to 0 at EOF (last_char != 0)
    u1, u2 = pairs(30, 80, with_short=False)
    write(os.path.join(d, "s5_R1.fa"), fasta(u1, wrap=0))
    write(os.path.join(d, "s5_R2.fa"), fasta(u2[:12], wrap=0))
    batch = ["PE\ts1_R1.fastq.gz\ts1_R2.fastq.gz", "#a comment\tx", "PEI\ts2_interleaved.fa", "SE\ts3_single.fa.gz",
             "pe\ts4_R1.fa\ts4_R2.fastq", "XX\tnope.fa", "ipe\ts2_interleaved.fa", "se\ts1_R1.fastq.gz", "PE\ts5_R1.fa\ts5_R2.fa",
             "PE\ts1_R1.fastq.gz", "SE"]
    write(os.path.join(d, "batch.txt"), "\n".join(batch) + "\n")
    runs = {
        "batch": ["-r", "ref.fa", "-a", "informative.txt.gz", "-B", "batch.txt", "-o", "out.tmp.gz"],
        "single_pe": ["-r", "ref.fa", "-a", "informative_plain.txt", "-b", "s1_R1.fastq.gz", "-c", "s1_R2.fastq.gz", "-t", "PE", "-o", "out.tmp.gz"],
        "single_se_default": ["-r", "ref.fa", "-a", "informative.txt.gz", "-b", "s3_single.fa.gz", "-o", "out.tmp.gz"],
        "single_pei": ["-r", "ref.fa", "-a", "informative.txt.gz", "-b", "s2_interleaved.fa", "-t", "PEI", "-o", "out.tmp.gz"],
    }
    for name, args in runs.items():
        rc, out, err = run("strain_detect", args, d)
        assert rc == 0, (name, rc, err)
        gzdata = open(os.path.join(d, "out.tmp.gz"), "rb").read()
        write(os.path.join(d, "expected_%s.hits.txt.gz" % name), gzip.decompress(gzdata), gz=True)
        write(os.path.join(d, "expected_%s.hits.gz.md5" % name), __import__("hashlib").md5(gzdata).hexdigest() + "\n")
        write(os.path.join(d, "expected_%s.stdout" % name), out)
        write(os.path.join(d, "expected_%s.stderr" % name), err)
        os.remove(os.path.join(d, "out.tmp.gz"))
    # -g background filter: background = the interleaved + single-end sets
    write(os.path.join(d, "background.txt"), "s2_interleaved.fa\ns3_single.fa.gz\n")
    for name, args in {"bg_batch": ["-r", "ref.fa", "-a", "informative.txt.gz", "-g", "background.txt", "-B", "batch.txt", "-o", "out.tmp.gz"],
                       "bg_single": ["-r", "ref.fa", "-a", "informative_plain.txt", "-g", "background.txt", "-b", "s1_R1.fastq.gz", "-c",
                                     "s1_R2.fastq.gz", "-t", "PE", "-o", "out.tmp.gz"]}.items():
        rc, out, err = run("strain_detect", args, d)
        assert rc == 0, (name, rc, err)
        gzdata = open(os.path.join(d, "out.tmp.gz"), "rb").read()
        write(os.path.join(d, "expected_%s.hits.txt.gz" % name), gzip.decompress(gzdata), gz=True)
        write(os.path.join(d, "expected_%s.stdout" % name), out)
        write(os.path.join(d, "expected_%s.stderr" % name), err)
        os.remove(os.path.join(d, "out.tmp.gz"))
    # PE2 runs out while its stale length is >= 31 -> error exit
    v1, v2 = pairs(20, 80, with_short=False)
    write(os.path.join(d, "s6_R1.fastq"), fastq(v1))
    write(os.path.join(d, "s6_R2.fastq"), fastq(v2[:8]))
    rc, out, err = run("strain_detect", ["-r", "ref.fa", "-a", "informative.txt.gz", "-b", "s6_R1.fastq", "-c", "s6_R2.fastq", "-t", "PE", "-o", "out.tmp.gz"], d)
    assert rc == 1, rc
    write(os.path.join(d, "expected_pe2_short.stderr"), err)
    if os.path.exists(os.path.join(d, "out.tmp.gz")):
        os.remove(os.path.join(d, "out.tmp.gz"))
    for name, args in {"usage_missing": ["-r", "ref.fa"], "usage_bad_type": ["-r", "ref.fa", "-a", "informative.txt.gz", "-b", "s3_single.fa.gz", "-t", "QQ", "-o", "o.gz"],
                       "usage_pe_needs_c": ["-r", "ref.fa", "-a", "informative.txt.gz", "-b", "s3_single.fa.gz", "-t", "PE", "-o", "o.gz"]}.items():
        rc, out, err = run("strain_detect", args, d)
        assert rc == 1
        write(os.path.join(d, "expected_%s.stdout" % name), out)
        write(os.path.join(d, "expected_%s.stderr" % name), err)
    return d


def case_iupac(base):
    """bytes other than ACGTN are kept, complemented through COMPLEMENT[] and hashed as strings"""
    r = random.Random(424242)
    d = os.path.join(base, "iupac")
    os.makedirs(d, exist_ok=True)
    g = list(rnd(r, 1500))
    for p, c in [(100, "R"), (101, "Y"), (400, "K"), (700, "M"), (701, "-"), (900, "r"), (1100, "E"), (1300, "S"), (1301, "N"), (1450, "W")]:
        g[p] = c
    g = "".join(g)
    contigs = [("x1", g), ("x2", rnd(r, 200) + "B" + rnd(r, 200)), ("x3_plain", rnd(r, 300))]
    write(os.path.join(d, "ref.fa"), fasta(contigs, wrap=60))
    same = [("y1", g[50:800]), ("y2", revcomp(g[850:1499]).replace("Y", "R")), ("y3", contigs[1][1].lower()), ("y4", g[1080:1130]), ("y5", rnd(r, 100) + "K" + rnd(r, 100))]
    write(os.path.join(d, "a1.fa"), fasta(same, wrap=70))
    reads = []
    for i in range(150):
        st = r.randrange(0, len(g) - 100)
        x = g[st:st + 100]
        reads.append(("q%d" % i, revcomp(x) if i % 3 == 0 else x))
    write(os.path.join(d, "b1.fastq"), fastq(reads))
    write(os.path.join(d, "listA.txt"), "a1.fa\n")
    write(os.path.join(d, "listB.txt"), "b1.fastq\n")
    write(os.path.join(d, "listC.txt"), "ref.fa\na1.fa\n")
    rc, out, err = run("kmer_scrub_count", ["-r", "ref.fa", "-A", "listA.txt", "-B", "listB.txt", "-C", "listC.txt"], d)
    assert rc == 0, err
    write(os.path.join(d, "expected_count.tsv"), out)
    write(os.path.join(d, "expected_count.stderr"), err)
    # informative list: a few plain and a few IUPAC-containing k-mers, some reverse-complemented
    rows = [l.split(b"\t")[0].decode("latin-1") for l in out.split(b"\n")[1:] if l]
    odd = [k for k in rows if any(c not in "ACGT" for c in k)]
    plain = [k for k in rows if all(c in "ACGT" for c in k)]
    assert len(odd) > 50
    inf = odd[::3] + plain[::40] + [revcomp(plain[5])]
    write(os.path.join(d, "informative.txt"), "\n".join(inf) + "\n")
    pe1 = [("p%d/1" % i, s) for i, (_, s) in enumerate(reads[:60])]
    pe2 = [("p%d/2" % i, revcomp(s) if "E" not in s else s) for i, (_, s) in enumerate(reads[60:120])]
    write(os.path.join(d, "r1.fastq"), fastq(pe1))
    write(os.path.join(d, "r2.fastq"), fastq(pe2))
    write(os.path.join(d, "batch.txt"), "PE\tr1.fastq\tr2.fastq\nSE\tb1.fastq\n")
    rc, out, err = run("strain_detect", ["-r", "ref.fa", "-a", "informative.txt", "-B", "batch.txt", "-o", "out.tmp.gz"], d)
    assert rc == 0, err
    gzdata = open(os.path.join(d, "out.tmp.gz"), "rb").read()
    write(os.path.join(d, "expected_detect.hits.txt.gz"), gzip.decompress(gzdata), gz=True)
    write(os.path.join(d, "expected_detect.stdout"), out)
    os.remove(os.path.join(d, "out.tmp.gz"))
    return d


if __name__ == "__main__":
    if not os.path.exists(os.path.join(REF, "kmer_scrub_count")):
        sys.exit("oracle/_ref is not built: run `make -C oracle` in the dev container first")
    base = os.path.join(HERE, "cases")
    os.makedirs(base, exist_ok=True)
    print(case_count(base))
    print(case_detect(base))
    print(case_iupac(base))
    total = sum(os.path.getsize(os.path.join(dp, f)) for dp, _, fs in os.walk(base) for f in fs)
    print("golden bytes:", total)
