#!/usr/bin/env python
"""Generate the golden fixtures of the kmer_scrub_filter step by running the UNMODIFIED reference
script /root/reference/scripts/kmer_scrub_filter.py on small count tables (the format kmer_scrub_count prints,
/root/reference/src/kmer_scrub_count.c:134-156, gzip-compressed as test/example.sh:4 does).

Run from the repo root in the dev container:   python tests/golden/make_golden_filter.py
Tables, argument lists and the script's stdout / stderr / exit code are committed under
tests/golden/cases/filter/, so the same vectors check the oracle restatement and the CUDA path on machines where
/root/reference does not exist."""
import gzip
import json
import os
import random
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "cases", "filter")
SCRIPT = "/root/reference/scripts/kmer_scrub_filter.py"
HEADER = "#kmer\treference_count\tpangenome_count\tmetagenome_count\tdrug_count\n"


def kmers(r, n):
    seen, out = set(), []
    while len(out) < n:
        k = "".join(r.choice("ACGT") for _ in range(31))
        if k not in seen:
            seen.add(k)
            out.append(k)
    return out


def table(r, keys, drug=False, pan_rate=0.3, meta_rate=0.2, hi=6, ties=True):
    """rows with plenty of equal values (ties decide who is scrubbed first: row order) and a few heavy hitters"""
    lines = [HEADER]
    for k in keys:
        rk = random.Random(k)                                  # the reference count belongs to the key: the same in every table of a strain
        ref = 1 if rk.random() < 0.97 else rk.randint(2, 5)
        pan = r.randint(1, hi) if r.random() < pan_rate else 0
        meta = r.randint(1, hi * 3) if r.random() < meta_rate else 0
        if not ties:
            pan *= r.randint(1, 1000)
            meta *= r.randint(1, 1000)
        if r.random() < 0.01:
            meta += r.randint(100, 5000)
        row = [k, ref, pan, meta]
        if drug:
            row.append(r.randint(1, 3) if r.random() < 0.15 else 0)
        lines.append("\t".join(str(x) for x in row) + "\n")
    return "".join(lines)


def write_gz(name, text):
    with gzip.GzipFile(os.path.join(OUT, name), "wb", mtime=0) as f:
        f.write(text.encode())


def main():
    os.makedirs(OUT, exist_ok=True)
    r = random.Random(20261018)
    keys = kmers(r, 3000)
    write_gz("t_plain.tsv.gz", table(r, keys))
    write_gz("t_noties.tsv.gz", table(r, keys, ties=False))
    write_gz("t_drug.tsv.gz", table(r, keys, drug=True))
    write_gz("t_drug_heavy.tsv.gz", table(r, keys[:400], drug=True).replace("\t0\n", "\t2\n"))      # drug scrub leaves too little
    write_gz("t_empty_counts.tsv.gz", HEADER + "".join(k + "\t1\t0\t0\n" for k in keys[:500]))
    write_gz("t_only_header.tsv.gz", HEADER)
    write_gz("t_tiny.tsv.gz", HEADER + "".join(k + "\t1\t%d\t%d\n" % (i % 3, (i * 7) % 5) for i, k in enumerate(keys[:7])))
    # the same strain counted against three different list sets: identical key columns, counts add up
    for j in range(3):
        write_gz("t_multi%d.tsv.gz" % j, table(r, keys[:1500], pan_rate=0.2 + 0.1 * j))
    open(os.path.join(OUT, "list3.txt"), "w").write("".join("t_multi%d.tsv.gz\n" % j for j in range(3)))
    open(os.path.join(OUT, "list2.txt"), "w").write("t_multi0.tsv.gz\nt_multi1.tsv.gz\n")
    # third file with another key set: the script's consistency check fires (only from the third file on)
    write_gz("t_other.tsv.gz", table(r, kmers(r, 1500)))
    open(os.path.join(OUT, "list_bad.txt"), "w").write("t_multi0.tsv.gz\nt_multi1.tsv.gz\nt_other.tsv.gz\n")
    # second file with another key order / subset: allowed by the script, hashes are merged by key
    sub = keys[:1500][::-1][:900]
    write_gz("t_reordered.tsv.gz", table(r, sub))
    open(os.path.join(OUT, "list_reordered.txt"), "w").write("t_multi0.tsv.gz\nt_reordered.tsv.gz\n")
    # duplicate keys inside one table (dict semantics: first position, last reference count, counts add up)
    dup = table(r, keys[:300] + keys[100:200] + keys[:50])
    write_gz("t_dups.tsv.gz", dup)
    # keys that are not plain ACGT 31-mers (IUPAC rows of SURVEY D6, odd lengths): still just dict keys
    odd = [k[:10] + "R" + k[11:] for k in keys[:40]] + [k[:20] for k in keys[40:60]] + keys[60:400]
    write_gz("t_odd_keys.tsv.gz", table(r, odd))
    # DOS and old-Mac line ends: the script reads in text mode (universal newlines)
    small = table(r, keys[:600])
    write_gz("t_crlf.tsv.gz", small.replace("\n", "\r\n"))
    write_gz("t_cr.tsv.gz", small.replace("\n", "\r"))

    cases = {
        "plain_default": ["-s", "t_plain.tsv.gz"],
        "plain_m01": ["-s", "t_plain.tsv.gz", "-m", "0.01"],
        "plain_m0": ["-s", "t_plain.tsv.gz", "-m", "0.0"],
        "plain_m1": ["-s", "t_plain.tsv.gz", "-m", "1.0"],
        "plain_m05_long": ["--scrub_count_file", "t_plain.tsv.gz", "--min_fraction", "0.5"],
        "plain_m0333": ["-s", "t_plain.tsv.gz", "-m", "0.3333"],
        "plain_independent": ["-s", "t_plain.tsv.gz", "-i"],
        "plain_independent_m30": ["-s", "t_plain.tsv.gz", "-i", "-m", "0.3"],
        "plain_independent_m90": ["-s", "t_plain.tsv.gz", "--independent", "-m", "0.9"],
        "noties_m02": ["-s", "t_noties.tsv.gz", "-m", "0.2"],
        "noties_independent": ["-s", "t_noties.tsv.gz", "-m", "0.5", "-i"],
        "drug_default": ["-s", "t_drug.tsv.gz"],
        "drug_m02": ["-s", "t_drug.tsv.gz", "-m", "0.2"],
        "drug_independent": ["-s", "t_drug.tsv.gz", "-m", "0.1", "-i"],
        "drug_too_few": ["-s", "t_drug_heavy.tsv.gz", "-m", "0.3"],
        "empty_counts": ["-s", "t_empty_counts.tsv.gz", "-m", "0.1"],
        "empty_counts_independent": ["-s", "t_empty_counts.tsv.gz", "-m", "0.1", "-i"],
        "only_header": ["-s", "t_only_header.tsv.gz"],
        "tiny": ["-s", "t_tiny.tsv.gz", "-m", "0.3"],
        "tiny_independent": ["-s", "t_tiny.tsv.gz", "-m", "0.3", "-i"],
        "multi3": ["-l", "list3.txt", "-m", "0.05"],
        "multi2_independent": ["--scrub_count_list", "list2.txt", "-m", "0.2", "-i"],
        "multi_bad": ["-l", "list_bad.txt"],
        "multi_reordered": ["-l", "list_reordered.txt", "-m", "0.1"],
        "dups": ["-s", "t_dups.tsv.gz", "-m", "0.1"],
        "odd_keys": ["-s", "t_odd_keys.tsv.gz", "-m", "0.1"],
        "crlf": ["-s", "t_crlf.tsv.gz", "-m", "0.1"],
        "cr_only": ["-s", "t_cr.tsv.gz", "-m", "0.1"],
        "crlf_independent": ["-s", "t_crlf.tsv.gz", "-m", "0.5", "-i"],
        "no_input": [],
        "both_inputs": ["-s", "t_tiny.tsv.gz", "-l", "list2.txt", "-m", "0.3"],
    }
    index = {}
    for name, argv in cases.items():
        p = subprocess.run([sys.executable, "-W", "ignore", SCRIPT] + argv, cwd=OUT, capture_output=True)
        open(os.path.join(OUT, "expected_%s.stdout" % name), "wb").write(p.stdout)
        err = p.stderr
        if p.returncode and b"Traceback" in err:            # an uncaught exception: only its last line is the script's own words
            err = b"<traceback>\n" + err.strip().split(b"\n")[-1] + b"\n"
        open(os.path.join(OUT, "expected_%s.stderr" % name), "wb").write(err)
        index[name] = {"argv": argv, "rc": p.returncode}
        print(name, p.returncode, len(p.stdout), len(p.stderr))
    json.dump(index, open(os.path.join(OUT, "cases.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
