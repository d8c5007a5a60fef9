#!/usr/bin/env python
"""The reference's own example (test/example.sh, BASELINE config #1) through the three drop-in executables, checked
against the digests of the unmodified reference (the scrubbed-k-mer digest was taken in the dev container
with the unmodified scripts/kmer_scrub_filter.py).  The reference's test data is NOT part of this repository: point
--test-dir at a copy of its test/ directory.
Usage: python tools/example_check.py --test-dir /path/to/strainer2/test"""
import argparse
import gzip
import hashlib
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STRAIN = "strains/Bacteroides_ovatus_1001283st1_B8_1001283B150210_160208"
WANT = {
    "count table md5": "75989a9bc31ef0b6f53a5112a60920bd",          # 6,698,541 lines
    "scrubbed k-mers md5 (-m 0.01)": "fe981fa571be70e602875ac3463ecdac",   # 2 header lines + 66,986 k-mers
    "kmer_hits text md5": "e1799e705d4f693240573da32540efcc",        # 1,130 lines
    "kmer_hits.gz md5": "997c3e1b8c1272a736168909c6be359b",          # same zlib, level 9, same byte stream
}


def md5(b):
    return hashlib.md5(b).hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--test-dir", required=True)
    args = ap.parse_args()
    d = os.path.abspath(args.test_dir)
    bin_dir = os.path.join(ROOT, "strainer2_b200", "bin")
    tmp = tempfile.mkdtemp(prefix="s2example_")
    env = dict(os.environ, S2_STATS="1")
    got = {}
    t = time.time()
    p = subprocess.run([os.path.join(bin_dir, "kmer_scrub_count"), "-r", STRAIN + ".fna.gz", "-A", "genomes_to_scrub.txt", "-B", "metagenomes_to_scrub.txt",
                        "-p", os.path.join(tmp, "progress")], cwd=d, env=env, capture_output=True)
    print(f"STEP1 kmer_scrub_count rc={p.returncode} {time.time() - t:.2f}s {p.stderr.decode().strip()[-250:]}", flush=True)
    got["count table md5"] = md5(p.stdout)
    with gzip.GzipFile(os.path.join(tmp, "counts.gz"), "wb", compresslevel=6) as f:
        f.write(p.stdout)
    t = time.time()
    q = subprocess.run([os.path.join(bin_dir, "kmer_scrub_filter"), "-s", os.path.join(tmp, "counts.gz"), "-m", "0.01"], cwd=d, env=env, capture_output=True)
    print(f"STEP2 kmer_scrub_filter rc={q.returncode} {time.time() - t:.2f}s {q.stderr.decode().strip()[-250:]}", flush=True)
    got["scrubbed k-mers md5 (-m 0.01)"] = md5(q.stdout)
    with gzip.GzipFile(os.path.join(tmp, "scrubbed.gz"), "wb", compresslevel=9) as f:
        f.write(q.stdout)
    t = time.time()
    r = subprocess.run([os.path.join(bin_dir, "strain_detect"), "-r", STRAIN + ".fna.gz", "-a", os.path.join(tmp, "scrubbed.gz"), "-B", "target_metagenomes.txt",
                        "-o", os.path.join(tmp, "kmer_hits.gz")], cwd=d, env=env, capture_output=True)
    print(f"STEP3 strain_detect rc={r.returncode} {time.time() - t:.2f}s {r.stderr.decode().strip()[-250:]}", flush=True)
    raw = open(os.path.join(tmp, "kmer_hits.gz"), "rb").read()
    got["kmer_hits text md5"] = md5(gzip.decompress(raw))
    got["kmer_hits.gz md5"] = md5(raw)
    ok = True
    for k, v in WANT.items():
        print(f"{k}: {got[k]} {'== reference' if got[k] == v else '!= reference ' + v}", flush=True)
        ok = ok and got[k] == v
    print("example.sh steps 1-3 byte-identical to the reference:", ok, flush=True)
    subprocess.run(["rm", "-rf", tmp])
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
