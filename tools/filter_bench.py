#!/usr/bin/env python
"""kmer_scrub_filter at the size of a real strain table: a synthetic count table of --rows rows
(the shape of config #1's 6.7 M-row table) through strainer2_b200/bin/kmer_scrub_filter and through the CPU oracle
(the reference script's algorithm, oracle/scrub_filter_oracle.py - test infrastructure, used here as the checker and
as the timed CPU baseline); outputs are compared byte for byte.
Usage: python tools/filter_bench.py [--rows 5000000] [--min-fraction 0.04]"""
import argparse
import gzip
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=5_000_000)
    ap.add_argument("--min-fraction", default="0.04")
    ap.add_argument("--oracle", type=int, default=1)
    args = ap.parse_args()
    rng = np.random.default_rng(7)
    n = args.rows
    t0 = time.time()
    codes = rng.integers(0, 4, size=(n, 31), dtype=np.uint8)
    keys = np.frombuffer(b"ACGT", dtype=np.uint8)[codes]
    _, first = np.unique(keys.view("S31").ravel(), return_index=True)
    keys = keys[np.sort(first)]                                     # distinct, original order
    n = keys.shape[0]
    pan = (rng.random(n) < 0.2) * rng.integers(1, 40, n)
    meta = (rng.random(n) < 0.02) * rng.integers(1, 400, n)
    tmp = tempfile.mkdtemp(prefix="s2filter_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    path = os.path.join(tmp, "table.tsv.gz")
    lines = [b"#kmer\treference_count\tpangenome_count\tmetagenome_count\tdrug_count\n"]
    ks = keys.view("S31").ravel()
    step = 200_000
    with gzip.open(path, "wb", compresslevel=6) as f:
        f.write(lines[0])
        for s in range(0, n, step):
            f.write(b"".join(b"%s\t1\t%d\t%d\n" % (ks[i], pan[i], meta[i]) for i in range(s, min(n, s + step))))
    print(f"# table: {n} rows, {os.path.getsize(path) / 1e6:.0f} MB gzip, written in {time.time() - t0:.0f}s", flush=True)
    exe = os.path.join(ROOT, "strainer2_b200", "bin", "kmer_scrub_filter")
    for extra in ([], ["-i"]):
        argv = ["-s", "table.tsv.gz", "-m", args.min_fraction] + extra
        t1 = time.time()
        p = subprocess.run([exe] + argv, cwd=tmp, capture_output=True)
        t_ours = time.time() - t1
        q = subprocess.run([exe] + argv, cwd=tmp, capture_output=True, env=dict(os.environ, S2_STATS="1"))      # once more for the phase times
        print("  " + "".join(l for l in q.stderr.decode().splitlines(True) if l.startswith("[s2 filter]")).strip(), flush=True)
        line = f"kmer_scrub_filter {' '.join(argv)}: rc={p.returncode} {t_ours:.2f}s, {len(p.stdout)} bytes out"
        if args.oracle:
            from oracle import scrub_filter_oracle as fo
            t1 = time.time()
            rc, out, err = fo.run(argv, cwd=tmp)
            t_cpu = time.time() - t1
            line += f"; CPU restatement of the reference script {t_cpu:.1f}s; identical stdout: {out == p.stdout}, stderr: {err == p.stderr}"
        print(line, flush=True)
    subprocess.run(["rm", "-rf", tmp])


if __name__ == "__main__":
    main()
