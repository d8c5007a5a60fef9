#!/usr/bin/env python
"""S2_GPUS in the executables on a multi-GPU box: kmer_scrub_count (files sharded over table replicas + one all-reduce
per counter column) and strain_detect (batch lines sharded over labelled replicas, no collective) must print the same
bytes whatever the number of GPUs.  Usage (2+ GPUs): python tools/multigpu_cli_check.py [--gpus 2]"""
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
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--reads", type=int, default=300_000)
    args = ap.parse_args()
    import bench
    from strainer2_b200 import synth
    tmp = tempfile.mkdtemp(prefix="s2mg_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    strain = bench.make_strain()
    synth.write_fasta(os.path.join(tmp, "strain.fa"), strain, gz=False)
    rng = synth.rng_for(4, 0)
    clean = [np.where(c == ord("N"), ord("A"), c).astype(np.uint8) for c in strain]
    metas = []
    for i in range(6):
        reads = synth.sample_reads(rng, clean + [synth.random_bases(rng, 2_000_000)], args.reads, 150, sub_rate=0.005, n_rate=1e-5)
        name = "m%d.fastq.gz" % i
        if i % 2:
            synth.write_reads_fastq(os.path.join(tmp, name), reads)          # ordinary gzip: host inflate
        else:
            synth.write_bgzf(os.path.join(tmp, name), synth.fastq_bytes(reads))   # BGZF: GPU ingest
        metas.append(name)
    genomes = []
    for i in range(12):
        name = "g%d.fa.gz" % i
        synth.write_bgzf(os.path.join(tmp, name), synth.fasta_bytes(bench.make_genome(strain, i), 80))
        genomes.append(name)
    open(os.path.join(tmp, "A.txt"), "w").write("".join(g + "\n" for g in genomes))
    open(os.path.join(tmp, "B.txt"), "w").write("".join(m + "\n" for m in metas))
    c0 = bytes(strain[0]).replace(b"N", b"A")
    with open(os.path.join(tmp, "inf.txt"), "wb") as f:
        for i in range(0, len(c0) - 31, 100):
            f.write(c0[i:i + 31] + b"\n")
    open(os.path.join(tmp, "batch.txt"), "w").write("".join("SE\t%s\n" % m for m in metas[:4]) + "PE\t%s\t%s\n" % (metas[4], metas[5]))
    # the batch tool: three strains in one pass through a union table
    strains = ["strain.fa"]
    for i in (1, 2):
        synth.write_fasta(os.path.join(tmp, "strain%d.fa" % i), synth.genome(synth.rng_for(5, i), 2_000_000, 8, n_runs=2), gz=False)
        strains.append("strain%d.fa" % i)
    open(os.path.join(tmp, "R.txt"), "w").write("".join(s + "\n" for s in strains))
    bin_dir = os.path.join(ROOT, "strainer2_b200", "bin")
    results = {}
    for n in sorted({1, args.gpus}):
        env = dict(os.environ, S2_STATS="1", S2_GPUS=str(n), S2_THREADS="8")
        t = time.time()
        p = subprocess.run([os.path.join(bin_dir, "kmer_scrub_count"), "-r", "strain.fa", "-A", "A.txt", "-B", "B.txt"], cwd=tmp, env=env, capture_output=True)
        print(f"kmer_scrub_count S2_GPUS={n}: rc={p.returncode} wall={time.time() - t:.2f}s {p.stderr.decode().strip()[-300:]}", flush=True)
        t = time.time()
        q = subprocess.run([os.path.join(bin_dir, "strain_detect"), "-r", "strain.fa", "-a", "inf.txt", "-B", "batch.txt", "-o", "hits%d.gz" % n], cwd=tmp, env=env,
                           capture_output=True)
        print(f"strain_detect    S2_GPUS={n}: rc={q.returncode} wall={time.time() - t:.2f}s {q.stderr.decode().strip()[-300:]}", flush=True)
        t = time.time()
        out_dir = os.path.join(tmp, "batch%d" % n)
        os.makedirs(out_dir)
        b = subprocess.run([os.path.join(bin_dir, "kmer_scrub_count_batch"), "-R", "R.txt", "-A", "A.txt", "-B", "B.txt", "-C", "R.txt", "-O", out_dir], cwd=tmp,
                           env=env, capture_output=True)
        print(f"kmer_scrub_count_batch S2_GPUS={n}: rc={b.returncode} wall={time.time() - t:.2f}s {b.stderr.decode().strip()[-300:]}", flush=True)
        tables = tuple(open(os.path.join(out_dir, f), "rb").read() for f in sorted(os.listdir(out_dir)))
        results[n] = (p.returncode, p.stdout, q.returncode, q.stdout, gzip.open(os.path.join(tmp, "hits%d.gz" % n)).read(), b.returncode, tables)
    a, b = results[1], results[args.gpus]
    print("count tables identical:", a[:2] == b[:2], len(a[1]), "bytes; kmer_hits identical:", a[2:5] == b[2:5], len(a[4]), "bytes;",
          "batch tables identical:", a[5:] == b[5:], [len(x) for x in a[6]], "bytes", flush=True)
    subprocess.run(["rm", "-rf", tmp])
    return 0 if a == b else 1


if __name__ == "__main__":
    sys.exit(main())
