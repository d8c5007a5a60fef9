#!/usr/bin/env python
"""BASELINE config #5 (down-scaled): kmer_scrub_count_batch - many strain tables in ONE pass over the inputs through a
union table (two-phase scan once the fingerprints outgrow L2) - on BGZF metagenomes (GPU ingest) and on the same reads as
ordinary .gz (host inflate).  Prints the executables' S2_STATS lines.
Usage: python tools/batch_bench.py [--strains 16] [--metas 4] [--reads 2000000]"""
import argparse
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _strain(a):
    from strainer2_b200 import synth
    path, i = a
    synth.write_fasta(path, synth.genome(synth.rng_for(5, i), 5_000_000, 40, n_runs=5), gz=False)
    return path


def _meta(a):
    from strainer2_b200 import synth
    path, i, n_reads, strain_paths = a
    import strainer2_b200 as s2   # noqa: F401  (library must be built)
    rng = synth.rng_for(3, 100 + i)
    src = [synth.random_bases(rng, 5_000_000) for _ in range(4)]
    # 2 % of the reads come from the first two strains
    strains = [synth.genome(synth.rng_for(5, k), 5_000_000, 40, n_runs=5) for k in range(2)]
    clean = [np.where(c == ord("N"), ord("A"), c).astype(np.uint8) for g in strains for c in g]
    r1 = synth.sample_reads(rng, clean, n_reads // 50, 150, sub_rate=0.005, n_rate=1e-5)
    r2 = synth.sample_reads(rng, src, n_reads - n_reads // 50, 150, sub_rate=0.005, n_rate=1e-5)
    reads = np.concatenate([r1, r2])
    rng.shuffle(reads)
    synth.write_bgzf(path + ".bgz", synth.fastq_bytes(reads))
    synth.write_reads_fastq(path + ".gz", reads)
    return path


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--strains", type=int, default=16)
    ap.add_argument("--metas", type=int, default=4)
    ap.add_argument("--reads", type=int, default=2_000_000)
    args = ap.parse_args()
    tmp = tempfile.mkdtemp(prefix="s2batch_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    t0 = time.time()
    with mp.Pool(min(16, os.cpu_count() or 1)) as pool:
        strains = pool.map(_strain, [(os.path.join(tmp, "strain%d.fa" % i), i) for i in range(args.strains)])
        metas = pool.map(_meta, [(os.path.join(tmp, "m%d.fastq" % i), i, args.reads, None) for i in range(args.metas)])
    open(os.path.join(tmp, "R.txt"), "w").write("".join(s + "\n" for s in strains))
    open(os.path.join(tmp, "A.txt"), "w").write("")
    open(os.path.join(tmp, "Bz.txt"), "w").write("".join(m + ".bgz\n" for m in metas))
    open(os.path.join(tmp, "Bg.txt"), "w").write("".join(m + ".gz\n" for m in metas))
    print(f"# {args.strains} strains x 5 Mb, {args.metas} metagenomes x {args.reads} reads generated in {time.time() - t0:.0f}s", flush=True)
    exe = os.path.join(ROOT, "strainer2_b200", "bin", "kmer_scrub_count_batch")
    outs = {}
    for name, lst in (("bgzf_gpu_ingest", "Bz.txt"), ("gz_host_inflate", "Bg.txt")):
        out = os.path.join(tmp, "out_" + name)
        os.makedirs(out)
        t = time.time()
        p = subprocess.run([exe, "-R", "R.txt", "-A", "A.txt", "-B", lst, "-O", out], cwd=tmp, env=dict(os.environ, S2_STATS="1"), capture_output=True)
        print(f"{name}: rc={p.returncode} wall={time.time() - t:.2f}s {p.stderr.decode().strip()[-420:]}", flush=True)
        outs[name] = [open(os.path.join(out, f), "rb").read() for f in sorted(os.listdir(out))]
    print("per-strain tables identical between the two runs:", outs["bgzf_gpu_ingest"] == outs["gz_host_inflate"], len(outs["gz_host_inflate"]), "tables", flush=True)
    subprocess.run(["rm", "-rf", tmp])


if __name__ == "__main__":
    main()
