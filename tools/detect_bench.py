#!/usr/bin/env python
"""strain_detect on files (GPU box): the same synthetic metagenomes as BGZF FASTQ, ordinary .gz FASTQ and ordinary .gz
two-line FASTA (the reference's own target format, test/target_metagenomes.txt), through the drop-in executable with
S2_STATS=1; prints the phase times, where the files went (GPU ingest / host reader) and whether the outputs agree.
Usage: python tools/detect_bench.py [--metas 4] [--reads 1000000] [--trace 1]"""
import argparse
import gzip
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _gen(a):
    import numpy as np
    import bench
    from strainer2_b200 import synth
    tmp, index, n_reads = a
    strain = bench.make_strain()
    rng = synth.rng_for(3, index)
    clean = [np.where(c == ord("N"), ord("A"), c).astype(np.uint8) for c in strain]
    others = [synth.random_bases(rng, 5_000_000) for _ in range(6)]
    r1 = synth.sample_reads(rng, clean, n_reads // 100, 150, sub_rate=0.005, n_rate=1e-5)
    r2 = synth.sample_reads(rng, others, n_reads - n_reads // 100, 150, sub_rate=0.005, n_rate=1e-5)
    reads = np.concatenate([r1, r2])
    rng.shuffle(reads)
    fq = synth.fastq_bytes(reads)
    synth.write_bgzf(os.path.join(tmp, "m%d.fastq.bgz" % index), fq)
    open(os.path.join(tmp, "m%d.fastq.gz" % index), "wb").write(gzip.compress(fq, 6))
    fa = b"".join(b">r%d 1\n%s\n" % (i, x.tobytes()) for i, x in enumerate(reads))
    open(os.path.join(tmp, "m%d.fasta.gz" % index), "wb").write(gzip.compress(fa, 6))
    return reads.size


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--metas", type=int, default=4)
    ap.add_argument("--reads", type=int, default=1_000_000)
    ap.add_argument("--trace", type=int, default=0)
    args = ap.parse_args()
    import bench
    from strainer2_b200 import synth
    tmp = tempfile.mkdtemp(prefix="s2det_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    t0 = time.time()
    strain = bench.make_strain()
    synth.write_fasta(os.path.join(tmp, "strain.fa"), strain, gz=False)
    with mp.Pool(min(args.metas, os.cpu_count() or 1)) as pool:
        bases = sum(pool.map(_gen, [(tmp, i, args.reads) for i in range(args.metas)]))
    c0 = bytes(strain[0]).replace(b"N", b"A")
    with open(os.path.join(tmp, "inf.txt"), "wb") as f:
        for i in range(0, len(c0) - 31, 100):
            f.write(c0[i:i + 31] + b"\n")
    print(f"# {args.metas} metagenomes x {args.reads} reads ({bases / 1e6:.0f} Mbases) generated in {time.time() - t0:.1f}s under {tmp}", flush=True)
    exe = os.path.join(ROOT, "strainer2_b200", "bin", "strain_detect")
    outs = {}
    for ext in ("fastq.bgz", "fastq.gz", "fasta.gz"):
        names = ["m%d.%s" % (i, ext) for i in range(args.metas)]
        lines = ["SE\t%s\n" % n for n in names[:max(1, args.metas - 2)]]
        if args.metas >= 2:
            lines.append("PE\t%s\t%s\n" % (names[-2], names[-1]))
        open(os.path.join(tmp, "batch.txt"), "w").write("".join(lines))
        for env_extra in ({}, {"S2_GPU_INGEST": "0"}):
            env = dict(os.environ, S2_STATS="1", **env_extra)
            if args.trace and not env_extra:
                env["S2_INGEST_TRACE"] = str(args.trace)
            t1 = time.time()
            p = subprocess.run([exe, "-r", "strain.fa", "-a", "inf.txt", "-B", "batch.txt", "-o", "hits.gz"], cwd=tmp, env=env, capture_output=True)
            wall = time.time() - t1
            err = p.stderr.decode(errors="replace").strip().splitlines()
            stat = [l for l in err if l.startswith("[s2 detect]")]
            print(f"{ext} {env_extra or 'defaults'} rc={p.returncode} wall={wall:.2f}s {stat[-1] if stat else err[-3:]}", flush=True)
            if args.trace and not env_extra:
                for l in err:
                    if "detect " in l and l.startswith("[s2 ingest]"):
                        print("   ", l, flush=True)
            text = gzip.open(os.path.join(tmp, "hits.gz")).read() if p.returncode == 0 else b""
            for n in names:
                text = text.replace(n.encode(), n.split(".")[0].encode())
            outs[(ext, bool(env_extra))] = text
    ref = outs[("fastq.gz", True)]
    print("outputs identical over formats and paths (file names aside):", all(v == ref for v in outs.values()), len(ref), "bytes", flush=True)
    subprocess.run(["rm", "-rf", tmp])


if __name__ == "__main__":
    main()
