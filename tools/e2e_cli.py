#!/usr/bin/env python
"""End-to-end run of the drop-in executables on files (GPU box): generates down-scaled config #2 / #3
inputs in a scratch directory, runs strainer2_b200/bin/kmer_scrub_count with S2_STATS=1, and (optionally)
the compiled reference on a subset, then compares the two tables byte for byte on that subset.
Usage: python tools/e2e_cli.py [--genomes 200] [--gz-genomes 40] [--metas 4] [--reads 2000000] [--ref-genomes 4]"""
import argparse
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _gen_genome(a):
    import bench
    from strainer2_b200 import synth
    path, index, gz = a
    strain = bench.make_strain()
    if path.endswith(".bgz"):
        import io
        buf = io.BytesIO()
        for i, c in enumerate(bench.make_genome(strain, index)):
            b = c.tobytes()
            buf.write(b">seq%d\n" % i + b"\n".join(b[j:j + 80] for j in range(0, len(b), 80)) + b"\n")
        synth.write_bgzf(path, buf.getvalue())
    else:
        synth.write_fasta(path, bench.make_genome(strain, index), gz=gz)
    return path


def _gen_meta(a):
    import numpy as np
    import bench
    from strainer2_b200 import synth
    path, index, n_reads = a
    strain = bench.make_strain()
    rng = synth.rng_for(3, index)
    clean = [np.where(c == ord("N"), ord("A"), c).astype(np.uint8) for c in strain]
    others = [synth.random_bases(rng, 5_000_000) for _ in range(6)]
    r1 = synth.sample_reads(rng, clean, n_reads // 100, 150, sub_rate=0.005, n_rate=1e-5)
    r2 = synth.sample_reads(rng, others, n_reads - n_reads // 100, 150, sub_rate=0.005, n_rate=1e-5)
    reads = np.concatenate([r1, r2])
    rng.shuffle(reads)
    if path.endswith(".bgz"):
        synth.write_bgzf(path, synth.fastq_bytes(reads))             # block gzip: the GPU ingest path
    else:
        synth.write_reads_fastq(path, reads)
    return path


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genomes", type=int, default=200)
    ap.add_argument("--gz-genomes", type=int, default=40)
    ap.add_argument("--metas", type=int, default=4)
    ap.add_argument("--reads", type=int, default=2_000_000)
    ap.add_argument("--ref-genomes", type=int, default=4)
    ap.add_argument("--threads", default="")
    args = ap.parse_args()
    import bench
    from strainer2_b200 import synth
    tmp = tempfile.mkdtemp(prefix="s2e2e_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    t0 = time.time()
    strain = bench.make_strain()
    synth.write_fasta(os.path.join(tmp, "strain.fa"), strain, gz=False)
    jobs = [(os.path.join(tmp, f"g{i}.fa" + (".gz" if i < args.gz_genomes else "")), i, i < args.gz_genomes) for i in range(args.genomes)]
    jobs += [(os.path.join(tmp, f"g{i}.fa.bgz"), i, False) for i in range(args.genomes)]
    jobs += [(os.path.join(tmp, f"z{i}.fa.gz"), i, True) for i in range(args.genomes)]
    mjobs = [(os.path.join(tmp, f"m{i}.fastq.gz"), i, args.reads) for i in range(args.metas)]
    mjobs += [(os.path.join(tmp, f"m{i}.fastq.bgz"), i, args.reads) for i in range(args.metas)]
    with mp.Pool(min(24, os.cpu_count() or 1)) as pool:
        A = pool.map(_gen_genome, jobs)
        B = pool.map(_gen_meta, mjobs)
    Az = [a for a in A if a.endswith(".bgz")]
    Ag = [a for a in A if os.path.basename(a).startswith("z")]
    A = [a for a in A if a not in Az and a not in Ag]
    open(os.path.join(tmp, "A.txt"), "w").write("".join(a + "\n" for a in A))
    open(os.path.join(tmp, "Az.txt"), "w").write("".join(a + "\n" for a in Az))
    open(os.path.join(tmp, "Ag.txt"), "w").write("".join(a + "\n" for a in Ag))
    Bz = [b for b in B if b.endswith(".bgz")]
    B = [b for b in B if not b.endswith(".bgz")]
    open(os.path.join(tmp, "B.txt"), "w").write("".join(b + "\n" for b in B))
    open(os.path.join(tmp, "Bz.txt"), "w").write("".join(b + "\n" for b in Bz))
    open(os.path.join(tmp, "empty.txt"), "w").write("")
    print(f"# generated {len(A)} genomes ({args.gz_genomes} gz), {len(B)} metagenomes x {args.reads} reads in {time.time() - t0:.1f}s "
          f"under {tmp}", flush=True)
    exe = os.path.join(ROOT, "strainer2_b200", "bin", "kmer_scrub_count")
    env = dict(os.environ, S2_STATS="1")
    runs = [("genomes_only", ["-r", "strain.fa", "-A", "A.txt", "-B", "empty.txt"]),
            ("genomes_all_gz_host_inflate", ["-r", "strain.fa", "-A", "Ag.txt", "-B", "empty.txt"]),
            ("genomes_all_bgzf_gpu_ingest", ["-r", "strain.fa", "-A", "Az.txt", "-B", "empty.txt"]),
            ("metagenomes_only", ["-r", "strain.fa", "-A", "empty.txt", "-B", "B.txt"]),
            ("metagenomes_bgzf_gpu_ingest", ["-r", "strain.fa", "-A", "empty.txt", "-B", "Bz.txt"])]
    for th in ([t for t in args.threads.split(",") if t] or [""]):
        for name, a in runs:
            e = dict(env)
            if th:
                e["S2_THREADS"] = th
            t1 = time.time()
            p = subprocess.run([exe] + a, cwd=tmp, env=e, stdout=open(os.path.join(tmp, name + ".tsv"), "wb"), stderr=subprocess.PIPE)
            print(f"{name} threads={th or 'default'} rc={p.returncode} wall={time.time() - t1:.2f}s {p.stderr.decode().strip()}", flush=True)
    a = subprocess.run(["cmp", os.path.join(tmp, "metagenomes_only.tsv"), os.path.join(tmp, "metagenomes_bgzf_gpu_ingest.tsv")])
    print("tables from .gz (host inflate) and .bgz (GPU ingest) identical:", a.returncode == 0, flush=True)
    # strain_detect on the same metagenomes: informative = every 100th k-mer of the strain's first contig
    c0 = bytes(strain[0]).replace(b"N", b"A")
    with open(os.path.join(tmp, "inf.txt"), "wb") as f:
        for i in range(0, len(c0) - 31, 100):
            f.write(c0[i:i + 31] + b"\n")
    open(os.path.join(tmp, "batch.txt"), "w").write("".join(f"SE\t{b}\n" for b in B[:2]) + (f"PE\t{B[2]}\t{B[3]}\n" if len(B) >= 4 else ""))
    dexe = os.path.join(ROOT, "strainer2_b200", "bin", "strain_detect")
    t1 = time.time()
    p = subprocess.run([dexe, "-r", "strain.fa", "-a", "inf.txt", "-B", "batch.txt", "-o", "hits.gz"], cwd=tmp, env=env, capture_output=True)
    print(f"strain_detect rc={p.returncode} wall={time.time() - t1:.2f}s {p.stderr.decode().strip()[-400:]}", flush=True)
    if len(Bz) >= 4:
        open(os.path.join(tmp, "batchz.txt"), "w").write("".join(f"SE\t{b}\n" for b in Bz[:2]) + f"PE\t{Bz[2]}\t{Bz[3]}\n")
        t1 = time.time()
        p = subprocess.run([dexe, "-r", "strain.fa", "-a", "inf.txt", "-B", "batchz.txt", "-o", "hitsz.gz"], cwd=tmp, env=env, capture_output=True)
        print(f"strain_detect_bgzf_gpu_ingest rc={p.returncode} wall={time.time() - t1:.2f}s {p.stderr.decode().strip()[-400:]}", flush=True)
        import gzip
        a = gzip.open(os.path.join(tmp, "hits.gz")).read().replace(b".fastq.gz", b".fastq.bgz")
        print("kmer_hits from .gz and .bgz identical (file names aside):", a == gzip.open(os.path.join(tmp, "hitsz.gz")).read(), flush=True)
    ref = os.path.join(ROOT, "oracle", "_ref", "kmer_scrub_count")
    if args.ref_genomes and os.path.exists(ref):
        open(os.path.join(tmp, "A_small.txt"), "w").write("".join(a + "\n" for a in A[:args.ref_genomes]))
        a = ["-r", "strain.fa", "-A", "A_small.txt", "-B", "empty.txt"]
        t1 = time.time()
        r = subprocess.run([ref] + a, cwd=tmp, capture_output=True)
        t_ref = time.time() - t1
        t1 = time.time()
        o = subprocess.run([exe] + a, cwd=tmp, env=env, capture_output=True)
        print(f"subset of {args.ref_genomes} genomes: reference {t_ref:.2f}s, ours {time.time() - t1:.2f}s, tables identical: {r.stdout == o.stdout} "
              f"({len(r.stdout)} bytes)", flush=True)
    subprocess.run(["rm", "-rf", tmp])


if __name__ == "__main__":
    main()
