#!/usr/bin/env python
"""Throughput of the GPU-side ingest (hardware DEFLATE + record splitting + count scan) on file IMAGES that sit in
pinned host memory: BGZF FASTA genomes (config #2 shape) and a BGZF FASTQ read file (config #3 shape).  Timed by
wall clock around synchronous s2_ingest_count_mem() calls (everything is inside: H2D of the compressed bytes,
inflate, indexing, validation, scan).  Also checks the counters against the host-parsed flat batches.
Usage: python tools/ingest_bench.py [--genomes 16] [--reads 400000] [--reps 3]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genomes", type=int, default=64)
    ap.add_argument("--distinct", type=int, default=4, help="distinct genome images (cycled)")
    ap.add_argument("--reads", type=int, default=400_000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--check", type=int, default=1)
    ap.add_argument("--bench-mix", type=int, default=0, help="also time one step of bench.py's file-image batch (64 genomes, 5 %% relatives)")
    args = ap.parse_args()
    import strainer2_b200 as s2
    from strainer2_b200 import synth
    import bench

    t0 = time.time()
    strain = bench.make_strain()
    ctx = s2.Context(0, batch_bytes=64 << 20, n_lanes=2)
    table = s2.StrainTable(ctx, synth.contigs_to_flat(strain), n_cols=4)
    rng = synth.rng_for(2, 99)
    images, flats, n_bases = [], [], []
    for i in range(args.distinct):
        contigs = [synth.mutate(c, 0.02, rng) for c in strain] if i == 0 else synth.genome(rng, 5_000_000, 40)
        text = synth.fasta_bytes(contigs, 80)
        z = synth.bgzf_bytes(text)
        pb = s2.PinnedBuffer(len(z))
        pb.array[:] = np.frombuffer(z, dtype=np.uint8)
        images.append((pb, len(z), len(text)))
        flats.append(synth.contigs_to_flat(contigs))
        n_bases.append(sum(c.size for c in contigs))
    clean = [np.where(c == ord("N"), ord("A"), c).astype(np.uint8) for c in strain]
    reads = synth.sample_reads(rng, clean + synth.genome(rng, 5_000_000, 4), args.reads, 150, sub_rate=0.005, n_rate=1e-5)
    rtext = synth.fastq_bytes(reads)
    rz = synth.bgzf_bytes_parallel(rtext)
    rpb = s2.PinnedBuffer(len(rz))
    rpb.array[:] = np.frombuffer(rz, dtype=np.uint8)
    print(f"# inputs ready in {time.time() - t0:.1f}s: genome image {images[0][1] / 1e6:.2f} MB for {images[0][2] / 1e6:.2f} MB of FASTA; "
          f"reads image {len(rz) / 1e6:.1f} MB for {len(rtext) / 1e6:.1f} MB of FASTQ", flush=True)

    if args.check:
        for i, (pb, n, _) in enumerate(images):
            table.clear_counts(1); table.clear_counts(2)
            want = ctx.scan_count(table, flats[i], 1)
            rc, b, l = ctx.ingest_count_mem(table, (pb.ptr, n), 2)
            st = ctx.sync()
            assert rc == 0 and b == n_bases[i], (rc, b, n_bases[i])
            assert st.hits == want.hits and np.array_equal(table.counts(1), table.counts(2)), "genome image %d differs" % i
        table.clear_counts(1); table.clear_counts(2)
        want = ctx.scan_count(table, synth.reads_to_flat(reads), 1)
        rc, b, l = ctx.ingest_count_mem(table, (rpb.ptr, len(rz)), 2)
        st = ctx.sync()
        assert rc == 0 and b == reads.size and l == reads.shape[0] * 120, (rc, b, l)
        assert st.hits == want.hits and np.array_equal(table.counts(1), table.counts(2)), "reads image differs"
        print("# counters equal the host-parsed batches", flush=True)

    for rep in range(args.reps):
        t = time.perf_counter()
        tot = 0
        for i in range(args.genomes):
            pb, n, _ = images[i % len(images)]
            rc, b, l = ctx.ingest_count_mem(table, (pb.ptr, n), 2)
            assert rc == 0
            tot += b
        ctx.sync()
        dt = time.perf_counter() - t
        print(f"genomes  rep {rep}: {args.genomes} files, {tot / 1e6:.0f} Mbases in {dt * 1e3:.2f} ms = {tot / dt / 1e9:.2f} Gbases/s "
              f"({dt / max(1, args.genomes) * 1e6:.0f} us per file)", flush=True)
    for rep in range(args.reps):
        t = time.perf_counter()
        rc, b, l = ctx.ingest_count_mem(table, (rpb.ptr, len(rz)), 2)
        ctx.sync()
        dt = time.perf_counter() - t
        assert rc == 0
        print(f"reads150 rep {rep}: {b / 1e6:.0f} Mbases ({len(rtext) / 1e6:.0f} MB text) in {dt * 1e3:.2f} ms = {b / dt / 1e9:.2f} Gbases/s, "
              f"{len(rtext) / dt / 1e9:.1f} GB/s of text", flush=True)
    if hasattr(ctx, "ingest_count_mem_batch"):
        ptrs = [images[i % len(images)][0].ptr for i in range(args.genomes)]
        sizes = [images[i % len(images)][1] for i in range(args.genomes)]
        for rep in range(args.reps):
            t = time.perf_counter()
            rcs, b, l = ctx.ingest_count_mem_batch(table, ptrs, sizes, 2)
            ctx.sync()
            dt = time.perf_counter() - t
            assert not any(rcs), rcs
            print(f"genomes batch rep {rep}: {args.genomes} files, {b / 1e6:.0f} Mbases in {dt * 1e3:.2f} ms = {b / dt / 1e9:.2f} Gbases/s", flush=True)
    if args.bench_mix:
        bench_mix(ctx, table, strain, args.reps)


def bench_mix(ctx, table, strain, reps):
    import strainer2_b200 as s2
    import bench
    imgs = bench.make_file_images(strain, 0, 64)
    arena = s2.PinnedBuffer(sum(len(z) for z in imgs))
    ptrs, sizes, at = [], [], 0
    for z in imgs:
        arena.array[at:at + len(z)] = np.frombuffer(z, dtype=np.uint8)
        ptrs.append(arena.ptr + at); sizes.append(len(z))
        at += len(z)
    for rep in range(reps):
        t = time.perf_counter()
        rcs, b, l = ctx.ingest_count_mem_batch(table, ptrs, sizes, 2)
        dt = time.perf_counter() - t
        assert not any(rcs)
        print(f"bench-mix batch rep {rep}: 64 files, {b / 1e6:.0f} Mbases, {sum(sizes) / 1e6:.1f} MB compressed in {dt * 1e3:.2f} ms = {b / dt / 1e9:.2f} Gbases/s", flush=True)


if __name__ == "__main__":
    main()
