#!/usr/bin/env python
"""Measure every compiled shape of the scan kernel on the same device-resident batches (one GPU).
Prints one line per (workload, shape): G lookups/s from CUDA events over K launches, and checks that
all shapes report identical hit counts.  Usage: python tools/scan_sweep.py [--steps 5] [--load 0.5]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--load", type=float, default=0.0)
    ap.add_argument("--variants", default="")
    ap.add_argument("--genomes", type=int, default=32)
    ap.add_argument("--reads", type=int, default=1_000_000)
    ap.add_argument("--union-strains", type=int, default=0,
                    help="also build ONE table from this many 5 Mb strains (config #5 shape: HBM-resident fingerprints)")
    args = ap.parse_args()
    import torch
    import strainer2_b200 as s2
    from strainer2_b200 import synth, lib
    import bench

    strain = bench.make_strain()
    ctx = s2.Context(0, batch_bytes=64 << 20, n_lanes=2)
    table = s2.StrainTable(ctx, synth.contigs_to_flat(strain), n_cols=4, load_factor=args.load)
    print(f"# table: {table.n_keys} keys, {table.n_slots} slots, probe bytes {table.probe_bytes}", flush=True)
    work = {}
    flat, bases, lookups = bench.make_batch(strain, 0, args.genomes)
    work["genomes"] = (torch.from_numpy(flat).cuda(), lookups)
    rng = synth.rng_for(3, 0)
    clean = [np.where(c == ord("N"), ord("A"), c).astype(np.uint8) for c in strain]
    others = [synth.random_bases(rng, 5_000_000) for _ in range(20)]
    r_strain = synth.sample_reads(rng, clean, args.reads // 100, 150, sub_rate=0.005, n_rate=1e-5)
    r_other = synth.sample_reads(rng, others, args.reads - args.reads // 100, 150, sub_rate=0.005, n_rate=1e-5)
    reads = np.concatenate([r_strain, r_other])
    rng.shuffle(reads)
    work["reads150"] = (torch.from_numpy(synth.reads_to_flat(reads)).cuda(), reads.shape[0] * 120)
    rel = np.concatenate([synth.contigs_to_flat([synth.mutate(c, 0.01, rng) for c in strain]) for _ in range(8)])
    work["relatives_1pct"] = (torch.from_numpy(rel).cuda(), sum(c.size - 30 for c in strain) * 8)

    tables = {"single": table}
    if args.union_strains:
        rng_u = synth.rng_for(5, 0)
        flat_u = np.concatenate([synth.contigs_to_flat(synth.genome(rng_u, 5_000_000, 40)) for _ in range(args.union_strains - 1)]
                                + [synth.contigs_to_flat(strain)])
        tu = s2.StrainTable(ctx, flat_u, n_cols=2, load_factor=args.load)
        print(f"# union table: {tu.n_keys} keys, {tu.n_slots} slots, probe bytes {tu.probe_bytes}, HBM bytes {tu.hbm_bytes}", flush=True)
        tables["union"] = tu
        del flat_u
    n_var = lib.s2_tune_scan_variant(ctx.h, -1)
    sel = [int(v) for v in args.variants.split(",")] if args.variants else list(range(n_var))
    results = []
    for tname, table in tables.items():
      for wname0, (dev, lookups) in work.items():
        wname = wname0 if tname == "single" else tname + ":" + wname0
        ref_hits = None
        for v in sel:
            assert lib.s2_tune_scan_variant(ctx.h, v) > 0
            name = lib.s2_tune_scan_variant_name(v).decode()
            for _ in range(3):
                ctx.scan_count_enqueue(table, dev, 1)
            ctx.sync()
            ctx.kernel_time(reset=True)
            ctx.event_record(0)
            for _ in range(args.steps):
                ctx.scan_count_enqueue(table, dev, 1)
            ctx.event_record(1)
            st = ctx.sync()
            ms = ctx.event_elapsed_ms(0, 1) / args.steps
            hits = st.hits // args.steps
            if ref_hits is None:
                ref_hits = hits
            ok = hits == ref_hits
            gl = lookups / ms / 1e6
            results.append({"workload": wname, "variant": v, "name": name, "ms": ms, "Glookups_s": gl,
                            "hit_rate": st.hits / max(1, st.valid_windows), "hits_equal": ok})
            print(f"{wname:16s} v{v} {name:10s} {ms:8.3f} ms  {gl:8.1f} Glookups/s  hit_rate {st.hits / max(1, st.valid_windows):.4f} "
                  f"{'ok' if ok else 'HITS DIFFER'}", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(results, open(os.path.join(ROOT, "gpurun_out", "scan_sweep.json"), "w"), indent=1)
    for t in tables.values():
        t.free()
    ctx.close()


if __name__ == "__main__":
    main()
