"""files_gz leg of bench.py alone (config-2 genomes as ordinary .gz images through s2_ingest_submit_mem_batch / s2_ingest_wait),
for sweeping the gz stage's knobs (S2_GZ_SUB_KB, S2_GZ_BATCH_MB, S2_INGEST_PIPES ...) - one process per setting, the images
cached under /dev/shm.  Usage: python tools/gz_sweep.py [--genomes 2000] [--distinct 160] [--steps 4] [--kind genomes|reads]"""
import argparse
import os
import pickle
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genomes", type=int, default=2000)
    ap.add_argument("--distinct", type=int, default=160)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--cache", default="/dev/shm/s2_gz_sweep.pkl")
    a = ap.parse_args()
    import strainer2_b200 as s2
    from strainer2_b200 import synth
    strain = bench.make_strain()
    if os.path.exists(a.cache):
        images = pickle.load(open(a.cache, "rb"))
    else:
        images = bench.make_file_images(strain, 0, a.distinct, "gz", min(32, len(os.sched_getaffinity(0))))
        pickle.dump(images, open(a.cache, "wb"))
    ctx = s2.Context(0, batch_bytes=64 << 20, n_lanes=4)
    table = s2.StrainTable(ctx, synth.contigs_to_flat(strain), n_cols=4)
    arena = bench.Arena(s2, images)
    ptrs, sizes = arena.cycle(a.genomes)
    secs, bases, ok = bench.time_ingest_jobs(ctx, table, 3, ptrs, sizes, a.steps)
    knobs = {k: v for k, v in os.environ.items() if k.startswith("S2_")}
    print(f"{knobs} files_gz {bases / secs / 1e9:.2f} Gbases/s ({secs / a.steps * 1e3:.1f} ms per step of {a.genomes} genomes, all on the GPU: {ok})", flush=True)


if __name__ == "__main__":
    main()
