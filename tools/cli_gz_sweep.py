"""cli_gz leg of bench.py alone: kmer_scrub_count (the drop-in executable) on 2,000 ordinary .gz genomes in /dev/shm (160 distinct,
cycled), one pair of runs per environment setting ("K=V,K=V" per argument; "-" = defaults).  The images are cached."""
import os
import pickle
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    import strainer2_b200 as s2
    strain = bench.make_strain()
    cache = "/dev/shm/s2_gz_sweep.pkl"
    if os.path.exists(cache):
        images = pickle.load(open(cache, "rb"))
    else:
        images = bench.make_file_images(strain, 0, bench.DISTINCT_GENOMES, "gz", min(32, len(os.sched_getaffinity(0))))
        pickle.dump(images, open(cache, "wb"))
    for st in sys.argv[1:] or ["-"]:
        env = {} if st == "-" else dict(kv.split("=", 1) for kv in st.split(","))
        runs = [bench.run_cli_leg(s2, strain, images, bench.GENOMES_PER_JOB, "/dev/shm", env_extra=env) for _ in range(3)]
        print(st, " | ".join("scan %.3f s wall %.2f s" % (r.get("scan_phase_s", -1), r["process_wall_s"]) for r in runs),
              "md5 equal:", len(set(r["stdout_md5"] for r in runs)) == 1, flush=True)


if __name__ == "__main__":
    main()
