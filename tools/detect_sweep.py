"""config4_detect leg of bench.py alone: strain_detect (the drop-in executable) on 64 files of 100,000 150-bp reads in /dev/shm,
one run per environment setting given on the command line ("K=V,K=V" per argument; "-" = defaults).  The inputs are cached.
Usage: python tools/detect_sweep.py [--ext fastq.bgz|fastq.gz] [--files 64] - S2_INGEST_PIPES=6 ..."""
import argparse
import gzip
import os
import re
import sys
import time
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ext", default="fastq.bgz")
    ap.add_argument("--files", type=int, default=64)
    ap.add_argument("--dir", default="/dev/shm/s2_detect_sweep")
    ap.add_argument("settings", nargs="*", default=["-"])
    a = ap.parse_args()
    import strainer2_b200 as s2
    synth = bench._synth()
    strain = bench.make_strain()
    d = a.dir
    if not os.path.exists(os.path.join(d, "ready." + a.ext)):
        os.makedirs(d, exist_ok=True)
        rb = [bench.make_reads(strain, b, bench.READS_PER_BATCH) for b in range(bench.READ_BATCHES)]
        pieces = [synth.fastq_bytes(r[k:k + bench.READS_PER_FILE]) for r in rb for k in range(0, bench.READS_PER_BATCH, bench.READS_PER_FILE)]
        with ThreadPoolExecutor(max_workers=min(16, len(os.sched_getaffinity(0)))) as ex:
            images = list(ex.map(synth.bgzf_bytes if a.ext.endswith("bgz") else (lambda t: gzip.compress(t, 6)), pieces))
        synth.write_fasta(os.path.join(d, "strain.fa"), strain, gz=False)
        c0 = bytes(strain[0]).replace(b"N", b"A")
        with open(os.path.join(d, "inf.txt"), "wb") as f:
            for i in range(0, len(c0) - 31, 100):
                f.write(c0[i:i + 31] + b"\n")
        for i, z in enumerate(images):
            open(os.path.join(d, "m%d.%s" % (i, a.ext)), "wb").write(z)
        open(os.path.join(d, "n_images." + a.ext), "w").write(str(len(images)))
        open(os.path.join(d, "ready." + a.ext), "w").write("1")
    n_img = int(open(os.path.join(d, "n_images." + a.ext)).read())
    open(os.path.join(d, "batch.txt"), "w").write("".join("SE\tm%d.%s\n" % (i % n_img, a.ext) for i in range(a.files)))
    for st in a.settings:
        env = {"S2_STATS": "1"}
        if st != "-":
            env.update(dict(kv.split("=", 1) for kv in st.split(",")))
        t0 = time.perf_counter()
        p = s2.run_strain_detect(["-r", "strain.fa", "-a", "inf.txt", "-B", "batch.txt", "-o", "hits.gz"], cwd=d, env=env, timeout=900)
        wall = time.perf_counter() - t0
        err = p.stderr.decode(errors="replace")
        m = re.search(r"\[s2 detect\].*", err)
        print(f"{st} rc={p.returncode} wall={wall:.2f}s {m.group(0) if m else err[-300:]}", flush=True)
        if "S2_INGEST_TRACE" in env:
            for ln in [x for x in err.splitlines() if "[s2 ingest] detect" in x][-16:]:
                print("    " + ln, flush=True)


if __name__ == "__main__":
    main()
