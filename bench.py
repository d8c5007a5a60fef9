#!/usr/bin/env python
"""bench.py - throughput of the k-mer scan hot path (BASELINE.json: Gbases/s scanned, k-mer lookups/s).

A *step* is one pass of the count scan (GEN_calculate_kmer_count's window loop,
/root/reference/src/genome_compare.c:203-229) over ONE WHOLE config-#2 job: 2,000 synthetic 5 Mb genomes (10 Gbases)
against the table of a synthetic 5 Mb strain (BASELINE.json configs[1], SURVEY 8d #2: 40 contigs, 10 N-runs; 5 % of the
genomes are relatives of the strain, substitution rate 0.5-5 %).  The 2,000 genomes are cycled from 160 distinct ones
(4 device batches of 40 genomes = 200 Mbases each, every launch's input larger than the 126 MB L2), so a default run keeps
the GPU busy for seconds, not milliseconds.  At N GPUs every rank holds a replica of the strain table and scans its own
2,000 genomes (file sharding, weak scaling, no data-path collective); the per-rank count vectors are summed with ONE NCCL
all-reduce at the end of the timed region.

  value  : whole-job Gbases/s with the batches already resident in HBM (CUDA events around the K x 50 launches on the
           launching stream, max over ranks)
  e2e    : the same metric through the C ABI with HOST buffers, host<->device copies inside the timed region.  ONE fixed
           form is reported: `files` = s2_ingest_submit_mem_batch() / s2_ingest_wait() on the step's 2,000 genomes as
           BGZF-compressed FASTA file images in pinned host memory (what the reference reads from disk): the compressed
           bytes cross PCIe, the GPU inflates (hardware engine), splits records, validates and scans, the verdicts come
           back; step i+1 is submitted before step i is waited for.  `e2e_forms` also holds: files_gz (the same genomes as
           ORDINARY single-member .gz images - the reference's own input format - through the chunk-parallel GPU gunzip),
           files_one_synchronous_call_per_step, flat (parsed batches, 1 byte per base over PCIe) and cli (the drop-in
           executable on files in /dev/shm: process wall and scan phase).
  roofline : the scan kernel.  For a single 5 Mb strain the 20 MB fingerprint table is L2 resident, so the bound is the
           SM -> L2 request path ("l2_request": one 32-byte sector request per lookup, one request per clock and SM), not
           HBM: the line carries the algorithmic bytes (33 B per lookup) against the measured HBM copy peak
           (`frac`, can exceed 1), the DRAM traffic ncu measured (`traffic`, `hbm_frac_measured`) and the fraction of the
           request-path ceiling (`l2_request_frac`).  The honest HBM case is the workload `config5_union64`.
  workloads : sub-records for config #2 (genomes), config #3 (150-bp reads: device resident, BGZF FASTQ images, ordinary
           .gz FASTQ images) and config #5 (64-strain union table, 1.28 GB of fingerprints, two-phase scan), each with its
           own roofline.
  cpu_baseline : the reference's own scan function timed on this box's host cores (rank 0, N=1)

`--impl reference` times the UNMODIFIED reference scan (GEN_calculate_kmer_count from oracle/_ref, falling back to the
oracle port) on all host cores for the same workload; it reads uncompressed FASTA from a temporary directory (zlib
inflate would only slow it further), the GPU arm's e2e reads BGZF images from pinned memory - both stated in `config`.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

STRAIN_BP = 5_000_000
GENOME_BP = 5_000_000
ALG_BYTES_PER_LOOKUP = 33.0
METRIC = "Gbases/s scanned (k-mer lookups/s) at 1/2/4/8 B200 vs host-CPU reference"
WORKLOAD = "config2: 5 Mb strain vs synthetic 5 Mb genomes (-A path), 5% relatives"
GENOMES_PER_JOB = 2000          # config #2
DISTINCT_GENOMES = 160          # cycled: 4 batches of 40
GENOMES_PER_BATCH = 40
READ_LEN = 150
READS_PER_FILE = 100_000        # a metagenome of config #3 is 200 such files' worth of reads (20 M)
READS_PER_BATCH = 400_000
READ_BATCHES = 4


def _synth():
    """strainer2_b200/synth.py without importing the package (whose __init__ loads libstrainer2_b200.so): the reference arm
    and its worker processes must not have the product library in their address space"""
    import importlib.util
    if "s2_synth_standalone" in sys.modules:
        return sys.modules["s2_synth_standalone"]
    spec = importlib.util.spec_from_file_location("s2_synth_standalone", os.path.join(ROOT, "strainer2_b200", "synth.py"))
    m = importlib.util.module_from_spec(spec)
    sys.modules["s2_synth_standalone"] = m
    spec.loader.exec_module(m)
    return m


# ------------------------------------------------------------------------------------------------
# synthetic workload
# ------------------------------------------------------------------------------------------------
def make_strain():
    synth = _synth()
    rng = synth.rng_for(2, 0)
    return synth.genome(rng, STRAIN_BP, 40, n_runs=10)


def make_genome(strain, index):
    """genome number `index` of config #2: every 20th is a relative of the strain"""
    synth = _synth()
    rng = synth.rng_for(2, 1 + index)
    if index % 20 == 7:
        rate = float(rng.uniform(0.005, 0.05))
        return [synth.mutate(c, rate, rng) for c in strain]
    return synth.genome(rng, GENOME_BP, 40)


def make_file_images(strain, first_index, n_genomes, fmt="bgzf", threads=None):
    """the same genomes as FASTA (80 columns) file images, what the reference would gzopen(): BGZF, or ordinary
    single-member gzip (level 6) - the format of every file the reference ships"""
    import gzip
    from concurrent.futures import ThreadPoolExecutor
    synth = _synth()

    def one(g):
        text = synth.fasta_bytes(make_genome(strain, first_index + g), 80)
        return synth.bgzf_bytes(text) if fmt == "bgzf" else gzip.compress(text, 6)
    with ThreadPoolExecutor(max_workers=threads or min(16, os.cpu_count() or 4)) as ex:      # zlib releases the GIL
        return list(ex.map(one, range(n_genomes)))


def make_reads(strain, index, n_reads):
    """reads of config #3: 1 % from a relative of the strain, the rest from unrelated genomes; 0.5 % substitutions, N at 1e-5"""
    synth = _synth()
    rng = synth.rng_for(3, index)
    clean = [np.where(c == ord("N"), ord("A"), c).astype(np.uint8) for c in strain]
    others = [synth.random_bases(rng, 5_000_000) for _ in range(6)]
    r1 = synth.sample_reads(rng, clean, n_reads // 100, READ_LEN, sub_rate=0.005, n_rate=1e-5)
    r2 = synth.sample_reads(rng, others, n_reads - n_reads // 100, READ_LEN, sub_rate=0.005, n_rate=1e-5)
    reads = np.concatenate([r1, r2])
    rng.shuffle(reads)
    return reads


def make_batch(strain, first_index, n_genomes):
    """flat device-format stream (contigs separated by '\\n'), bases, lookups"""
    synth = _synth()
    parts, bases, lookups = [], 0, 0
    for g in range(n_genomes):
        contigs = make_genome(strain, first_index + g)
        for c in contigs:
            bases += c.size
            if c.size >= 31:
                lookups += c.size - 30
        parts.append(synth.contigs_to_flat(contigs))
    return np.concatenate(parts), bases, lookups


# ------------------------------------------------------------------------------------------------
# NUMA: a rank's pinned host buffers should live next to its GPU
# ------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(device):
    """run this process on the CPUs of the NUMA node the GPU hangs off, so that the pinned buffers allocated afterwards
    (first touch) are local to the GPU's PCIe root; returns the node number or None.  Best effort: any failure is ignored."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(device)
        if hasattr(pr, "pci_bus_id"):                             # CUDA's own numbering of the device
            bus = "%04x:%02x:%02x.0" % (getattr(pr, "pci_domain_id", 0), pr.pci_bus_id, getattr(pr, "pci_device_id", 0))
        else:
            out = subprocess.run(["nvidia-smi", "-i", str(device), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                                 capture_output=True, text=True, timeout=20).stdout.strip().lower()
            if not out:
                return None
            bus = out[-12:] if len(out) >= 12 else out            # 00000000:1B:00.0 -> 0000:1b:00.0
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
# clocks sampling (recipe: B200_PROFILING.md)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nme, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU reference: the reference's own GEN_calculate_kmer_count through oracle/_ref/libref_prims.so
# ------------------------------------------------------------------------------------------------
def _ref_worker(args):
    """one host core: build the reference's BIO_hash from the strain once, then time its scan function
    on `files` (runs in a child process; returns (kind, seconds per file list pass, bases))"""
    strain_path, files, repeats = args
    so = os.path.join(ROOT, "oracle", "_ref", "libref_prims.so")
    if os.path.exists(so):
        R = C.CDLL(so)
        R.BIO_initHash.restype = C.c_void_p
        R.BIO_initHash.argtypes = [C.c_int]
        R.GEN_hash_sequences_set_count_vec.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        R.GEN_calculate_kmer_count.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_uint]
        h = R.BIO_initHash(8000000)                                        # src/kmer_scrub_count.c:87
        R.GEN_hash_sequences_set_count_vec(strain_path.encode(), 31, h, 1, 1, 0, 4)   # :89
        scan = lambda f: R.GEN_calculate_kmer_count(f.encode(), 31, h, 1)  # noqa: E731   genome_compare.c:179
        kind = "reference"
    else:
        from oracle import pyoracle as ou
        t = ou.OracleTable(4)
        t.build(strain_path)
        scan = lambda f: t.count_file(f, 1)  # noqa: E731
        kind = "port"
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        for f in files:
            scan(f)
        times.append(time.perf_counter() - t0)
    return kind, times


def write_cpu_inputs(tmp, strain, n_files, first_index=0):
    synth = _synth()
    sp = os.path.join(tmp, "strain.fa")
    synth.write_fasta(sp, strain, gz=False)
    files, bases = [], 0
    for i in range(n_files):
        g = make_genome(strain, first_index + i)
        p = os.path.join(tmp, f"genome{first_index + i}.fa")
        synth.write_fasta(p, g, gz=False)
        files.append(p)
        bases += sum(c.size for c in g)
    return sp, files, bases


def cpu_baseline_single_core(strain, n_files=3):
    """rank 0, N=1: one host core, `n_files` genomes of the same workload (about 10-20 s)"""
    import multiprocessing as mp
    with tempfile.TemporaryDirectory() as tmp:
        sp, files, bases = write_cpu_inputs(tmp, strain, n_files)
        with mp.get_context("spawn").Pool(1) as pool:
            kind, times = pool.map(_ref_worker, [(sp, files, 1)])[0]
    return {"value": bases / 1e9 / times[0], "unit": "Gbases/s", "cores": 1, "kind": kind,
            "sample": f"{n_files} synthetic 5 Mb genomes ({bases} bases) of the config-2 workload, scan only "
                      f"(table build excluded), {times[0]:.2f} s"}


def run_reference_arm(args):
    """--impl reference: all host cores, one reference process per core (README.md:47's recipe), each step =
    every core scans one 5 Mb genome of the workload."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    workers = max(1, min(cores, 64))
    try:
        avail = int([l for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0].split()[1]) // 1024
        workers = max(1, min(workers, avail // 1500))
    except Exception:
        pass
    strain = make_strain()
    total_passes = args.warmup + args.steps
    with tempfile.TemporaryDirectory() as tmp:
        sp, files, bases = write_cpu_inputs(tmp, strain, workers)
        per_file = bases // workers
        with mp.get_context("spawn").Pool(workers) as pool:
            res = pool.map(_ref_worker, [(sp, [files[w]], total_passes) for w in range(workers)])
    kind = res[0][0]
    # step time = slowest core in that step (all cores run concurrently, one genome each)
    step_times = [max(r[1][s] for r in res) for s in range(total_passes)][args.warmup:]
    total_t = sum(step_times)
    value = workers * per_file * args.steps / 1e9 / total_t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Gbases/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "step": f"{workers} host processes x 1 genome of 5 Mb each", "strain_bp": STRAIN_BP,
                   "input_format": {"reference_arm": "uncompressed FASTA files in a temporary directory (no zlib inflate on the reference's side)",
                                    "gpu_arm_e2e": "BGZF-compressed FASTA file images in pinned host memory"}},
        "kmer_lookups_per_s": value * 1e9 * (GENOME_BP - 40 * 30) / GENOME_BP,
        "cpu_baseline": {"value": value, "unit": "Gbases/s", "cores": workers, "kind": kind,
                         "sample": f"{workers} genomes of 5 Mb per step, one per host core, scan function only "
                                   f"(GEN_calculate_kmer_count incl. file parse), table build excluded"},
        "e2e": {"value": value, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def load_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        v = json.load(open(p))
        return float(v["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic(workload_key):
    try:
        v = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return v.get(workload_key)
    except Exception:
        return None


class Arena:
    """file images back to back in ONE pinned host buffer (like files read into memory) + their pointers and sizes"""

    def __init__(self, s2, images):
        self.buf = s2.PinnedBuffer(max(1, sum(len(z) for z in images)))
        self.ptrs, self.sizes, at = [], [], 0
        for z in images:
            self.buf.array[at:at + len(z)] = np.frombuffer(z, dtype=np.uint8)
            self.ptrs.append(self.buf.ptr + at); self.sizes.append(len(z))
            at += len(z)
        self.bytes = at

    def cycle(self, n):
        k = len(self.ptrs)
        return [self.ptrs[i % k] for i in range(n)], [self.sizes[i % k] for i in range(n)]

    def free(self):
        self.buf.free()


def time_ingest_jobs(ctx, table, col, ptrs, sizes, steps, warm=1):
    """K steps through s2_ingest_submit_mem_batch() / s2_ingest_wait(), step i+1 submitted before step i is waited for;
    every step's images cross PCIe and every step's verdicts + totals are read back.  -> (seconds, bases, all handled)"""
    for _ in range(warm):
        ctx.ingest_count_mem_batch(table, ptrs, sizes, col)
    ctx.sync()
    import torch
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    bases, ok = 0, True
    job = ctx.ingest_submit_mem_batch(table, ptrs, sizes, col)
    for i in range(steps):
        nxt = ctx.ingest_submit_mem_batch(table, ptrs, sizes, col) if i + 1 < steps else None
        rcs, b, _ = ctx.ingest_wait(job)
        ok = ok and not any(rcs)
        bases += b
        job = nxt
    torch.cuda.synchronize()
    return time.perf_counter() - t0, bases, ok


def run_cli_leg(s2, strain, images, n_files, tmp_root, env_extra=None, threads=None):
    """the drop-in executable on FILES (images written to tmp_root, the list cycles them up to n_files paths):
    -> dict with process wall, scan phase, bases"""
    import re
    import shutil
    tmp = tempfile.mkdtemp(prefix="s2bench_", dir=tmp_root)
    try:
        synth = _synth()
        synth.write_fasta(os.path.join(tmp, "strain.fa"), strain, gz=False)
        names = []
        for i, z in enumerate(images):
            names.append("g%d.fa.gz" % i)
            open(os.path.join(tmp, names[-1]), "wb").write(z)
        open(os.path.join(tmp, "A.txt"), "w").write("".join(names[i % len(names)] + "\n" for i in range(n_files)))
        open(os.path.join(tmp, "B.txt"), "w").write("")
        env = {"S2_STATS": "1"}
        if threads:
            env["S2_THREADS"] = str(threads)
        env.update(env_extra or {})
        t0 = time.perf_counter()
        p = s2.run_kmer_scrub_count(["-r", "strain.fa", "-A", "A.txt", "-B", "B.txt"], cwd=tmp, env=env, timeout=900)
        wall = time.perf_counter() - t0
        err = p.stderr.decode(errors="replace")
        m = re.search(r"table build ([0-9.]+)s \| scan ([0-9.]+)s \| all-reduce ([0-9.]+)s \| row order \+ format \+ write ([0-9.]+)s", err)
        b = re.search(r"bases=(\d+)", err)
        c = re.search(r"CUDA contexts up after ([0-9.]+)s", err)
        g = re.search(r"files: (\d+) inflated \+ split on the GPU .*?, (\d+) through host", err)
        out = {"rc": p.returncode, "process_wall_s": wall, "stdout_bytes": len(p.stdout), "stdout_md5": __import__("hashlib").md5(p.stdout).hexdigest()}
        if m and b:
            scan = float(m.group(2))
            out.update({"scan_phase_s": scan, "cuda_context_s": float(c.group(1)) if c else None, "table_build_s": float(m.group(1)),
                        "allreduce_s": float(m.group(3)), "print_s": float(m.group(4)), "bases": int(b.group(1)),
                        "value": int(b.group(1)) / 1e9 / max(scan, 1e-9), "unit": "Gbases/s (scan phase)",
                        "value_process_wall": int(b.group(1)) / 1e9 / wall})
        if g:
            out["files_gpu_ingest"], out["files_host_reader"] = int(g.group(1)), int(g.group(2))
        return out
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def run_detect_leg(s2, strain, images, n_files, reads_per_file, tmp_root, ext):
    """strain_detect (the drop-in executable) on FILES of 150-bp reads: config #4's shape, scaled - every 100th k-mer of the
    strain's first contig is informative, one SE batch line per file.  -> dict with the phase wall and what went where"""
    import re
    import shutil
    tmp = tempfile.mkdtemp(prefix="s2bench_det_", dir=tmp_root)
    try:
        synth = _synth()
        synth.write_fasta(os.path.join(tmp, "strain.fa"), strain, gz=False)
        c0 = bytes(strain[0]).replace(b"N", b"A")
        with open(os.path.join(tmp, "inf.txt"), "wb") as f:
            for i in range(0, len(c0) - 31, 100):
                f.write(c0[i:i + 31] + b"\n")
        names = []
        for i, z in enumerate(images):
            names.append("m%d.%s" % (i, ext))
            open(os.path.join(tmp, names[-1]), "wb").write(z)
        open(os.path.join(tmp, "batch.txt"), "w").write("".join("SE\t%s\n" % names[i % len(names)] for i in range(n_files)))
        t0 = time.perf_counter()
        p = s2.run_strain_detect(["-r", "strain.fa", "-a", "inf.txt", "-B", "batch.txt", "-o", "hits.gz"], cwd=tmp, env={"S2_STATS": "1"}, timeout=900)
        wall = time.perf_counter() - t0
        err = p.stderr.decode(errors="replace")
        m = re.search(r"bytes=(\d+) .*files_gpu_ingest=(\d+) files_host_reader=(\d+) detect_phase=([0-9.]+)s", err)
        out = {"rc": p.returncode, "process_wall_s": wall, "files": n_files, "reads_per_file": reads_per_file,
               "kmer_hits_gz_bytes": os.path.getsize(os.path.join(tmp, "hits.gz")) if os.path.exists(os.path.join(tmp, "hits.gz")) else None}
        if m:
            bases, phase = int(m.group(1)), float(m.group(4))
            out.update({"bases": bases, "detect_phase_s": phase, "value": bases / 1e9 / max(phase, 1e-9), "unit": "Gbases/s (batch list, wall)",
                        "files_gpu_ingest": int(m.group(2)), "files_host_reader": int(m.group(3))})
        return out
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--load-factor", type=float, default=0.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--legs", default="gz,reads,union,cli,flat", help="optional legs to run besides config #2 device-resident + files")
    ap.add_argument("--union-strains", type=int, default=64)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    legs = set(x for x in args.legs.split(",") if x)

    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import strainer2_b200 as s2
    from strainer2_b200 import synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_node = bind_to_gpu_numa_node(local) if world > 1 else None
    n_cores = len(os.sched_getaffinity(0))
    gen_threads = max(2, min(16, n_cores // max(1, world)))
    warm = max(args.warmup, 3)
    K = args.steps

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max_sum(values):
        v = torch.tensor(values, dtype=torch.float64, device=dev)
        if dist is None:
            return list(values), list(values)
        mx = v.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = v.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        return [float(x) for x in mx], [float(x) for x in sm]

    # ---- strain table replica on this rank's GPU -------------------------------------------------
    strain = make_strain()
    ctx = s2.Context(local, batch_bytes=64 << 20, n_lanes=4)
    t0 = time.perf_counter()
    table = s2.StrainTable(ctx, synth.contigs_to_flat(strain), n_cols=4, load_factor=args.load_factor)
    build_s = time.perf_counter() - t0

    # ---- config #2 inputs: 160 distinct genomes per rank (file sharding: rank r owns its own genomes) ---------------
    first = 1000 * rank + 10_000 * (world > 1)
    n_batches = DISTINCT_GENOMES // GENOMES_PER_BATCH
    launches_per_step = GENOMES_PER_JOB // GENOMES_PER_BATCH
    batches = [make_batch(strain, first + b * GENOMES_PER_BATCH, GENOMES_PER_BATCH) for b in range(n_batches)]
    dev_batches = [torch.from_numpy(f).to(dev) for f, _, _ in batches]
    step_bases = sum(batches[i % n_batches][1] for i in range(launches_per_step))
    step_lookups = sum(batches[i % n_batches][2] for i in range(launches_per_step))
    bgzf_images = make_file_images(strain, first, DISTINCT_GENOMES, "bgzf", gen_threads)
    arena = Arena(s2, bgzf_images)
    job_ptrs, job_sizes = arena.cycle(GENOMES_PER_JOB)

    from strainer2_b200 import multigpu
    ar_buf = torch.empty(table.n_keys, dtype=torch.int32, device=dev) if dist is not None else None

    def allreduce_counts():
        """the ONE collective of the path: dense first-occurrence-order count vector, SUM over ranks"""
        if dist is None:
            return
        table.gather_counts_dev(1, ar_buf)                 # slot order -> first-occurrence order (same on all ranks)
        multigpu.allreduce_counts_(ar_buf, dist)           # NCCL over NVLink; uint32 wrap-around add == int32 add
        table.scatter_counts_dev(1, ar_buf)
        torch.cuda.synchronize()

    # ---- (1) device-resident: K x 50 launches between two CUDA events on the launching stream ---------
    for i in range(warm * 4):
        ctx.scan_count_enqueue(table, dev_batches[i % n_batches], 1)
    ctx.sync()
    allreduce_counts()                                     # warm-up of the collective (NCCL channel set-up)
    ctx.kernel_time(reset=True)
    table.clear_counts(1)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    ctx.event_record(0)
    for _ in range(K):
        for i in range(launches_per_step):
            ctx.scan_count_enqueue(table, dev_batches[i % n_batches], 1)
    ctx.event_record(1)
    stats = ctx.sync()
    dev_ms = ctx.event_elapsed_ms(0, 1)
    t_ar = time.perf_counter()
    allreduce_counts()
    ar_ms = (time.perf_counter() - t_ar) * 1e3 if dist is not None else 0.0
    barrier()
    kernel_ms, launches = ctx.kernel_time(reset=True)
    total_ms = dev_ms + ar_ms
    # after the all-reduce every rank's column 1 holds the job's counters: their sum is the hits of ALL ranks
    col1_sum = int(table.counts(1).astype(np.uint64).sum())
    _, sm = reduce_max_sum([float(stats.hits)])
    allreduce_sum_check = col1_sum == int(round(sm[0]))

    # ---- (2) e2e, the fixed form: BGZF file images through the GPU ingest ---------------------------------------------
    table.clear_counts(3)
    barrier()
    files_s, files_bases, files_ok = time_ingest_jobs(ctx, table, 3, job_ptrs, job_sizes, K)
    files_stats = ctx.sync()
    barrier()
    # what one rank's job looks like in column 3 (warm-up job + K jobs) against column 1 of a single-rank run
    files_parity_ok = None
    if dist is None:
        files_parity_ok = bool(np.array_equal((K + 1) * (table.counts(1).astype(np.uint64) // K), table.counts(3).astype(np.uint64))) and files_ok
    forms = {}
    # ---- (3) the same with ONE synchronous call per step ----------------------------------------------------------------
    t0 = time.perf_counter()
    sync_bases = 0
    for _ in range(max(1, K // 2)):
        rcs, b, _ = ctx.ingest_count_mem_batch(table, job_ptrs, job_sizes, 3)
        sync_bases += b
    torch.cuda.synchronize()
    files_sync_s = time.perf_counter() - t0
    barrier()
    # ---- (4) H2D only: what the host -> device fabric gives N ranks copying their pinned arenas at the same time ------
    pin = torch.empty(256 << 20, dtype=torch.uint8, pin_memory=True)
    dst = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(2):
        dst.copy_(pin, non_blocking=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(8):
        dst.copy_(pin, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    h2d_ms = e0.elapsed_time(e1)
    barrier()
    del pin, dst
    # ---- (5) parsed flat batches in pinned memory (1 byte per base over PCIe) ----------------------------------------------
    flat_ms = flat_hits = None
    if "flat" in legs:
        pinned = []
        for f, _, _ in batches:
            pb = s2.PinnedBuffer(f.size)
            pb.array[:] = f
            pinned.append(pb)
        table.clear_counts(2)
        for i in range(2):
            ctx.scan_count_ptr(table, pinned[i % n_batches].ptr, pinned[i % n_batches].n, 2)
        table.clear_counts(2)
        barrier()
        t0 = time.perf_counter()
        flat_hits = 0
        for i in range(launches_per_step):                 # ONE step: this form is PCIe-bound at 1 byte per base
            flat_hits += ctx.scan_count_ptr(table, pinned[i % n_batches].ptr, pinned[i % n_batches].n, 2).hits
        torch.cuda.synchronize()
        flat_ms = (time.perf_counter() - t0) * 1e3
        barrier()
        for pb in pinned:
            pb.free()
    # ---- (6) ordinary .gz images of the same genomes through the chunk-parallel GPU gunzip -------------------------------
    gz_s = gz_bases = gz_ok = gz_arena = None
    if "gz" in legs:
        gz_images = make_file_images(strain, first, DISTINCT_GENOMES, "gz", gen_threads)
        gz_arena = Arena(s2, gz_images)
        gp, gs = gz_arena.cycle(GENOMES_PER_JOB)
        barrier()
        gz_s, gz_bases, gz_ok = time_ingest_jobs(ctx, table, 3, gp, gs, max(1, K // 2))
        ctx.sync()
        barrier()
    clocks = sampler.stop()                                # sampled through the timed regions above

    mx, sm = reduce_max_sum([total_ms, files_s * 1e3, files_sync_s * 1e3, h2d_ms, flat_ms or 0.0, (gz_s or 0.0) * 1e3,
                             float(K * step_bases), float(K * step_lookups), float(files_bases), float(sync_bases), float(gz_bases or 0)])
    total_ms, files_ms, files_sync_ms, h2d_ms_max, flat_ms_max, gz_ms = mx[0], mx[1], mx[2], mx[3], mx[4], mx[5]
    all_bases, all_lookups, all_files_bases, all_sync_bases, all_gz_bases = sm[6], sm[7], sm[8], sm[9], sm[10]

    # ---- optional workloads: config #3 (reads) and config #5 (union table) ----------------------------------------------
    workloads = {}
    peak, peak_src = load_peak()
    sm_clock_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
    n_sm = 148

    def roofline_l2(lookups_per_launch, per_launch_ms, key):
        achieved = lookups_per_launch * ALG_BYTES_PER_LOOKUP / (per_launch_ms * 1e-3) / 1e9
        traffic = load_traffic(key)
        r = {"bound": "l2_request", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
             "frac_algorithmic_vs_hbm": achieved / peak, "traffic": traffic, "peak_source": peak_src,
             "kernel": "s2_scan_kernel<COUNT>", "avg_launch_ms": per_launch_ms, "alg_bytes_per_lookup": ALG_BYTES_PER_LOOKUP,
             "lookups_per_launch": lookups_per_launch,
             "l2_request_frac": lookups_per_launch / (per_launch_ms * 1e-3) / (n_sm * sm_clock_hz),
             "note": "the strain's 20 MB fingerprint table is L2 resident: every lookup is one 32-byte sector request to L2, an SM sends one "
                     "per clock; frac compares the ALGORITHMIC bytes with the HBM copy peak and can exceed 1, hbm_frac_measured is the DRAM "
                     "traffic ncu measured over the same time"}
        if traffic:
            r["hbm_frac_measured"] = traffic / (per_launch_ms * 1e-3) / 1e9 / peak
        return r

    if "reads" in legs:
        rb = [make_reads(strain, 100 * rank + b, READS_PER_BATCH) for b in range(READ_BATCHES)]
        rdev = [torch.from_numpy(synth.reads_to_flat(r)).to(dev) for r in rb]
        r_launches = 20_000_000 // READS_PER_BATCH                       # one metagenome of config #3 per step
        r_bases = READS_PER_BATCH * READ_LEN
        r_lookups = READS_PER_BATCH * (READ_LEN - 30)
        for i in range(8):
            ctx.scan_count_enqueue(table, rdev[i % READ_BATCHES], 2)
        ctx.sync(); ctx.kernel_time(reset=True)
        barrier()
        ctx.event_record(2)
        for _ in range(K):
            for i in range(r_launches):
                ctx.scan_count_enqueue(table, rdev[i % READ_BATCHES], 2)
        ctx.event_record(3)
        r_stats = ctx.sync()
        r_ms = ctx.event_elapsed_ms(2, 3)
        r_kernel_ms, r_n = ctx.kernel_time(reset=True)
        barrier()
        # the same reads as FASTQ files of 100,000 reads: BGZF images, and ordinary .gz images
        import gzip
        from concurrent.futures import ThreadPoolExecutor
        pieces = [synth.fastq_bytes(r[k:k + READS_PER_FILE]) for r in rb for k in range(0, READS_PER_BATCH, READS_PER_FILE)]
        with ThreadPoolExecutor(max_workers=gen_threads) as ex:
            r_bgzf = list(ex.map(synth.bgzf_bytes, pieces))
            r_gz = list(ex.map(lambda t: gzip.compress(t, 6), pieces)) if "gz" in legs else None
        files_per_step = 20_000_000 // READS_PER_FILE
        ra = Arena(s2, r_bgzf)
        rp, rs = ra.cycle(files_per_step)
        r_files_s, r_files_bases, r_files_ok = time_ingest_jobs(ctx, table, 2, rp, rs, max(1, K // 2))
        ctx.sync(); barrier()
        ra.free()
        r_gz_s = r_gz_bases = r_gz_ok = None
        if r_gz:
            rga = Arena(s2, r_gz)
            rp2, rs2 = rga.cycle(files_per_step)
            r_gz_s, r_gz_bases, r_gz_ok = time_ingest_jobs(ctx, table, 2, rp2, rs2, max(1, K // 2))
            ctx.sync(); barrier()
            rga.free()
        mxr, smr = reduce_max_sum([r_ms, r_files_s * 1e3, (r_gz_s or 0.0) * 1e3, float(K * r_launches * r_bases), float(r_files_bases), float(r_gz_bases or 0)])
        workloads["config3_reads"] = {
            "workload": "config3: 5 Mb strain vs 150-bp reads (1 % from a relative of the strain), one step = one metagenome of 20 M reads "
                        f"cycled from {READ_BATCHES} distinct batches of {READS_PER_BATCH} reads",
            "value": smr[3] / 1e9 / (mxr[0] * 1e-3), "unit": "Gbases/s",
            "kmer_lookups_per_s": smr[3] / READ_LEN * (READ_LEN - 30) / (mxr[0] * 1e-3),
            "hit_rate": r_stats.hits / max(1, r_stats.valid_windows),
            "e2e": {"value": smr[4] / 1e9 / (mxr[1] * 1e-3), "unit": "Gbases/s", "form": "files",
                    "what": f"{files_per_step} BGZF FASTQ file images of {READS_PER_FILE} reads per step through s2_ingest_submit_mem_batch() / s2_ingest_wait()",
                    "h2d_bytes_per_step": int(sum(rs)), "all_files_handled_on_gpu": r_files_ok},
            "e2e_gz": None if r_gz_s is None else {
                "value": smr[5] / 1e9 / (mxr[2] * 1e-3), "unit": "Gbases/s", "form": "files_gz",
                "what": "the same files as ordinary single-member .gz (gzip -6) through the chunk-parallel GPU gunzip",
                "h2d_bytes_per_step": int(sum(rs2)), "all_files_handled_on_gpu": r_gz_ok},
            "roofline": roofline_l2(r_lookups, r_kernel_ms / max(1, r_n), "config3_reads"),
            "gpu_launches": int(r_n),
        }
        del rdev
        if "cli" in legs and rank == 0 and world == 1:
            # config #4's shape through the strain_detect executable: one metagenome's worth - 200 files of 100,000 reads (cycled from
            # the 16 distinct ones; 64 files were over in 0.12 s, most of it the first file of each pipeline: profiles/r2z_detect_sweep.txt)
            try:
                shm_d = "/dev/shm" if os.path.isdir("/dev/shm") else None
                det = run_detect_leg(s2, strain, r_bgzf, 200, READS_PER_FILE, shm_d, "fastq.bgz")
                det_gz = run_detect_leg(s2, strain, r_gz, 200, READS_PER_FILE, shm_d, "fastq.gz") if r_gz else None
                workloads["config4_detect"] = {
                    "workload": "config4: strain_detect (pass 1 on the GPU, pairing loop replayed on the host) of 1,250 informative k-mers over 200 files of "
                                f"{READS_PER_FILE} 150-bp reads, through the drop-in executable on files in /dev/shm",
                    "value": det.get("value"), "unit": "Gbases/s (wall time of the batch list)", "cli": det, "cli_gz": det_gz,
                }
            except Exception as e:
                workloads["config4_detect"] = {"error": repr(e)}

    if "union" in legs and world == 1:
        # config #5's shape: ONE table of 64 strains (320 M keys, 1.28 GB of fingerprints: HBM resident), genome batches
        rng_u = synth.rng_for(5, 0)
        flat_u = np.concatenate([synth.contigs_to_flat(synth.genome(rng_u, 5_000_000, 40)) for _ in range(args.union_strains - 1)]
                                + [synth.contigs_to_flat(strain)])
        tu = s2.StrainTable(ctx, flat_u, n_cols=2)
        del flat_u
        # one batch = all distinct genomes of the job (the 1.28 GB of fingerprints are swept once per batch whatever its size:
        # the larger the batch, the more probes every fetched sector serves - 4 per sector at 200 MB, 16 at 800 MB)
        u_batch = torch.cat(dev_batches)
        u_batch_bases = sum(b[1] for b in batches)
        u_batch_lookups = sum(b[2] for b in batches)
        for i in range(2):
            ctx.scan_count_enqueue(tu, u_batch, 1)
        ctx.sync(); ctx.kernel_time(reset=True)
        ctx.event_record(2)
        u_launches = max(1, launches_per_step // n_batches) * max(1, K // 2)
        for i in range(u_launches):
            ctx.scan_count_enqueue(tu, u_batch, 1)
        ctx.event_record(3)
        u_stats = ctx.sync()
        u_ms = ctx.event_elapsed_ms(2, 3)
        u_kernel_ms, u_n = ctx.kernel_time(reset=True)
        u_lookups = u_batch_lookups * u_launches
        u_bases = u_batch_bases * u_launches
        del u_batch
        per_batch_ms = u_ms / u_launches
        lookups_per_batch = u_lookups / u_launches
        ach = lookups_per_batch * ALG_BYTES_PER_LOOKUP / (per_batch_ms * 1e-3) / 1e9
        traffic = load_traffic("config5_union64")
        workloads["config5_union64"] = {
            "workload": f"config5: ONE union table of {args.union_strains} strains ({int(tu.n_keys)} keys, {int(tu.probe_bytes)} bytes of fingerprints, HBM "
                        f"resident) scanned with the config-2 genomes in batches of {n_batches * GENOMES_PER_BATCH} ({u_batch_bases // 10**6} Mbases per batch); two-phase scan (radix partition by hash, then partition-wise probe)",
            "value": u_bases / 1e9 / (u_ms * 1e-3), "unit": "Gbases/s", "kmer_lookups_per_s": u_lookups / (u_ms * 1e-3),
            "strain_lookups_per_s": args.union_strains * u_lookups / (u_ms * 1e-3),
            "hit_rate": u_stats.hits / max(1, u_stats.valid_windows),
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                         "peak_source": peak_src, "kernel": "s2_partition_kernel + s2_probe_all_kernel (one batch = both launches)",
                         "avg_launch_ms": per_batch_ms, "alg_bytes_per_lookup": ALG_BYTES_PER_LOOKUP, "lookups_per_launch": lookups_per_batch,
                         "hbm_frac_measured": (traffic / (per_batch_ms * 1e-3) / 1e9 / peak) if traffic else None},
            "gpu_launches": int(u_n),
        }
        tu.free()

    # ---- the drop-in executable on files -------------------------------------------------------------------------------------
    cli = cli_gz = cli_multi = None
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else None
    if "cli" in legs and rank == 0 and world == 1:
        try:
            # (the box's host side is shared and noisy - context creation alone varies between 0.25 and 2 s: the better of two
            # runs is reported, the other one's scan phase beside it)
            def best_of_two(images):
                r1 = run_cli_leg(s2, strain, images, GENOMES_PER_JOB, shm)
                r2 = run_cli_leg(s2, strain, images, GENOMES_PER_JOB, shm)
                a, b = (r1, r2) if r1.get("scan_phase_s", 1e9) <= r2.get("scan_phase_s", 1e9) else (r2, r1)
                a["other_run_scan_phase_s"] = b.get("scan_phase_s")
                a["runs_agree"] = r1.get("stdout_md5") == r2.get("stdout_md5")
                return a
            cli = best_of_two(bgzf_images)
            if "gz" in legs:
                cli_gz = best_of_two(gz_images)
        except Exception as e:
            cli = {"error": repr(e)}
    if world > 1:
        # the executable's own multi-GPU path (S2_GPUS=N: replicas, files dealt over the GPUs, all-reduce over peer memory)
        # against the same run on one GPU: byte-identical count tables
        if rank == 0:
            try:
                small = bgzf_images[:24]
                one = run_cli_leg(s2, strain, small, 48, shm, {"S2_GPUS": "1"})
                many = run_cli_leg(s2, strain, small, 48, shm, {"S2_GPUS": str(world)})
                cli_multi = {"parity": one["rc"] == 0 and many["rc"] == 0 and one["stdout_md5"] == many["stdout_md5"] and one["stdout_bytes"] > 1000,
                             "one_gpu": one, "n_gpus": many}
            except Exception as e:
                cli_multi = {"parity": False, "error": repr(e)}
        barrier()

    if rank == 0:
        per_launch_ms = kernel_ms / max(1, launches)
        lookups_per_launch = K * step_lookups / max(1, launches)
        value = all_bases / 1e9 / (total_ms * 1e-3)
        h2d_gbs = world * 8 * (256 << 20) / 1e9 / (h2d_ms_max * 1e-3)
        bgzf_bytes_per_base = sum(job_sizes) / step_bases
        forms["files"] = {"value": all_files_bases / 1e9 / (files_ms * 1e-3), "unit": "Gbases/s",
                          "h2d_bytes_per_step": int(sum(job_sizes)), "d2h_bytes_per_step": 32 * ((sum(job_sizes) >> 24) + 1),
                          "what": "s2_ingest_submit_mem_batch() / s2_ingest_wait() on the step's 2,000 genomes as BGZF FASTA file images in pinned host "
                                  "memory, the next step submitted before the previous one is waited for: H2D of the compressed bytes + hardware "
                                  "inflate + record splitting + validation + scan kernel + D2H of the verdicts",
                          "counts_equal_device_path": files_parity_ok}
        forms["files_one_synchronous_call_per_step"] = {"value": all_sync_bases / 1e9 / (files_sync_ms * 1e-3), "unit": "Gbases/s",
                                                        "what": "the same through s2_ingest_count_mem_batch(): the pipeline drains between steps"}
        if gz_s is not None:
            forms["files_gz"] = {"value": all_gz_bases / 1e9 / (gz_ms * 1e-3), "unit": "Gbases/s", "h2d_bytes_per_step": int(sum(gs)),
                                 "what": "the same 2,000 genomes as ORDINARY single-member .gz images (gzip -6, the reference's own input format): "
                                         "H2D + chunk-parallel GPU gunzip (block finder, speculative decode, chain, translate, CRC-32) + record splitting + scan",
                                 "all_files_handled_on_gpu": gz_ok}
        if flat_ms is not None:
            forms["flat"] = {"value": world * step_bases / 1e9 / (flat_ms_max * 1e-3), "unit": "Gbases/s",
                             "h2d_bytes_per_step": int(sum(batches[i % n_batches][0].size for i in range(launches_per_step))), "d2h_bytes_per_step": 16 * launches_per_step,
                             "what": "s2_scan_count() on parsed flat batches in pinned memory: H2D (1 byte per base) + scan kernel + D2H of each launch's hit statistics"}
        if cli:
            forms["cli"] = dict(cli, what="strainer2_b200/bin/kmer_scrub_count on 2,000 BGZF FASTA files in /dev/shm (160 distinct, cycled), default threads")
        if cli_gz:
            forms["cli_gz"] = dict(cli_gz, what="the same with ordinary .gz files")
        line = {
            "metric": METRIC, "value": value, "unit": "Gbases/s", "n_gpus": world, "steps": K,
            "warmup": warm, "ms_per_step": total_ms / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "step": f"one whole config-2 job per rank: {GENOMES_PER_JOB} genomes of 5 Mb = {step_bases} bases, {launches_per_step} launches "
                               f"cycled from {DISTINCT_GENOMES} distinct genomes ({n_batches} device batches of {GENOMES_PER_BATCH})",
                       "bases_per_step": int(step_bases), "strain_keys": int(table.n_keys),
                       "table_probe_bytes": int(table.probe_bytes), "table_hbm_bytes": int(table.hbm_bytes),
                       "l2": "inputs larger than L2 (200 MB per launch, 4 alternating batches; the file images come from host memory every step)",
                       "parallelism": f"file-shard x{world}, replicated table, 1 NCCL all-reduce at the end",
                       "input_format": {"value": "parsed batches resident in HBM", "e2e": "BGZF-compressed FASTA file images in pinned host memory",
                                        "reference_arm": "uncompressed FASTA files in a temporary directory"},
                       "numa": "ranks bound to their GPU's NUMA node" if numa_node is not None else "unbound"},
            "kmer_lookups_per_s": all_lookups / (total_ms * 1e-3),
            "hits": int(stats.hits), "hit_rate": stats.hits / max(1, stats.valid_windows),
            "table_build_s": build_s,
            "allreduce_ms": ar_ms,
            "allreduce_sum_check": allreduce_sum_check,
            "timed_region_s": total_ms * 1e-3,
            "e2e": dict(forms["files"], form="files"),
            "e2e_forms": forms,
            "e2e_ceiling": {"h2d_only_gbs": h2d_gbs, "implied_e2e_ceiling_gbases_per_s": h2d_gbs / bgzf_bytes_per_base,
                            "bgzf_bytes_per_base": bgzf_bytes_per_base,
                            "what": f"{world} rank(s) copying 8 x 256 MB from pinned host memory to their GPU at the same time, no kernels: what the "
                                    "host -> device fabric gives; the e2e forms cannot exceed it"},
            "cli_multi_gpu_parity": None if cli_multi is None else cli_multi.get("parity"),
            "cli_multi_gpu": cli_multi,
            "gpu_launches": int(launches),
            "roofline": roofline_l2(lookups_per_launch, per_launch_ms, "config2"),
            "workloads": workloads,
            "clocks": clocks,
        }
        line["workloads"]["config2_genomes"] = {"workload": WORKLOAD, "value": value, "unit": "Gbases/s", "kmer_lookups_per_s": line["kmer_lookups_per_s"],
                                                "e2e": line["e2e"], "roofline": line["roofline"]}
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline_single_core(strain)
            except Exception as e:  # the baseline is a reported extra, never the measurement itself
                line["cpu_baseline"] = {"value": None, "unit": "Gbases/s", "cores": 1, "kind": "unavailable", "sample": repr(e)}
        print(json.dumps(line))
    arena.free()
    if gz_arena is not None:
        gz_arena.free()
    table.free()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
