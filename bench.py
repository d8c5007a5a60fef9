#!/usr/bin/env python
"""bench.py - throughput of the k-mer scan hot path (BASELINE.json: Gbases/s scanned, k-mer lookups/s).

A *step* is one pass of the count scan (GEN_calculate_kmer_count's window loop,
/root/reference/src/genome_compare.c:203-229) over one batch of synthetic genomes.

Workload (BASELINE.json configs[1]): synthetic 5 Mb strain genome (40 contigs, 10 N-runs)
scrubbed against synthetic 5 Mb genomes, 5 % of them relatives of the strain (substitution rate
0.5-5 %), 95 % independent random; one step = one batch of `--genomes-per-step` genomes
(default 64 = 320 Mbases, larger than the 126 MB L2, so consecutive steps never find their input
cached; two distinct batches alternate).  At N GPUs every rank holds a replica of the strain table,
scans its own batches (file sharding, weak scaling, no data-path collective) and the per-rank count
vectors are summed with ONE NCCL all-reduce at the end of the timed region.

  value  : whole-job Gbases/s with the batches already resident in HBM (CUDA events around K launches
           on the launching stream, max over ranks)
  e2e    : same metric through the C-ABI with HOST buffers, host<->device copies inside the timed region.  Two
           forms are measured and the faster one is reported as `e2e` (both are in `e2e_forms`):
             files : s2_ingest_submit_mem_batch() / s2_ingest_wait() on the step's genomes as BGZF-compressed FASTA
                     file images in pinned host memory - what the reference reads from disk; the compressed bytes
                     cross PCIe, the GPU inflates (hardware engine), splits records, validates and scans; the
                     verdicts come back; step i+1 is submitted before step i is waited for (double buffering)
             flat  : s2_scan_count() on parsed flat batches in pinned memory (1 byte per base over PCIe) + D2H of
                     the step's hit statistics
  roofline : the scan kernel against the measured HBM copy peak (MEASURED_PEAKS.json), algorithmic
           bytes = 33 B per k-mer lookup (1 B base + one 32-byte fingerprint bucket)
  cpu_baseline : the reference's own scan function timed on this box's host cores (rank 0, N=1)

`--impl reference` times the UNMODIFIED reference scan (GEN_calculate_kmer_count from oracle/_ref,
falling back to the oracle port) on all host cores for the same workload.
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


# ------------------------------------------------------------------------------------------------
# synthetic workload
# ------------------------------------------------------------------------------------------------
def make_strain():
    from strainer2_b200 import synth
    rng = synth.rng_for(2, 0)
    return synth.genome(rng, STRAIN_BP, 40, n_runs=10)


def make_genome(strain, index):
    """genome number `index` of config #2: every 20th is a relative of the strain"""
    from strainer2_b200 import synth
    rng = synth.rng_for(2, 1 + index)
    if index % 20 == 7:
        rate = float(rng.uniform(0.005, 0.05))
        return [synth.mutate(c, rate, rng) for c in strain]
    return synth.genome(rng, GENOME_BP, 40)


def make_file_images(strain, first_index, n_genomes):
    """the same genomes as BGZF-compressed FASTA (80 columns) file images, what the reference would gzopen()"""
    from concurrent.futures import ThreadPoolExecutor
    from strainer2_b200 import synth

    def one(g):
        return synth.bgzf_bytes(synth.fasta_bytes(make_genome(strain, first_index + g), 80))
    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 4)) as ex:      # zlib releases the GIL
        return list(ex.map(one, range(n_genomes)))


def make_batch(strain, first_index, n_genomes):
    """flat device-format stream (contigs separated by '\\n'), bases, lookups"""
    from strainer2_b200 import synth
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
    from strainer2_b200 import synth
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
                   "step": f"{workers} host processes x 1 genome of 5 Mb each", "strain_bp": STRAIN_BP},
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genomes-per-step", type=int, default=64)
    ap.add_argument("--load-factor", type=float, default=0.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import strainer2_b200 as s2

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

    # ---- strain table replica on this rank's GPU -------------------------------------------------
    from strainer2_b200 import synth
    strain = make_strain()
    ctx = s2.Context(local, batch_bytes=64 << 20, n_lanes=4)
    t0 = time.perf_counter()
    table = s2.StrainTable(ctx, synth.contigs_to_flat(strain), n_cols=4, load_factor=args.load_factor)
    build_s = time.perf_counter() - t0

    # ---- two distinct batches per rank (file sharding: rank r owns genomes r, r+world, ...) --------
    G = args.genomes_per_step
    batches = []
    for b in range(2):
        flat, bases, lookups = make_batch(strain, 1000 * rank + 100 * b + 10_000 * (world > 1), G)
        batches.append((flat, bases, lookups))
    dev_batches = [torch.from_numpy(f).to(dev) for f, _, _ in batches]
    pinned = []
    for f, _, _ in batches:
        pb = s2.PinnedBuffer(f.size)
        pb.array[:] = f
        pinned.append(pb)
    # the genomes of both batches once more, as the files the reference would read: BGZF-compressed FASTA images,
    # back to back in one pinned buffer per batch (like files read into memory)
    img_bufs, img_ptrs, img_sizes = [], [], []
    for b in range(2):
        imgs = make_file_images(strain, 1000 * rank + 100 * b + 10_000 * (world > 1), G)
        arena = s2.PinnedBuffer(sum(len(z) for z in imgs))
        ptrs, sizes, at = [], [], 0
        for z in imgs:
            arena.array[at:at + len(z)] = np.frombuffer(z, dtype=np.uint8)
            ptrs.append(arena.ptr + at); sizes.append(len(z))
            at += len(z)
        img_bufs.append(arena); img_ptrs.append(ptrs); img_sizes.append(sizes)
        del imgs
    sizes = img_sizes[0]
    step_bases = [b for _, b, _ in batches]
    step_lookups = [l for _, _, l in batches]

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

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

    # ---- (1) device-resident: K launches between two CUDA events on the launching stream ---------
    for i in range(max(args.warmup, 3)):
        ctx.scan_count_enqueue(table, dev_batches[i % 2], 1)
    ctx.sync()
    allreduce_counts()                                     # warm-up of the collective (NCCL channel set-up)
    ctx.kernel_time(reset=True)
    table.clear_counts(1)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    ctx.event_record(0)
    for i in range(args.steps):
        ctx.scan_count_enqueue(table, dev_batches[i % 2], 1)
    ctx.event_record(1)
    stats = ctx.sync()
    dev_ms = ctx.event_elapsed_ms(0, 1)
    t_ar = time.perf_counter()
    allreduce_counts()
    ar_ms = (time.perf_counter() - t_ar) * 1e3 if dist is not None else 0.0
    barrier()
    kernel_ms, launches = ctx.kernel_time(reset=True)
    my_bases = sum(step_bases[i % 2] for i in range(args.steps))
    my_lookups = sum(step_lookups[i % 2] for i in range(args.steps))
    total_ms = dev_ms + ar_ms

    # ---- (2) end to end: pinned host batches through s2_scan_count (H2D + kernel + stats D2H) ----
    for i in range(2):
        ctx.scan_count_ptr(table, pinned[i % 2].ptr, pinned[i % 2].n, 2)
    table.clear_counts(2)
    barrier()
    t0 = time.perf_counter()
    e2e_hits = 0
    for i in range(args.steps):
        st = ctx.scan_count_ptr(table, pinned[i % 2].ptr, pinned[i % 2].n, 2)      # returns the step's result
        e2e_hits += st.hits
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    # ---- (3) end to end from FILE IMAGES: BGZF FASTA in pinned host memory through the GPU ingest ----
    for i in range(2):
        ctx.ingest_count_mem_batch(table, img_ptrs[i % 2], img_sizes[i % 2], 3)
    ctx.sync()
    table.clear_counts(3)
    barrier()
    t0 = time.perf_counter()
    files_bases = 0
    # double buffered like any input pipeline: step i+1 is submitted (its H2D copies start) before step i's verdicts
    # and totals are read back; every step's inputs cross PCIe and every step's result is read inside the timed region
    job = ctx.ingest_submit_mem_batch(table, img_ptrs[0], img_sizes[0], 3)
    for i in range(args.steps):
        nxt = ctx.ingest_submit_mem_batch(table, img_ptrs[(i + 1) % 2], img_sizes[(i + 1) % 2], 3) if i + 1 < args.steps else None
        rcs, b, _ = ctx.ingest_wait(job)                   # the step's verdicts + totals
        assert not any(rcs)
        files_bases += b
        job = nxt
    torch.cuda.synchronize()
    files_s = time.perf_counter() - t0
    files_stats = ctx.sync()
    barrier()
    # the same with ONE synchronous call per step (the pipeline fills and drains every step): reported beside it
    t0 = time.perf_counter()
    sync_bases = 0
    for i in range(args.steps):
        rcs, b, _ = ctx.ingest_count_mem_batch(table, img_ptrs[i % 2], img_sizes[i % 2], 3)
        sync_bases += b
    torch.cuda.synchronize()
    files_sync_s = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop()                                # sampled through all timed regions
    parity_ok = bool(np.array_equal(table.counts(2), table.counts(1))) if dist is None else None
    # the file-image path saw the same genomes in the same alternation as the flat path, twice (double buffered, then one
    # synchronous call per step): its counter column is exactly twice the flat path's
    files_parity_ok = bool(np.array_equal(2 * table.counts(2).astype(np.uint64), table.counts(3).astype(np.uint64))) and files_bases == my_bases == sync_bases

    # ---- reduce over ranks ----------------------------------------------------------------------
    vals = torch.tensor([total_ms, e2e_s * 1e3, float(my_bases), float(my_lookups), kernel_ms, float(launches),
                         files_s * 1e3, float(files_bases), files_sync_s * 1e3], dtype=torch.float64, device=dev)
    if dist is not None:
        mx = vals.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        total_ms, e2e_ms, files_ms, files_sync_ms = float(mx[0]), float(mx[1]), float(mx[6]), float(mx[8])
        all_bases, all_lookups, all_files_bases = float(sm[2]), float(sm[3]), float(sm[7])
    else:
        e2e_ms, files_ms, files_sync_ms = e2e_s * 1e3, files_s * 1e3, files_sync_s * 1e3
        all_bases, all_lookups, all_files_bases = float(my_bases), float(my_lookups), float(files_bases)

    if rank == 0:
        peak, peak_src = load_peak()
        per_launch_ms = kernel_ms / max(1, launches)
        lookups_per_launch = my_lookups / args.steps
        achieved = lookups_per_launch * ALG_BYTES_PER_LOOKUP / (per_launch_ms * 1e-3) / 1e9
        value = all_bases / 1e9 / (total_ms * 1e-3)
        forms = {
            "files": {"value": all_files_bases / 1e9 / (files_ms * 1e-3), "unit": "Gbases/s",
                      "h2d_bytes_per_step": int(sum(sizes)), "d2h_bytes_per_step": 32 * ((sum(sizes) >> 24) + 1),
                      "what": "s2_ingest_submit_mem_batch() / s2_ingest_wait() on the step's genomes as BGZF FASTA file images in pinned host "
                              "memory, the next step submitted before the previous one is waited for: H2D of the compressed bytes + hardware "
                              "inflate + record splitting + validation + scan kernel + D2H of the verdicts",
                      "text_bytes_per_step": int(step_bases[0] + step_bases[0] // 80 + 50 * G),
                      "counts_equal_flat_path": files_parity_ok},
            "files_one_synchronous_call_per_step": {"value": all_files_bases / 1e9 / (files_sync_ms * 1e-3), "unit": "Gbases/s",
                                                    "what": "the same through s2_ingest_count_mem_batch(): the copy -> inflate -> kernels pipeline fills and drains every step"},
            "flat": {"value": all_bases / 1e9 / (e2e_ms * 1e-3), "unit": "Gbases/s",
                     "h2d_bytes_per_step": int(pinned[0].n), "d2h_bytes_per_step": 16,
                     "what": "s2_scan_count() on pinned host batches: H2D + scan kernel + D2H of the step's hit statistics"},
        }
        best = "files" if forms["files"]["value"] >= forms["flat"]["value"] else "flat"
        line = {
            "metric": METRIC, "value": value, "unit": "Gbases/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "genomes_per_step": G, "bases_per_step": step_bases[0], "strain_keys": int(table.n_keys),
                       "table_probe_bytes": int(table.probe_bytes), "table_hbm_bytes": int(table.hbm_bytes),
                       "l2": "inputs larger than L2 (320 MB per step, two alternating batches; the file images come from host memory every step)",
                       "parallelism": f"file-shard x{world}, replicated table, 1 NCCL all-reduce at the end",
                       "numa": "ranks bound to their GPU's NUMA node" if numa_node is not None else "unbound"},
            "kmer_lookups_per_s": all_lookups / (total_ms * 1e-3),
            "hits": int(stats.hits), "hit_rate": stats.hits / max(1, stats.valid_windows),
            "table_build_s": build_s,
            "allreduce_ms": ar_ms,
            "e2e": dict(forms[best], form=best),
            "e2e_forms": forms,
            "e2e_counts_equal_device_path": parity_ok,
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": load_traffic("config2"), "peak_source": peak_src,
                         "kernel": "s2_scan_kernel<COUNT>", "avg_launch_ms": per_launch_ms,
                         "alg_bytes_per_lookup": ALG_BYTES_PER_LOOKUP, "lookups_per_launch": lookups_per_launch},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline_single_core(strain)
            except Exception as e:  # the baseline is a reported extra, never the measurement itself
                line["cpu_baseline"] = {"value": None, "unit": "Gbases/s", "cores": 1, "kind": "unavailable", "sample": repr(e)}
        print(json.dumps(line))
    for pb in pinned + img_bufs:
        pb.free()
    table.free()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
