"""Thin Python mirror of the C ABI (include/strainer2_b200.h).  No compute happens here.

Names follow the reference's domain: a *strain table* built from the ``-r`` genome
(GEN_hash_sequences_set_count_vec, /root/reference/src/genome_compare.c:967-1030), *count scans* of
genome / metagenome batches into counter columns (GEN_calculate_kmer_count, :179-236) and *detect
scans* (quantify_hits_PE pass 1, src/strain_detect.c:465-539).
"""
import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import lib, check, S2Error, ScanStatsStruct

K = 31
BIN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bin")


@dataclass
class ScanStats:
    hits: int
    valid_windows: int


def _buf(x):
    """-> (pointer as int, n_bytes, on_device, keepalive).  Host: bytes / bytearray / numpy uint8.
    Device: anything with .is_cuda/.data_ptr() (a torch uint8 CUDA tensor)."""
    if hasattr(x, "is_cuda"):
        if not x.is_cuda:
            x = x.contiguous().numpy()
        else:
            x = x.contiguous()
            return x.data_ptr(), x.numel() * x.element_size(), 1, x
    if isinstance(x, (bytes, bytearray, memoryview)):
        x = np.frombuffer(x, dtype=np.uint8)
    x = np.ascontiguousarray(x, dtype=np.uint8)
    return x.ctypes.data, x.size, 0, x


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


class Context:
    """One GPU (s2_ctx): streams, pinned batch ring, scratch.  Fails loudly without an sm_100 GPU."""

    def __init__(self, device=0, batch_bytes=0, n_lanes=0):
        self.h = lib.s2_init(device, batch_bytes, n_lanes)
        if not self.h:
            raise S2Error("s2_init: " + _lib.last_error())
        self.device = device

    def close(self):
        if self.h:
            lib.s2_shutdown(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def sm_count(self):
        return lib.s2_ctx_sm_count(self.h)

    # -- count scan ---------------------------------------------------------------------------
    def scan_count(self, table, bases, col):
        p, n, dev, keep = _buf(bases)
        st = ScanStatsStruct()
        check(lib.s2_scan_count(self.h, table.h, p, n, col, dev, C.byref(st)), "s2_scan_count")
        return ScanStats(st.hits, st.valid_windows)

    def scan_count_enqueue(self, table, dev_tensor, col):
        """device-resident batch, no host synchronisation (bench: K launches between two events)"""
        check(lib.s2_scan_count_enqueue(self.h, table.h, dev_tensor.data_ptr(),
                                        dev_tensor.numel() * dev_tensor.element_size(), col), "s2_scan_count_enqueue")

    def scan_count_ptr(self, table, ptr, n_bytes, col, on_device=0):
        st = ScanStatsStruct()
        check(lib.s2_scan_count(self.h, table.h, ptr, n_bytes, col, on_device, C.byref(st)), "s2_scan_count")
        return ScanStats(st.hits, st.valid_windows)

    def event_record(self, which):
        check(lib.s2_event_record(self.h, which), "s2_event_record")

    def event_elapsed_ms(self, a, b):
        ms = C.c_double()
        check(lib.s2_event_elapsed_ms(self.h, a, b, C.byref(ms)), "s2_event_elapsed_ms")
        return ms.value

    def batch_acquire(self):
        cap = C.c_uint64()
        p = lib.s2_batch_acquire(self.h, C.byref(cap))
        if not p:
            raise S2Error("s2_batch_acquire: " + _lib.last_error())
        return p, cap.value

    def batch_submit_count(self, table, ptr, n_bytes, col):
        check(lib.s2_batch_submit_count(self.h, table.h, ptr, n_bytes, col), "s2_batch_submit_count")

    def sync(self):
        st = ScanStatsStruct()
        check(lib.s2_sync(self.h, C.byref(st)), "s2_sync")
        return ScanStats(st.hits, st.valid_windows)

    def ingest_count_file(self, table, path, col):
        """GPU-side ingest of one BGZF / plain strict-FASTQ file -> (rc, bases, lookups); rc 1 = not handled"""
        b, l = C.c_uint64(), C.c_uint64()
        rc = lib.s2_ingest_count_file(self.h, table.h, os.fsencode(path), col, C.byref(b), C.byref(l))
        if rc < 0:
            raise S2Error("s2_ingest_count_file: " + _lib.last_error())
        return rc, b.value, l.value

    def ingest_count_mem(self, table, image, col):
        """the same for a file image in host memory (bytes / numpy uint8 / (pointer, length)) -> (rc, bases, lookups)"""
        if isinstance(image, tuple):
            ptr, n, keep = image[0], image[1], None
        else:
            keep = np.frombuffer(image, dtype=np.uint8) if not isinstance(image, np.ndarray) else np.ascontiguousarray(image, dtype=np.uint8)
            ptr, n = keep.ctypes.data, keep.size
        b, l = C.c_uint64(), C.c_uint64()
        rc = lib.s2_ingest_count_mem(self.h, table.h, ptr, n, col, C.byref(b), C.byref(l))
        if rc < 0:
            raise S2Error("s2_ingest_count_mem: " + _lib.last_error())
        return rc, b.value, l.value

    def ingest_count_mem_batch(self, table, ptrs, sizes, col):
        """many file images (host pointers + lengths) in one call, small files grouped -> (rc_each, bases, lookups)"""
        n = len(ptrs)
        P = (C.c_void_p * n)(*ptrs)
        S = (C.c_uint64 * n)(*sizes)
        R = (C.c_int * n)()
        b, l = C.c_uint64(), C.c_uint64()
        if lib.s2_ingest_count_mem_batch(self.h, table.h, P, S, n, col, R, C.byref(b), C.byref(l)) < 0:
            raise S2Error("s2_ingest_count_mem_batch: " + _lib.last_error())
        return list(R), b.value, l.value

    def ingest_submit_mem_batch(self, table, ptrs, sizes, col):
        """asynchronous form: -> job handle for ingest_wait(); the images must stay valid until then"""
        n = len(ptrs)
        P = (C.c_void_p * n)(*ptrs)
        S = (C.c_uint64 * n)(*sizes)
        h = lib.s2_ingest_submit_mem_batch(self.h, table.h, P, S, n, col)
        if not h:
            raise S2Error("s2_ingest_submit_mem_batch: " + _lib.last_error())
        return (h, n)

    def ingest_wait(self, job):
        """-> (rc_each, bases, lookups) of a submitted job"""
        h, n = job
        R = (C.c_int * n)()
        b, l = C.c_uint64(), C.c_uint64()
        if lib.s2_ingest_wait(h, R, C.byref(b), C.byref(l)) < 0:
            raise S2Error("s2_ingest_wait: " + _lib.last_error())
        return list(R), b.value, l.value

    def ingest_count_files(self, table, paths, col):
        """many files in one call, small files grouped -> (rc_each, bases, lookups)"""
        n = len(paths)
        P = (C.c_char_p * n)(*[os.fsencode(p) for p in paths])
        R = (C.c_int * n)()
        b, l = C.c_uint64(), C.c_uint64()
        if lib.s2_ingest_count_files(self.h, table.h, P, n, col, R, C.byref(b), C.byref(l)) < 0:
            raise S2Error("s2_ingest_count_files: " + _lib.last_error())
        return list(R), b.value, l.value

    def ingest_reset(self):
        """drop the context's ingest pipelines (the next call re-reads S2_INGEST_CHUNK_MB / S2_INGEST_TEXT_MB / S2_INGEST_PIPES)"""
        lib.s2_ingest_reset(self.h)

    def kernel_time(self, reset=False):
        ms, n = C.c_double(), C.c_uint64()
        check(lib.s2_kernel_time(self.h, C.byref(ms), C.byref(n), 1 if reset else 0), "s2_kernel_time")
        return ms.value, n.value

    # -- detect scan --------------------------------------------------------------------------
    def scan_detect(self, table, bases, rec_off, inf_cap=None):
        """-> (read_hits[n_rec], read_inf[n_rec], inf_pos[ascending], stats)"""
        p, n, dev, keep = _buf(bases)
        rec_off = np.ascontiguousarray(rec_off, dtype=np.uint64)
        n_rec = rec_off.size - 1
        hits = np.zeros(max(n_rec, 1), dtype=np.uint32)
        inf = np.zeros(max(n_rec, 1), dtype=np.uint32)
        cap = inf_cap if inf_cap is not None else max(1024, n // 64)
        while True:
            pos = np.zeros(cap, dtype=np.uint64)
            n_inf = C.c_uint64()
            st = ScanStatsStruct()
            check(lib.s2_scan_detect(self.h, table.h, p, n, _ptr(rec_off, C.c_uint64), n_rec,
                                     _ptr(hits, C.c_uint32), _ptr(inf, C.c_uint32), _ptr(pos, C.c_uint64),
                                     cap, C.byref(n_inf), dev, C.byref(st)), "s2_scan_detect")
            if n_inf.value <= cap:
                break
            cap = n_inf.value
        return hits[:n_rec], inf[:n_rec], pos[:n_inf.value], ScanStats(st.hits, st.valid_windows)

    def pack_2bit(self, bases):
        p, n, dev, keep = _buf(bases)
        nch = (n + 15) // 16
        words = np.zeros(nch, dtype=np.uint32)
        masks = np.zeros(nch, dtype=np.uint16)
        check(lib.s2_pack_2bit(self.h, p, n, dev, _ptr(words, C.c_uint32), _ptr(masks, C.c_uint16)), "s2_pack_2bit")
        return words, masks


class StrainTable:
    """Device-resident table of the strain's canonical 31-mers (the BIO_hash replacement)."""

    def __init__(self, ctx, bases, n_cols=4, load_factor=0.0):
        p, n, dev, keep = _buf(bases)
        self.ctx = ctx
        self.h = lib.s2_table_build(ctx.h, p, n, n_cols, load_factor, dev)
        if not self.h:
            raise S2Error("s2_table_build: " + _lib.last_error())
        self.n_cols = n_cols

    def free(self):
        if self.h:
            lib.s2_table_free(self.h)
            self.h = None

    @property
    def n_keys(self):
        return lib.s2_table_n_keys(self.h)

    @property
    def n_slots(self):
        return lib.s2_table_n_slots(self.h)

    @property
    def hbm_bytes(self):
        return lib.s2_table_hbm_bytes(self.h)

    @property
    def probe_bytes(self):
        return lib.s2_table_probe_bytes(self.h)

    def export(self, with_pos=False):
        """keys (uint64, A0 C1 G2 T3) and djb2 (uint32) [and first_pos (uint32)], all in first-occurrence order."""
        n = self.n_keys
        keys = np.zeros(n, dtype=np.uint64)
        djb2 = np.zeros(n, dtype=np.uint32)
        pos = np.zeros(n, dtype=np.uint32)
        check(lib.s2_table_export(self.h, _ptr(keys, C.c_uint64), _ptr(djb2, C.c_uint32), _ptr(pos, C.c_uint32)),
              "s2_table_export")
        return (keys, djb2, pos) if with_pos else (keys, djb2)

    def counts(self, col):
        out = np.zeros(self.n_keys, dtype=np.uint32)
        check(lib.s2_table_counts_fetch(self.h, col, _ptr(out, C.c_uint32)), "s2_table_counts_fetch")
        return out

    def set_counts(self, col, values):
        v = np.ascontiguousarray(values, dtype=np.uint32)
        assert v.size == self.n_keys
        check(lib.s2_table_counts_store(self.h, col, _ptr(v, C.c_uint32)), "s2_table_counts_store")

    def clear_counts(self, col):
        check(lib.s2_table_counts_clear(self.h, col), "s2_table_counts_clear")

    def gather_counts_dev(self, col, dev_tensor):
        check(lib.s2_table_counts_gather_dev(self.h, col, dev_tensor.data_ptr()), "s2_table_counts_gather_dev")

    def scatter_counts_dev(self, col, dev_tensor):
        check(lib.s2_table_counts_scatter_dev(self.h, col, dev_tensor.data_ptr()), "s2_table_counts_scatter_dev")

    def format_to(self, path, order, n_print_cols):
        """print_hash_counts formatted on the device (s2_table_format) into `path`"""
        order = np.ascontiguousarray(order, dtype=np.uint32)
        fp = _libc.fopen(os.fsencode(path), b"w")
        if not fp:
            raise OSError(f"cannot open {path}")
        try:
            check(lib.s2_table_format(self.h, _ptr(order, C.c_uint32), n_print_cols, fp), "s2_table_format")
        finally:
            _libc.fclose(fp)

    def flag(self, kmers):
        k = np.ascontiguousarray(kmers, dtype=np.uint64)
        found = np.zeros(max(k.size, 1), dtype=np.uint8)
        check(lib.s2_table_flag(self.h, _ptr(k, C.c_uint64), k.size, _ptr(found, C.c_uint8)), "s2_table_flag")
        return found[:k.size].astype(bool)

    def lookup(self, kmers):
        k = np.ascontiguousarray(kmers, dtype=np.uint64)
        out = np.zeros(max(k.size, 1), dtype=np.uint32)
        check(lib.s2_table_lookup(self.h, _ptr(k, C.c_uint64), k.size, _ptr(out, C.c_uint32)), "s2_table_lookup")
        return out[:k.size]


class Reader:
    """FASTA/FASTQ (plain or gzip) records with the reference parser's semantics (src/kseq.h:171-211)."""

    def __init__(self, path):
        self.h = lib.s2_reader_open(os.fsencode(path))
        if not self.h:
            raise S2Error(_lib.last_error())

    def next(self):
        """-> (ret, seq bytes).  ret = length, -1 (EOF) or -2 (truncated quality)."""
        s = C.c_char_p()
        ret = lib.s2_reader_next(self.h, C.byref(s))
        n = lib.s2_reader_len(self.h)
        seq = C.string_at(s, n) if s else b""
        return ret, seq

    @property
    def damaged(self):
        """True once zlib has reported corrupt gzip data (the reference never returns from such a file)"""
        return bool(lib.s2_reader_damaged(self.h))

    def close(self):
        if self.h:
            lib.s2_reader_close(self.h)
            self.h = None


def load_flat(path):
    """All records of a file as the flat stream s2_table_build / s2_scan_count take."""
    r = Reader(path)
    parts = []
    while True:
        ret, seq = r.next()
        if ret < 0:
            break
        parts.append(seq)
    damaged = r.damaged
    r.close()
    if damaged:
        raise S2Error("damaged gzip data in %s" % path)
    return flatten_records(parts)[0]


def flatten_records(records):
    """records (iterable of bytes) -> (flat uint8 array with a '\\n' after every record, rec_off uint64[n+1])"""
    off = [0]
    for s in records:
        off.append(off[-1] + len(s) + 1)
    flat = np.frombuffer(b"".join(s + b"\n" for s in records), dtype=np.uint8).copy() if records else np.zeros(0, np.uint8)
    return flat, np.asarray(off, dtype=np.uint64)


class PinnedBuffer:
    """page-locked host memory (cudaHostAlloc) exposed as a numpy uint8 array"""

    def __init__(self, n_bytes):
        self.ptr = lib.s2_pinned_alloc(n_bytes)
        if not self.ptr:
            raise S2Error("s2_pinned_alloc: " + _lib.last_error())
        self.n = n_bytes
        self.array = np.ctypeslib.as_array((C.c_uint8 * n_bytes).from_address(self.ptr))

    def free(self):
        if self.ptr:
            self.array = None
            lib.s2_pinned_free(self.ptr)
            self.ptr = None


# ---- codecs / host helpers --------------------------------------------------------------------
def encode_2bit(dna: bytes) -> int:
    return lib.s2_encode_2bit(dna, len(dna))


def decode_2bit(v: int, length: int) -> bytes:
    out = C.create_string_buffer(33)
    lib.s2_decode_2bit(v, length, out)
    return out.value


def kmer_from_ascii(s: bytes):
    out = C.c_uint64()
    if lib.s2_kmer_from_ascii(s, C.byref(out)) != 0:
        return None
    return out.value


def kmer_to_ascii(k: int) -> bytes:
    out = C.create_string_buffer(32)
    lib.s2_kmer_to_ascii(int(k), out)
    return out.value


def roworder_emulate(djb2, initial_capacity=0):
    d = np.ascontiguousarray(djb2, dtype=np.uint32)
    order = np.zeros(max(d.size, 1), dtype=np.uint32)
    cap = C.c_uint32()
    check(lib.s2_roworder_emulate(_ptr(d, C.c_uint32), d.size, initial_capacity, _ptr(order, C.c_uint32),
                                  C.byref(cap)), "s2_roworder_emulate")
    return order[:d.size], cap.value


_libc = C.CDLL(None)
_libc.fopen.restype = C.c_void_p
_libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
_libc.fclose.argtypes = [C.c_void_p]


def format_count_table(path, keys, order, cols, n_threads=4):
    """print_hash_counts (src/kmer_scrub_count.c:134-156) into `path`; cols = list of 3 or 4 uint32 arrays."""
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    order = np.ascontiguousarray(order, dtype=np.uint32)
    cols = [np.ascontiguousarray(c, dtype=np.uint32) for c in cols]
    arr = (C.POINTER(C.c_uint32) * 4)()
    for i, c in enumerate(cols):
        arr[i] = _ptr(c, C.c_uint32)
    fp = _libc.fopen(os.fsencode(path), b"w")
    if not fp:
        raise OSError(f"cannot open {path}")
    try:
        check(lib.s2_format_count_table(fp, _ptr(keys, C.c_uint64), _ptr(order, C.c_uint32), keys.size, arr,
                                        len(cols), n_threads), "s2_format_count_table")
    finally:
        _libc.fclose(fp)


# ---- the drop-in executables ------------------------------------------------------------------
def _run(exe, args, cwd=None, env=None, timeout=None):
    path = os.path.join(BIN_DIR, exe)
    if not os.path.exists(path):
        raise S2Error(f"{path} is missing - run __graft_entry__.build()")
    e = dict(os.environ)
    if env:
        e.update(env)
    return subprocess.run([path] + list(args), cwd=cwd, env=e, capture_output=True, timeout=timeout)


def run_kmer_scrub_count(args, cwd=None, env=None, timeout=None):
    """Run strainer2_b200/bin/kmer_scrub_count with reference-compatible argv; -> CompletedProcess."""
    return _run("kmer_scrub_count", args, cwd, env, timeout)


def run_strain_detect(args, cwd=None, env=None, timeout=None):
    return _run("strain_detect", args, cwd, env, timeout)


def run_kmer_scrub_filter(args, cwd=None, env=None, timeout=None):
    """Run strainer2_b200/bin/kmer_scrub_filter with the argv of scripts/kmer_scrub_filter.py; -> CompletedProcess."""
    return _run("kmer_scrub_filter", args, cwd, env, timeout)


def py_float_repr(x: float) -> str:
    buf = C.create_string_buffer(40)
    lib.s2_py_float_repr(float(x), buf)
    return buf.value.decode()


def run_kmer_scrub_count_batch(args, cwd=None, env=None, timeout=None):
    """many strains against the same lists in one pass (-R strains.txt -A -B [-C] -O outdir)"""
    return _run("kmer_scrub_count_batch", args, cwd, env, timeout)
