"""ctypes declarations for every symbol of include/strainer2_b200.h (kept in the same order)."""
import ctypes as C
import os

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libstrainer2_b200.so")


class S2Error(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(or `make -C strainer2_b200/csrc`).  There is no CPU fallback for the k-mer scan path.")

lib = C.CDLL(LIB_PATH)

c_u8p = C.POINTER(C.c_uint8)
c_u16p = C.POINTER(C.c_uint16)
c_u32p = C.POINTER(C.c_uint32)
c_u64p = C.POINTER(C.c_uint64)


class ScanStatsStruct(C.Structure):
    _fields_ = [("hits", C.c_uint64), ("valid_windows", C.c_uint64)]


# name -> (restype, argtypes); tests/test_abi.py checks this list against the header
SIGNATURES = {
    "s2_abi_version": (C.c_int, []),
    "s2_last_error": (C.c_char_p, []),
    "s2_device_count": (C.c_int, []),
    "s2_init": (C.c_void_p, [C.c_int, C.c_uint64, C.c_int]),
    "s2_shutdown": (None, [C.c_void_p]),
    "s2_ctx_device": (C.c_int, [C.c_void_p]),
    "s2_ctx_sm_count": (C.c_int, [C.c_void_p]),
    "s2_table_build": (C.c_void_p, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_double, C.c_int]),
    "s2_table_free": (None, [C.c_void_p]),
    "s2_table_n_keys": (C.c_uint64, [C.c_void_p]),
    "s2_table_n_slots": (C.c_uint64, [C.c_void_p]),
    "s2_table_hbm_bytes": (C.c_uint64, [C.c_void_p]),
    "s2_table_probe_bytes": (C.c_uint64, [C.c_void_p]),
    "s2_table_export": (C.c_int, [C.c_void_p, c_u64p, c_u32p, c_u32p]),
    "s2_table_counts_fetch": (C.c_int, [C.c_void_p, C.c_int, c_u32p]),
    "s2_table_counts_store": (C.c_int, [C.c_void_p, C.c_int, c_u32p]),
    "s2_table_counts_clear": (C.c_int, [C.c_void_p, C.c_int]),
    "s2_table_counts_gather_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "s2_table_counts_scatter_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "s2_tables_allreduce": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int]),
    "s2_table_flag": (C.c_int, [C.c_void_p, c_u64p, C.c_uint64, c_u8p]),
    "s2_table_counts_by_key": (C.c_int, [C.c_void_p, C.c_int, c_u64p, C.c_uint64, c_u32p]),
    "s2_table_unflag": (C.c_int, [C.c_void_p, c_u64p, C.c_uint64]),
    "s2_table_lookup": (C.c_int, [C.c_void_p, c_u64p, C.c_uint64, c_u32p]),
    "s2_scan_count": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int,
                                C.POINTER(ScanStatsStruct)]),
    "s2_scan_count_enqueue": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]),
    "s2_pinned_alloc": (C.c_void_p, [C.c_uint64]),
    "s2_pinned_free": (None, [C.c_void_p]),
    "s2_event_record": (C.c_int, [C.c_void_p, C.c_int]),
    "s2_event_elapsed_ms": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "s2_batch_acquire": (C.c_void_p, [C.c_void_p, c_u64p]),
    "s2_batch_submit_count": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]),
    "s2_batch_release": (C.c_int, [C.c_void_p, C.c_void_p]),
    "s2_sync": (C.c_int, [C.c_void_p, C.POINTER(ScanStatsStruct)]),
    "s2_ingest_count_file": (C.c_int, [C.c_void_p, C.c_void_p, C.c_char_p, C.c_int, c_u64p, c_u64p]),
    "s2_ingest_count_mem": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, c_u64p, c_u64p]),
    "s2_ingest_count_mem_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), c_u64p, C.c_int, C.c_int,
                                            C.POINTER(C.c_int), c_u64p, c_u64p]),
    "s2_ingest_count_files": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_char_p), C.c_int, C.c_int,
                                        C.POINTER(C.c_int), c_u64p, c_u64p]),
    "s2_ingest_submit_mem_batch": (C.c_void_p, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), c_u64p, C.c_int, C.c_int]),
    "s2_ingest_submit_files": (C.c_void_p, [C.c_void_p, C.c_void_p, C.POINTER(C.c_char_p), C.c_int, C.c_int]),
    "s2_ingest_wait": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), c_u64p, c_u64p]),
    "s2_ingest_detect_file": (C.c_int, [C.c_void_p, C.c_void_p, C.c_char_p, C.c_void_p]),
    "s2_ingest_detect_free": (None, [C.c_void_p]),
    "s2_ingest_thread_cleanup": (None, []),
    "s2_ingest_warm": (C.c_int, [C.c_void_p, C.c_int]),
    "s2_ingest_warm_gz": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64]),
    "s2_ingest_reset": (None, [C.c_void_p]),
    "s2_ingest_engine_failed": (C.c_int, []),
    "s2_gz_writer_open": (C.c_void_p, [C.c_char_p, C.c_int]),
    "s2_gz_writer_write": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "s2_gz_writer_close": (C.c_int, [C.c_void_p]),
    "s2_scrub_joint": (C.c_int, [C.c_void_p, c_u64p, c_u64p, C.POINTER(C.c_uint8), C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64,
                                 C.POINTER(C.c_uint8)]),
    "s2_scrub_histogram": (C.c_int, [C.c_void_p, c_u64p, C.c_uint64, c_u64p]),
    "s2_scrub_count_above": (C.c_int, [C.c_void_p, c_u64p, C.c_uint64, C.c_uint64, c_u64p]),
    "s2_py_float_repr": (None, [C.c_double, C.c_char_p]),
    "s2_kmer_scrub_filter_main": (C.c_int, [C.c_int, C.POINTER(C.c_char_p)]),
    "s2_scan_detect": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, c_u64p, C.c_uint32, c_u32p,
                                 c_u32p, c_u64p, C.c_uint64, c_u64p, C.c_int, C.POINTER(ScanStatsStruct)]),
    "s2_kernel_time": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), c_u64p, C.c_int]),
    "s2_tune_scan_variant": (C.c_int, [C.c_void_p, C.c_int]),
    "s2_tune_scan_variant_name": (C.c_char_p, [C.c_int]),
    "s2_encode_2bit": (C.c_uint64, [C.c_char_p, C.c_int]),
    "s2_decode_2bit": (None, [C.c_uint64, C.c_int, C.c_char_p]),
    "s2_pack_2bit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, c_u32p, c_u16p]),
    "s2_kmer_from_ascii": (C.c_int, [C.c_char_p, c_u64p]),
    "s2_kmer_to_ascii": (None, [C.c_uint64, C.c_char_p]),
    "s2_roworder_emulate": (C.c_int, [c_u32p, C.c_uint64, C.c_uint32, c_u32p, c_u32p]),
    "s2_format_count_table": (C.c_int, [C.c_void_p, c_u64p, c_u32p, C.c_uint64, C.POINTER(c_u32p), C.c_int, C.c_int]),
    "s2_table_format": (C.c_int, [C.c_void_p, c_u32p, C.c_int, C.c_void_p]),
    "s2_reader_open": (C.c_void_p, [C.c_char_p]),
    "s2_reader_next": (C.c_int64, [C.c_void_p, C.POINTER(C.c_char_p)]),
    "s2_reader_len": (C.c_uint64, [C.c_void_p]),
    "s2_reader_damaged": (C.c_int, [C.c_void_p]),
    "s2_reader_close": (None, [C.c_void_p]),
    "s2_kmer_scrub_count_main": (C.c_int, [C.c_int, C.POINTER(C.c_char_p)]),
    "s2_strain_detect_main": (C.c_int, [C.c_int, C.POINTER(C.c_char_p)]),
    "s2_kmer_scrub_count_batch_main": (C.c_int, [C.c_int, C.POINTER(C.c_char_p)]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _f = getattr(lib, _name)          # AttributeError here = the library does not export the symbol
    _f.restype = _res
    _f.argtypes = _args


def last_error() -> str:
    return (lib.s2_last_error() or b"").decode("utf-8", "replace")


def check(rc, what=""):
    if rc != 0:
        raise S2Error(f"{what}: {last_error()}")
