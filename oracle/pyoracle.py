"""TEST INFRASTRUCTURE ONLY: Python access to the CPU oracle (oracle/liboracle.so, oracle/oracle_cli)
and, where it was built, the compiled reference (oracle/_ref).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module; the product package
(strainer2_b200/) never does."""
import ctypes as C
import gzip
import os
import re
import subprocess

import numpy as np

ORACLE_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(ORACLE_DIR)
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so", "oracle_cli"])
        L = C.CDLL(path)
        L.s2o_djb2.restype = C.c_uint32
        L.s2o_djb2.argtypes = [C.c_char_p]
        L.s2o_complement.restype = C.c_int
        L.s2o_complement.argtypes = [C.c_int]
        L.s2o_orient.restype = C.c_void_p
        L.s2o_orient.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        L.s2o_encode_2bit.restype = C.c_uint64
        L.s2o_encode_2bit.argtypes = [C.c_char_p, C.c_int]
        L.s2o_decode_2bit.argtypes = [C.c_uint64, C.c_int, C.c_char_p]
        L.s2o_table_new.restype = C.c_void_p
        L.s2o_table_new.argtypes = [C.c_uint, C.c_int]
        L.s2o_table_free.argtypes = [C.c_void_p]
        L.s2o_table_size.restype = C.c_uint
        L.s2o_table_size.argtypes = [C.c_void_p]
        L.s2o_table_capacity.restype = C.c_uint
        L.s2o_table_capacity.argtypes = [C.c_void_p]
        L.s2o_build.restype = C.c_int
        L.s2o_build.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int]
        L.s2o_count_file.restype = C.c_int
        L.s2o_count_file.argtypes = [C.c_void_p, C.c_char_p, C.c_uint, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.s2o_print_counts.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def orient(window: bytes) -> bytes:
    """canonical spelling of a 31-byte window by the reference's rule (orient_string)"""
    L = lib()
    scratch = C.create_string_buffer(64)
    w = C.create_string_buffer(window, len(window) + 1)
    p = L.s2o_orient(w, scratch, len(window))
    return C.string_at(p, len(window))


_libc = C.CDLL(None)
_libc.fopen.restype = C.c_void_p
_libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
_libc.fclose.argtypes = [C.c_void_p]


class OracleTable:
    """the oracle's string table (BIO_hash restatement) driven from Python"""

    def __init__(self, vec_size=4, capacity=8000000):
        self.L = lib()
        self.h = self.L.s2o_table_new(capacity, vec_size)

    def build(self, ref_file, default=1, incr=1, idx=0):
        assert self.L.s2o_build(self.h, os.fsencode(ref_file), default, incr, idx) == 0

    def count_file(self, path, col):
        nb, nw = C.c_uint64(0), C.c_uint64(0)
        assert self.L.s2o_count_file(self.h, os.fsencode(path), col, C.byref(nb), C.byref(nw)) == 0
        return nb.value, nw.value

    @property
    def size(self):
        return self.L.s2o_table_size(self.h)

    def table_text(self, with_C, tmp_path):
        fp = _libc.fopen(os.fsencode(tmp_path), b"w")
        self.L.s2o_print_counts(self.h, 1 if with_C else 0, fp)
        _libc.fclose(fp)
        return open(tmp_path, "rb").read()

    def free(self):
        if self.h:
            self.L.s2o_table_free(self.h)
            self.h = None


def parse_table(text: bytes):
    """count table bytes -> (list of kmer bytes in row order, uint32 array [n, ncols])"""
    rows = [l.split(b"\t") for l in text.split(b"\n")[1:] if l]
    kmers = [r[0] for r in rows]
    vals = np.array([[int(x) & 0xFFFFFFFF for x in r[1:]] for r in rows], dtype=np.uint32) if rows else np.zeros((0, 3), np.uint32)
    return kmers, vals


def oracle_cli(args, cwd=None):
    return subprocess.run([os.path.join(ORACLE_DIR, "oracle_cli")] + list(args), cwd=cwd, capture_output=True)


def ref_run(exe, args, cwd=None):
    return subprocess.run([os.path.join(REF_DIR, exe)] + list(args), cwd=cwd, capture_output=True)


def have_ref():
    return os.path.exists(os.path.join(REF_DIR, "kmer_scrub_count"))


def mask_progress(text: str) -> str:
    lines = text.splitlines()
    if not lines:
        return ""
    return "\n".join([lines[0]] + [re.sub(r"\t.*$", "\t<T>", l) for l in lines[1:]]) + "\n"


def gunzip(path) -> bytes:
    with gzip.open(path, "rb") as f:
        return f.read()
