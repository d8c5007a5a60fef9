"""strainer2_b200 - B200-native k-mer scan path of strainer2 (kmer_scrub_count / strain_detect).

The product is ``libstrainer2_b200.so`` (hand-written sm_100a CUDA kernels behind the C ABI declared in
``include/strainer2_b200.h``) plus the drop-in executables in ``strainer2_b200/bin``.  This package is
only the ctypes binding used by the tests, ``bench.py`` and ``__graft_entry__``; it contains no
compute and NO fallback: importing it without the built library raises.
"""
from ._lib import lib, LIB_PATH, S2Error            # noqa: F401  (raises if the .so is missing)
from .api import (                                  # noqa: F401
    K, Context, StrainTable, Reader, ScanStats, PinnedBuffer,
    encode_2bit, decode_2bit, kmer_from_ascii, kmer_to_ascii,
    roworder_emulate, format_count_table, load_flat, flatten_records,
    run_kmer_scrub_count, run_strain_detect, run_kmer_scrub_count_batch, run_kmer_scrub_filter, py_float_repr, BIN_DIR,
)
