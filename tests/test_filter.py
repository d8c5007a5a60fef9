"""kmer_scrub_filter: the drop-in executable and its GPU selection against the reference script.

CPU tests: the oracle restatement (oracle/scrub_filter_oracle.py) against the golden vectors that
tests/golden/make_golden_filter.py produced with the UNMODIFIED /root/reference/scripts/kmer_scrub_filter.py;
Python's str(float) as the product prints it; the runs of the executable that end before any selection.
GPU tests: every golden case through strainer2_b200/bin/kmer_scrub_filter (stdout / stderr / exit code bytes), and
randomized tables - heavy ties, duplicates across the 8-bit digits of the radix select - against the oracle."""
import gzip
import json
import os
import random

import numpy as np
import pytest

from oracle import scrub_filter_oracle as fo

HERE = os.path.dirname(os.path.abspath(__file__))
CASES_DIR = os.path.join(HERE, "golden", "cases", "filter")
CASES = json.load(open(os.path.join(CASES_DIR, "cases.json")))
NO_GPU_NEEDED = ["no_input", "only_header", "multi_bad", "drug_too_few"]       # end before / without a selection


def _expected(name):
    return (CASES[name]["rc"], open(os.path.join(CASES_DIR, "expected_%s.stdout" % name), "rb").read(),
            open(os.path.join(CASES_DIR, "expected_%s.stderr" % name), "rb").read())


def _norm_err(err):
    """an uncaught Python exception: only its last line is the script's own words (as in make_golden_filter.py)"""
    if b"Traceback" in err:
        return b"<traceback>\n" + err.strip().split(b"\n")[-1] + b"\n"
    return err


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_script(name):
    rc, out, err = fo.run(CASES[name]["argv"], cwd=CASES_DIR)
    assert (rc, out, err) == _expected(name)


def test_python_float_repr():
    import strainer2_b200 as s2
    r = random.Random(5)
    xs = [0.0, 1.0, 0.5, 0.1, 0.04, 1 / 3, 2 / 3, 0.6886666666666666, 3000.0, 6698540.0, 1e15, 1e16, 1.5e16, 123456789012345680.0,
          1e-4, 1e-5, 0.00012345, 9.999e-5, 1e22, 1e23, 5e-324, 1.7976931348623157e308, 0.30000000000000004, 100.0, 1e-7, 123.456]
    xs += [r.random() for _ in range(300)] + [r.random() * 10 ** r.randint(-12, 20) for _ in range(300)]
    xs += [1 - (h / 3000.0) for h in range(0, 3000, 7)]
    for x in xs:
        assert s2.py_float_repr(x) == repr(float(x)), x


@pytest.mark.parametrize("name", NO_GPU_NEEDED)
def test_executable_runs_that_select_nothing(name):
    import strainer2_b200 as s2
    p = s2.run_kmer_scrub_filter(CASES[name]["argv"], cwd=CASES_DIR)
    assert (p.returncode, p.stdout, _norm_err(p.stderr)) == _expected(name)


def test_executable_usage_errors():
    import strainer2_b200 as s2
    p = s2.run_kmer_scrub_filter(["-x"], cwd=CASES_DIR)
    assert p.returncode == 2 and p.stdout == b"" and b"usage:" in p.stderr
    p = s2.run_kmer_scrub_filter(["-s", "t_tiny.tsv.gz", "-m", "1.5"], cwd=CASES_DIR)
    assert p.returncode == 1 and p.stdout == b""


# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ctx():
    import strainer2_b200 as s2
    c = s2.Context(0, batch_bytes=1 << 20, n_lanes=1)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_executable_matches_reference_script(name):
    import strainer2_b200 as s2
    p = s2.run_kmer_scrub_filter(CASES[name]["argv"], cwd=CASES_DIR)
    want = _expected(name)
    assert p.returncode == want[0]
    assert _norm_err(p.stderr) == want[2]
    assert p.stdout == want[1]


def _random_table(r, n, style):
    keys = set()
    while len(keys) < n:
        keys.add("".join(r.choice("ACGT") for _ in range(31)))
    keys = list(keys)
    r.shuffle(keys)
    rows = ["#kmer\treference_count\tpangenome_count\tmetagenome_count\tdrug_count\n"]
    for k in keys:
        if style == "ties":                   # a handful of distinct values: the cut falls inside a large tie group
            pan, meta = r.choice([0, 0, 0, 1, 1, 2]), r.choice([0, 0, 1, 2, 3])
        elif style == "wide":                 # values spread over many binades and digits
            pan = r.randint(0, 1) * r.randint(1, 10 ** r.randint(0, 9))
            meta = r.randint(0, 1) * r.randint(1, 10 ** r.randint(0, 12))
        else:                                 # a single heavy column
            pan, meta = 0, r.randint(0, 40)
        row = [k, "1", str(pan), str(meta)]
        if style == "wide":
            row.append(str(r.choice([0, 0, 0, 0, 0, 0, 0, 0, 0, 1])))
        rows.append("\t".join(row) + "\n")
    return "".join(rows)


@pytest.mark.gpu
@pytest.mark.parametrize("style,n,fractions", [("ties", 200_000, ["0.04", "0.5", "0.73", "0.999"]), ("wide", 150_000, ["0.01", "0.2", "0.35"]),
                                               ("meta_only", 120_000, ["0.1", "0.9"])])
def test_executable_matches_oracle_on_random_tables(tmp_path, style, n, fractions):
    import strainer2_b200 as s2
    r = random.Random(hash(style) & 0xFFFF)
    path = os.path.join(str(tmp_path), "t.tsv.gz")
    with gzip.open(path, "wt", compresslevel=1) as f:
        f.write(_random_table(r, n, style))
    for m in fractions:
        for extra in ([], ["-i"]):
            argv = ["-s", "t.tsv.gz", "-m", m] + extra
            rc, out, err = fo.run(argv, cwd=str(tmp_path))
            p = s2.run_kmer_scrub_filter(argv, cwd=str(tmp_path))
            assert p.returncode == rc, (argv, p.stderr[-300:])
            assert _norm_err(p.stderr) == err, argv
            assert p.stdout == out, argv


@pytest.mark.gpu
def test_scrub_joint_abi_keeps_exactly_the_rows_of_a_stable_sort(ctx):
    """s2_scrub_joint against numpy: value = max(pan / pan_sum, meta / meta_sum), stable descending order, top n_scrub alive rows go"""
    import ctypes as C
    from strainer2_b200 import lib
    rng = np.random.default_rng(3)
    n = 1_000_003
    pan = (rng.integers(0, 4, n) * rng.integers(0, 2, n)).astype(np.uint64)
    meta = (rng.integers(0, 1000, n) * (rng.random(n) < 0.2)).astype(np.uint64)
    alive = (rng.random(n) < 0.9).astype(np.uint8)
    ps, ms = int(pan.sum()), int(meta.sum())
    val = np.maximum(np.where(pan > 0, pan / float(ps), 0.0), np.where(meta > 0, meta / float(ms), 0.0))
    order = np.argsort(-val[alive == 1], kind="stable")
    alive_idx = np.flatnonzero(alive)
    u64p, u8p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint8)
    for n_scrub in (0, 1, 1234, int(alive.sum()) // 2, int(alive.sum()) - 1, int(alive.sum())):
        want = alive.copy()
        want[alive_idx[order[:n_scrub]]] = 0
        keep = np.full(n, 7, dtype=np.uint8)
        rc = lib.s2_scrub_joint(ctx.h, pan.ctypes.data_as(u64p), meta.ctypes.data_as(u64p), alive.ctypes.data_as(u8p), n, ps, ms, n_scrub,
                                keep.ctypes.data_as(u8p))
        assert rc == 0
        assert np.array_equal(keep, want), n_scrub
