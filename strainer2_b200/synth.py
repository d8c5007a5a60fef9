"""Deterministic synthetic genomes / reads for tests and bench.py (SURVEY.md section 8d shapes).

There is no network and no dataset in the image: every benchmark and every large parity test runs
on sequences generated here (numpy PCG64, fixed seeds).  Nothing in this module computes k-mers.
"""
import gzip
import os

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
SEED0 = 0x5712A1E2


def rng_for(config: int, index: int = 0):
    return np.random.Generator(np.random.PCG64(SEED0 + 1000 * config + index))


def random_bases(rng, n: int) -> np.ndarray:
    """uniform i.i.d. ACGT as ASCII bytes"""
    return ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)]


def mutate(seq: np.ndarray, rate: float, rng) -> np.ndarray:
    """per-base substitutions at `rate` (always to a different base)"""
    out = seq.copy()
    m = rng.random(seq.size) < rate
    n = int(m.sum())
    if n:
        idx = np.nonzero(m)[0]
        lut = np.zeros(256, dtype=np.uint8)
        lut[ACGT] = np.arange(4, dtype=np.uint8)
        cur = lut[out[idx]]
        out[idx] = ACGT[(cur + rng.integers(1, 4, size=n, dtype=np.uint8)) % 4]
    return out


def sprinkle(seq: np.ndarray, rate: float, rng, byte=ord("N")) -> np.ndarray:
    out = seq.copy()
    out[rng.random(seq.size) < rate] = byte
    return out


def genome(rng, total: int, n_contigs: int, n_runs: int = 0):
    """list of contigs (uint8 arrays); optional N-runs of length U[1,100]"""
    per = total // n_contigs
    contigs = []
    for _ in range(n_contigs):
        contigs.append(random_bases(rng, per))
    for _ in range(n_runs):
        c = contigs[int(rng.integers(0, n_contigs))]
        ln = int(rng.integers(1, 101))
        st = int(rng.integers(0, max(1, c.size - ln)))
        c[st:st + ln] = ord("N")
    return contigs


def sample_reads(rng, sources, n_reads: int, read_len: int, sub_rate: float = 0.0, n_rate: float = 0.0):
    """reads drawn uniformly from `sources` (list of uint8 arrays), both strands -> (n_reads, read_len) uint8"""
    src = np.concatenate(sources)
    # avoid reads spanning two sources: draw starts inside each source
    sizes = np.array([s.size for s in sources])
    offs = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    ok = sizes >= read_len
    w = np.where(ok, sizes - read_len + 1, 0).astype(np.float64)
    which = rng.choice(len(sources), size=n_reads, p=w / w.sum())
    start = offs[which] + (rng.random(n_reads) * (sizes[which] - read_len + 1)).astype(np.int64)
    idx = start[:, None] + np.arange(read_len)[None, :]
    reads = src[idx]
    rev = rng.random(n_reads) < 0.5
    comp = np.zeros(256, dtype=np.uint8)
    comp[:] = np.arange(256, dtype=np.uint8)
    for a, b in zip(b"ACGTN", b"TGCAN"):
        comp[a] = b
    reads[rev] = comp[reads[rev][:, ::-1]]
    if sub_rate:
        flat = mutate(reads.reshape(-1), sub_rate, rng)
        reads = flat.reshape(n_reads, read_len)
    if n_rate:
        reads = sprinkle(reads.reshape(-1), n_rate, rng).reshape(n_reads, read_len)
    return reads


def reads_to_flat(reads: np.ndarray) -> np.ndarray:
    """(n, L) reads -> flat stream with '\\n' after every read (the device batch format)"""
    n, L = reads.shape
    out = np.full((n, L + 1), ord("\n"), dtype=np.uint8)
    out[:, :L] = reads
    return out.reshape(-1)


def contigs_to_flat(contigs) -> np.ndarray:
    parts = []
    for c in contigs:
        parts.append(np.asarray(c, dtype=np.uint8))
        parts.append(np.array([ord("\n")], dtype=np.uint8))
    return np.concatenate(parts) if parts else np.zeros(0, np.uint8)


def _open(path, gz):
    return gzip.open(path, "wb", compresslevel=6) if gz else open(path, "wb")


def write_fasta(path, records, wrap=80, gz=None, names=None, newline=b"\n"):
    gz = path.endswith(".gz") if gz is None else gz
    with _open(path, gz) as f:
        for i, r in enumerate(records):
            b = bytes(r) if not isinstance(r, np.ndarray) else r.tobytes()
            f.write(b">" + (names[i] if names else b"seq%d" % i) + newline)
            if wrap:
                for j in range(0, len(b), wrap):
                    f.write(b[j:j + wrap] + newline)
            else:
                f.write(b + newline)


def write_fastq(path, records, gz=None, names=None, newline=b"\n"):
    gz = path.endswith(".gz") if gz is None else gz
    with _open(path, gz) as f:
        for i, r in enumerate(records):
            b = bytes(r) if not isinstance(r, np.ndarray) else r.tobytes()
            f.write(b"@" + (names[i] if names else b"read%d" % i) + newline + b + newline + b"+" + newline
                    + b"I" * len(b) + newline)


def write_reads_fastq(path, reads: np.ndarray, gz=None):
    """fast FASTQ writer for (n, L) read matrices"""
    gz = path.endswith(".gz") if gz is None else gz
    n, L = reads.shape
    qual = b"I" * L
    with _open(path, gz) as f:
        step = 20000
        for s in range(0, n, step):
            chunk = reads[s:s + step]
            f.write(b"".join(b"@r%d\n%s\n+\n%s\n" % (s + i, chunk[i].tobytes(), qual) for i in range(chunk.shape[0])))


def bgzf_bytes(data: bytes, block: int = 65280, level: int = 6) -> bytes:
    """BGZF (bgzip) container: independent gzip members of <= 64 KB of text each, sizes in the 'BC' extra
    field, empty EOF member at the end.  A valid multi-member .gz for zlib (what the reference uses); the
    block structure is what lets the hardware decompression engine inflate it in parallel."""
    import struct
    import zlib
    out = []
    for off in list(range(0, len(data), block)) + [None]:
        chunk = b"" if off is None else data[off:off + block]
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        comp = co.compress(chunk) + co.flush()
        bsize = len(comp) + 25                                   # total member size - 1
        out.append(b"\x1f\x8b\x08\x04" + b"\x00" * 4 + b"\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize))
        out.append(comp + struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
    return b"".join(out)


def bgzf_bytes_parallel(data: bytes, block: int = 65280, level: int = 6, threads: int = 16) -> bytes:
    """bgzf_bytes() on a thread pool (zlib releases the GIL): members are independent, so pieces cut at multiples of the
    block size compress separately and concatenate; the empty EOF member comes once, at the end"""
    from concurrent.futures import ThreadPoolExecutor
    piece = block * 256
    eof = bgzf_bytes(b"", block, level)
    parts = [data[o:o + piece] for o in range(0, len(data), piece)]
    with ThreadPoolExecutor(max_workers=threads) as ex:
        out = list(ex.map(lambda d: bgzf_bytes(d, block, level)[:-len(eof)], parts))
    return b"".join(out) + eof


def write_bgzf(path, data: bytes, block: int = 65280, level: int = 6):
    with open(path, "wb") as f:
        f.write(bgzf_bytes(data, block, level))


def fasta_bytes(records, wrap=80) -> bytes:
    out = []
    for i, r in enumerate(records):
        b = bytes(r) if not isinstance(r, np.ndarray) else r.tobytes()
        out.append(b">seq%d\n" % i)
        if wrap:
            out.extend(b[j:j + wrap] + b"\n" for j in range(0, len(b), wrap))
        else:
            out.append(b + b"\n")
    return b"".join(out)


def fastq_bytes(reads: np.ndarray) -> bytes:
    n, L = reads.shape
    qual = b"I" * L
    return b"".join(b"@r%d\n%s\n+\n%s\n" % (i, reads[i].tobytes(), qual) for i in range(n))


def ensure_dir(p):
    os.makedirs(p, exist_ok=True)
    return p
