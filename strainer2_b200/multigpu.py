"""File-sharded multi-GPU plumbing for the count path.

One process per GPU (torch.distributed).  Input files are dealt to ranks, every rank scans its own
shard into its replica of the strain table, and the per-rank counter columns are summed with ONE
all-reduce over the dense first-occurrence-order vector.  Nothing here touches k-mers: the vectors come
from s2_table_counts_gather_dev / go back through s2_table_counts_scatter_dev.

The reference has no counterpart (it is single-threaded; README.md:47 suggests one process per strain).
"""
import numpy as np


def shard_files(paths, sizes, world_size):
    """greedy longest-first deal of files to ranks; returns a list of lists of indices into `paths`.
    Deterministic, so every rank computes the same plan without communicating."""
    order = sorted(range(len(paths)), key=lambda i: (-int(sizes[i]), i))
    load = [0] * world_size
    plan = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        plan[r].append(i)
        load[r] += int(sizes[i])
    for p in plan:
        p.sort()                       # each rank walks its shard in list order
    return plan


def allreduce_counts_(vec, dist=None):
    """in-place SUM all-reduce of a dense uint32 counter vector held as an int32 torch tensor (CPU/gloo in
    the tests, CUDA/NCCL in bench.py).  Two's-complement int32 addition is the reference's `unsigned int`
    wrap-around addition bit for bit, and it is associative, so the result does not depend on sharding."""
    import torch
    assert vec.dtype == torch.int32
    if dist is None:
        import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM)
        # NCCL runs on torch's own stream and only torch's CURRENT stream waits for it; the library's gather / scatter
        # kernels run on the context's streams, which know nothing of either: wait here, on the host, before the
        # caller scatters the sums (without this the scatter raced the collective - bench.py's allreduce_sum_check
        # caught it at N = 2, profiles/r2u_bench_n2.json)
        if vec.is_cuda:
            torch.cuda.current_stream(vec.device).synchronize()
    return vec


def u32_as_i32(a: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint32).view(np.int32)


def i32_as_u32(a: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32).view(np.uint32)
