"""world_size-2 test of the multi-GPU host logic on CPU (gloo): file sharding + the one all-reduce of the
dense count vectors.  The per-rank vectors are produced by the oracle here (no GPU in this container);
on the GPU the same vectors come from s2_table_counts_gather_dev."""
import os
import subprocess
import sys
import textwrap

import numpy as np

import oracle_util as ou

ROOT = ou.ROOT

WORKER = textwrap.dedent('''
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, {root!r})
    sys.path.insert(0, os.path.join({root!r}, "tests"))
    import oracle_util as ou
    from strainer2_b200 import multigpu

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = {case!r}
    files = [l.strip() for l in open(os.path.join(d, "listA.txt")) if l.strip()]
    sizes = [os.path.getsize(os.path.join(d, f)) for f in files]
    plan = multigpu.shard_files(files, sizes, world)
    t = ou.OracleTable(4)
    t.build(os.path.join(d, "ref.fa.gz"))
    for i in plan[rank]:
        t.count_file(os.path.join(d, files[i]), 1)
    kmers, vals = ou.parse_table(t.table_text(False, os.path.join({tmp!r}, "r%d.tsv" % rank)))
    vec = torch.from_numpy(multigpu.u32_as_i32(vals[:, 1].copy()))
    # push one counter over 2^31 on rank 0 and near 2^32 overall to exercise the wrap-around claim
    if rank == 0: vec[0] += np.int32(-2**31)
    else: vec[0] += np.int32(2**31 - 5)
    multigpu.allreduce_counts_(vec)
    if rank == 0:
        np.save(os.path.join({tmp!r}, "sum.npy"), multigpu.i32_as_u32(vec.numpy()))
        open(os.path.join({tmp!r}, "plan.txt"), "w").write(repr(plan))
    dist.destroy_process_group()
''')


def test_two_ranks_sum_to_the_single_process_result(golden_dir, tmp_path):
    case = os.path.join(golden_dir, "count_edge")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, case=case, tmp=str(tmp_path)))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env))
    assert all(p.wait(timeout=300) == 0 for p in procs)
    got = np.load(tmp_path / "sum.npy")
    want_k, want_v = ou.parse_table(open(os.path.join(case, "expected_AB.tsv"), "rb").read())
    want = want_v[:, 1].copy()
    want[0] = (int(want[0]) + 2**31 + 2**31 - 5) % 2**32          # the injected wrap-around
    assert np.array_equal(got, want)
    plan = eval(open(tmp_path / "plan.txt").read())
    assert sorted(plan[0] + plan[1]) == list(range(7)) and plan[0] and plan[1]


def test_shard_plan_is_balanced_and_deterministic():
    from strainer2_b200 import multigpu
    sizes = [5, 1, 9, 3, 7, 7, 2, 8]
    plan = multigpu.shard_files([str(i) for i in range(8)], sizes, 3)
    assert plan == multigpu.shard_files([str(i) for i in range(8)], sizes, 3)
    loads = [sum(sizes[i] for i in p) for p in plan]
    assert max(loads) - min(loads) <= max(sizes)
    assert sorted(sum(plan, [])) == list(range(8))
