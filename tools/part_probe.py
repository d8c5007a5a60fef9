#!/usr/bin/env python
"""one partitioned count scan against a union table (for `ncu --metrics gpu__time_duration.sum`)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import strainer2_b200 as s2
from strainer2_b200 import synth
import bench
n_strains = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n_genomes = int(sys.argv[2]) if len(sys.argv) > 2 else 32           # genomes of 5 Mb per batch (one launch sequence)
strain = bench.make_strain()
rng = synth.rng_for(5, 0)
flat = np.concatenate([synth.contigs_to_flat(synth.genome(rng, 5_000_000, 40)) for _ in range(n_strains - 1)] + [synth.contigs_to_flat(strain)])
ctx = s2.Context(0, batch_bytes=64 << 20, n_lanes=2)
t = s2.StrainTable(ctx, flat, n_cols=2)
batch, bases, lookups = bench.make_batch(strain, 0, n_genomes)
dev = torch.from_numpy(batch).cuda()
ctx.scan_count(t, dev, 1)          # warm-up: allocates the partition pool
ctx.kernel_time(reset=True)
for i in range(5):
    st = ctx.scan_count(t, dev, 1)
ms, n = ctx.kernel_time(reset=True)
print(f"A={os.environ.get('S2_PART_A', '2')} B={os.environ.get('S2_PART_B', '2')} genomes_per_batch={n_genomes} strains={n_strains} probe_bytes={t.probe_bytes} hits={st.hits} valid={st.valid_windows} avg_ms={ms / n:.3f} Glookups/s={lookups / (ms / n) / 1e6:.1f}")
