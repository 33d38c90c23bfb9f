#!/usr/bin/env python3
"""Bloom insert at full size, split by pass (CUDA events of the library's own profiler), for both ranking kernels:
   python scripts/gpu_insert_split.py [genome_mbp]
NTS_BF_BIN=1 -> bf_bin_kernel (shared-memory atomicAdd ranking), 2 -> bf_rank_bin_kernel (private counters)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from ntsynt_b200 import device, synth
mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 3000
ctx = device.Context(0)
ctx.prof_enable(True)
wl = synth.Workload(2, int(mbp * 1e6), 1.0)
gens = [wl.materialize(ctx, g) for g in range(2)]
# optional second argument: size the filter for a genome of that many Mbp (a contig-sharded rank inserts its share of
# the k-mers into a filter sized for the whole genome)
fmbp = float(sys.argv[2]) if len(sys.argv) > 2 else None
nbytes = device.BloomFilter.size_for(int(fmbp * 1e6) if fmbp else gens[0].total_bases, 0.025)
common, level = ctx.bloom(nbytes), ctx.bloom(nbytes)
pops = {}
for rank in ("1", "1"):
    os.environ["NTS_BF_BIN"] = rank
    for rep in range(2):
        ctx.prof_reset()
        common.build_common(level, gens, 24)
        p = ctx.prof()
    pops.setdefault(rank, common.popcount())
    print(f"NTS_BF_BIN={rank}: " + "  ".join(f"{k} {v[0]:.2f}" for k, v in p.items() if v[0] > 0), flush=True)
assert len(set(pops.values())) == 1, pops
print("popcount", pops)
