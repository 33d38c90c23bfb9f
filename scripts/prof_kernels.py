#!/usr/bin/env python3
"""Minimal driver for ncu captures of the two dominant kernels at full size:
   materialise G synthetic genomes, one Bloom insert per genome, AND, one round-0 sketch of the first two.
   python scripts/prof_kernels.py [genome_mbp] [genomes] [divergence]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ntsynt_b200 import device, synth, pipeline
mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 3000
G = int(sys.argv[2]) if len(sys.argv) > 2 else 2              # 5 genomes at d = 12: the query-everything sketch (QALL)
d = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
ctx = device.Context(0)
wl = synth.Workload(G, int(mbp * 1e6), d)
gens = [wl.materialize(ctx, g) for g in range(G)]
bf = pipeline.build_common_bf(ctx, gens, [wl.file_name(g) for g in range(G)], 24)
for g in gens[:2]:
    mx = ctx.sketch(g, 24, 1000, common=bf)
    print("minimizers", len(mx))
