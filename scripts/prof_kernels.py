#!/usr/bin/env python3
"""Minimal driver for ncu captures of the two dominant kernels at full size:
   materialise G synthetic genomes, one Bloom insert per genome, AND, one round-0 sketch per genome."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ntsynt_b200 import device, synth, pipeline
mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 3000
ctx = device.Context(0)
wl = synth.Workload(2, int(mbp * 1e6), 1.0)
gens = [wl.materialize(ctx, g) for g in range(2)]
bf = pipeline.build_common_bf(ctx, gens, [wl.file_name(g) for g in range(2)], 24)
for g in gens:
    mx = ctx.sketch(g, 24, 1000, common=bf)
    print("minimizers", len(mx))
