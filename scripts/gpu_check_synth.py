#!/usr/bin/env python3
"synthetic workload sanity + first timing at a given size"
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ntsynt_b200 import device, synth, pipeline
from ntsynt_b200.synteny import SyntenyEngine

mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 100
G = int(sys.argv[2]) if len(sys.argv) > 2 else 2
d = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
k, w = 24, 1000
ctx = device.Context(0)
ctx.prof_enable(True)
t = time.time()
wl = synth.Workload(G, int(mbp * 1e6), d)
gens = [wl.materialize(ctx, g) for g in range(G)]
print("materialize s", time.time() - t, [g.total_bases for g in gens], ctx.prof()["synth"])
seq = gens[0].contig_ascii(0, 0, 200000)
print("base comp", {c: seq.count(c.encode()) for c in "ACGTN"})
names = [pipeline.tsv_name(wl.file_name(g), k, w) for g in range(G)]
order = pipeline.processing_order(names)
for rep in range(2):
    ctx.prof_reset()
    t0 = time.time()
    bf = pipeline.build_common_bf(ctx, gens, [wl.file_name(g) for g in range(G)], k)
    t1 = time.time()
    be = pipeline.CudaBackend(ctx, [gens[i] for i in order], [names[i] for i in order], [wl.names] * G,
                              [[int(x) for x in gens[i].lengths] for i in order], k, common=bf)
    eng = SyntenyEngine(be, k, w, [250, 100], 50000, "100000", 1000, write_files=False, quiet=True)
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable()
    out = eng.run()
    pr.disable()
    t2 = time.time()
    print(f"rep{rep}: bf {t1-t0:.3f}s graph+sketch {t2-t1:.3f}s total {t2-t0:.3f}s  bp/s {sum(g.total_bases for g in gens)/(t2-t0):.3e}")
    print("  stats", eng.stats, be.timing, "blocks", out.count("\n") // G)
    print("  prof", {k_: (round(v[0], 3), v[2]) for k_, v in ctx.prof().items() if v[2]})
    if rep == 1:
        pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
    bf.close()
print(out[:600])
