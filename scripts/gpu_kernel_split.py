#!/usr/bin/env python3
"kernel-level experiments at full size: sketch with / without Bloom query, insert, fill, combine"
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ntsynt_b200 import device, synth, pipeline
mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 3000
ctx = device.Context(0)
ctx.prof_enable(True)
wl = synth.Workload(2, int(mbp * 1e6), 1.0)
gens = [wl.materialize(ctx, g) for g in range(2)]
bf = pipeline.build_common_bf(ctx, gens, [wl.file_name(g) for g in range(2)], 24)
print("common fpr", bf.fpr())
for label, kw in (("with BF w=1000", dict(common=bf)), ("no BF w=1000", dict()), ("with BF w=250", dict(common=bf)),
                  ("no BF w=250", dict())):
    w = 1000 if "1000" in label else 250
    for rep in range(2):
        ctx.prof_reset()
        mx = ctx.sketch(gens[0], 24, w, **kw)
        p = ctx.prof()
        n = len(mx)
        mx.close()
    print(f"{label}: sketch {p['sketch'][0]:.2f} ms  post {p['sketch_post'][0]:.2f} ms  minimizers {n}")
