"""sketch_sparse_kernel time against the candidate density (NTS_SKETCH_LAMBDA) at full size"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ntsynt_b200 import device, synth, pipeline
ctx = device.Context(0)
wl = synth.Workload(2, 3_000_000_000, 1.0)
gens = [wl.materialize(ctx, g) for g in range(2)]
bf = pipeline.build_common_bf(ctx, gens, [wl.file_name(g) for g in range(2)], 24)
ref = None
for lam in (32, 24, 20, 16, 12):
    os.environ["NTS_SKETCH_LAMBDA"] = str(lam)
    best = 1e9
    for _ in range(3):
        e0 = ctx.sketch_escalated
        ctx.sync(); t0 = time.perf_counter()
        mx = ctx.sketch(gens[0], 24, 1000, common=bf)
        ctx.sync(); best = min(best, time.perf_counter() - t0)
        esc = ctx.sketch_escalated - e0
        n = len(mx)
        if ref is None:
            ref = [x.copy() for x in mx.to_numpy()]
        mx.close()
    print(f"lambda {lam}: {best * 1e3:.2f} ms, {n} minimizers, {esc} dense sub-tiles escalated")
