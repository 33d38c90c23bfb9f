"""cProfile of the host side of one full-size hot-path step (run on the GPU box):
    python scripts/prof_host.py [genome_mbp] [genomes] [divergence] > gpurun_out/prof_host.txt"""
import cProfile
import io
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ntsynt_b200 import device, pipeline, synth  # noqa: E402
from ntsynt_b200.synteny import SyntenyEngine  # noqa: E402

mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 3000.0
K, W, G = 24, 1000, (int(sys.argv[2]) if len(sys.argv) > 2 else 2)
d = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
ps = bench.presets(d)
ctx = device.Context(0)
ctx.prof_enable(True)
wl = synth.Workload(G, int(mbp * 1e6), d, seed=20260117)
file_names = [wl.file_name(g) for g in range(G)]
names = [pipeline.tsv_name(f, K, W) for f in file_names]
order = pipeline.processing_order(names)
gens = [wl.materialize(ctx, g) for g in range(G)]
common = pipeline.build_common_bf(ctx, gens, file_names, K)


def step():
    be = pipeline.CudaBackend(ctx, [gens[i] for i in order], [names[i] for i in order], [wl.names] * G,
                              [[int(x) for x in gens[i].lengths] for i in order], K, common=common)
    eng = SyntenyEngine(be, K, W, ps["w_rounds"], ps["indel"], ps["merge"], ps["block_size"], write_files=False, quiet=True)
    text = eng.run()
    be.close()
    return eng


for _ in range(2):
    step()

# where do the numpy calls of the graph stage spend their time?  (caller line, calls, total ms, largest operand)
import collections
import time
import numpy as _np
_acc = collections.defaultdict(lambda: [0, 0.0, 0])


def _wrap(name):
    orig = getattr(_np, name)

    def f(*a, **k):
        t0 = time.perf_counter()
        r = orig(*a, **k)
        fr = sys._getframe(1)
        e = _acc[(name, os.path.basename(fr.f_code.co_filename), fr.f_lineno)]
        e[0] += 1; e[1] += time.perf_counter() - t0
        e[2] = max(e[2], max((x.size for x in a if isinstance(x, _np.ndarray)), default=0))
        return r
    setattr(_np, name, f)
    return orig


_origs = {n: _wrap(n) for n in ("searchsorted", "unique", "union1d", "concatenate", "flatnonzero", "isin", "argsort", "fromiter",
                                "zeros", "repeat", "cumsum", "bincount", "array", "sort", "where")}
eng = step()
for n, o in _origs.items():
    setattr(_np, n, o)
for key, (cnt, tot, size) in sorted(_acc.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{key[0]:14s} {key[1]}:{key[2]:<5d} calls {cnt:5d}  {tot * 1e3:8.2f} ms  max operand {size}")
print({k: (round(v * 1e3, 1) if k.startswith("t_") else v) for k, v in eng.stats.items() if k != "dev_runs"})
ctx.prof_reset()
step()
print("kernel ms per step:", {k: round(v[0], 2) for k, v in ctx.prof().items() if v[0] > 0})
pr = cProfile.Profile()
pr.enable()
eng = step()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(70)
print(s.getvalue())
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(40)
print(s.getvalue())
print({k: (round(v * 1e3, 1) if k.startswith("t_") else v) for k, v in eng.stats.items()})
