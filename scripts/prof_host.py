"""cProfile of the host side of one full-size hot-path step (run on the GPU box):
    python scripts/prof_host.py [genome_mbp] > gpurun_out/prof_host.txt"""
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
K, W, G, d = 24, 1000, (int(sys.argv[2]) if len(sys.argv) > 2 else 2), 1.0
ps = bench.presets(d)
ctx = device.Context(0)
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
