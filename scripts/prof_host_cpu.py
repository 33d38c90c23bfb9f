#!/usr/bin/env python3
"""cProfile of the host side of the graph stage WITHOUT a GPU: seeded rearranged genomes, sketches from the CPU oracle,
the join restated in numpy in the device-resident form (tests/backends.py).  Per-call overheads are the same as on the
GPU box; array sizes are ~100x smaller.   python scripts/prof_host_cpu.py [G] [Mbp] [n_inv]"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth_small  # noqa: E402
from backends import OracleBackend  # noqa: E402
from ntsynt_b200.synteny import SyntenyEngine  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
mbp = float(sys.argv[2]) if len(sys.argv) > 2 else 12.0
n_inv = int(sys.argv[3]) if len(sys.argv) > 3 else 120
wd = f"/tmp/hp/g{G}_{mbp:g}_{n_inv}"
os.makedirs(wd, exist_ok=True)
paths = [os.path.join(wd, f"g{chr(65 + i)}.fa") for i in range(G)]
if not all(os.path.exists(p) for p in paths):
    L = int(mbp * 1e6)
    gens = synth_small.make_genomes(7, G, contig_lens=(L // 2, L // 3, L // 6), sub=0.01, indel=0.00001, n_inv=n_inv,
                                    n_trans=n_inv // 3, n_dup=20, n_nruns=5)
    for p, recs in zip(paths, gens):
        synth_small.write_fasta(p, recs)
k, w = 24, 1000
tsv = [f"{os.path.basename(p)}.k{k}.w{w}.tsv" for p in paths]
order = sorted(range(G), key=lambda i: tsv[i], reverse=True)
be = OracleBackend([paths[i] for i in order], [tsv[i] for i in order], k, lean="dev")
sk = be.sketch
cache = {}


def cached_sketch(a, w_, masks):        # the oracle sketch is not what is being profiled
    key = (a, w_, None if masks is None else tuple((tuple(s.tolist()), tuple(e.tolist())) for s, e in masks))
    if key not in cache:
        cache[key] = sk(a, w_, masks)
    return cache[key]


be.sketch = cached_sketch


def step():
    eng = SyntenyEngine(be, k, w, [250, 100], 50000, "100000", 1000, write_files=False, quiet=True)
    eng.run()
    return eng


eng = step()
t0 = time.perf_counter()
for _ in range(3):
    eng = step()
print("ms per run", (time.perf_counter() - t0) / 3 * 1e3, "blocks", eng.outputs["final"].count("\n") // G, "V", eng.V0)
print({k_[2:]: round(v * 1e3, 1) for k_, v in eng.stats.items() if k_.startswith("t_")})
pr = cProfile.Profile()
pr.enable()
step()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)
