#!/usr/bin/env python3
"""Can the sketch kernel run in the shadow of the Bloom insert's apply pass?  Two contexts (two streams) on one GPU:
context A repeats a Bloom insert (bin + apply), context B repeats an unfiltered round-0 sketch (its hashing phase is
what would be moved).  Times alone and together."""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ntsynt_b200 import device, synth
mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 3000
A, B = device.Context(0), device.Context(0)
wl = synth.Workload(1, int(mbp * 1e6), 1.0)
ga, gb = wl.materialize(A, 0), wl.materialize(B, 0)
nbytes = device.BloomFilter.size_for(ga.total_bases, 0.025)
bf = A.bloom(nbytes)
A.prof_enable(True)


def inserts(n, out):
    t0 = time.perf_counter()
    for _ in range(n):
        bf.set_genome(ga, 24)
    out["insert_ms"] = (time.perf_counter() - t0) * 1e3 / n


def sketches(n, out):
    t0 = time.perf_counter()
    for _ in range(n):
        B.sketch(gb, 24, 1000).close()
    out["sketch_ms"] = (time.perf_counter() - t0) * 1e3 / n


for f in (inserts, sketches):
    f(2, {})
alone = {}
inserts(6, alone); sketches(12, alone)
print("alone   ", {k: round(v, 2) for k, v in alone.items()})
both = {}
A.prof_reset()
ta = threading.Thread(target=inserts, args=(6, both)); tb = threading.Thread(target=sketches, args=(24, both))
t0 = time.perf_counter(); ta.start(); tb.start(); ta.join(); t_ins = time.perf_counter() - t0; tb.join()
print("together", {k: round(v, 2) for k, v in both.items()}, "(6 inserts took", round(t_ins * 1e3, 1), "ms)")
p = A.prof()
print({k: round(v[0] / 6, 2) for k, v in p.items() if v[0] > 0})
