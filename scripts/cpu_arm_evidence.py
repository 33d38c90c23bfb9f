#!/usr/bin/env python3
"""Evidence for the reference arm's bp/s extrapolation (run on the GPU box's host cores; nothing of the CUDA library is
loaded): the CPU restatement of the reference path (bench.cpu_path: oracle/ C + OpenMP Bloom filter and sketch with the
reference's threading structure, pure-Python graph stage) on 96 / 192 / 384 Mbp samples per genome and, with --full, on
the whole 2 x 3 Gbp workload with its 14.8 GB filter.  One JSON line per run.
    python scripts/cpu_arm_evidence.py [--full] [--samples 96 192 384]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ntsynt_b200 import synth_layout  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--full", action="store_true")
ap.add_argument("--samples", nargs="*", type=float, default=[96, 192, 384])
ap.add_argument("--genome-mbp", type=float, default=3000.0)
a = ap.parse_args()
G, d = 2, 1.0
lay = synth_layout.Layout(G, int(a.genome_mbp * 1e6), d, seed=20260117)
names = [lay.file_name(g) for g in range(G)]
threads = os.cpu_count() or 1
runs = [(s, f"first {s:g} Mbp of each genome") for s in a.samples] + ([(a.genome_mbp, "the whole workload")] if a.full else [])
for smbp, what in runs:
    t0 = time.perf_counter()
    recs = bench.cpu_sample_records(lay, G, smbp)
    t_gen = time.perf_counter() - t0
    dt, tot, text = bench.cpu_path(recs, names, d, threads, log=lambda m: print(f"[{what}] {m}", file=sys.stderr, flush=True))
    print(json.dumps({"sample": what, "bases": tot, "seconds": round(dt, 2), "bp_per_s": round(tot / dt), "cores": threads,
                      "blocks": text.count("\n") // G, "blocks_sha1": bench.sha1_text(text),
                      "generate_s": round(t_gen, 1)}), flush=True)
    del recs
assert "ntsynt_b200._lib" not in sys.modules
