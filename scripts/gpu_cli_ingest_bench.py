#!/usr/bin/env python3
"""bin/ntSynt from FASTA FILES at scale (the ingest pipeline in front of the path): two synthetic genomes are written as
60-column FASTA (plain and .gz) under /tmp, then `bin/ntSynt --benchmark` runs on them.
    python scripts/gpu_cli_ingest_bench.py [genome_mbp]"""
import gzip
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from ntsynt_b200 import device, synth  # noqa: E402

mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 1000.0
wd = f"/tmp/cli_ingest_{int(mbp)}"
os.makedirs(wd, exist_ok=True)
ctx = device.Context(0)
wl = synth.Workload(2, int(mbp * 1e6), 1.0)
paths = []
for g in range(2):
    gen = wl.materialize(ctx, g)
    p = os.path.join(wd, wl.file_name(g))
    t0 = time.perf_counter()
    with open(p, "wb") as fh:
        for c in range(gen.n_contigs):
            seq = np.frombuffer(gen.contig_ascii(c), dtype=np.uint8)
            n = len(seq) - len(seq) % 60
            lines = np.empty((n // 60, 61), dtype=np.uint8)
            lines[:, :60] = seq[:n].reshape(-1, 60)
            lines[:, 60] = 10
            fh.write(b">" + gen.names[c].encode() + b"\n")
            fh.write(lines.tobytes())
            if len(seq) > n:
                fh.write(seq[n:].tobytes() + b"\n")
    gen.close()
    paths.append(p)
    print(f"wrote {p}: {os.path.getsize(p) / 1e9:.2f} GB in {time.perf_counter() - t0:.1f} s", flush=True)
ctx.close()
presets = ["-d", "1", "-k24", "-w", "1000"]
for label, files in (("plain FASTA", paths),):
    for rep in range(2):
        res = subprocess.run([sys.executable, os.path.join(ROOT, "bin", "ntSynt"), *files, *presets, "--prefix", os.path.join(wd, "run"),
                              "--benchmark"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, cwd=wd)
        line = [x for x in res.stdout.splitlines() if x.startswith("ingest + Bloom filter")]
        print(label, "run", rep, "rc", res.returncode, line[-1] if line else res.stdout[-800:], flush=True)
if len(sys.argv) > 2 and sys.argv[2] == "gz":
    gz = []
    for p in paths:
        t0 = time.perf_counter()
        with open(p, "rb") as fi, gzip.open(p + ".gz", "wb", compresslevel=1) as fo:
            while True:
                buf = fi.read(1 << 26)
                if not buf:
                    break
                fo.write(buf)
        gz.append(p + ".gz")
        print(f"gzip -1 {p}: {time.perf_counter() - t0:.1f} s", flush=True)
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bin", "ntSynt"), *gz, *presets, "--prefix", os.path.join(wd, "rungz"),
                          "--benchmark"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, cwd=wd)
    line = [x for x in res.stdout.splitlines() if x.startswith("ingest + Bloom filter")]
    print(".fa.gz (single gzip stream per file)", "rc", res.returncode, line[-1] if line else res.stdout[-800:], flush=True)
