#!/usr/bin/env python3
"""Host-only: .gz FASTA ingest, zlib against the library's own decoder (csrc/nts_inflate.cu) at 1..N threads.
    python scripts/cpu_gz_ingest_bench.py [Mbp] [max threads]
Makes Mbp of random ACGT as 60-column FASTA, compresses it once with zlib level 6 (what `gzip` writes), then times
zlib.decompress, nts_gz_inflate_mt and fasta.read_fasta (inflate + scan + 2-bit pack)."""
import os
import sys
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ntsynt_b200 import fasta  # noqa: E402

mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 200.0
max_threads = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
rng = np.random.default_rng(1)
n = int(mbp * 1e6) // 60 * 60
rows = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), n).reshape(-1, 60)
buf = np.empty((rows.shape[0], 61), dtype=np.uint8)
buf[:, :60], buf[:, 60] = rows, 10
text = b">chr1\n" + buf.tobytes()
t0 = time.perf_counter()
c = zlib.compressobj(6, zlib.DEFLATED, 31)
raw = c.compress(text) + c.flush()
print(f"{len(text) / 1e6:.0f} MB of text -> {len(raw) / 1e6:.1f} MB .gz ({time.perf_counter() - t0:.0f} s to compress)")
t0 = time.perf_counter()
out = zlib.decompress(raw, 31)
dt = time.perf_counter() - t0
assert out == text
print(f"zlib                      {len(text) / 1e6 / dt:7.0f} MB/s")
th = 1
while th <= max_threads:
    t0 = time.perf_counter()
    got = fasta.inflate_gz_native(raw, threads=th)
    dt = time.perf_counter() - t0
    assert bytes(got) == text
    print(f"nts_gz_inflate_mt {th:2d} thr   {len(text) / 1e6 / dt:7.0f} MB/s")
    th *= 2
path = "/tmp/_nts_gz_bench.fa.gz"
with open(path, "wb") as fh:
    fh.write(raw)
for mode in ("zlib", "native"):
    os.environ["NTS_GZ_INFLATE"] = mode
    t0 = time.perf_counter()
    g = fasta.read_fasta(path, threads=max_threads)
    dt = time.perf_counter() - t0
    print(f"read_fasta ({mode:6s})       {g.total_bases / 1e6 / dt:7.0f} Mbp/s  ({dt:.2f} s)")
os.remove(path)
