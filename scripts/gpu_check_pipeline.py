#!/usr/bin/env python3
"full pipeline on the GPU vs the reference's golden block files"
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ntsynt_b200 import pipeline  # noqa: E402

DEMO = os.path.join(ROOT, "tests", "golden", "_ref_demo")


def run(files, k, gold):
    t = time.time()
    out, eng = pipeline.run_ntsynt([os.path.join(DEMO, f) for f in files], k=k, w=1000, w_rounds=(100, 10), indel=500,
                                   merge="3000", block_size=500, write_files=False)
    dt = time.time() - t
    g = open(os.path.join(DEMO, "expected_result", gold + ".synteny_blocks.tsv")).read()
    gp = open(os.path.join(DEMO, "expected_result", gold + ".pre-collinear-merge.synteny_blocks.tsv")).read()
    print(gold, "FINAL", out == g, "PRE", eng.outputs["pre_merge"] == gp, f"{dt:.2f}s", eng.stats, eng.backend_timing)


run(["celegans-chrII-III.fa.gz", "celegans-chrII-III.A.fa.gz"], 24, "celegans-A-ntSynt")
run(["celegans-chrII-III.fa.gz", "celegans-chrII-III.A.fa.gz", "celegans-chrII-III.B.fa.gz"], 20, "celegans-A-B-ntSynt")
run(["celegans-chrII-III.fa.gz", "celegans-chrII-III.A.fa.gz"], 24, "celegans-A-ntSynt")
