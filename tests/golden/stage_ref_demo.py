#!/usr/bin/env python3
"""Stage the reference's demo FASTAs and golden outputs where the GPU box can see them.

/root/reference does not exist on the GPU box, so the three C. elegans demo genomes
(tests/*.fa.gz, ~25 MB) and the golden outputs are COPIED (data only, no source code) into
tests/golden/_ref_demo/, which is git-ignored (kept out of history) but not gpurun-ignored
(travels with the snapshot).  __graft_entry__.build() calls this when the reference is
present; tests that need the full demo skip when the directory is absent.
"""
import os
import shutil
import sys

REF = os.environ.get("NTSYNT_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref_demo")

FILES = ["celegans-chrII-III.fa.gz", "celegans-chrII-III.A.fa.gz", "celegans-chrII-III.B.fa.gz"]
EXPECTED = [
    "celegans-A-ntSynt.synteny_blocks.tsv", "celegans-A-ntSynt.pre-collinear-merge.synteny_blocks.tsv",
    "celegans-A-B-ntSynt.synteny_blocks.tsv", "celegans-A-B-ntSynt.pre-collinear-merge.synteny_blocks.tsv",
    "celegans-chrII-III.fa.k24.w1000.tsv", "celegans-chrII-III.A.fa.k24.w1000.tsv",
    "celegans-chrII-III.fa.k20.w1000.tsv", "celegans-chrII-III.A.fa.k20.w1000.tsv",
    "celegans-chrII-III.B.fa.k20.w1000.tsv",
    "celegans-chrII-III.fa.fai", "celegans-chrII-III.A.fa.fai", "celegans-chrII-III.B.fa.fai",
]


def stage(force=False):
    src_tests = os.path.join(REF, "tests")
    if not os.path.isdir(src_tests):
        return None
    os.makedirs(os.path.join(DST, "expected_result"), exist_ok=True)
    for f in FILES:
        d = os.path.join(DST, f)
        if force or not os.path.exists(d):
            shutil.copyfile(os.path.join(src_tests, f), d)
    for f in EXPECTED:
        d = os.path.join(DST, "expected_result", f)
        if force or not os.path.exists(d):
            shutil.copyfile(os.path.join(src_tests, "expected_result", f), d)
    return DST


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv))
