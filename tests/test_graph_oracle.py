"""Pins oracle/graph_oracle.py (CPU restatement of the graph stage) against fixtures made by the
reference's own code and against the reference's golden block files.  CPU only."""
import json
import gzip
import os

import pytest

from conftest import MINI, mini_expected, mini_fastas
from oracle import sketch_oracle as so
from oracle.graph_oracle import GraphOracle


def run_oracle(paths, k, w, w_rounds, indel, merge, z, fpr=0.025):
    genomes = [(os.path.basename(p)[:-3] if p.endswith(".gz") else os.path.basename(p), so.read_fasta(p)) for p in paths]
    bits = so.common_bf(genomes, k, fpr)
    go = GraphOracle([(f"{n}.k{k}.w{w}.tsv", r) for n, r in genomes], k, w, w_rounds, indel, merge, z, bits)
    go.run()
    return go


@pytest.mark.parametrize("tag", ["AB", "ABC"])
def test_mini_against_reference_fixture(tag, mini_params):
    p = mini_params
    go = run_oracle(mini_fastas(tag), p["k"], p["w"], p["w_rounds"], p["indel"], p["merge"], p["block_size"])
    assert go.outputs["final"] == mini_expected(tag, "synteny_blocks.tsv")
    assert go.outputs["pre_merge"] == mini_expected(tag, "pre-collinear-merge.synteny_blocks.tsv")
    with gzip.open(os.path.join(MINI, tag, "mx_dot_edges.json.gz"), "rt") as fh:
        want = json.load(fh)
    got = sorted([sorted((u, v)) + [wt] for u, v, wt in go.round0_edges])
    assert got == want


@pytest.mark.slow
@pytest.mark.parametrize("k,names,gold", [
    (24, ["celegans-chrII-III.fa", "celegans-chrII-III.A.fa"], "celegans-A-ntSynt"),
    (20, ["celegans-chrII-III.fa", "celegans-chrII-III.A.fa", "celegans-chrII-III.B.fa"], "celegans-A-B-ntSynt")])
def test_reference_golden_blocks(demo_dir, k, names, gold):
    "tests/ntsynt_tests.py:40-52 (-d 0.5 --indel 500 --merge 3000), whole files compared"
    go = run_oracle([os.path.join(demo_dir, n + ".gz") for n in names], k, 1000, [100, 10], 500, "3000", 500)
    exp = os.path.join(demo_dir, "expected_result")
    assert go.outputs["final"] == open(os.path.join(exp, gold + ".synteny_blocks.tsv")).read()
    assert go.outputs["pre_merge"] == open(os.path.join(exp, gold + ".pre-collinear-merge.synteny_blocks.tsv")).read()
