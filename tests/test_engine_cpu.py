"""Host-side graph logic of the product (ntsynt_b200/synteny.py) exercised without a GPU: the engine
is fed by a TEST-ONLY backend (oracle sketches + numpy join, tests/backends.py).  The CUDA backend
is covered by the -m gpu tests."""
import os

import numpy as np
import pytest

import synth_small
from backends import OracleBackend, numpy_join
from conftest import mini_expected, mini_fastas
from ntsynt_b200.synteny import IntervalIndex, SyntenyEngine
from oracle import sketch_oracle as so
from oracle.graph_oracle import GraphOracle


def run_engine(paths, k, w, w_rounds, indel, merge, z, lean=False, native=True, **kw):
    bases = [os.path.basename(p)[:-3] if p.endswith(".gz") else os.path.basename(p) for p in paths]
    tsv = [f"{b}.k{k}.w{w}.tsv" for b in bases]
    order = sorted(range(len(paths)), key=lambda i: tsv[i], reverse=True)
    be = OracleBackend([paths[i] for i in order], [tsv[i] for i in order], k, lean=lean,
                       repeat_bits=kw.pop("repeat_bits", None), filter_mode=kw.pop("filter_mode", None))
    eng = SyntenyEngine(be, k, w, w_rounds, indel, merge, z, write_files=False, quiet=True, **kw)
    eng.native = native      # C++ walks (csrc/nts_hostgraph.cu) or their Python statements
    eng.run()
    return eng, be


@pytest.mark.parametrize("lean", [False, True, "dev"])
@pytest.mark.parametrize("tag", ["AB", "ABC"])
def test_mini_against_reference_fixture(tag, mini_params, lean):
    p = mini_params
    eng, _ = run_engine(mini_fastas(tag), p["k"], p["w"], p["w_rounds"], p["indel"], p["merge"], p["block_size"], lean)
    assert eng.outputs["final"] == mini_expected(tag, "synteny_blocks.tsv")
    assert eng.outputs["pre_merge"] == mini_expected(tag, "pre-collinear-merge.synteny_blocks.tsv")


def _repeat_bits():
    import gzip
    from conftest import MINI
    with gzip.open(os.path.join(MINI, "repeat_bits.bin.gz"), "rb") as fh:
        return np.frombuffer(fh.read(), dtype=np.uint8).copy()


@pytest.mark.parametrize("lean", [False, "dev"])
@pytest.mark.parametrize("tag,mode", [("ABC_filter", "Filter"), ("ABC_indexlr", "Indexlr"), ("ABC", None)])
def test_repeat_filter_modes_and_interarrivals_against_reference_fixture(tag, mode, mini_params, lean):
    "--filter Filter | Indexlr and --interarrivals (bin/ntsynt_synteny.py:172-187,557-564,601-609): files made by the reference's own code"
    p = mini_params
    eng, _ = run_engine(mini_fastas("ABC"), p["k"], p["w"], p["w_rounds"], p["indel"], p["merge"], p["block_size"], lean,
                        repeat_bits=_repeat_bits() if mode else None, filter_mode=mode, interarrivals=True)
    assert eng.outputs["final"] == mini_expected(tag, "synteny_blocks.tsv")
    assert eng.outputs["pre_merge"] == mini_expected(tag, "pre-collinear-merge.synteny_blocks.tsv")
    got = sorted(int(x) for x in eng.outputs["interarrivals"].split())
    assert got == [int(x) for x in mini_expected(tag, "interarrivals.sorted.txt.gz").split()]


def test_dev_overlap_check_against_reference_cases():
    "check_non_overlapping (bin/ntsynt_synteny.py:234-253): warnings recorded from the reference's own function"
    import json
    import types
    from conftest import GOLDEN
    from ntsynt_b200.synteny import Block
    with open(os.path.join(GOLDEN, "overlap_cases.json"), encoding="utf-8") as fh:
        cases = json.load(fh)["cases"]
    assert sum(len(c["warnings"]) for c in cases) > 20
    for c in cases:
        G, k = c["G"], 24
        eng = SyntenyEngine.__new__(SyntenyEngine)
        eng.G, eng.k, eng.z, eng.outputs = G, k, c["z"], {}
        eng.names = [f"asm{a}" for a in range(G)]
        eng.be = types.SimpleNamespace(contig_names=[["ctg0", "ctg1"]] * G)
        blocks = [Block(None, [int(r[a][0][3:]) for a in range(G)], ["+"] * G, 0, 0, [r[a][1] for a in range(G)],
                        [r[a][2] - k for a in range(G)], 5) for r in c["blocks"]]
        got = eng._check_non_overlapping(blocks)
        assert [list(w) for w in got] == c["warnings"]


@pytest.mark.slow
@pytest.mark.parametrize("k,names,gold", [
    (24, ["celegans-chrII-III.fa", "celegans-chrII-III.A.fa"], "celegans-A-ntSynt"),
    (20, ["celegans-chrII-III.fa", "celegans-chrII-III.A.fa", "celegans-chrII-III.B.fa"], "celegans-A-B-ntSynt")])
def test_reference_golden_blocks(demo_dir, k, names, gold):
    eng, _ = run_engine([os.path.join(demo_dir, n + ".gz") for n in names], k, 1000, [100, 10], 500, "3000", 500)
    exp = os.path.join(demo_dir, "expected_result")
    assert eng.outputs["final"] == open(os.path.join(exp, gold + ".synteny_blocks.tsv")).read()
    assert eng.outputs["pre_merge"] == open(os.path.join(exp, gold + ".pre-collinear-merge.synteny_blocks.tsv")).read()


@pytest.mark.parametrize("seed,G,presets", [(101, 2, "low"), (102, 3, "low"), (103, 4, "mid"), (104, 2, "mid"), (105, 5, "low")])
def test_engine_equals_graph_oracle_on_rearranged_genomes(tmp_path, seed, G, presets):
    "seeded genomes with inversions, translocations, duplications, N runs; d<1 and 1<=d<=10 style presets"
    gens = synth_small.make_genomes(seed, G, contig_lens=(100000, 70000, 30000), sub=0.004 * (1 + seed % 3), n_inv=4,
                                    n_trans=3, n_dup=3)
    paths = []
    for i, recs in enumerate(gens):
        p = str(tmp_path / f"g{chr(65 + i)}.fa")
        synth_small.write_fasta(p, recs)
        paths.append(p)
    k, w = 16, 40
    w_rounds, indel, merge, z = ([20, 5], 300, "400", 200) if presets == "low" else ([25, 10], 2000, "2w", 400)
    lean = ["dev", True, False][seed % 3]
    eng, be = run_engine(paths, k, w, w_rounds, indel, merge, z, lean=lean)
    go = GraphOracle([(os.path.basename(p) + f".k{k}.w{w}.tsv", so.read_fasta(p)) for p in paths], k, w, w_rounds,
                     indel, merge, z, be.bits)
    go.run()
    assert eng.outputs["final"] == go.outputs["final"]
    assert eng.outputs["pre_merge"] == go.outputs["pre_merge"]
    assert eng.outputs["final"].count("\n") >= G * 3
    # the Python statements of the two native walks give the same blocks and the same intermediate counts
    eng_py, _ = run_engine(paths, k, w, w_rounds, indel, merge, z, lean=lean, native=False)
    assert eng_py.outputs == eng.outputs
    assert eng_py.stats["simplified_vertices"] == eng.stats["simplified_vertices"] and eng_py.stats["paths"] == eng.stats["paths"]
    # and the device-resident form gives what the dense forms give
    if lean != "dev":
        eng_dev, be_dev = run_engine(paths, k, w, w_rounds, indel, merge, z, lean="dev")
        assert eng_dev.outputs == eng.outputs
        assert be_dev.calls.get("runs_to_blocks", 0) == 1 + len(w_rounds)
        assert be_dev.calls.get("refine_filter", 0) == len(w_rounds)
        assert eng_dev.stats["new_raw"] == eng.stats["new_raw"] and eng_dev.stats["new_common"] == eng.stats["new_common"]


def test_interval_index_matches_bruteforce():
    rng = np.random.default_rng(5)
    s = rng.integers(0, 1000, 40)
    e = s + rng.integers(1, 60, 40)
    ii = IntervalIndex(s, e)
    a = rng.integers(0, 1100, 500)
    b = a + rng.integers(1, 30, 500)
    want = np.array([any(si < bi and ai < ei for si, ei in zip(s, e)) for ai, bi in zip(a, b)])
    assert np.array_equal(ii.overlaps(a.astype(np.int64), b.astype(np.int64)), want)


def test_numpy_join_semantics():
    "dedup within an assembly, intersection across assemblies, numbering by the orienting assembly"
    t0 = (np.array([5, 7, 7, 9, 11], dtype=np.uint64), np.arange(5, dtype=np.uint32) * 10, np.zeros(5, dtype=np.uint32))
    t1 = (np.array([11, 9, 5, 13], dtype=np.uint64), np.arange(4, dtype=np.uint32) * 10, np.zeros(4, dtype=np.uint32))
    H, POS, CTG, RANK, link, deg = numpy_join([t0, t1], 1)
    assert list(H) == [11, 9, 5]            # order of assembly 1; 7 is duplicated in assembly 0, 13 is not common
    assert list(POS[0]) == [40, 30, 0] and list(POS[1]) == [0, 10, 20]
    assert list(RANK[0]) == [2, 1, 0]
    assert list(link) == [1, 1, 0]


def test_block_masks_match_a_bruteforce_union():
    "get_synteny_bed_lists + slop + maskfasta (bin/ntsynt_synteny.py:117-157): per contig, the union of the shrunk block extents"
    import types
    from ntsynt_b200.synteny import Block
    rng = np.random.default_rng(11)
    G, k, prev_w, n_ctg = 3, 24, 100, 6
    lens = [[int(x) for x in rng.integers(20000, 60000, n_ctg)] for _ in range(G)]
    eng = SyntenyEngine.__new__(SyntenyEngine)
    eng.G, eng.k = G, k
    eng.be = types.SimpleNamespace(contig_names=[[f"c{c}" for c in range(n_ctg)]] * G, contig_lengths=lens)
    blocks = []
    for _ in range(400):
        ctg = [int(c) for c in rng.integers(0, n_ctg - 1, G)]          # the last contig never gets a block
        fp = [int(rng.integers(0, lens[a][ctg[a]])) for a in range(G)]
        lp = [int(f + rng.integers(-3000, 3000)) for f in fp]          # some shorter than the threshold, some reach past the end
        blocks.append(Block(None, ctg, ["+"] * G, 0, 0, fp, lp, 5))
    masks = eng._masks_for(blocks, prev_w)
    thr, shrink = max(2 * prev_w, prev_w + k + 1), prev_w + k
    assert len(masks) == G
    for a in range(G):
        assert len(masks[a]) == n_ctg and len(masks[a][n_ctg - 1][0]) == 0
        for c in range(n_ctg):
            want = np.zeros(lens[a][c] + 4000, dtype=bool)
            for b in blocks:
                if b.ctg[a] != c:
                    continue
                s, e = min(b.first_pos[a], b.last_pos[a]), max(b.first_pos[a], b.last_pos[a]) + k
                if e - s > thr:
                    s2, e2 = max(s + shrink, 0), min(e - shrink, lens[a][c])
                    if s2 < e2:
                        want[s2:e2] = True
            got = np.zeros_like(want)
            ss, ee = masks[a][c]
            assert ss.dtype == np.uint64 and ee.dtype == np.uint64
            assert (ss[1:] > ee[:-1]).all()                             # sorted, disjoint, not even touching
            for s, e in zip(ss.tolist(), ee.tolist()):
                got[s:e] = True
            assert np.array_equal(got, want)


def test_randomised_cases_engine_equals_graph_oracle():
    """the three vertex-storage forms of the engine against oracle/graph_oracle.py on cases that draw the genome count (2-6),
    rearrangements, N runs, k, w, the rounds (none to three), --indel, --collinear-merge, -z and --simplify-graph at random
    (scripts/fuzz_graph_stage.py engine)"""
    import importlib.util
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts", "fuzz_graph_stage.py")
    spec = importlib.util.spec_from_file_location("fuzz_graph_stage", path)
    fz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fz)
    assert all(fz.check_seed("engine", seed) for seed in range(1000, 1012))
