"""Pins the CPU oracle (oracle/) against the reference's own golden vectors.  CPU only."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, mini_expected, mini_fastas, parse_sketch_tsv
from oracle import sketch_oracle as so


def test_hash_known_answers():
    "h1 of 1964 k-mers sampled from the reference's golden indexlr TSVs (real btllib output)"
    with open(os.path.join(GOLDEN, "hash_kats.json"), encoding="utf-8") as fh:
        kats = json.load(fh)["kats"]
    assert len(kats) > 1500
    for seq, k, h1 in kats:
        h0 = so.kmer_hash(seq.encode())
        assert so.ext_hash(h0, 1, k) == int(h1)


def test_documented_kat():
    "SURVEY F.3: k=24 GAAAAACTGATTTTTGAGCAGAAA -> 16504492209254430087"
    assert so.ext_hash(so.kmer_hash(b"GAAAAACTGATTTTTGAGCAGAAA"), 1, 24) == 16504492209254430087


def test_rolling_equals_direct():
    rng = np.random.default_rng(7)
    seq = bytes(rng.choice(np.frombuffer(b"ACGTNacgt", dtype=np.uint8), 5000, p=[.22, .22, .22, .22, .02, .025, .025, .025, .025]))
    for k in (5, 20, 24, 33, 64):
        h0, valid = so.hash_seq(seq, k)
        for p in range(0, len(seq) - k + 1, 37):
            d = so.kmer_hash(seq[p:p + k])
            assert (d is not None) == bool(valid[p])
            if d is not None:
                assert d == int(h0[p])


def test_reverse_complement_invariance():
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    rng = np.random.default_rng(3)
    for _ in range(50):
        s = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 24))
        assert so.kmer_hash(s) == so.kmer_hash(s.translate(comp)[::-1])


def test_bf_size_formula():
    "n = 29 058 289 -> 143 467 640 bytes (SURVEY A.2; +8 bytes breaks golden parity)"
    assert so.bf_bytes(29058289, 0.025) == 143467640


def test_cascade_equals_and():
    "src/ntsynt_make_common_bf.cpp:136-160 with one hash function is an AND of per-genome arrays"
    recs = [so.read_fasta(p) for p in mini_fastas("ABC")]
    genomes = [(f"g{i}", r) for i, r in enumerate(recs)]
    a = so.common_bf(genomes, 24, 0.025, cascade=False)
    b = so.common_bf(genomes, 24, 0.025, cascade=True)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("tag", ["AB", "ABC"])
def test_mini_sketch_fixture(tag, mini_params):
    "oracle sketch == committed fixture (which the reference's graph stage consumed)"
    paths = mini_fastas(tag)
    genomes = [(os.path.basename(p)[:-3], so.read_fasta(p)) for p in paths]
    bits = so.common_bf(genomes, mini_params["k"], mini_params["fpr"])
    for name, recs in genomes:
        want = mini_expected(tag, f"{name}.k{mini_params['k']}.w{mini_params['w']}.tsv.gz")
        got = "".join(so.sketch_tsv_lines(recs, mini_params["k"], mini_params["w"], bits))
        assert got == want


@pytest.mark.slow
@pytest.mark.parametrize("k,names", [(24, ["celegans-chrII-III.fa", "celegans-chrII-III.A.fa"]),
                                     (20, ["celegans-chrII-III.fa", "celegans-chrII-III.A.fa", "celegans-chrII-III.B.fa"])])
def test_reference_golden_sketches(demo_dir, k, names):
    "all 295 028 golden minimizers of tests/expected_result/*.k{24,20}.w1000.tsv, byte for byte"
    genomes = [(n, so.read_fasta(os.path.join(demo_dir, n + ".gz"))) for n in names]
    bits = so.common_bf(genomes, k, 0.025)
    for n, recs in genomes:
        with open(os.path.join(demo_dir, "expected_result", f"{n}.k{k}.w1000.tsv"), encoding="utf-8") as fh:
            want = fh.read()
        assert "".join(so.sketch_tsv_lines(recs, k, 1000, bits)) == want


def test_window_rule_properties():
    "every emitted position is the rightmost minimum of some window; windows span N runs"
    rng = np.random.default_rng(11)
    seq = bytearray(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 30000).tobytes())
    seq[5000:5200] = b"N" * 200
    seq = bytes(seq)
    k, w = 16, 50
    h1, pos = so.minimize(seq, k, w)
    h0, valid = so.hash_seq(seq, k)
    vpos = np.nonzero(valid)[0]
    keys = h0[vpos]
    want = []
    for j in range(w - 1, len(vpos)):
        win = keys[j - w + 1:j + 1]
        a = j - w + 1 + (len(win) - 1 - int(np.argmin(win[::-1])))
        if not want or want[-1] != a:
            want.append(a)
    assert np.array_equal(pos, vpos[want].astype(np.uint64))
    # the unpinned alternative (windows restart after a gap) is a test-only switch of the oracle: it may
    # only differ from the spanning rule in windows that straddle the N run
    _, posb = so.minimize(seq, k, w, restart_on_gap=True)
    far = lambda p: p[(p < 5000 - 2 * w) | (p > 5200 + 2 * w)]     # noqa: E731
    assert np.array_equal(far(pos), far(posb))
