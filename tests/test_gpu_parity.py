"""Parity of the CUDA path (through the C-ABI) against the CPU oracle and the golden fixtures.
Run on the B200 box:  python -m pytest tests -m gpu"""
import gzip
import json
import os

import numpy as np
import pytest

import synth_small
from backends import numpy_edges, numpy_join
from conftest import MINI, mini_expected, mini_fastas, parse_sketch_tsv
from ntsynt_b200 import _lib, device, fasta, pipeline, synth
from ntsynt_b200.synteny import SyntenyEngine
from oracle import sketch_oracle as so
from oracle.graph_oracle import GraphOracle

pytestmark = pytest.mark.gpu


def _upload(ctx, records):
    return ctx.upload(fasta.pack_records(records))


def _tricky_records(seed=3):
    "N runs, IUPAC codes, lower case, a contig shorter than k, one shorter than k+w-1, an all-N contig"
    g = synth_small.make_genomes(seed, 1, contig_lens=(90000, 40000), n_nruns=6, lowercase=True)[0]
    rng = np.random.default_rng(seed)
    extra = bytearray(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 3000).tobytes())
    extra[100:103] = b"RYK"
    extra[2000:2001] = b"n"
    return g + [("short", b"ACGTACGTAC"), ("mid", bytes(extra[:60])), ("iupac", bytes(extra)), ("allN", b"N" * 500),
                ("empty", b"")]


@pytest.mark.parametrize("k", [24, 20, 5, 33, 64])
def test_nthash_kernel_matches_oracle(cuda_ctx, k):
    recs = _tricky_records()
    g = _upload(cuda_ctx, recs)
    for c, (_, seq) in enumerate(recs):
        h0, valid = cuda_ctx.hash_contig(g, c, k)
        oh, ov = so.hash_seq(seq, k)
        assert np.array_equal(valid, ov)
        assert np.array_equal(h0[ov == 1], oh[ov == 1])


def test_hash_known_answers_on_device(cuda_ctx):
    "h1 of k-mers sampled from the reference's golden indexlr output, computed by the sketch kernel"
    with open(os.path.join(os.path.dirname(MINI), "hash_kats.json"), encoding="utf-8") as fh:
        kats = json.load(fh)["kats"]
    for k in (24, 20):
        sel = [(s, h) for s, kk, h in kats if kk == k][:400]
        g = _upload(cuda_ctx, [(f"r{i}", s.encode()) for i, (s, _) in enumerate(sel)])
        h1, pos, ctg = cuda_ctx.sketch(g, k, 1).to_numpy()       # w = 1: every k-mer is its own minimizer
        assert list(ctg) == list(range(len(sel))) and not pos.any()
        assert [int(x) for x in h1] == [int(h) for _, h in sel]


@pytest.mark.parametrize("tag", ["AB", "ABC"])
def test_bloom_filter_bits_match_oracle(cuda_ctx, tag, mini_params):
    k = mini_params["k"]
    recs = [so.read_fasta(p) for p in mini_fastas(tag)]
    names = [os.path.basename(p)[:-3] for p in mini_fastas(tag)]
    nbytes = device.BloomFilter.size_for(sum(len(s) for _, s in recs[0]), 0.025)
    assert nbytes == so.bf_bytes(sum(len(s) for _, s in recs[0]), 0.025)
    gens = [_upload(cuda_ctx, r) for r in recs]
    per = []
    for g, r in zip(gens, recs):
        bf = cuda_ctx.bloom(nbytes)
        bf.insert_genome(g, k)
        want = so.genome_bits(r, k, nbytes)
        assert np.array_equal(bf.to_numpy(), want)
        assert bf.popcount() == int(np.unpackbits(want).sum())
        per.append(bf)
    common = pipeline.build_common_bf(cuda_ctx, gens, names, k)
    assert np.array_equal(common.to_numpy(), so.common_bf(list(zip(names, recs)), k, 0.025, cascade=True))
    # idempotence / commutativity of the merge
    a = cuda_ctx.bloom(nbytes).from_numpy(per[0].to_numpy())
    a.iand(per[1]); a.iand(per[1])
    b = cuda_ctx.bloom(nbytes).from_numpy(per[1].to_numpy())
    b.iand(per[0])
    assert np.array_equal(a.to_numpy(), b.to_numpy())
    # repeat filter (bin/ntsynt_make_repeat_bfs.py)
    rep, scratch = cuda_ctx.bloom(nbytes), cuda_ctx.bloom(nbytes)
    for g in gens:
        rep.insert_repeats(scratch, g, k)
    assert np.array_equal(rep.to_numpy(), so.repeat_bf(list(zip(names, recs)), k, nbytes))


@pytest.mark.parametrize("tag", ["AB", "ABC"])
def test_sketch_matches_reference_consumed_fixture(cuda_ctx, tag, mini_params):
    k, w = mini_params["k"], mini_params["w"]
    paths = mini_fastas(tag)
    names = [os.path.basename(p)[:-3] for p in paths]
    packed = [fasta.read_fasta(p) for p in paths]
    gens = [cuda_ctx.upload(p) for p in packed]
    common = pipeline.build_common_bf(cuda_ctx, gens, names, k)
    for n, pk, g in zip(names, packed, gens):
        want = parse_sketch_tsv(mini_expected(tag, f"{n}.k{k}.w{w}.tsv.gz"))
        h1, pos, ctg = cuda_ctx.sketch(g, k, w, common=common).to_numpy()
        for c, cname in enumerate(pk.names):
            sel = ctg == c
            assert [int(x) for x in h1[sel]] == want[cname][0]
            assert [int(x) for x in pos[sel]] == want[cname][1]


@pytest.mark.parametrize("w", [1, 2, 10, 100, 999, 1000, 2500, 5000])
def test_sketch_window_sizes_and_masks(cuda_ctx, w):
    "ragged inputs, N runs, masks (refinement rounds), windows larger than some contigs"
    recs = _tricky_records(seed=9)
    k = 24
    g = _upload(cuda_ctx, recs)
    nbytes = so.bf_bytes(sum(len(s) for _, s in recs), 0.025)
    bits = so.genome_bits(recs, k, nbytes)
    bits[::3] = 0                                    # drop a third of the bytes: a sparse "common" filter
    bf = cuda_ctx.bloom(nbytes).from_numpy(bits)
    rng = np.random.default_rng(w)
    masks, masked = [], []
    for _, seq in recs:
        L = len(seq)
        s = np.sort(rng.integers(0, max(L, 1), 12)) if L else np.zeros(0, dtype=np.int64)
        iv, last = [], 0
        for a in s:
            a = max(int(a), last)
            b = min(a + int(rng.integers(1, 4000)), L)
            if a < b:
                iv.append((a, b)); last = b
        masks.append((np.array([x for x, _ in iv], dtype=np.uint64), np.array([y for _, y in iv], dtype=np.uint64)))
        buf = bytearray(seq)
        for a, b in iv:
            buf[a:b] = b"N" * (b - a)
        masked.append(bytes(buf))
    for use_bf in (True, False):
        for use_mask in (False, True):
            h1, pos, ctg = cuda_ctx.sketch(g, k, w, common=bf if use_bf else None,
                                           masks=masks if use_mask else None).to_numpy()
            for c, (_, seq) in enumerate(recs):
                oh1, opos = so.minimize(masked[c] if use_mask else seq, k, w, bits if use_bf else None)
                sel = ctg == c
                assert np.array_equal(h1[sel], oh1), (w, use_bf, use_mask, c)
                assert np.array_equal(pos[sel].astype(np.uint64), opos)


def test_join_links_degrees_and_edges_match_numpy(cuda_ctx, mini_params):
    k, w = mini_params["k"], mini_params["w"]
    paths = mini_fastas("ABC")
    names = [os.path.basename(p)[:-3] for p in paths]
    gens = [cuda_ctx.upload(fasta.read_fasta(p)) for p in paths]
    common = pipeline.build_common_bf(cuda_ctx, gens, names, k)
    tables = [cuda_ctx.sketch(g, k, w, common=common) for g in gens]
    host = [t.to_numpy() for t in tables]
    for order in (0, 2):
        mg = device.MinimizerGraph(cuda_ctx, tables, order)
        H, POS, CTG, RANK, link, deg = mg.vertices()
        nH, nPOS, nCTG, nRANK, nlink, ndeg = numpy_join(host, order)
        assert np.array_equal(H, nH) and np.array_equal(POS, nPOS) and np.array_equal(CTG, nCTG)
        assert np.array_equal(RANK, nRANK) and np.array_equal(link, nlink) and np.array_equal(deg, ndeg)
        u, v, sup = mg.edges()
        want = numpy_edges(nRANK.astype(np.int64), nCTG)
        assert list(zip(u.tolist(), v.tolist(), sup.tolist())) == want
    # the edge multiset equals the reference's round-0 graph (.mx.dot of the fixture run)
    with gzip.open(os.path.join(MINI, "ABC", "mx_dot_edges.json.gz"), "rt") as fh:
        dot = json.load(fh)
    mg = device.MinimizerGraph(cuda_ctx, tables, 0)
    H = mg.vertices()[0]
    u, v, sup = mg.edges()
    got = sorted([sorted((str(int(H[a])), str(int(H[b])))) + [bin(int(s)).count("1")] for a, b, s in zip(u, v, sup)])
    assert got == dot


@pytest.mark.parametrize("tag", ["AB", "ABC"])
def test_pipeline_blocks_match_reference_fixture(tag, mini_params):
    p = mini_params
    out, eng = pipeline.run_ntsynt(mini_fastas(tag), k=p["k"], w=p["w"], w_rounds=p["w_rounds"], indel=p["indel"],
                                   merge=p["merge"], block_size=p["block_size"], write_files=False)
    assert out == mini_expected(tag, "synteny_blocks.tsv")
    assert eng.outputs["pre_merge"] == mini_expected(tag, "pre-collinear-merge.synteny_blocks.tsv")


@pytest.mark.parametrize("k,names,gold", [
    (24, ["celegans-chrII-III.fa", "celegans-chrII-III.A.fa"], "celegans-A-ntSynt"),
    (20, ["celegans-chrII-III.fa", "celegans-chrII-III.A.fa", "celegans-chrII-III.B.fa"], "celegans-A-B-ntSynt")])
def test_pipeline_reproduces_reference_goldens(demo_dir, k, names, gold):
    "tests/ntsynt_tests.py:40-52 end to end on the GPU, whole files compared"
    out, eng = pipeline.run_ntsynt([os.path.join(demo_dir, n + ".gz") for n in names], k=k, w=1000, w_rounds=(100, 10),
                                   indel=500, merge="3000", block_size=500, write_files=False)
    exp = os.path.join(demo_dir, "expected_result")
    assert out == open(os.path.join(exp, gold + ".synteny_blocks.tsv")).read()
    assert eng.outputs["pre_merge"] == open(os.path.join(exp, gold + ".pre-collinear-merge.synteny_blocks.tsv")).read()


def test_golden_demo_sketches_on_device(cuda_ctx, demo_dir):
    "tests/expected_result/*.k24.w1000.tsv from the GPU (hash, BF geometry, AND, tie-break all pinned)"
    names = ["celegans-chrII-III.fa", "celegans-chrII-III.A.fa"]
    packed = [fasta.read_fasta(os.path.join(demo_dir, n + ".gz")) for n in names]
    gens = [cuda_ctx.upload(p) for p in packed]
    common = pipeline.build_common_bf(cuda_ctx, gens, names, 24)
    for n, pk, g in zip(names, packed, gens):
        with open(os.path.join(demo_dir, "expected_result", f"{n}.k24.w1000.tsv"), encoding="utf-8") as fh:
            want = parse_sketch_tsv(fh.read())
        h1, pos, ctg = cuda_ctx.sketch(g, 24, 1000, common=common).to_numpy()
        for c, cname in enumerate(pk.names):
            sel = ctg == c
            assert [int(x) for x in h1[sel]] == want[cname][0] and [int(x) for x in pos[sel]] == want[cname][1]


@pytest.mark.parametrize("seed,G", [(201, 2), (202, 3), (203, 4)])
def test_pipeline_equals_graph_oracle_on_rearranged_genomes(tmp_path, seed, G):
    gens = synth_small.make_genomes(seed, G, contig_lens=(100000, 70000, 30000), sub=0.006, n_inv=4, n_trans=3, n_dup=3)
    paths = []
    for i, recs in enumerate(gens):
        p = str(tmp_path / f"g{chr(65 + i)}.fa")
        synth_small.write_fasta(p, recs)
        paths.append(p)
    k, w, w_rounds, indel, merge, z = 16, 40, [20, 5], 300, "400", 200
    out, eng = pipeline.run_ntsynt(paths, k=k, w=w, w_rounds=w_rounds, indel=indel, merge=merge, block_size=z,
                                   write_files=False)
    genomes = [(os.path.basename(p), so.read_fasta(p)) for p in paths]
    bits = so.common_bf(genomes, k, 0.025)
    go = GraphOracle([(n + f".k{k}.w{w}.tsv", r) for n, r in genomes], k, w, w_rounds, indel, merge, z, bits)
    go.run()
    assert out == go.outputs["final"] and eng.outputs["pre_merge"] == go.outputs["pre_merge"]


def test_synthetic_generator_matches_host_formula_and_is_deterministic(cuda_ctx):
    wl = synth.Workload(2, 2_000_000, 1.0, n_contigs=4)
    g0 = wl.materialize(cuda_ctx, 0)
    again = wl.materialize(cuda_ctx, 0)
    for c in range(4):
        assert np.array_equal(g0.contig_words(c), again.contig_words(c))
    lengths, segs = wl.segments(0)
    f = _lib.lib.nts_synth_ancestor_base
    text = g0.contig_ascii(1)
    s1 = segs[segs["dst_contig"] == 1]
    seg = s1[(s1["anc_contig"] >= 0) & (s1["strand"] > 0) & (s1["len"] > 2000)][0]
    start, anc, ac = int(seg["dst_start"]), int(seg["anc_start"]), int(seg["anc_contig"])
    want = bytes(b"ACGT"[f(wl.seed, wl.n_repeat_fam, wl.repeat_slot_prob, ac, anc + i)] for i in range(2000))
    got = text[start:start + 2000]
    diff = sum(a != b for a, b in zip(want, got))
    assert diff < 40            # only the d/200 substitutions differ
    off, st, ln = g0.nruns()
    assert len(st) > 0 and text[int(st[int(off[1])]):int(st[int(off[1])]) + 5] == b"NNNNN"


def test_full_size_properties(cuda_ctx):
    "size-independent checks at 2 x 150 Mbp: sortedness, window property on samples, AND monotonicity"
    k, w = 24, 1000
    wl = synth.Workload(2, 150_000_000, 1.0)
    gens = [wl.materialize(cuda_ctx, g) for g in range(2)]
    nbytes = device.BloomFilter.size_for(gens[0].total_bases, 0.025)
    a, b = cuda_ctx.bloom(nbytes), cuda_ctx.bloom(nbytes)
    a.insert_genome(gens[0], k); b.insert_genome(gens[1], k)
    pa, pb = a.popcount(), b.popcount()
    a.iand(b)
    pc = a.popcount()
    assert 0 < pc <= min(pa, pb)
    h1, pos, ctg = cuda_ctx.sketch(gens[0], k, w, common=a).to_numpy()
    key = ctg.astype(np.int64) * (1 << 32) + pos
    assert (np.diff(key) > 0).all()                                   # sorted, distinct
    dens = len(h1) / gens[0].total_bases
    assert 0.5 / w < dens < 2.5 / w
    # one contig slice checked against the oracle
    seq = gens[0].contig_ascii(3, 0, 400000)
    bits = a.to_numpy()
    oh1, opos = so.minimize(seq, k, w, bits)
    sel = (ctg == 3) & (pos < 400000 - 2 * w)
    n = int(sel.sum())
    assert n > 100 and np.array_equal(h1[sel], oh1[:n]) and np.array_equal(pos[sel].astype(np.uint64), opos[:n])


@pytest.mark.parametrize("w", [128, 500, 1000, 4000])
def test_sparse_and_dense_sketch_kernels_agree(cuda_ctx, w):
    """sketch_sparse_kernel (candidates below a hash threshold + nearest-smaller scans, escalating unresolved
    tiles) against sketch_kernel (van Herk over every slot) on 2 x 40 Mbp; also a thinned filter that leaves
    windows without any survivor, which must take the escalation path"""
    k = 24
    wl = synth.Workload(2, 40_000_000, 1.0)
    gens = [wl.materialize(cuda_ctx, g) for g in range(2)]
    nbytes = device.BloomFilter.size_for(gens[0].total_bases, 0.025)
    a, b = cuda_ctx.bloom(nbytes), cuda_ctx.bloom(nbytes)
    a.insert_genome(gens[0], k); b.insert_genome(gens[1], k)
    a.iand(b)
    thin = a.to_numpy().copy()
    thin[len(thin) // 16:] = 0                       # 15/16 of the hash space never passes: many windows without a survivor
    thin_bf = cuda_ctx.bloom(nbytes).from_numpy(thin)
    for bf, must_escalate in ((a, False), (thin_bf, True), (thin_bf, False), (None, False)):
        e0, q0 = cuda_ctx.sketch_escalated, cuda_ctx.sketch_queried_all
        if must_escalate:                            # pin the candidate density: the sampled pass rate would raise it
            os.environ["NTS_SKETCH_LAMBDA"] = "24"   # (or choose the dense kernel outright)
        try:
            sparse = [x.copy() for x in cuda_ctx.sketch(gens[1], k, w, common=bf).to_numpy()]
        finally:
            os.environ.pop("NTS_SKETCH_LAMBDA", None)
        esc = cuda_ctx.sketch_escalated - e0
        os.environ["NTS_SKETCH_DENSE"] = "1"
        try:
            dense = cuda_ctx.sketch(gens[1], k, w, common=bf).to_numpy()
        finally:
            del os.environ["NTS_SKETCH_DENSE"]
        assert len(sparse[0]) > 0
        for x, y in zip(sparse, dense):
            assert np.array_equal(x, y)
        if must_escalate:
            assert esc > 0
        elif bf is thin_bf:
            # < 10 % of the k-mers pass: every slot looked up, survivors listed, nothing handed to the dense selector
            assert cuda_ctx.sketch_queried_all - q0 == 1 and esc == 0
        elif bf is a and w >= 500:
            n_tiles_dense = gens[1].total_bases / (8960 - w)
            assert esc < 0.02 * n_tiles_dense        # the sparse kernel did the work


def test_async_upload_gives_the_same_filter_and_sketch(cuda_ctx):
    "nts_genome_upload_async: consumers order themselves after the copy (page-locked source)"
    k, w = 24, 200
    recs = _tricky_records(seed=21)
    packed = fasta.pack_records(recs) if hasattr(fasta, "pack_records") else None
    if packed is None:
        pytest.skip("no in-memory packer")
    pin = device.PinnedU64(len(packed.words))
    pin.array[:] = packed.words
    sync_g = cuda_ctx.upload(packed)
    packed.words = pin.array
    async_g = cuda_ctx.upload(packed, async_copy=True)
    nbytes = so.bf_bytes(sum(len(s) for _, s in recs), 0.025)
    a, b = cuda_ctx.bloom(nbytes), cuda_ctx.bloom(nbytes)
    b.build_common(None, [async_g], k)              # first consumer of the async genome
    a.insert_genome(sync_g, k)
    assert np.array_equal(a.to_numpy(), b.to_numpy())
    for x, y in zip(cuda_ctx.sketch(sync_g, k, w, common=a).to_numpy(), cuda_ctx.sketch(async_g, k, w, common=b).to_numpy()):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("G,mbp,d", [(2, 16, 1.0), (3, 12, 1.3), (5, 8, 12.0)])
def test_baseline_configs_at_reduced_size_equal_the_oracle(cuda_ctx, G, mbp, d):
    """BASELINE.json configs 2 / 3 / 4 (2 genomes d=1; 3 genomes d=1.3; 5 genomes d=12 with w_rounds 500 250) on the
    bench's own synthetic generator at a size the CPU oracle finishes in seconds: k=24, w=1000, presets of
    bin/ntSynt:89-99.  Filter bits, round-0 sketches and both block files must equal the oracle's."""
    import bench
    k, w = 24, 1000
    ps = bench.presets(d)
    wl = synth.Workload(G, int(mbp * 1e6), d)
    gens = [wl.materialize(cuda_ctx, g) for g in range(G)]
    files = [wl.file_name(g) for g in range(G)]
    names = [pipeline.tsv_name(f, k, w) for f in files]
    order = pipeline.processing_order(names)
    recs = [[(wl.names[c], g.contig_ascii(c)) for c in range(g.n_contigs)] for g in gens]
    bits = so.common_bf(list(zip(files, recs)), k, 0.025)
    common = pipeline.build_common_bf(cuda_ctx, gens, files, k)
    assert np.array_equal(common.to_numpy(), bits)
    for g, rc in zip(gens[:2], recs[:2]):
        h1, pos, ctg = cuda_ctx.sketch(g, k, w, common=common).to_numpy()
        for c in (0, len(rc) - 1):
            oh1, opos = so.minimize(rc[c][1], k, w, bits)
            assert np.array_equal(h1[ctg == c], oh1) and np.array_equal(pos[ctg == c].astype(np.uint64), opos)
    be = pipeline.CudaBackend(cuda_ctx, [gens[i] for i in order], [names[i] for i in order], [wl.names] * G,
                              [[int(x) for x in gens[i].lengths] for i in order], k, common=common)
    eng = SyntenyEngine(be, k, w, ps["w_rounds"], ps["indel"], ps["merge"], ps["block_size"], write_files=False, quiet=True)
    go = GraphOracle(list(zip(names, recs)), k, w, ps["w_rounds"], ps["indel"], ps["merge"], ps["block_size"], bits)
    got = want = None
    try:
        got = eng.run()
    except SystemExit:
        got = "no paths"
    try:
        want = go.run()
    except SystemExit:
        want = "no paths"
    be.close()
    assert got == want
    if got != "no paths":
        assert eng.outputs["pre_merge"] == go.outputs["pre_merge"]


@pytest.mark.parametrize("tag", ["d1.3_G3", "d12_G5"])
def test_presets_cuda_path_equals_reference_made_fixture(cuda_ctx, tag):
    """the d >= 1 presets of bin/ntSynt:89-99 (w_rounds 250 100 / 500 250, --indel 50000 / 100000) with 3 and 5
    genomes: the CUDA path against block files written by the reference's own bin/ntsynt_run.py
    (tests/golden/make_golden.py presets(); tests/preset_cases.py regenerates the seeded genomes)"""
    import preset_cases as pc
    p = pc.CASES[tag]["params"]
    names = pc.names(tag)
    packed = [fasta.pack_records(r) for r in pc.checked_genomes(tag)]
    out, eng = pipeline.run_ntsynt(names, k=p["k"], w=p["w"], w_rounds=p["w_rounds"], indel=p["indel"], merge=p["merge"],
                                   block_size=p["block_size"], write_files=False, ctx=cuda_ctx, packed=packed)
    assert out == pc.expected(tag)
    assert eng.outputs["pre_merge"] == pc.expected(tag, "pre-collinear-merge.synteny_blocks.tsv")


def test_oracle_generator_equals_device_genome(cuda_ctx):
    "the CPU arm's sample (oracle/synth_oracle.c) is the same workload the GPU arm materialises: base for base, Ns included"
    from ntsynt_b200 import synth_layout
    wl = synth.Workload(2, 3_000_000, 1.3, n_contigs=5)
    lay = synth_layout.Layout(2, 3_000_000, 1.3, n_contigs=5)
    for g in range(2):
        dev = wl.materialize(cuda_ctx, g)
        recs = so.synth_records(lay, g)
        assert [n for n, _ in recs] == dev.names
        for c, (_, seq) in enumerate(recs):
            assert seq == dev.contig_ascii(c), (g, c)
        part = so.synth_records(lay, g, per_contig=12345)
        assert all(p[1] == r[1][:12345] for p, r in zip(part, recs))


@pytest.mark.parametrize("G", [2, 3])
def test_common_filter_kept_as_a_pair_gives_the_same_sketches(cuda_ctx, G):
    """nts_bf_build_common_lazy leaves the last cascade level apart (common AND level = the common filter of
    src/ntsynt_make_common_bf.cpp:136-160); nts_sketch2 looks candidates up in both.  Same bits, same minimizers, in the
    sparse kernel, the dense kernel (small w, masked round) and with the partitioned insert forced."""
    k = 24
    wl = synth.Workload(G, 30_000_000, 1.0)
    gens = [wl.materialize(cuda_ctx, g) for g in range(G)]
    nbytes = device.BloomFilter.size_for(gens[0].total_bases, 0.025)
    full, lvl = cuda_ctx.bloom(nbytes), cuda_ctx.bloom(nbytes)
    full.build_common(lvl, gens, k)
    want_bits = full.to_numpy().copy()
    for force in ("0", "1"):
        os.environ["NTS_BF_PARTITION"] = force
        try:
            first, last = cuda_ctx.bloom(nbytes), cuda_ctx.bloom(nbytes)
            last.from_numpy(np.full(nbytes, 0x5A, dtype=np.uint8))               # stale contents must not matter
            apart = first.build_common(last, gens, k, lazy=True)
        finally:
            del os.environ["NTS_BF_PARTITION"]
        assert apart is True
        assert np.array_equal(first.to_numpy() & last.to_numpy(), want_bits)
        assert not np.array_equal(first.to_numpy(), want_bits)                   # the AND really was left out
        masks = [(np.array([1000, 500000], dtype=np.uint64), np.array([200000, 900000], dtype=np.uint64))
                 if c == 0 else (np.zeros(0, dtype=np.uint64), np.zeros(0, dtype=np.uint64)) for c in range(gens[0].n_contigs)]
        for w, mk in ((1000, None), (250, None), (40, masks), (1000, masks)):
            a = cuda_ctx.sketch(gens[-1], k, w, common=full, masks=mk).to_numpy()
            b = cuda_ctx.sketch(gens[-1], k, w, common=first, common2=last, masks=mk).to_numpy()
            assert len(a[0]) > 0
            for x, y in zip(a, b):
                assert np.array_equal(x, y)
        first.close(); last.close()
    one = cuda_ctx.bloom(nbytes)
    assert one.build_common(None, gens[:1], k, lazy=True) is False              # a single genome: nothing to leave apart


def test_insert_that_starts_during_the_upload_gives_the_same_bits(cuda_ctx):
    """nts_genome_upload_async copies in growing chunks with an event after each; the first partitioned insert bins the
    tiles whose bases have arrived while the rest is still on its way (staged launches of the binning kernel).  Same
    bits as a synchronous upload, for SET, AND and the lazy pair, and the sketch that follows sees the whole genome."""
    k, w = 24, 500
    wl = synth.Workload(2, 48_000_000, 1.0)
    sync_g = [wl.materialize(cuda_ctx, g) for g in range(2)]
    nbytes = device.BloomFilter.size_for(sync_g[0].total_bases, 0.025)
    want, lvl = cuda_ctx.bloom(nbytes), cuda_ctx.bloom(nbytes)
    want.build_common(lvl, sync_g, k)
    want_bits = want.to_numpy().copy()
    want_mx = [x.copy() for x in cuda_ctx.sketch(sync_g[1], k, w, common=want).to_numpy()]
    pins = []
    for g in sync_g:
        pk = g.to_packed()
        pin = device.PinnedU64(len(pk.words)); pin.array[:] = pk.words; pk.words = pin.array
        pins.append((pk, pin))
    os.environ["NTS_UPLOAD_STAGE_MIN_WORDS"] = "1"
    os.environ["NTS_BF_PARTITION"] = "1"
    try:
        for rep in range(3):                                   # (timing-dependent paths: a few rounds)
            n0 = cuda_ctx.part_inserts
            fresh = [cuda_ctx.upload(pk, async_copy=True) for pk, _ in pins]
            got, l2 = cuda_ctx.bloom(nbytes), cuda_ctx.bloom(nbytes)
            if rep == 2:
                assert got.build_common(l2, fresh, k, lazy=True) is True
                bits = got.to_numpy() & l2.to_numpy()
                mx = cuda_ctx.sketch(fresh[1], k, w, common=got, common2=l2).to_numpy()
            else:
                got.build_common(l2, fresh, k)
                bits = got.to_numpy()
                mx = cuda_ctx.sketch(fresh[1], k, w, common=got).to_numpy()
            assert cuda_ctx.part_inserts - n0 == 2
            assert np.array_equal(bits, want_bits)
            for x, y in zip(mx, want_mx):
                assert np.array_equal(x, y)
            for f in fresh:
                f.close()
            got.close(); l2.close()
    finally:
        del os.environ["NTS_UPLOAD_STAGE_MIN_WORDS"]
        del os.environ["NTS_BF_PARTITION"]
