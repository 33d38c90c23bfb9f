"""C-ABI surface and host-side logic that needs no GPU."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import MINI, mini_fastas
from ntsynt_b200 import _lib, device, fasta, synth
from oracle import sketch_oracle as so


def test_library_exports_every_declared_symbol():
    syms = _lib.declared_symbols()
    assert len(syms) >= 50
    missing = [s for s in syms if not hasattr(_lib.lib, s)]
    assert not missing, missing


def test_every_declared_symbol_has_a_ctypes_signature():
    assert sorted(_lib._SIGS) == _lib.declared_symbols()


def test_no_cpu_fallback_without_a_device():
    n = C.c_int()
    rc = _lib.lib.nts_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(_lib.NtsError):
        device.Context(0)


def test_bf_size_matches_reference_formula():
    for n, fpr in [(29058289, 0.025), (3_000_000_000, 0.025), (123456, 0.01), (1, 0.5)]:
        assert int(_lib.lib.nts_bf_bytes(n, fpr)) == so.bf_bytes(n, fpr)
    assert int(_lib.lib.nts_bf_bytes(29058289, 0.025)) == 143467640


def test_pack_unpack_roundtrip_and_n_runs():
    seq = b"ACGTNNNNacgtRYACGT" * 7 + b"N"
    g = fasta.pack_records([("c1", seq), ("c2", b"ACGT" * 50)])
    assert g.total_bases == len(seq) + 200
    text = g.contig_text(0)
    want = seq.upper().replace(b"R", b"A").replace(b"Y", b"A").replace(b"N", b"A")
    assert text == want
    runs = [(int(s), int(l)) for s, l in zip(g.nrun_start[:int(g.nrun_off[1])], g.nrun_len[:int(g.nrun_off[1])])]
    exp, i = [], 0
    while i < len(seq):
        if seq[i:i + 1].upper() not in (b"A", b"C", b"G", b"T"):
            j = i
            while j < len(seq) and seq[j:j + 1].upper() not in (b"A", b"C", b"G", b"T"):
                j += 1
            exp.append((i, j - i)); i = j
        else:
            i += 1
    assert runs == exp
    assert g.kmer_text(1, 3, 5) == "TACGT"
    assert int(g.word_off[1]) % 2 == 0


def test_fai_writer_matches_samtools_goldens(demo_dir, tmp_path):
    for n in ["celegans-chrII-III.fa", "celegans-chrII-III.A.fa", "celegans-chrII-III.B.fa"]:
        import gzip, shutil
        plain = tmp_path / n
        with gzip.open(os.path.join(demo_dir, n + ".gz"), "rb") as fi, open(plain, "wb") as fo:
            shutil.copyfileobj(fi, fo)
        g = fasta.read_fasta(str(plain))
        out = tmp_path / (n + ".fai")
        fasta.write_fai(g, str(out))
        assert out.read_text() == open(os.path.join(demo_dir, "expected_result", n + ".fai")).read()


def test_fai_writer_mini_fixture(tmp_path):
    import gzip, shutil
    plain = tmp_path / "miniA.fa"
    with gzip.open(mini_fastas("AB")[0], "rb") as fi, open(plain, "wb") as fo:
        shutil.copyfileobj(fi, fo)
    g = fasta.read_fasta(str(plain))
    out = tmp_path / "miniA.fa.fai"
    fasta.write_fai(g, str(out))
    assert out.read_text() == open(os.path.join(MINI, "AB", "miniA.fa.fai")).read()


def test_synth_segment_tables_tile_every_contig():
    wl = synth.Workload(2, 3_000_000, 1.0, n_contigs=6)
    for g in range(2):
        lengths, segs = wl.segments(g)
        assert segs.dtype.itemsize == C.sizeof(_lib.SynthSeg)
        for c in range(6):
            s = segs[segs["dst_contig"] == c]
            assert s["dst_start"][0] == 0
            assert (s["dst_start"][1:] == s["dst_start"][:-1] + s["len"][:-1]).all()
            assert int(s["len"].sum()) == int(lengths[c])
        assert (segs["strand"] == -1).any() and (segs["anc_contig"] == -1).any() and (segs["anc_contig"] == -2).any()
    a = wl.segments(0)[1]
    b = synth.Workload(2, 3_000_000, 1.0, n_contigs=6).segments(0)[1]
    assert a.tobytes() == b.tobytes()          # deterministic


def test_ancestor_formula_base_composition():
    f = _lib.lib.nts_synth_ancestor_base
    bases = np.array([f(1234, 0, 0.0, 0, p) for p in range(20000)])
    frac = np.bincount(bases, minlength=4) / len(bases)
    assert abs(frac[0] - 0.295) < 0.02 and abs(frac[1] - 0.205) < 0.02 and abs(frac[3] - 0.295) < 0.02


def test_native_graph_walks_on_tiny_inputs():
    "nts_host_walk_paths / nts_host_simplify: empty inputs, one three-run path through two sparse edges, one bubble"
    import ctypes as C
    from ntsynt_b200._lib import check, lib, ptr
    i64 = np.int64
    # runs [0..2] [3..5] [6..8]; sparse edges 2-6 and 8-3: path 0..2, 6..8, 3..5
    nbr = np.array([[-1, 1], [0, 2], [1, 6], [8, 4], [3, 5], [4, -1], [2, 7], [6, 8], [7, 3]], dtype=np.int32)
    starts, ends = np.array([0, 3, 6], dtype=i64), np.array([2, 5, 8], dtype=i64)
    sv = np.array([2, 3, 6, 8], dtype=i64)
    opos = np.arange(100, 109, dtype=i64)
    lo, hi, off = (np.zeros(16, dtype=i64) for _ in range(3))
    dr = np.zeros(16, dtype=np.int8)
    n_p, n_s = C.c_int64(), C.c_int64()
    check(lib.nts_host_walk_paths(nbr.ctypes.data_as(C.POINTER(C.c_int32)), 9, ptr(starts, C.c_int64), ptr(ends, C.c_int64), 3,
                                  ptr(sv, C.c_int64), 4, ptr(opos, C.c_int64), ptr(lo, C.c_int64), ptr(hi, C.c_int64),
                                  dr.ctypes.data_as(C.POINTER(C.c_int8)), ptr(off, C.c_int64), 16, C.byref(n_p), C.byref(n_s)))
    assert (n_p.value, n_s.value) == (1, 3)
    assert list(zip(lo[:3].tolist(), hi[:3].tolist(), dr[:3].tolist())) == [(0, 2, 1), (6, 8, 1), (3, 5, 1)]
    check(lib.nts_host_walk_paths(nbr.ctypes.data_as(C.POINTER(C.c_int32)), 9, ptr(starts, C.c_int64), ptr(ends, C.c_int64), 3,
                                  ptr(sv, C.c_int64), 0, ptr(opos, C.c_int64), ptr(lo, C.c_int64), ptr(hi, C.c_int64),
                                  dr.ctypes.data_as(C.POINTER(C.c_int8)), ptr(off, C.c_int64), 16, C.byref(n_p), C.byref(n_s)))
    assert (n_p.value, n_s.value) == (0, 0)
    # bubble: assembly 0 lists 0 1 2 3, assembly 1 lists 0 1 3 (vertex 2 missing there is impossible after the join, so
    # use ranks that put 2 elsewhere): a0 = [0,1,2,3,4], a1 = [0,1,3,4,2] -> 1 and 3 have three neighbours
    rank = np.array([[0, 1, 2, 3, 4], [0, 1, 4, 2, 3]], dtype=np.uint32)
    inv = np.array([[0, 1, 2, 3, 4], [0, 1, 3, 4, 2]], dtype=np.uint32)
    ctg = np.zeros((2, 5), dtype=np.int32)
    cand = np.array([1, 3], dtype=i64)
    bs, bt, rm = (np.zeros(16, dtype=i64) for _ in range(3))
    n_out = C.c_int64()
    check(lib.nts_host_simplify(ptr(cand, C.c_int64), 2, ptr(rank, C.c_uint32), ptr(inv, C.c_uint32),
                                ctg.ctypes.data_as(C.POINTER(C.c_int32)), 5, 5, 2, ptr(bs, C.c_int64), ptr(bt, C.c_int64),
                                ptr(rm, C.c_int64), 16, C.byref(n_out)))
    assert n_out.value == 1 and (bs[0], bt[0], rm[0]) == (1, 3, 2)      # edge 1-3 closes the triangle over vertex 2
    check(lib.nts_host_simplify(ptr(cand, C.c_int64), 0, ptr(rank, C.c_uint32), ptr(inv, C.c_uint32),
                                ctg.ctypes.data_as(C.POINTER(C.c_int32)), 5, 5, 2, ptr(bs, C.c_int64), ptr(bt, C.c_int64),
                                ptr(rm, C.c_int64), 16, C.byref(n_out)))
    assert n_out.value == 0


def test_reference_arm_runs_without_the_cuda_library():
    "bench.py --impl reference: oracle-made sample + oracle path; the product library is never loaded (bench.py asserts it)"
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--genome-mbp", "60", "--cpu-sample-mbp", "6"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                         timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0


def _run(cmd, cwd):
    import subprocess
    res = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-3000:]
    return res.stdout


def test_block_stats_utility(tmp_path):
    import shutil
    import sys
    BIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bin")
    "analysis_scripts/denovo_synteny_block_stats.py: same ten numbers (computed here by the script's own definitions)"
    names = ["miniA.fa", "miniB.fa", "miniC.fa"]
    shutil.copyfile(os.path.join(MINI, "ABC", "synteny_blocks.tsv"), tmp_path / "b.tsv")
    for n in names:
        shutil.copyfile(os.path.join(MINI, "ABC", n + ".fai"), tmp_path / (n + ".fai"))
    out = _run([sys.executable, os.path.join(BIN, "denovo_synteny_block_stats.py"), "--tsv", "b.tsv", "--fai", *[n + ".fai" for n in names]],
              tmp_path).splitlines()
    assert out[0].split("\t")[:2] == ["Number_blocks", "Number_blocks_all_asm"] and len(out[1].split("\t")) == 10
    with open(os.path.join(MINI, "ABC", "block_stats.txt"), encoding="utf-8") as fh:
        assert out == fh.read().splitlines()


def test_bf_file_formats_round_trip(tmp_path):
    """<prefix>.bf (src/ntsynt_make_common_bf.cpp:164): btllib's header layout by default (unpinned: no .bf fixture in
    the reference), our own container on request; both read back, and foreign files are refused"""
    import numpy as np
    from ntsynt_b200 import io
    bits = np.random.default_rng(5).integers(0, 256, 4096, dtype=np.uint8)
    for fmt in ("btllib", "native"):
        path = str(tmp_path / f"x.{fmt}.bf")
        io.save_bf(path, bits, 24, fmt=fmt)
        got, k = io.load_bf_bytes(path)
        assert k == 24 and np.array_equal(got, bits)
    head = open(str(tmp_path / "x.btllib.bf"), "rb").read(200)
    assert head.startswith(b"[BTLKmerBloomFilter_v") and b"[HeaderEnd]\n" in head and b"hash_num = 1" in head
    (tmp_path / "junk.bf").write_bytes(b"not a filter")
    with pytest.raises(ValueError):
        io.load_bf_bytes(str(tmp_path / "junk.bf"))
    (tmp_path / "cut.bf").write_bytes(open(str(tmp_path / "x.btllib.bf"), "rb").read()[:-10])
    with pytest.raises(ValueError):
        io.load_bf_bytes(str(tmp_path / "cut.bf"))


def test_repeat_bf_size_argument():
    "bin/ntsynt_make_repeat_bfs.py:10-24"
    import argparse
    from ntsynt_b200 import cli
    ap = argparse.ArgumentParser()
    assert [cli.parse_bf_size(x, ap) for x in ("17B", "3k", "5M", "2G")] == [17, 3000, 5000000, 2000000000]
    with pytest.raises(SystemExit):
        cli.parse_bf_size("12", ap)


def test_sort_blocks_utility(tmp_path):
    "visualization_scripts/sort_ntsynt_blocks.py: fixture written by the reference's own script (order C, A, B)"
    import gzip
    import shutil
    import sys
    BIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bin")
    shutil.copyfile(os.path.join(MINI, "ABC", "synteny_blocks.tsv"), tmp_path / "b.tsv")
    out = _run([sys.executable, os.path.join(BIN, "sort_ntsynt_blocks.py"), "--synteny_blocks", "b.tsv", "--sort_order",
                "miniC.fa", "miniA.fa", "miniB.fa"], tmp_path)
    with gzip.open(os.path.join(MINI, "ABC", "sorted_CAB.tsv.gz"), "rt") as fh:
        assert out == fh.read()
    for n in ("miniC.fa", "miniA.fa", "miniB.fa"):
        (tmp_path / (n + ".fai")).write_text("x\t1\t0\t1\t2\n")
    out2 = _run([sys.executable, os.path.join(BIN, "sort_ntsynt_blocks.py"), "--synteny_blocks", "b.tsv", "--fais", "--sort_order",
                 "miniC.fa.fai", "miniA.fa.fai", "miniB.fa.fai"], tmp_path)
    assert out2 == out
