"""Native FASTA ingest (csrc/nts_fasta.cu) against the line-by-line Python statement of the format
(ntsynt_b200.fasta.read_fasta_python) and the reference's own .fai goldens."""
import gzip
import os

import numpy as np
import pytest

import synth_small
from conftest import mini_fastas
from ntsynt_b200 import fasta


def same(a, b):
    assert a.names == b.names
    assert np.array_equal(a.lengths, b.lengths) and np.array_equal(a.word_off, b.word_off)
    assert np.array_equal(a.words, b.words)
    assert np.array_equal(a.nrun_off, b.nrun_off) and np.array_equal(a.nrun_start, b.nrun_start)
    assert np.array_equal(a.nrun_len, b.nrun_len)
    assert [tuple(r) for r in a.fai] == [tuple(r) for r in b.fai]


@pytest.mark.parametrize("tag", ["AB", "ABC"])
def test_native_reader_equals_python_reader_on_fixtures(tag):
    for p in mini_fastas(tag):
        same(fasta.read_fasta(p), fasta.read_fasta_python(p))
        same(fasta.read_fasta(p, threads=1), fasta.read_fasta(p, threads=5))


@pytest.mark.parametrize("width,crlf,trailing_nl", [(60, False, True), (70, True, True), (61, False, False), (1, False, True)])
def test_native_reader_on_ragged_text(tmp_path, width, crlf, trailing_nl):
    "N runs across line and piece boundaries, IUPAC, lower case, empty records, CRLF, no final newline, odd headers"
    rng = np.random.default_rng(width)
    big = bytearray(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 9_500_000).tobytes())   # > 2 pieces of 4 Mbp
    for a, n in ((0, 17), (4194290, 40), (4194304 * 2 - 3, 3), (9_499_990, 10), (123456, 70000)):
        big[a:a + n] = b"N" * n
    big[5_000_000:5_000_020] = b"acgtnRYKMacgtacgtacg"
    recs = [("chrBig desc text", bytes(big)), ("empty", b""), ("\ttabbed\theader", b"ACGTNNNNACGT"),
            ("short", b"ACG"), ("allN", b"N" * 1000)]
    recs += [(n, s) for n, s in synth_small.make_genomes(3, 1, contig_lens=(30000, 12000), n_nruns=2, lowercase=True)[0]]
    eol = b"\r\n" if crlf else b"\n"
    chunks = [b"; a comment line before the first record" + eol]
    for name, seq in recs:
        chunks.append(b">" + name.encode() + eol)
        for i in range(0, len(seq), width):
            chunks.append(seq[i:i + width] + eol)
    text = b"".join(chunks)
    if not trailing_nl:
        text = text[:-len(eol)]
    p = tmp_path / "ragged.fa"
    p.write_bytes(text)
    same(fasta.read_fasta(str(p)), fasta.read_fasta_python(str(p)))
    pz = tmp_path / "ragged.fa.gz"
    with gzip.open(pz, "wb", compresslevel=1) as fh:
        fh.write(text)
    same(fasta.read_fasta(str(pz), threads=3), fasta.read_fasta_python(str(p)))


def test_non_uniform_lines_fall_back_to_one_piece(tmp_path):
    seq = bytes(np.random.default_rng(1).choice(np.frombuffer(b"ACGTN", dtype=np.uint8), 5_000_000).tobytes())
    lines, i, w = [b">r1"], 0, 50
    while i < len(seq):
        lines.append(seq[i:i + w]); i += w; w = 50 + (i % 7)          # ragged widths
    lines.append(b"")                                                 # blank line inside
    lines.append(b">r2")
    lines.append(b"ACGT")
    p = tmp_path / "nonuni.fa"
    p.write_bytes(b"\n".join(lines) + b"\n")
    same(fasta.read_fasta(str(p)), fasta.read_fasta_python(str(p)))


def test_fai_rows_match_reference_goldens(demo_dir):
    exp = os.path.join(demo_dir, "expected_result")
    for name in ("celegans-chrII-III.fa", "celegans-chrII-III.A.fa", "celegans-chrII-III.B.fa"):
        gold = os.path.join(exp, name + ".fai")
        if not os.path.exists(gold):
            pytest.skip("reference demo data not staged")
        pk = fasta.read_fasta(os.path.join(demo_dir, name + ".gz"))
        rows = "".join("\t".join(str(x) for x in r) + "\n" for r in pk.fai)
        assert rows == open(gold, encoding="utf-8").read()
