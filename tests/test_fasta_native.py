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


def _bgzf(data, bs=4000):
    "bgzip's container: independent gzip members that carry their compressed size in a 'BC' extra field, + the EOF block"
    import struct
    import zlib
    out = []
    for i in range(0, len(data), bs):
        chunk = data[i:i + bs]
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        payload = c.compress(chunk) + c.flush()
        out.append(b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, len(payload) + 25) +
                   payload + struct.pack("<II", zlib.crc32(chunk), len(chunk)))
    out.append(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
    return b"".join(out)


def test_gz_inputs_plain_concatenated_and_bgzf(tmp_path):
    """.gz inputs (src/ntsynt_make_common_bf.cpp:32-36 reads them through btllib's SeqReader): a plain gzip stream,
    concatenated members and a BGZF file (inflated by several threads) give the same packed genome as the text"""
    import gzip
    rng = np.random.default_rng(3)
    recs = [(f"c{i}", bytes(rng.choice(np.frombuffer(b"ACGTNacgtn", dtype=np.uint8), int(n)))) for i, n in enumerate((70001, 5, 0, 333333))]
    txt = b"".join(b">" + n.encode() + b" desc\n" + b"\n".join(s[i:i + 70] for i in range(0, len(s), 70)) + b"\n" for n, s in recs)
    want = fasta.parse_fasta_bytes(txt)
    forms = {"plain.fa.gz": gzip.compress(txt), "cat.fa.gz": gzip.compress(txt[:1234]) + gzip.compress(txt[1234:]) + b"\0\0",
             "bgzf.fa.gz": _bgzf(txt), "text.fa": txt}
    assert len(fasta._bgzf_blocks(forms["bgzf.fa.gz"])) > 64 and fasta._bgzf_blocks(forms["plain.fa.gz"]) is None
    paths = []
    for name, raw in forms.items():
        (tmp_path / name).write_bytes(raw)
        paths.append(str(tmp_path / name))
    got = dict(fasta.read_fastas(paths, threads=4))            # concurrent readers, results in input order
    assert sorted(got) == [0, 1, 2, 3]
    for i in range(4):
        g = got[i]
        assert g.names == want.names and np.array_equal(g.lengths, want.lengths) and np.array_equal(g.words, want.words)
        assert np.array_equal(g.nrun_start, want.nrun_start) and np.array_equal(g.nrun_len, want.nrun_len)
    with pytest.raises(ValueError):
        fasta.inflate_gz(forms["plain.fa.gz"][:-20])


def _scan(data, mt_threads=None):
    import ctypes as C
    from ntsynt_b200._lib import check, lib, ptr
    cap = 1 << 16
    name_off, n_bases, seq_off, seq_end = (np.zeros(cap, dtype=np.uint64) for _ in range(4))
    name_len, lb, lw = (np.zeros(cap, dtype=np.uint32) for _ in range(3))
    uni = np.zeros(cap, dtype=np.uint8)
    nrec = C.c_uint64()
    keep = np.frombuffer(data, dtype=np.uint8)
    args = [C.c_void_p(keep.ctypes.data), len(data), cap, ptr(name_off, C.c_uint64), ptr(name_len, C.c_uint32), ptr(n_bases, C.c_uint64),
            ptr(seq_off, C.c_uint64), ptr(seq_end, C.c_uint64), ptr(lb, C.c_uint32), ptr(lw, C.c_uint32), ptr(uni, C.c_uint8), C.byref(nrec)]
    if mt_threads is None:
        check(lib.nts_fasta_scan(*args))
    else:
        check(lib.nts_fasta_scan_mt(*args, mt_threads))
    R = int(nrec.value)
    return [x[:R].tolist() for x in (name_off, name_len, n_bases, seq_off, seq_end, lb, lw, uni)]


@pytest.mark.parametrize("eol", [b"\n", b"\r\n"])
def test_threaded_scan_equals_serial_scan(eol):
    "nts_fasta_scan_mt (headers by slice, bodies in 8 MB pieces) against the line-by-line scan on > 16 MB of awkward text"
    rng = np.random.default_rng(len(eol))
    acgt = np.frombuffer(b"ACGTN", dtype=np.uint8)
    chunks = [b"junk before the first record" + eol]

    def rec(name, n, width, extra=()):
        seq = bytes(rng.choice(acgt, n))
        chunks.append(b">" + name + eol)
        lines = [seq[i:i + width] for i in range(0, n, width)]
        for at, what in extra:
            lines.insert(at, what)
        chunks.extend(l + eol for l in lines)
    rec(b"big1 has > inside the header", 19_000_000, 60)                      # uniform, spans several pieces
    rec(b"", 100, 60)                                                         # empty name
    rec(b"blank_first", 1000, 70, extra=[(0, b""), (0, b"")])                 # empty lines before the first sequence line
    rec(b"odd_in_the_middle", 9_000_000, 80, extra=[(60_000, b"ACG")])        # a short line far inside: not uniform
    rec(b"blank_at_the_end", 5000, 50, extra=[(100, b""), (101, b"")])        # blank lines after the last line: still uniform
    rec(b"short_then_more", 200, 10, extra=[(5, b"A"), (6, b"")])
    chunks.append(b">no_sequence" + eol + b">last_no_newline" + eol + b"ACGTACGT")
    data = b"".join(chunks)
    assert len(data) > (1 << 24)
    want = _scan(data)
    assert len(want[0]) == 8 and want[7][0] == 1 and want[7][3] == 0
    for t in (2, 5, 16):
        assert _scan(data, t) == want
