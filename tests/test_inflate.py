"""nts_gz_inflate (csrc/nts_inflate.cu), the decompressor in front of the FASTA reader (btllib::SeqReader pipes .gz input
through one: src/ntsynt_make_common_bf.cpp:32-36,125,143), against zlib: every block type and compressor strategy, several
members, header fields, exact / short output buffers, truncation, corruption, garbage -- host code, no GPU."""
import ctypes as C
import glob
import gzip
import io
import os
import zlib

import numpy as np
import pytest

from ntsynt_b200 import fasta
from ntsynt_b200._lib import lib

HERE = os.path.dirname(os.path.abspath(__file__))


def native(raw, cap, crc=1):
    raw = bytes(raw)
    a = np.frombuffer(raw, dtype=np.uint8) if raw else np.zeros(1, np.uint8)
    out = np.full(cap + 8, 0xEE, dtype=np.uint8)
    n = C.c_uint64()
    rc = lib.nts_gz_inflate(C.c_void_p(a.ctypes.data), len(raw), C.c_void_p(out.ctypes.data), cap, C.byref(n), crc)
    assert n.value <= cap and (out[cap:] == 0xEE).all()            # never writes past the capacity it was given
    return rc, bytes(out[:n.value])


def _cases():
    rng = np.random.default_rng(0)
    cases = []
    for n in (0, 1, 2, 100, 5000, 70000, 400000):
        cases.append(bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), n)))
        cases.append(bytes(rng.integers(0, 256, n, dtype=np.uint8)))                       # incompressible: stored blocks
        cases.append(bytes(rng.choice(np.frombuffer(b"ACGTN\n", dtype=np.uint8), n, p=[.24, .24, .24, .24, .02, .02])))
    cases.append(b"A" * 300000)                                                            # matches at distance 1
    cases.append(b"ACGTTGCA" * 50000)
    cases.append(b"ACG" * 70000)                                                           # distance 3: byte-wise overlap copy
    cases.append(bytes(rng.integers(0, 256, 300, dtype=np.uint8)) * 1000)                  # long matches
    cases.append(b"".join(bytes([i % 251]) * (i % 7 + 1) for i in range(60000)))           # many symbols: long codes
    return cases


def test_every_block_type_and_strategy_matches_zlib():
    for data in _cases():
        for level in (0, 1, 6, 9):
            for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE):
                c = zlib.compressobj(level, zlib.DEFLATED, 31, 9, strategy)
                raw = c.compress(data) + c.flush()
                rc, out = native(raw, len(data), crc=level & 1)
                assert rc == 0 and out == data, (len(data), level, strategy, rc)


def test_members_header_fields_and_capacity():
    data = _cases()[17]
    raw = gzip.compress(data[:1000]) + gzip.compress(data[1000:]) + b"\0\0\0"              # concatenated members, zero padding
    assert native(raw, len(data)) == (0, data)
    bio = io.BytesIO()
    with gzip.GzipFile(filename="some_name.fa", mode="wb", fileobj=bio, mtime=5) as fh:   # FNAME field
        fh.write(data)
    assert native(bio.getvalue(), len(data)) == (0, data)
    hdr = b"\x1f\x8b\x08\x1e" + b"\0" * 6 + b"\x03\0abc" + b"name\0" + b"comment\0" + b"\x12\x34"   # FEXTRA FNAME FCOMMENT FHCRC
    body = gzip.compress(data)[10:]
    assert native(hdr + body, len(data)) == (0, data)
    raw = gzip.compress(data)
    assert native(raw, len(data))[0] == 0
    assert native(raw, len(data) - 1)[0] == 1 and native(raw, 10)[0] == 1 and native(raw, 0)[0] == 1     # "does not fit"
    assert native(gzip.compress(b""), 0) == (0, b"")
    # the wrapper sizes the buffer itself (several members: it has to grow it) and returns an array
    assert bytes(fasta.inflate_gz_native(gzip.compress(data[:100000]) + gzip.compress(data[100000:]))) == data
    assert bytes(fasta.inflate_gz_native(gzip.compress(b""))) == b""


def test_truncated_corrupt_and_random_input_is_refused_not_crashed_on():
    rng = np.random.default_rng(1)
    data = _cases()[15]
    raw = gzip.compress(data)
    for cut in list(range(1, 40)) + [len(raw) // 2, len(raw) - 5]:
        assert native(raw[:-cut], len(data))[0] < 0, cut
    with pytest.raises(ValueError):
        fasta.inflate_gz_native(raw[:-20])
    detected = 0
    for _ in range(300):
        b = bytearray(raw)
        b[int(rng.integers(10, len(raw)))] ^= 1 << int(rng.integers(0, 8))
        rc, out = native(bytes(b), len(data) + 1000)
        assert rc != 0 or out == data                      # a flipped bit is either noticed (CRC, length, code) or harmless
        detected += rc != 0
    assert detected >= 290
    b = bytearray(raw); b[-6] ^= 1                          # the CRC field itself
    assert native(bytes(b), len(data), crc=1)[0] < 0 and native(bytes(b), len(data), crc=0)[0] == 0
    for trial in range(2000):
        g = b"\x1f\x8b\x08\x00\0\0\0\0\0\x03" + bytes(rng.integers(0, 256, int(rng.integers(0, 400)), dtype=np.uint8))
        native(g, 100000, crc=trial & 1)
    assert native(b"\x1f\x8b\x08", 10)[0] < 0 and native(b"plain text, not gzip at all", 100)[0] < 0 and native(b"", 10)[0] < 0


def test_every_gz_fixture_inflates_like_zlib():
    files = sorted(glob.glob(os.path.join(HERE, "golden", "**", "*.gz"), recursive=True))
    assert len(files) >= 4
    for f in files:
        raw = open(f, "rb").read()
        want = gzip.decompress(raw)
        assert bytes(fasta.inflate_gz_native(raw)) == want, f


def native_mt(raw, cap, threads, crc=1):
    raw = bytes(raw)
    a = np.frombuffer(raw, dtype=np.uint8)
    out = np.full(cap + 8, 0xEE, dtype=np.uint8)
    n = C.c_uint64()
    rc = lib.nts_gz_inflate_mt(C.c_void_p(a.ctypes.data), len(raw), C.c_void_p(out.ctypes.data), cap, C.byref(n), crc, threads)
    assert n.value <= cap and (out[cap:] == 0xEE).all()
    return rc, bytes(out[:n.value])


def _fasta_text(rng, n, nruns=()):
    a = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), n)
    for s_, l_ in nruns:
        a[s_:s_ + l_] = ord("N")
    rows = a[: n // 60 * 60].reshape(-1, 60)
    buf = np.empty((rows.shape[0], 61), dtype=np.uint8)
    buf[:, :60], buf[:, 60] = rows, 10
    return b">chr1 test\n" + buf.tobytes()


def test_one_member_decoded_by_several_threads(monkeypatch):
    """nts_gz_inflate_mt: chunks that start without their history (16-bit symbols, resolved afterwards).  Small chunk sizes
    bring the machinery down to test size: boundaries found by trial, chunks longer / shorter than a block, N runs that
    expand a thousand times, streams with no dynamic block at all (stored, fixed Huffman: every searching chunk is
    dropped), sync-flush markers, several members, short buffers, truncation, corruption"""
    rng = np.random.default_rng(5)
    cases = {"dna": _fasta_text(rng, 1_500_000), "n_runs": _fasta_text(rng, 3_000_000, [(100000, 1_000_000), (2_000_000, 700_000)]),
             "random": bytes(rng.integers(0, 256, 600_000, dtype=np.uint8)),
             "soft_masked": bytes(rng.choice(np.frombuffer(b"ACGTacgtN\n", dtype=np.uint8), 1_000_000))}
    for name, data in cases.items():
        for level, strategy in ((1, 0), (6, 0), (9, 0), (0, 0), (6, zlib.Z_FIXED)):
            c = zlib.compressobj(level, zlib.DEFLATED, 31, 9, strategy)
            raw = c.compress(data) + c.flush()
            for chunk, th in ((4096, 3), (20000, 2), (20000, 8), (65536, 4)):
                monkeypatch.setenv("NTS_GZ_CHUNK_BYTES", str(chunk))
                assert native_mt(raw, len(data), th) == (0, data), (name, level, strategy, chunk, th)
    monkeypatch.setenv("NTS_GZ_CHUNK_BYTES", "20000")
    data = cases["dna"]
    raw = gzip.compress(data)
    assert native_mt(gzip.compress(data[:500_000]) + gzip.compress(data[500_000:]) + b"\0\0", len(data), 4) == (0, data)
    assert native_mt(raw, len(data) - 1, 4)[0] == 1 and native_mt(raw, 1000, 4)[0] == 1 and native_mt(raw, len(data) // 2, 4)[0] == 1
    for cut in (1, 5, 9, 100, len(raw) // 3, len(raw) - 100):
        assert native_mt(raw[:-cut], len(data), 4)[0] < 0, cut
    for _ in range(60):
        b = bytearray(raw)
        b[int(rng.integers(10, len(raw)))] ^= 1 << int(rng.integers(0, 8))
        rc, out = native_mt(bytes(b), len(data) + 5000, 4)
        assert rc != 0 or out == data
    c = zlib.compressobj(6, zlib.DEFLATED, 31)                   # pigz-like stream: flush markers (empty stored blocks) inside
    parts = []
    for i in range(0, len(data), 131072):
        parts.append(c.compress(data[i:i + 131072]))
        parts.append(c.flush(zlib.Z_FULL_FLUSH if i % 3 else zlib.Z_SYNC_FLUSH))
    parts.append(c.flush())
    assert native_mt(b"".join(parts), len(data), 8) == (0, data)
    # and through the reader
    assert bytes(fasta.inflate_gz(raw, threads=4)) == data


def test_bgzf_members_decoded_independently():
    "bgzip files: every member names its compressed size, so the members are found without decoding and inflated in parallel"
    from test_fasta_native import _bgzf
    rng = np.random.default_rng(8)
    data = _fasta_text(rng, 2_000_000, [(300000, 400000)])
    raw = _bgzf(data, 65280)
    blocks = fasta._bgzf_blocks(raw)
    assert len(blocks) > 30
    for th in (1, 3, 8):
        assert bytes(fasta.inflate_gz(raw, threads=th)) == data
    b = bytearray(raw)
    a0, a1 = blocks[len(blocks) // 2]
    b[(a0 + a1) // 2] ^= 0x10                                      # damage inside one member: its CRC (or its codes) notice
    with pytest.raises(ValueError):
        fasta.inflate_gz(bytes(b), threads=4)
    b = bytearray(raw)
    b[blocks[3][1] - 4] ^= 1                                       # a member's ISIZE: the member no longer fills its place
    with pytest.raises(ValueError):
        fasta.inflate_gz(bytes(b), threads=4)
