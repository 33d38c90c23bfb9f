"""FASTA ingest for the B200 path: records -> 2-bit packed contigs + N runs + .fai rows.

Replaces btllib::SeqReader (src/ntsynt_make_common_bf.cpp:32-36,125,143; inside indexlr) and
`samtools faidx` (bin/ntsynt_run_pipeline.smk:48-53).  Record id = header up to the first
whitespace, as indexlr prints it.
"""
import ctypes as C
import gzip

import numpy as np

from ._lib import check, lib, ptr


class PackedGenome:
    "host-side packed genome (what nts_genome_upload consumes)"

    def __init__(self, names, lengths, word_off, words, nrun_off, nrun_start, nrun_len, fai=None, path=None):
        self.names = list(names)
        self.lengths = np.asarray(lengths, dtype=np.uint64)
        self.word_off = np.asarray(word_off, dtype=np.uint64)
        self.words = words
        self.nrun_off = np.asarray(nrun_off, dtype=np.uint64)
        self.nrun_start = np.asarray(nrun_start, dtype=np.uint64)
        self.nrun_len = np.asarray(nrun_len, dtype=np.uint64)
        self.fai = fai or []
        self.path = path

    @property
    def total_bases(self):
        return int(self.lengths.sum())

    def kmer_text(self, contig, pos, k):
        "ASCII of the k-mer at (contig, pos) -- the :seq field of indexlr --seq"
        buf = C.create_string_buffer(int(k))
        off = int(self.word_off[contig])
        view = self.words[off:]
        check(lib.nts_unpack_ascii(ptr(view, C.c_uint64), int(pos), int(k), buf))
        return buf.raw.decode()

    def contig_text(self, contig):
        n = int(self.lengths[contig])
        buf = C.create_string_buffer(n)
        view = self.words[int(self.word_off[contig]):]
        check(lib.nts_unpack_ascii(ptr(view, C.c_uint64), 0, n, buf))
        return buf.raw


def iter_fasta(path):
    """yield (name, seq_bytes, (offset, linebases, linewidth)) per record"""
    op = gzip.open if str(path).endswith(".gz") else open
    with op(path, "rb") as fh:
        name, chunks = None, []
        off = 0
        seq_off = lb = lw = 0
        for line in fh:
            if line[:1] == b">":
                if name is not None:
                    yield name, b"".join(chunks), (seq_off, lb, lw)
                fields = line[1:].split()
                name = fields[0].decode() if fields else ""
                chunks = []
                seq_off, lb, lw = off + len(line), 0, 0
            elif name is not None:
                stripped = line.rstrip(b"\r\n")
                if lb == 0:
                    lb, lw = len(stripped), len(line)
                chunks.append(stripped)
            off += len(line)
        if name is not None:
            yield name, b"".join(chunks), (seq_off, lb, lw)


def pack_records(records, path=None):
    "records: iterable of (name, seq_bytes[, fai_tuple]) -> PackedGenome"
    names, lengths, word_off, parts, fai = [], [], [], [], []
    nrun_off, nrs, nrl = [0], [], []
    woff = 0
    for rec in records:
        name, seq = rec[0], rec[1]
        n = len(seq)
        nw = int(lib.nts_packed_words(n))
        words = np.zeros(nw, dtype=np.uint64)
        cap = 1024
        while True:
            rs = np.zeros(cap, dtype=np.uint64)
            rl = np.zeros(cap, dtype=np.uint64)
            cnt = C.c_uint64()
            check(lib.nts_pack_ascii(seq, n, ptr(words, C.c_uint64), ptr(rs, C.c_uint64), ptr(rl, C.c_uint64), cap,
                                     C.byref(cnt)))
            if cnt.value <= cap:
                break
            cap = int(cnt.value)
        names.append(name)
        lengths.append(n)
        word_off.append(woff)
        parts.append(words)
        woff += nw
        nrs.append(rs[:cnt.value])
        nrl.append(rl[:cnt.value])
        nrun_off.append(nrun_off[-1] + int(cnt.value))
        if len(rec) > 2 and rec[2] is not None:
            fai.append((name, n, *rec[2]))
    words = np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint64)
    return PackedGenome(names, lengths, word_off, words, nrun_off,
                        np.concatenate(nrs) if nrs else np.zeros(0, dtype=np.uint64),
                        np.concatenate(nrl) if nrl else np.zeros(0, dtype=np.uint64), fai, path)


def read_fasta_python(path):
    "line-by-line reader + per-record packing (the statement of the format the native reader is tested against)"
    return pack_records(iter_fasta(path), path=path)


def parse_fasta_bytes(data, path=None, threads=0):
    """FASTA text (bytes) -> PackedGenome with the native reader: one memchr scan for the records, then the records
    (and the 4 Mbp pieces of long uniform-width records) packed by `threads` threads (0 = all cores)"""
    n = len(data)
    cap = 4096
    while True:
        name_off, n_bases, seq_off, seq_end = (np.zeros(cap, dtype=np.uint64) for _ in range(4))
        name_len, lb, lw = (np.zeros(cap, dtype=np.uint32) for _ in range(3))
        uni = np.zeros(cap, dtype=np.uint8)
        nrec = C.c_uint64()
        check(lib.nts_fasta_scan(data, n, cap, ptr(name_off, C.c_uint64), ptr(name_len, C.c_uint32), ptr(n_bases, C.c_uint64),
                                 ptr(seq_off, C.c_uint64), ptr(seq_end, C.c_uint64), ptr(lb, C.c_uint32), ptr(lw, C.c_uint32),
                                 ptr(uni, C.c_uint8), C.byref(nrec)))
        if nrec.value <= cap:
            break
        cap = int(nrec.value)
    R = int(nrec.value)
    n_bases, seq_off, seq_end, lb, lw, uni = n_bases[:R], seq_off[:R], seq_end[:R], lb[:R], lw[:R], uni[:R]
    names = [data[int(o):int(o) + int(l)].decode() for o, l in zip(name_off[:R], name_len[:R])]
    nw = ((n_bases + np.uint64(63)) // np.uint64(64)) * np.uint64(2)
    word_off = np.zeros(R + 1, dtype=np.uint64)
    np.cumsum(nw, out=word_off[1:])
    words = np.zeros(int(word_off[R]), dtype=np.uint64)
    nrun_off = np.zeros(R + 1, dtype=np.uint64)
    rcap = 65536
    z64 = np.zeros(1, dtype=np.uint64)
    while True:
        rs, rl = np.zeros(rcap, dtype=np.uint64), np.zeros(rcap, dtype=np.uint64)
        nruns = C.c_uint64()
        check(lib.nts_fasta_pack(data, R, ptr(n_bases if R else z64, C.c_uint64), ptr(seq_off if R else z64, C.c_uint64),
                                 ptr(seq_end if R else z64, C.c_uint64), ptr(lb if R else np.zeros(1, np.uint32), C.c_uint32),
                                 ptr(lw if R else np.zeros(1, np.uint32), C.c_uint32), ptr(uni if R else np.zeros(1, np.uint8), C.c_uint8),
                                 ptr(word_off, C.c_uint64), ptr(words if words.size else z64, C.c_uint64), ptr(nrun_off, C.c_uint64),
                                 ptr(rs, C.c_uint64), ptr(rl, C.c_uint64), rcap, C.byref(nruns), int(threads)))
        if nruns.value <= rcap:
            break
        rcap = int(nruns.value)
    fai = [(nm, int(nb), int(so_), int(a), int(b)) for nm, nb, so_, a, b in zip(names, n_bases, seq_off, lb, lw)]
    return PackedGenome(names, n_bases, word_off[:R], words, nrun_off, rs[:nruns.value], rl[:nruns.value], fai, path)


def read_fasta(path, threads=0):
    "FASTA file (.gz accepted) -> PackedGenome through the native reader (csrc/nts_fasta.cu)"
    op = gzip.open if str(path).endswith(".gz") else open
    with op(path, "rb") as fh:
        data = fh.read()
    return parse_fasta_bytes(data, path=path, threads=threads)


def write_fai(packed, out_path):
    "5-column .fai exactly as samtools faidx writes it (goldens: tests/expected_result/*.fai)"
    with open(out_path, "w", encoding="utf-8") as out:
        for row in packed.fai:
            out.write("\t".join(str(x) for x in row) + "\n")
