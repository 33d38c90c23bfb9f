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


def read_fasta(path):
    return pack_records(iter_fasta(path), path=path)


def write_fai(packed, out_path):
    "5-column .fai exactly as samtools faidx writes it (goldens: tests/expected_result/*.fai)"
    with open(out_path, "w", encoding="utf-8") as out:
        for row in packed.fai:
            out.write("\t".join(str(x) for x in row) + "\n")
