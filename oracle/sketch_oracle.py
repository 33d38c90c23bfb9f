"""ctypes front-end of oracle/ntsynt_oracle.c plus small pure-Python helpers.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Each function names the reference
call site it restates; the arithmetic is btllib's (not in /root/reference), pinned by
the reference's golden indexlr TSVs -- see the header of ntsynt_oracle.c.
"""
import ctypes as C
import gzip
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    "Compile libntsynt_oracle.so with the committed Makefile (gcc, no reference sources involved)."
    so = os.path.join(_HERE, "libntsynt_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("ntsynt_oracle.c", "synth_oracle.c", "Makefile")]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(x) for x in srcs):
        subprocess.check_call(["make", "-C", _HERE, "clean", "all"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        u64, sz, u8p, u64p, cp = C.c_uint64, C.c_size_t, C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.c_char_p
        L.orc_srol.restype = u64; L.orc_srol.argtypes = [u64]
        L.orc_sror.restype = u64; L.orc_sror.argtypes = [u64]
        L.orc_kmer_hash.restype = C.c_int; L.orc_kmer_hash.argtypes = [cp, C.c_uint, u64p]
        L.orc_ext_hash.restype = u64; L.orc_ext_hash.argtypes = [u64, C.c_uint, C.c_uint]
        L.orc_hash_seq.restype = sz; L.orc_hash_seq.argtypes = [cp, sz, C.c_uint, u64p, u8p]
        L.orc_bf_bytes.restype = u64; L.orc_bf_bytes.argtypes = [C.c_longlong, C.c_double]
        L.orc_bf_insert_seq.restype = None; L.orc_bf_insert_seq.argtypes = [u8p, u64, cp, sz, C.c_uint, C.c_int]
        L.orc_bf_cascade_seq.restype = None
        L.orc_bf_cascade_seq.argtypes = [u8p, u8p, u64, cp, sz, C.c_uint, C.c_int]
        L.orc_bf_repeat_seq.restype = None; L.orc_bf_repeat_seq.argtypes = [u8p, u8p, u64, cp, sz, C.c_uint]
        L.orc_bf_popcount.restype = u64; L.orc_bf_popcount.argtypes = [u8p, u64]
        L.orc_minimize2.restype = sz
        L.orc_minimize2.argtypes = [cp, sz, C.c_uint, C.c_uint, u8p, u64, u8p, u64, C.c_int, u64p, u64p, sz]
        L.orc_minimize.restype = sz
        L.orc_minimize.argtypes = [cp, sz, C.c_uint, C.c_uint, u8p, u8p, u64, C.c_int, u64p, u64p, sz]
        L.orc_common_bf_level1.restype = None
        L.orc_common_bf_level1.argtypes = [u8p, u64, C.POINTER(cp), C.POINTER(sz), C.c_int, C.c_uint, C.c_int]
        L.orc_common_bf_cascade.restype = None
        L.orc_common_bf_cascade.argtypes = [u8p, u8p, u64, C.POINTER(cp), C.POINTER(sz), C.c_int, C.c_uint, C.c_int]
        L.orc_sketch_records.restype = None
        L.orc_sketch_records.argtypes = [C.POINTER(cp), C.POINTER(sz), C.c_int, C.c_uint, C.c_uint, u8p, u64,
                                         C.POINTER(u64p), C.POINTER(u64p), C.POINTER(sz), C.POINTER(sz), C.c_int]
        L.orc_max_threads.restype = C.c_int
        L.orc_synth_contig.restype = None
        L.orc_synth_contig.argtypes = [C.c_void_p, sz, C.c_uint32, u64, u64, u64, C.c_double, C.c_uint32, C.c_double, C.c_char_p]
        _LIB = L
    return _LIB


def _u8p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8)) if a is not None else None


def _u64p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


# ----------------------------------------------------------------------------- FASTA
def read_fasta(path):
    """[(id, seq_bytes)] ; id = header up to first whitespace (btllib SeqReader / indexlr)."""
    op = gzip.open if str(path).endswith(".gz") else open
    recs, name, chunks = [], None, []
    with op(path, "rb") as fh:
        for line in fh:
            if line.startswith(b">"):
                if name is not None:
                    recs.append((name, b"".join(chunks)))
                name = line[1:].split()[0].decode() if len(line) > 1 and line[1:].split() else ""
                chunks = []
            else:
                chunks.append(line.strip())
    if name is not None:
        recs.append((name, b"".join(chunks)))
    return recs


# ----------------------------------------------------------------------------- hashing
def kmer_hash(kmer: bytes, k=None):
    "canonical ntHash2 h0 of one k-mer, or None if it holds a non-ACGT base (direct definition)."
    k = len(kmer) if k is None else k
    out = C.c_uint64()
    ok = lib().orc_kmer_hash(kmer, k, C.byref(out))
    return out.value if ok else None


def ext_hash(h0, i, k):
    return lib().orc_ext_hash(h0, i, k)


def hash_seq(seq: bytes, k):
    "(h0[n-k+1] uint64, valid[n-k+1] uint8) for every k-mer start of seq (rolling)."
    nk = max(len(seq) - k + 1, 0)
    h0 = np.zeros(nk, dtype=np.uint64)
    valid = np.zeros(nk, dtype=np.uint8)
    if nk:
        lib().orc_hash_seq(seq, len(seq), k, _u64p(h0), _u8p(valid))
    return h0, valid


# ----------------------------------------------------------------------------- Bloom filter
def bf_bytes(genome_size, fpr):
    "src/ntsynt_make_common_bf.cpp:28-40 + btllib's 8-byte round-up."
    return int(lib().orc_bf_bytes(int(genome_size), float(fpr)))


def genome_bits(records, k, nbytes):
    "bit array of one genome: every valid k-mer's h0 %% m set (cpp:122-131)."
    bits = np.zeros(nbytes, dtype=np.uint8)
    for _, seq in records:
        lib().orc_bf_insert_seq(_u8p(bits), nbytes * 8, seq, len(seq), k, 0)
    return bits


def common_bf(genomes, k, fpr=0.025, nbytes=None, cascade=False):
    """Common BF of `genomes` = list of (path_string, records), in ANY order.

    Follows src/ntsynt_make_common_bf.cpp:107 (sort paths), :116 (size from first sorted
    genome), :122-160 (cascade).  cascade=False computes the equivalent AND of per-genome
    arrays; cascade=True runs the literal cascade (used to prove the equivalence)."""
    genomes = sorted(genomes, key=lambda g: g[0])
    if nbytes is None:
        nbytes = bf_bytes(sum(len(s) for _, s in genomes[0][1]), fpr)
    m = nbytes * 8
    bits = genome_bits(genomes[0][1], k, nbytes)
    for _, recs in genomes[1:]:
        if cascade:
            nxt = np.zeros(nbytes, dtype=np.uint8)
            for _, seq in recs:
                lib().orc_bf_cascade_seq(_u8p(bits), _u8p(nxt), m, seq, len(seq), k, 0)
            bits = nxt
        else:
            bits &= genome_bits(recs, k, nbytes)
    return bits


def repeat_bf(genomes, k, nbytes):
    "bin/ntsynt_make_repeat_bfs.py:56-69"
    rep = np.zeros(nbytes, dtype=np.uint8)
    for _, recs in genomes:
        gb = np.zeros(nbytes, dtype=np.uint8)
        for _, seq in recs:
            lib().orc_bf_repeat_seq(_u8p(gb), _u8p(rep), nbytes * 8, seq, len(seq), k)
    return rep


# ----------------------------------------------------------------------------- minimizers
def minimize(seq: bytes, k, w, common=None, repeat=None, restart_on_gap=False):
    "indexlr on one record: (h1 uint64[], pos uint64[]) in emission order."
    n = len(seq)
    m = (common.size if common is not None else 0) * 8
    rm = (repeat.size if repeat is not None else 0) * 8          # the repeat filter has its own size
    cap = max(16, 4 * (n // max(w, 1)) + 64)
    while True:
        h1 = np.empty(cap, dtype=np.uint64)
        pos = np.empty(cap, dtype=np.uint64)
        cnt = lib().orc_minimize2(seq, n, k, w, _u8p(common), m, _u8p(repeat), rm, int(restart_on_gap),
                                 _u64p(h1), _u64p(pos), cap)
        if cnt <= cap:
            return h1[:cnt].copy(), pos[:cnt].copy()
        cap = cnt


def sketch_tsv_lines(records, k, w, common=None, repeat=None, with_seq=True):
    """indexlr --long --pos [--seq] output, one text line per record.  The :seq field is printed
    upper case: btllib's SeqReader folds case by default [RECALL; unpinned -- the demo FASTAs hold no lower
    case, and the graph stage ignores the field unless --filter Filter is used]."""
    for name, seq in records:
        h1, pos = minimize(seq, k, w, common, repeat)
        if with_seq:
            toks = [f"{int(h)}:{int(p)}:{seq[int(p):int(p) + k].decode().upper()}" for h, p in zip(h1, pos)]
        else:
            toks = [f"{int(h)}:{int(p)}" for h, p in zip(h1, pos)]
        yield name + "\t" + " ".join(toks) + "\n"


def write_sketch_tsv(path, records, k, w, common=None, repeat=None):
    with open(path, "w", encoding="utf-8") as out:
        for line in sketch_tsv_lines(records, k, w, common, repeat):
            out.write(line)


def synth_records(layout, g, per_contig=None):
    """ASCII records of synthetic genome g (SURVEY 8d workload; oracle/synth_oracle.c), from a workload layout object
    with .names, .seed, .d, .n_repeat_fam, .repeat_slot_prob and .segments(g) (ntsynt_b200.synth_layout based).
    per_contig: only the first that many bases of every contig."""
    lengths, segs = layout.segments(g)
    out = []
    for c, name in enumerate(layout.names):
        sel = np.ascontiguousarray(segs[segs["dst_contig"] == c])
        n = int(lengths[c]) if per_contig is None else min(int(per_contig), int(lengths[c]))
        buf = C.create_string_buffer(max(n, 1))
        if n:
            lib().orc_synth_contig(sel.ctypes.data_as(C.c_void_p), len(sel), c, n, int(layout.seed),
                                   int(layout.seed) * 1000003 + g + 1, float(layout.d) / 200.0, int(layout.n_repeat_fam),
                                   float(layout.repeat_slot_prob), buf)
        out.append((name, buf.raw[:n]))
    return out
