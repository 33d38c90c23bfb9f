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
    view = np.frombuffer(data, dtype=np.uint8) if n else np.zeros(1, dtype=np.uint8)    # bytes, mmap, ... : no copy
    dptr = C.c_void_p(view.ctypes.data)
    cap = 4096
    while True:
        name_off, n_bases, seq_off, seq_end = (np.zeros(cap, dtype=np.uint64) for _ in range(4))
        name_len, lb, lw = (np.zeros(cap, dtype=np.uint32) for _ in range(3))
        uni = np.zeros(cap, dtype=np.uint8)
        nrec = C.c_uint64()
        check(lib.nts_fasta_scan_mt(dptr, n, cap, ptr(name_off, C.c_uint64), ptr(name_len, C.c_uint32), ptr(n_bases, C.c_uint64),
                                    ptr(seq_off, C.c_uint64), ptr(seq_end, C.c_uint64), ptr(lb, C.c_uint32), ptr(lw, C.c_uint32),
                                    ptr(uni, C.c_uint8), C.byref(nrec), int(threads)))
        if nrec.value <= cap:
            break
        cap = int(nrec.value)
    R = int(nrec.value)
    n_bases, seq_off, seq_end, lb, lw, uni = n_bases[:R], seq_off[:R], seq_end[:R], lb[:R], lw[:R], uni[:R]
    names = [bytes(view[int(o):int(o) + int(l)]).decode() for o, l in zip(name_off[:R], name_len[:R])]
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
        check(lib.nts_fasta_pack(dptr, R, ptr(n_bases if R else z64, C.c_uint64), ptr(seq_off if R else z64, C.c_uint64),
                                 ptr(seq_end if R else z64, C.c_uint64), ptr(lb if R else np.zeros(1, np.uint32), C.c_uint32),
                                 ptr(lw if R else np.zeros(1, np.uint32), C.c_uint32), ptr(uni if R else np.zeros(1, np.uint8), C.c_uint8),
                                 ptr(word_off, C.c_uint64), ptr(words if words.size else z64, C.c_uint64), ptr(nrun_off, C.c_uint64),
                                 ptr(rs, C.c_uint64), ptr(rl, C.c_uint64), rcap, C.byref(nruns), int(threads)))
        if nruns.value <= rcap:
            break
        rcap = int(nruns.value)
    fai = [(nm, int(nb), int(so_), int(a), int(b)) for nm, nb, so_, a, b in zip(names, n_bases, seq_off, lb, lw)]
    return PackedGenome(names, n_bases, word_off[:R], words, nrun_off, rs[:nruns.value], rl[:nruns.value], fai, path)


def _bgzf_blocks(raw):
    """offsets [(start, end)] of the members of a BGZF file (bgzip: every gzip member carries its own compressed size in
    a 'BC' extra field), or None when `raw` is not BGZF from the first to the last byte"""
    out, off, n = [], 0, len(raw)
    while off < n:
        if n - off < 18 or raw[off:off + 4] != b"\x1f\x8b\x08\x04":
            return None
        xlen = int.from_bytes(raw[off + 10:off + 12], "little")
        x, xend, bsize = off + 12, off + 12 + xlen, None
        while x + 4 <= xend:
            slen = int.from_bytes(raw[x + 2:x + 4], "little")
            if raw[x:x + 2] == b"BC" and slen == 2:
                bsize = int.from_bytes(raw[x + 4:x + 6], "little") + 1
            x += 4 + slen
        if bsize is None or off + bsize > n:
            return None
        out.append((off, off + bsize))
        off += bsize
    return out


def inflate_gz(raw, threads=0):
    """gzip bytes -> plain bytes.  BGZF input (independent members) is inflated by `threads` threads, a batch of
    members each (zlib releases the GIL); anything else goes through the library's own decoder (inflate_gz_native), which
    also spreads one long member over the threads.  NTS_GZ_INFLATE=zlib: one streaming zlib inflate instead, as `gzip -dc`
    does -- which is how btllib's SeqReader reads .gz (one decompressor process per file)."""
    import os
    import zlib
    from concurrent.futures import ThreadPoolExecutor
    blocks = _bgzf_blocks(raw) if raw[:4] == b"\x1f\x8b\x08\x04" else None
    threads = threads or os.cpu_count() or 1
    if blocks and len(blocks) > 64 and os.environ.get("NTS_GZ_INFLATE", "native") != "zlib":
        # members decoded independently by the library's decoder, each at its place in one output buffer
        src = np.frombuffer(raw, dtype=np.uint8)
        off = np.array([a for a, _ in blocks] + [blocks[-1][1]], dtype=np.uint64)
        isz = sum(int.from_bytes(raw[b - 4:b], "little") for _, b in blocks)
        out = np.empty(isz + 8, dtype=np.uint8)
        got = C.c_uint64()
        rc = lib.nts_gz_inflate_members(C.c_void_p(src.ctypes.data), ptr(off, C.c_uint64), len(blocks), C.c_void_p(out.ctypes.data),
                                        isz, C.byref(got), 1, int(threads))
        if rc != 0:
            raise ValueError(lib.nts_last_error().decode() or "corrupt gzip stream")
        return out[:got.value]
    if blocks and len(blocks) > 64 and threads > 1:
        view = memoryview(raw)
        per = max(64, (len(blocks) + 8 * threads - 1) // (8 * threads))

        def job(i):
            # deflate payload of a member: after the 12 + xlen byte header, before the 8 byte CRC32 / ISIZE trailer
            parts = []
            for a, b in blocks[i:i + per]:
                xlen = int.from_bytes(view[a + 10:a + 12], "little")
                parts.append(zlib.decompress(view[a + 12 + xlen:b - 8], -15))
            return b"".join(parts)
        with ThreadPoolExecutor(max_workers=threads) as ex:
            return b"".join(ex.map(job, range(0, len(blocks), per)))
    if os.environ.get("NTS_GZ_INFLATE", "native") != "zlib":
        return inflate_gz_native(raw, threads=threads)
    out, data = [], memoryview(raw)
    while len(data):
        d = zlib.decompressobj(31)
        for i in range(0, len(data), 1 << 24):
            out.append(d.decompress(data[i:i + (1 << 24)]))
            if d.eof:
                break
        out.append(d.flush())
        if not d.eof:
            raise ValueError("truncated gzip stream")
        data = memoryview(d.unused_data).toreadonly() if d.unused_data else data[:0]
        while len(data) and data[0] == 0:          # zero padding between / after members is legal
            data = data[1:]
    return b"".join(out)


def inflate_gz_native(raw, verify_crc=True, threads=0):
    """gzip bytes -> plain bytes (a uint8 array) with the library's own decoder (csrc/nts_inflate.cu: nts_gz_inflate_mt), which
    follows concatenated members and checks every member's length and CRC-32; a long member is decoded by `threads`
    threads (0 = all cores).  The output buffer is sized from the ISIZE trailer -- exact for a file of one member below
    4 GB -- and grown when the decoder says it does not fit."""
    n = len(raw)
    src = np.frombuffer(raw, dtype=np.uint8) if n else np.zeros(1, dtype=np.uint8)
    if n < 18:
        raise ValueError("truncated gzip stream")
    cap = int.from_bytes(bytes(src[n - 4:n]), "little")
    if n > (1 << 30):
        while cap < n:                     # ISIZE is the size modulo 2^32
            cap += 1 << 32
    for _ in range(8):
        out = np.empty(cap + 8, dtype=np.uint8)
        got = C.c_uint64()
        rc = lib.nts_gz_inflate_mt(C.c_void_p(src.ctypes.data), n, C.c_void_p(out.ctypes.data), cap, C.byref(got),
                                   int(bool(verify_crc)), int(threads))
        if rc == 0:
            return out[:got.value]
        if rc != 1:
            raise ValueError(lib.nts_last_error().decode() or "corrupt gzip stream")
        cap = max(2 * cap, 6 * n)          # several members (ISIZE names the last one only)
    raise ValueError("gzip stream expands more than 700 times")


def read_fasta(path, threads=0):
    "FASTA file (.gz accepted) -> PackedGenome through the native reader (csrc/nts_fasta.cu)"
    import mmap
    with open(path, "rb") as fh:
        if fh.read(2) == b"\x1f\x8b":
            mm = mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ)        # the compressed bytes are read through a mapping
            try:
                text = inflate_gz(mm, threads)
            finally:
                try:
                    mm.close()
                except BufferError:        # the traceback of a failed inflate may still hold a view: the mapping goes with it
                    pass
            return parse_fasta_bytes(text, path=path, threads=threads)
        size = fh.seek(0, 2)
        if size == 0:
            return parse_fasta_bytes(b"", path=path, threads=threads)
        # plain text: the scanner and the packer read the page cache through a mapping (no 3 GB copy into a bytes object)
        with mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ) as mm:
            return parse_fasta_bytes(mm, path=path, threads=threads)


def read_fastas(paths, threads=0):
    """the files of a run read CONCURRENTLY (one reader per file -- each inflates and scans on its own and shares the
    packer's threads); yields (index, PackedGenome) in the order of `paths` as soon as each is ready, so that the caller
    can upload and insert genome i while the later ones are still being parsed"""
    import os
    from concurrent.futures import ThreadPoolExecutor
    n = len(paths)
    if n == 0:
        return
    total = threads or os.cpu_count() or 1
    workers = min(n, max(1, total // 2))
    with ThreadPoolExecutor(max_workers=workers) as ex:
        futs = [ex.submit(read_fasta, p, max(1, total // workers)) for p in paths]
        for i, f in enumerate(futs):
            yield i, f.result()


def write_fai(packed, out_path):
    "5-column .fai exactly as samtools faidx writes it (goldens: tests/expected_result/*.fai)"
    with open(out_path, "w", encoding="utf-8") as out:
        for row in packed.fai:
            out.write("\t".join(str(x) for x in row) + "\n")
