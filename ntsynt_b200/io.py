"""File formats at the drop-in boundary: Bloom filter files, indexlr sketch TSVs, .mx.dot."""
import struct

import numpy as np

BF_MAGIC = b"NTSB200BF1\n"
# btllib's KmerBloomFilter::save as remembered from btllib >= 1.5 (bloom_filter.hpp): a cpptoml table named after the
# signature with the keys below, the line "[HeaderEnd]", then the raw bit array.  NOT pinned by any fixture of the
# reference: there is no .bf file in the tree and btllib is not installable here.
BTL_KMER_SIGNATURE = "[BTLKmerBloomFilter_v6]"
BTL_HASH_FN = "ntHash_v2"


def save_bf(path, bloom, k, fmt="btllib"):
    """<prefix>.bf written by ntsynt_make_common_bf (src/ntsynt_make_common_bf.cpp:164) / ntsynt_make_repeat_bfs.py.
    fmt "btllib": btllib's KmerBloomFilter file layout (unpinned, see above) so that a stock `indexlr -s` can read it;
    fmt "native": our own container (magic, uint64 byte count, uint32 k).  The payload is the same in both: the raw
    bit array, byte = idx >> 3, bit = idx & 7 (btllib's in-memory layout)."""
    bits = bloom.to_numpy() if hasattr(bloom, "to_numpy") else np.asarray(bloom, dtype=np.uint8)
    with open(path, "wb") as fh:
        if fmt == "native":
            fh.write(BF_MAGIC)
            fh.write(struct.pack("<QI", bits.size, int(k)))
        elif fmt == "btllib":
            fh.write((f"{BTL_KMER_SIGNATURE}\nbytes = {bits.size}\nhash_num = 1\nhash_fn = \"{BTL_HASH_FN}\"\nk = {int(k)}\n"
                      "[HeaderEnd]\n").encode())
        else:
            raise ValueError(f"unknown Bloom filter file format {fmt!r}")
        fh.write(bits.tobytes())


def load_bf_bytes(path):
    "(bit array, k) from either file layout save_bf writes (btllib files: any key order, blank lines allowed)"
    with open(path, "rb") as fh:
        head = fh.read(len(BF_MAGIC))
        if head == BF_MAGIC:
            n, k = struct.unpack("<QI", fh.read(12))
        elif head.startswith(b"[BTL"):
            fh.seek(0)
            sig = fh.readline().decode().strip()
            if not (sig.startswith("[BTLKmerBloomFilter_v") or sig.startswith("[BTLBloomFilter_v")):
                raise ValueError(f"{path}: unsupported btllib Bloom filter signature {sig}")
            meta = {}
            while True:
                line = fh.readline()
                if not line:
                    raise ValueError(f"{path}: truncated btllib Bloom filter header")
                line = line.decode().strip()
                if line == "[HeaderEnd]":
                    break
                if "=" in line:
                    key, val = (x.strip() for x in line.split("=", 1))
                    meta[key] = val.strip('"')
            if int(meta.get("hash_num", "1")) != 1:
                raise ValueError(f"{path}: only Bloom filters with one hash function are supported (ntSynt uses 1)")
            n, k = int(meta["bytes"]), int(meta.get("k", 0))
        else:
            raise ValueError(f"{path}: not a Bloom filter file this library can read")
        bits = np.frombuffer(fh.read(n), dtype=np.uint8)
        if bits.size != n:
            raise ValueError(f"{path}: truncated Bloom filter file")
    return bits, k


def load_bf(ctx, path):
    bits, k = load_bf_bytes(path)
    return ctx.bloom(bits.size).from_numpy(bits), k


def write_sketch_tsv(out, packed, table_arrays, k, with_seq=True, with_pos=True):
    """indexlr --long [--pos] [--seq] output: one line per record in input order,
    `<id>\\t<h1>:<pos>:<kmer> ...`; records without minimizers give `<id>\\t`."""
    h1, pos, ctg = table_arrays
    bounds = np.searchsorted(ctg, np.arange(len(packed.names) + 1))
    for c, name in enumerate(packed.names):
        a, b = int(bounds[c]), int(bounds[c + 1])
        toks = []
        if b > a:
            text = packed.contig_text(c) if with_seq else None
            for h, p in zip(h1[a:b].tolist(), pos[a:b].tolist()):
                t = str(h)
                if with_pos:
                    t += f":{p}"
                if with_seq:
                    t += ":" + text[p:p + k].decode()
                toks.append(t)
        out.write(name + "\t" + " ".join(toks) + "\n")


def read_sketch_tsv(path):
    """parse an indexlr TSV -> (contig_names, h1 u64[], pos u32[], contig u32[]); every record line gets a contig
    index (subprojects/ntJoin/bin/ntjoin_utils.py:167-193 skips the empty ones later)."""
    names, hs, ps, cs = [], [], [], []
    with open(path, "r", encoding="utf-8") as fh:
        for line in fh:
            parts = line.rstrip("\n").split("\t")
            names.append(parts[0])
            if len(parts) > 1 and parts[1].strip():
                toks = parts[1].strip().split(" ")
                c = len(names) - 1
                for t in toks:
                    f = t.split(":")
                    hs.append(int(f[0])); ps.append(int(f[1])); cs.append(c)
    return names, np.array(hs, dtype=np.uint64), np.array(ps, dtype=np.uint32), np.array(cs, dtype=np.uint32)


DOT_COLOURS = ["red", "green", "blue", "purple", "orange", "turquoise", "pink", "yellow", "orchid", "salmon"]


def write_mx_dot(path, asm_names, contig_names, H, POS, CTG, edges):
    """<prefix>.mx.dot as print_graph writes it (subprojects/ntJoin/bin/ntjoin.py:23-65).  Vertex order is
    arbitrary in the reference (Python set order); edges keep build_graph's order.  edges = (u, v, support)."""
    G = len(asm_names)
    colours = DOT_COLOURS if G <= len(DOT_COLOURS) else ["red"] * G
    u, v, sup = edges
    with open(path, "w", encoding="utf-8") as out:
        out.write("graph G {\n")
        for i in range(len(H)):
            name = str(int(H[i]))
            lab = "\n".join(f"{asm_names[a]}_{(contig_names[a][int(CTG[a, i])], int(POS[a, i]))}" for a in range(G))
            out.write(f"\"{name}\" [label=\"{name}\n{lab}\"]\n")
        for a, b, s in zip(u.tolist(), v.tolist(), sup.tolist()):
            w = bin(s).count("1")
            if w == 1:
                col = colours[s.bit_length() - 1]
            elif w == 2:
                col = "lightgrey"
            else:
                col = "black"
            out.write(f"\"{int(H[a])}\" --\"{int(H[b])}\" [weight={w} color={col}]\n")
        out.write("}\n")
