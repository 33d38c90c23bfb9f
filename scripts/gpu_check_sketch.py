#!/usr/bin/env python3
"""Quick GPU-side parity + timing check of kernels (i)-(iii) against the oracle and goldens."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ntsynt_b200 import device, fasta  # noqa: E402
from oracle import sketch_oracle as so  # noqa: E402

DEMO = os.path.join(ROOT, "tests", "golden", "_ref_demo")


def parse_golden(path):
    out = {}
    for line in open(path):
        name, rest = line.rstrip("\n").split("\t")
        toks = rest.split(" ") if rest else []
        out[name] = ([int(t.split(":")[0]) for t in toks], [int(t.split(":")[1]) for t in toks])
    return out


def main():
    k, w = 24, 1000
    ctx = device.Context(0)
    names = ["celegans-chrII-III.fa", "celegans-chrII-III.A.fa"]
    t = time.time()
    packed = [fasta.read_fasta(os.path.join(DEMO, n + ".gz")) for n in names]
    print("ingest s", time.time() - t)
    gens = [ctx.upload(p) for p in packed]
    recs = [so.read_fasta(os.path.join(DEMO, n + ".gz")) for n in names]
    # (i) hashes
    h0, valid = ctx.hash_contig(gens[0], 0, k)
    oh, ov = so.hash_seq(recs[0][0][1], k)
    print("hash parity:", np.array_equal(valid, ov), np.array_equal(h0[ov == 1], oh[ov == 1]), len(h0))
    # (iii) bloom
    order = sorted(range(2), key=lambda i: names[i])
    nbytes = device.BloomFilter.size_for(packed[order[0]].total_bases, 0.025)
    print("bf bytes", nbytes, so.bf_bytes(packed[order[0]].total_bases, 0.025))
    bfs = []
    for g in gens:
        bf = ctx.bloom(nbytes)
        ctx.timer_start()
        bf.insert_genome(g, k)
        print("  insert ms", ctx.timer_stop())
        bfs.append(bf)
    ob = [so.genome_bits(r, k, nbytes) for r in recs]
    for i in range(2):
        print("bits parity", i, np.array_equal(bfs[i].to_numpy(), ob[i]), bfs[i].popcount(), int(np.unpackbits(ob[i]).sum()))
    ctx.timer_start()
    bfs[0].iand(bfs[1])
    print("  and ms", ctx.timer_stop())
    common = ob[0] & ob[1]
    print("common parity", np.array_equal(bfs[0].to_numpy(), common))
    # (ii) sketch vs golden
    for gi, n in enumerate(names):
        ctx.timer_start()
        mx = ctx.sketch(gens[gi], k, w, common=bfs[0])
        ms = ctx.timer_stop()
        h1, pos, ctg = mx.to_numpy()
        gold = parse_golden(os.path.join(DEMO, "expected_result", f"{n}.k{k}.w{w}.tsv"))
        ok = True
        for c, cname in enumerate(gens[gi].names):
            sel = ctg == c
            gh, gp = gold[cname]
            good = (len(gh) == int(sel.sum())) and np.array_equal(h1[sel], np.array(gh, dtype=np.uint64)) and \
                np.array_equal(pos[sel], np.array(gp, dtype=np.uint32))
            if not good:
                ok = False
                print("   MISMATCH contig", c, "got", int(sel.sum()), "want", len(gh))
                gpa = np.array(gp, dtype=np.int64)
                mine = pos[sel].astype(np.int64)
                n_ = min(len(gpa), len(mine))
                d = np.nonzero(gpa[:n_] != mine[:n_])[0]
                if len(d):
                    j = d[0]
                    print("   first diff at", j, "mine", mine[max(0, j - 2):j + 3], "gold", gpa[max(0, j - 2):j + 3])
        print(f"sketch {n}: parity={ok} count={len(h1)} ms={ms:.3f}")
    # masked + small w vs oracle
    rng = np.random.default_rng(1)
    for w2 in (100, 10, 250, 1):
        g = gens[1]
        masks = []
        seqs = []
        for c in range(g.n_contigs):
            L = int(g.lengths[c])
            n_iv = 200
            s = np.sort(rng.integers(0, L, n_iv))
            e = np.minimum(s + rng.integers(1, 200000, n_iv), L)
            # make disjoint & sorted
            keep_s, keep_e = [], []
            last = 0
            for a, b in zip(s, e):
                a = max(int(a), last)
                if a < b:
                    keep_s.append(a); keep_e.append(int(b)); last = int(b)
            masks.append((np.array(keep_s, dtype=np.uint64), np.array(keep_e, dtype=np.uint64)))
            seq = bytearray(recs[1][c][1])
            for a, b in zip(keep_s, keep_e):
                seq[a:b] = b"N" * (b - a)
            seqs.append(bytes(seq))
        ctx.timer_start()
        mx = ctx.sketch(g, k, w2, common=bfs[0], masks=masks)
        ms = ctx.timer_stop()
        h1, pos, ctg = mx.to_numpy()
        ok = True
        for c in range(g.n_contigs):
            oh1, opos = so.minimize(seqs[c], k, w2, common)
            sel = ctg == c
            good = np.array_equal(h1[sel], oh1) and np.array_equal(pos[sel].astype(np.uint64), opos)
            if not good:
                ok = False
                print("   masked MISMATCH w", w2, "contig", c, int(sel.sum()), len(oh1))
        print(f"masked sketch w={w2}: parity={ok} count={len(h1)} ms={ms:.3f}")
    # no-filter sketch
    mx = ctx.sketch(gens[0], k, 500)
    h1, pos, ctg = mx.to_numpy()
    ok = True
    for c in range(gens[0].n_contigs):
        oh1, opos = so.minimize(recs[0][c][1], k, 500, None)
        sel = ctg == c
        ok &= np.array_equal(h1[sel], oh1) and np.array_equal(pos[sel].astype(np.uint64), opos)
    print("nofilter sketch w=500 parity", ok, len(h1))
    print("launches", ctx.launches)


if __name__ == "__main__":
    main()
