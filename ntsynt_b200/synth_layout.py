"""Synthetic genome workloads (SURVEY.md 8d recipe) -- the host-side layout: ancestor contig sizes and N runs, and the
segment table of every genome (inversions, translocations, indels as copies of ancestor intervals).  Pure numpy: no
device library is loaded by importing this module (bench.py's CPU arm and the oracle's generator use it too).
"""
import numpy as np

# chr1..22, X, Y in Mbp (GRCh38, rounded) -- only the proportions matter
HUMAN_MBP = [248, 242, 198, 190, 182, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64, 47, 51,
             156, 57]
SEG_DTYPE = np.dtype([("dst_contig", "<u4"), ("anc_contig", "<i4"), ("dst_start", "<u8"), ("anc_start", "<u8"),
                      ("len", "<u8"), ("strand", "<i4"), ("pad", "<u4")])


def contig_names(n):
    base = [f"chr{i}" for i in range(1, 23)] + ["chrX", "chrY"]
    return [base[i] if i < len(base) else f"chrUn{i}" for i in range(n)]


def ancestor_layout(total_bp, n_contigs=24):
    w = np.array([HUMAN_MBP[i % len(HUMAN_MBP)] for i in range(n_contigs)], dtype=np.float64)
    lens = np.maximum((w / w.sum() * total_bp).astype(np.int64), 1000)
    return lens


def ancestor_nruns(seed, lengths, total_ref=3.0e9, n_small=40):
    "per contig: n_small N runs of 1-50 kbp and one 3 Mbp 'centromere', scaled to the genome size"
    rng = np.random.default_rng([seed, 0xA11])
    scale = float(np.sum(lengths)) / total_ref
    out = []
    for L in lengths:
        L = int(L)
        runs = []
        big = max(int(3e6 * scale), 50)
        n_sm = n_small if L > 200000 else max(int(n_small * L / 200000), 1)
        lens = np.concatenate([[big], np.maximum((rng.uniform(1e3, 5e4, n_sm) * max(scale, 0.02)).astype(np.int64), 20)])
        starts = np.sort(rng.integers(0, max(L - 1, 1), len(lens)))
        last = 0
        for s, n in zip(starts, lens):
            s = max(int(s), last + 100)
            e = min(s + int(n), L - 100)
            if e > s:
                runs.append((s, e - s))
                last = e
        out.append(runs)
    return out


# ---- piece algebra: a piece is (anc_contig, anc_start_of_first_output_base, len, strand)
def _slice(piece, off, n):
    c, s, _, st = piece
    return (c, s + off if st > 0 else s - off, n, st)


def _split_at(pieces, x):
    "split a piece list at output offset x -> (left, right)"
    left, right, pos = [], [], 0
    for p in pieces:
        n = p[2]
        if pos + n <= x:
            left.append(p)
        elif pos >= x:
            right.append(p)
        else:
            left.append(_slice(p, 0, x - pos))
            right.append(_slice(p, x - pos, n - (x - pos)))
        pos += n
    return left, right


def _revcomp(pieces):
    out = []
    for c, s, n, st in reversed(pieces):
        if c < 0:
            out.append((c, s, n, st))
        elif st > 0:
            out.append((c, s + n - 1, n, -1))
        else:
            out.append((c, s - n + 1, n, 1))
    return out


def _plen(pieces):
    return sum(p[2] for p in pieces)


def _apply_nruns(pieces, anc_nruns):
    out = []
    for p in pieces:
        c, s, n, st = p
        if c < 0 or not anc_nruns[c]:
            out.append(p)
            continue
        lo, hi = (s, s + n) if st > 0 else (s - n + 1, s + 1)       # ancestor interval [lo, hi)
        cuts = []
        for rs, rl in anc_nruns[c]:
            a, b = max(rs, lo), min(rs + rl, hi)
            if a < b:
                cuts.append((a, b))
        if not cuts:
            out.append(p)
            continue
        cur = lo
        parts = []
        for a, b in cuts:
            if a > cur:
                parts.append((cur, a, False))
            parts.append((a, b, True))
            cur = b
        if cur < hi:
            parts.append((cur, hi, False))
        if st < 0:
            parts = parts[::-1]
        for a, b, is_n in parts:
            if is_n:
                out.append((-2, 0, b - a, 1))
            else:
                out.append((c, a if st > 0 else b - 1, b - a, st))
    return out


def _indels(rng, pieces, rate, mean_len=3.0):
    "vectorised indels over one contig's piece list -> (anc_contig, anc_start, len, strand) arrays"
    pc = np.array([p[0] for p in pieces], dtype=np.int64)
    ps = np.array([p[1] for p in pieces], dtype=np.int64)
    pn = np.array([p[2] for p in pieces], dtype=np.int64)
    pst = np.array([p[3] for p in pieces], dtype=np.int64)
    cs = np.concatenate([[0], np.cumsum(pn)])
    L = int(cs[-1])
    n_ev = rng.poisson(L * rate) if rate > 0 else 0
    if n_ev == 0:
        return pc, ps, pn, pst
    x = np.sort(rng.integers(1, max(L - 1, 2), n_ev))
    x = x[pc[np.searchsorted(cs, x, side="right") - 1] >= 0]      # no indels inside N runs
    n_ev = len(x)
    ln = rng.geometric(1.0 / mean_len, n_ev).astype(np.int64)
    is_del = rng.random(n_ev) < 0.5
    dx, dl = x[is_del], ln[is_del]
    ix, il = x[~is_del], ln[~is_del]
    dend = np.minimum(dx + dl, L)
    B = np.unique(np.concatenate([cs, dx, dend, ix]))
    B = B[(B >= 0) & (B <= L)]
    a, b = B[:-1], B[1:]
    pi = np.searchsorted(cs, a, side="right") - 1
    # coverage by deletions
    cov = np.zeros(len(B), dtype=np.int64)
    np.add.at(cov, np.searchsorted(B, dx), 1)
    np.add.at(cov, np.searchsorted(B, dend), -1)
    deleted = np.cumsum(cov)[:-1] > 0
    off = a - cs[pi]
    seg_c = pc[pi]
    seg_s = np.where(pst[pi] > 0, ps[pi] + off, ps[pi] - off)
    seg_n = b - a
    seg_st = pst[pi]
    keep = ~deleted
    # merge kept intervals and insertions, ordered by (position, insertions first)
    key_pos = np.concatenate([a[keep], ix])
    key_kind = np.concatenate([np.ones(keep.sum(), dtype=np.int64), np.zeros(len(ix), dtype=np.int64)])
    order = np.lexsort((key_kind, key_pos))
    oc = np.concatenate([seg_c[keep], np.full(len(ix), -1, dtype=np.int64)])[order]
    os_ = np.concatenate([seg_s[keep], np.zeros(len(ix), dtype=np.int64)])[order]
    on = np.concatenate([seg_n[keep], il])[order]
    ost = np.concatenate([seg_st[keep], np.ones(len(ix), dtype=np.int64)])[order]
    return oc, os_, on, ost


def genome_segments(anc_seed, genome_index, anc_lengths, anc_nruns, divergence_pct, n_inv=50, n_trans=20,
                    total_ref=3.0e9):
    """segment table of genome `genome_index`: substitutions d/200 per base happen on the device;
    here: n_inv inversions (10 kbp-5 Mbp), n_trans translocations (100 kbp-2 Mbp), indels at rate d/2000
    with geometric lengths (mean 3).  Sizes scale with the genome size."""
    rng = np.random.default_rng([anc_seed, 0x6E0, genome_index])
    anc_lengths = [int(x) for x in anc_lengths]
    scale = sum(anc_lengths) / total_ref
    contigs = [[(c, 0, L, 1)] for c, L in enumerate(anc_lengths)]
    prob = np.array(anc_lengths, dtype=np.float64) / sum(anc_lengths)
    for _ in range(n_inv):
        c = int(rng.choice(len(contigs), p=prob))
        L = _plen(contigs[c])
        n = int(min(max(rng.uniform(1e4, 5e6) * scale, 200), L // 3))
        if n < 50:
            continue
        a = int(rng.integers(0, L - n))
        left, rest = _split_at(contigs[c], a)
        mid, right = _split_at(rest, n)
        contigs[c] = left + _revcomp(mid) + right
    for _ in range(n_trans):
        c1 = int(rng.choice(len(contigs), p=prob))
        c2 = int(rng.choice(len(contigs), p=prob))
        L1 = _plen(contigs[c1])
        n = int(min(max(rng.uniform(1e5, 2e6) * scale, 500), L1 // 4))
        if n < 100 or c1 == c2:
            continue
        a = int(rng.integers(0, L1 - n))
        left, rest = _split_at(contigs[c1], a)
        mid, right = _split_at(rest, n)
        contigs[c1] = left + right
        p = int(rng.integers(0, _plen(contigs[c2])))
        l2, r2 = _split_at(contigs[c2], p)
        contigs[c2] = l2 + mid + r2
    rows = []
    lengths = []
    for c, pieces in enumerate(contigs):
        pieces = _apply_nruns(pieces, anc_nruns)
        oc, os_, on, ost = _indels(rng, pieces, divergence_pct / 2000.0)
        seg = np.zeros(len(oc), dtype=SEG_DTYPE)
        seg["dst_contig"] = c
        seg["anc_contig"] = oc
        seg["anc_start"] = np.where(oc >= 0, os_, 0).astype(np.uint64)
        seg["len"] = on
        seg["strand"] = ost
        seg["dst_start"] = np.concatenate([[0], np.cumsum(on)[:-1]])
        rows.append(seg)
        lengths.append(int(on.sum()))
    return np.array(lengths, dtype=np.uint64), np.concatenate(rows)


class Layout:
    "G synthetic genomes derived from one ancestor: sizes, N runs and segment tables (host only)"

    def __init__(self, n_genomes, genome_bp, divergence_pct, seed=20260117, n_contigs=24, n_repeat_fam=64,
                 repeat_slot_prob=0.26, n_inv=50, n_trans=20):
        self.G, self.genome_bp, self.d, self.seed = n_genomes, int(genome_bp), float(divergence_pct), seed
        self.n_repeat_fam, self.repeat_slot_prob = n_repeat_fam, repeat_slot_prob
        self.anc_lengths = ancestor_layout(self.genome_bp, n_contigs)
        self.anc_nruns = ancestor_nruns(seed, self.anc_lengths)
        self.n_inv, self.n_trans = n_inv, n_trans
        self.names = contig_names(n_contigs)
        self._segs = {}

    def file_name(self, g):
        return f"synth_g{g}.fa"

    def segments(self, g):
        if g not in self._segs:
            self._segs[g] = genome_segments(self.seed, g, self.anc_lengths, self.anc_nruns, self.d, self.n_inv,
                                            self.n_trans)
        return self._segs[g]
