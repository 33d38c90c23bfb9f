"""Backends that feed ntsynt_b200.synteny.SyntenyEngine in tests.

OracleBackend is TEST-ONLY: it sketches with the CPU oracle and joins with numpy so that the
host-side graph logic can be exercised without a GPU (the driver's `-m "not gpu"` run).  The
product never uses it: ntsynt_b200.pipeline.CudaBackend is the only backend the package ships.
"""
import numpy as np

from oracle import sketch_oracle as so


def numpy_join(tables, order_asm):
    """numpy restatement of nts_graph_build (device kernel iv): dedup within each assembly, G-way
    intersection, vertex numbering by rank in the orienting assembly's filtered list."""
    G = len(tables)
    filt = []
    common = None
    for h1, pos, ctg in tables:
        u, cnt = np.unique(h1, return_counts=True)
        uniq = u[cnt == 1]
        common = uniq if common is None else np.intersect1d(common, uniq, assume_unique=True)
    for h1, pos, ctg in tables:
        keep = np.isin(h1, common)
        filt.append((h1[keep], pos[keep], ctg[keep]))
    V = len(common)
    H = filt[order_asm][0].copy()
    order = np.argsort(H, kind="stable")
    Hs = H[order]
    POS = np.zeros((G, V), dtype=np.uint32)
    CTG = np.zeros((G, V), dtype=np.uint32)
    RANK = np.zeros((G, V), dtype=np.uint32)
    INV = np.zeros((G, V), dtype=np.int64)
    for a, (h1, pos, ctg) in enumerate(filt):
        vid = order[np.searchsorted(Hs, h1)]
        POS[a, vid] = pos
        CTG[a, vid] = ctg
        RANK[a, vid] = np.arange(V)
        INV[a] = vid
    link = np.zeros(V, dtype=np.uint8)
    if V > 1:
        full = np.ones(V - 1, dtype=bool)
        for a in range(G):
            r0, r1 = RANK[a, :-1].astype(np.int64), RANK[a, 1:].astype(np.int64)
            full &= (np.abs(r0 - r1) == 1) & (CTG[a, :-1] == CTG[a, 1:])
        link[:-1] = full
    degree = np.zeros(V, dtype=np.uint8)
    nbsets = [set() for _ in range(V)]
    for a in range(G):
        vid = INV[a]
        same = CTG[a, vid[:-1]] == CTG[a, vid[1:]]
        for x, y in zip(vid[:-1][same], vid[1:][same]):
            nbsets[x].add(int(y)); nbsets[y].add(int(x))
    degree[:] = [len(s) for s in nbsets]
    return H, POS, CTG, RANK, link, degree


def numpy_pairs(POS):
    "per-pair (i, i+1) direction masks and |dpos| spread (graph_links_kernel restated)"
    G, V = POS.shape
    inc = np.zeros(V, dtype=np.uint32)
    dec = np.zeros(V, dtype=np.uint32)
    spread = np.zeros(V, dtype=np.uint32)
    if V > 1:
        d = POS[:, 1:].astype(np.int64) - POS[:, :-1].astype(np.int64)
        for a in range(G):
            inc[:-1] |= ((d[a] > 0).astype(np.uint32) << np.uint32(a))
            dec[:-1] |= ((d[a] < 0).astype(np.uint32) << np.uint32(a))
        ad = np.abs(d)
        spread[:-1] = (ad.max(axis=0) - ad.min(axis=0)).astype(np.uint32)
    return inc, dec, spread


def numpy_edges(RANK, CTG):
    "edge table in build_graph's first-insertion order: list of (u, v, support_mask)"
    G, V = RANK.shape
    seen = {}
    order = []
    for a in range(G):
        inv = np.empty(V, dtype=np.int64)
        inv[RANK[a]] = np.arange(V)
        for r in range(V - 1):
            u, v = int(inv[r]), int(inv[r + 1])
            if CTG[a, u] != CTG[a, v]:
                continue
            key = (min(u, v), max(u, v))
            if key in seen:
                seen[key][2] |= 1 << a
            else:
                seen[key] = [u, v, 1 << a]
                order.append(key)
    return [tuple(seen[k]) for k in order]


class _HostTable(tuple):
    def __new__(cls, t):
        return super().__new__(cls, t)

    def close(self):
        pass


class OracleBackend:
    def __init__(self, fasta_paths, tsv_names, k, fpr=0.025, common=True, lean=False, repeat_bits=None, filter_mode=None):
        """fasta_paths/tsv_names in the engine's processing order (reverse-sorted TSV names).
        lean=True hands the join to the engine in the form the CUDA backend uses (host-ready columns,
        sparse lists, lazily fetched pair masks; ntsynt_b200.device.MinimizerGraph.join_result)."""
        self.k = k
        self.lean = lean
        self.repeat_bits, self.filter_mode = repeat_bits, filter_mode     # bin/ntsynt_synteny.py:172-187
        self.names = list(tsv_names)
        self.records = [so.read_fasta(p) for p in fasta_paths]
        self.contig_names = [[n for n, _ in recs] for recs in self.records]
        self.contig_lengths = [[len(s) for _, s in recs] for recs in self.records]
        self.bits = None
        if common:
            import os
            genomes = [(os.path.basename(p)[:-3] if p.endswith(".gz") else os.path.basename(p), recs)
                       for p, recs in zip(fasta_paths, self.records)]
            self.bits = so.common_bf(genomes, k, fpr)
        self.n_sketch = 0

    def sketch(self, a, w, masks):
        self.n_sketch += 1
        hs, ps, cs = [], [], []
        for c, (_, seq) in enumerate(self.records[a]):
            if masks is not None and len(masks[c][0]):
                buf = bytearray(seq)
                for s, e in zip(masks[c][0], masks[c][1]):
                    s, e = int(s), int(e)
                    buf[s:e] = b"N" * (e - s)
                seq = bytes(buf)
            h1, pos = so.minimize(seq, self.k, w, self.bits, self.repeat_bits if self.filter_mode == "Indexlr" else None)
            if self.filter_mode == "Filter" and len(h1):
                # read_minimizers(tsv, repeat_bf): a minimizer whose k-mer is in the repeat filter is dropped
                m = self.repeat_bits.size * 8
                idx = np.array([so.kmer_hash(seq[int(p_):int(p_) + self.k].upper()) % m for p_ in pos], dtype=np.uint64)
                hit = (self.repeat_bits[(idx >> np.uint64(3)).astype(np.int64)] >> (idx & np.uint64(7)).astype(np.uint8)) & 1
                h1, pos = h1[hit == 0], pos[hit == 0]
            hs.append(h1); ps.append(pos.astype(np.uint32)); cs.append(np.full(len(h1), c, dtype=np.uint32))
        return np.concatenate(hs), np.concatenate(ps), np.concatenate(cs)

    def join(self, tables, order_asm):
        H, POS, CTG, RANK, link, degree = numpy_join(tables, order_asm)
        G, V = RANK.shape
        INV = np.zeros((G, V), dtype=np.uint32)
        for a in range(G):
            INV[a, RANK[a]] = np.arange(V, dtype=np.uint32)
        inc, dec, spread = numpy_pairs(POS)
        self._h_order = np.argsort(H, kind="stable")
        self._h_sorted = H[self._h_order]
        if self.lean == "dev":
            return self._device_form(H, POS, CTG, RANK, INV, link, degree, inc, dec, spread, order_asm)
        if self.lean:
            CI = np.zeros((G, V + 1), dtype=np.int32)
            CD = np.zeros((G, V + 1), dtype=np.int32)
            for a in range(G):
                np.cumsum((inc >> np.uint32(a)) & np.uint32(1), out=CI[a, 1:])
                np.cumsum((dec >> np.uint32(a)) & np.uint32(1), out=CD[a, 1:])

            def host(cap):
                rng = np.random.default_rng(V)                       # rows beyond V hold garbage on the device path too
                Hc = rng.integers(0, 1 << 62, cap).astype(np.uint64); Hc[:V] = H
                P = rng.integers(0, 1 << 30, (G, cap)).astype(np.int64); P[:, :V] = POS
                Cc = rng.integers(0, 50, (G, cap)).astype(np.int32); Cc[:, :V] = CTG
                nbr = rng.integers(-1, 100, (cap, 2)).astype(np.int32)
                conn = rng.integers(0, 2, cap).astype(np.uint8); conn[:V] = link
                ar = np.arange(V, dtype=np.int32)
                nbr[:V, 1] = np.where(link.astype(bool), ar + 1, -1)
                nbr[:V, 0] = -1
                if V > 1:
                    nbr[1:V, 0] = np.where(link[:-1].astype(bool), ar[:-1], -1)
                return Hc, P, Cc, nbr, conn

            def sparse(bp):
                return (np.flatnonzero(link[:max(V - 1, 0)] == 0).astype(np.int64), np.flatnonzero(degree == 3).astype(np.int64),
                        np.flatnonzero(spread[:max(V - 1, 0)] > bp).astype(np.int64))
            return dict(V=V, RANK=RANK, INV=INV, CI=CI, CD=CD, host=host, sparse=sparse,
                        pair_masks=lambda: (inc, dec, spread))
        return dict(H=H, POS=POS, CTG=CTG, RANK=RANK, INV=INV, link=link, degree=degree, incmask=inc, decmask=dec,
                    spread=spread)

    def _device_form(self, H, POS, CTG, RANK, INV, link, degree, inc, dec, spread, order_asm):
        """numpy restatement of the device-resident join result (ntsynt_b200.device.MinimizerGraph.join_result):
        gather / range_sums / neigh / links_nbr / set_links / runs / runs_to_blocks -- the kernels of csrc/nts_graph.cu
        restated for the CPU tests of the host logic"""
        G, V = RANK.shape
        CI = np.zeros((G, V + 1), dtype=np.int64)
        CD = np.zeros((G, V + 1), dtype=np.int64)
        for a in range(G):
            np.cumsum((inc >> np.uint32(a)) & np.uint32(1), out=CI[a, 1:])
            np.cumsum((dec >> np.uint32(a)) & np.uint32(1), out=CD[a, 1:])
        dlink = link.astype(np.uint8).copy()
        cols = {"h1": H, "pos": POS.astype(np.int64), "ctg": CTG.astype(np.int32), "rank": RANK, "inv": INV}
        self.calls = {}

        def count(name):
            self.calls[name] = self.calls.get(name, 0) + 1

        def gather(what, ids):
            count("gather")
            ids = np.asarray(ids, dtype=np.int64)
            assert ((ids >= 0) & (ids < V)).all()
            c = cols[what]
            return c[ids].copy() if c.ndim == 1 else c[:, ids].copy()

        def range_sums(lo, hi):
            count("range_sums")
            lo, hi = np.minimum(lo, V), np.minimum(hi, V)
            return CI[:, hi] - CI[:, lo], CD[:, hi] - CD[:, lo]

        def neigh(cand):
            count("neigh")
            n = len(cand)
            left, right, rk = (np.full((n, G), -1, dtype=np.int64) for _ in range(3))
            for a in range(G):
                r = RANK[a, cand].astype(np.int64)
                rk[:, a] = r
                ok = r > 0
                x = INV[a, r[ok] - 1].astype(np.int64)
                left[ok, a] = np.where(CTG[a, x] == CTG[a, cand[ok]], x, -1)
                ok = r + 1 < V
                x = INV[a, r[ok] + 1].astype(np.int64)
                right[ok, a] = np.where(CTG[a, x] == CTG[a, cand[ok]], x, -1)
            return left, right, rk

        def links_nbr(cap):
            rng = np.random.default_rng(V)
            nbr = rng.integers(-1, 100, (cap, 2)).astype(np.int32)
            conn = rng.integers(0, 2, cap).astype(np.uint8); conn[:V] = link
            ar = np.arange(V, dtype=np.int32)
            nbr[:V, 1] = np.where(link.astype(bool), ar + 1, -1)
            nbr[:V, 0] = -1
            if V > 1:
                nbr[1:V, 0] = np.where(link[:-1].astype(bool), ar[:-1], -1)
            return nbr, conn

        def set_links(idx, val):
            count("set_links")
            dlink[np.asarray(idx, dtype=np.int64)] = val

        def runs():
            count("runs")
            l = dlink[:max(V - 1, 0)].astype(bool)
            prev = np.r_[False, l]            # link[i-1]
            cur = np.r_[l, False]             # link[i]
            return np.flatnonzero(~prev & cur)[:V].astype(np.int64), np.flatnonzero(prev & ~cur).astype(np.int64)

        def sparse(bp):
            return (np.zeros(0, dtype=np.int64), np.flatnonzero(degree == 3).astype(np.int64),
                    np.flatnonzero(spread[:max(V - 1, 0)] > bp).astype(np.int64))

        def runs_to_blocks(starts, ends, bp, m_pct, min_mx):
            count("runs_to_blocks")
            big = np.flatnonzero(spread[:max(V - 1, 0)] > bp)
            out = dict(b_lo=[], b_hi=[], b_plus=[], b_dir=[], r_lo=[], r_hi=[], cuts=[])
            for s_, e_ in zip(starts.tolist(), ends.tolist()):
                assert e_ > s_ and dlink[s_:e_].all()
                x, y = int(POS[order_asm, s_]), int(POS[order_asm, e_])
                if x == y:
                    continue
                d = 1 if x < y else -1
                n = e_ - s_ + 1
                plus, bad = 0, False
                for a in range(G):
                    up, down = int(CI[a, e_] - CI[a, s_]), int(CD[a, e_] - CD[a, s_])
                    i_, d_ = (up, down) if d > 0 else (down, up)
                    if i_ == n - 1:
                        plus |= 1 << a
                    elif d_ == n - 1:
                        pass
                    else:
                        positive = i_ / float(n - 1) * 100
                        if positive >= m_pct:
                            plus |= 1 << a
                        elif 100 - positive >= m_pct:
                            pass
                        else:
                            bad = True
                if bad:
                    out["r_lo"].append(s_); out["r_hi"].append(e_)
                    continue
                cuts = big[(big >= s_) & (big < e_)].tolist()
                lo = s_
                for hi in cuts + [e_]:
                    if hi - lo + 1 >= min_mx:
                        out["b_lo"].append(lo); out["b_hi"].append(hi); out["b_plus"].append(plus); out["b_dir"].append(d)
                    else:
                        out["r_lo"].append(lo); out["r_hi"].append(hi)
                    lo = hi + 1
                out["cuts"].extend(cuts)
            rng = np.random.default_rng(len(starts))           # the device appends in no particular order
            pb, pr = rng.permutation(len(out["b_lo"])), rng.permutation(len(out["r_lo"]))
            res = {k: np.asarray(v, dtype=np.int8 if k == "b_dir" else np.uint32) for k, v in out.items()}
            for k in ("b_lo", "b_hi", "b_plus", "b_dir"):
                res[k] = res[k][pb]
            for k in ("r_lo", "r_hi"):
                res[k] = res[k][pr]
            return res
        def refine_filter(tables, seg_lo, seg_hi, term, x_key, x_vid, iv_start, iv_maxend, iv_off):
            "nts_graph_refine_filter restated with numpy (dedup, block filter, sub-lists, G-way intersection)"
            count("refine_filter")
            assert (np.diff(seg_lo) > 0).all() and (np.diff(term) > 0).all() and (np.diff(x_key.astype(np.float64)) >= 0).all()

            def hit(st, mx, a, b):
                n = np.searchsorted(st, b, side="left")
                res = np.zeros(len(a), dtype=bool)
                ok = n > 0
                res[ok] = mx[n[ok] - 1] > a[ok]
                return res
            kept, n_raw = [], []
            for a, (h1, pos, ctg) in enumerate(tables):
                _, inv_, cnt = np.unique(h1, return_inverse=True, return_counts=True)
                uniq = cnt[inv_] == 1 if len(h1) else np.zeros(0, dtype=bool)
                h1, pos, ctg = h1[uniq], pos[uniq].astype(np.int64), ctg[uniq].astype(np.int64)
                n_raw.append(len(h1))
                vid = self.lookup(h1).astype(np.int64)
                vid[vid == 0xFFFFFFFF] = -1
                if len(x_key):
                    j = np.searchsorted(x_key, h1)
                    j[j >= len(x_key)] = 0
                    m = (vid < 0) & (x_key[j] == h1)
                    vid[m] = x_vid[j[m]]
                internal = np.zeros(len(h1), dtype=bool)
                if len(seg_lo):
                    k = np.searchsorted(seg_lo, vid, side="right") - 1
                    ok = (vid >= 0) & (k >= 0)
                    internal[ok] = seg_hi[k[ok]] >= vid[ok]
                    internal &= ~np.isin(vid, term)
                st, mx = iv_start[int(iv_off[a]):int(iv_off[a + 1])], iv_maxend[int(iv_off[a]):int(iv_off[a + 1])]
                key = (ctg << np.int64(40)) + pos
                keep = ~internal & ~hit(st, mx, key, key + 1)
                kh, kp, kc, kk = h1[keep], pos[keep], ctg[keep], key[keep]
                cut = np.ones(len(kh), dtype=bool)
                if len(kh) > 1:
                    same = kc[1:] == kc[:-1]
                    cut[1:] = ~same | (same & hit(st, mx, kk[:-1], kk[1:]))
                kept.append((kh, kp, kc, np.cumsum(cut) - 1))
            allk = np.concatenate([x[0] for x in kept])
            uk, ck = np.unique(allk, return_counts=True) if len(allk) else (allk, allk)
            common = uk[ck == G] if len(allk) else allk
            out = []
            for kh, kp, kc, sub in kept:
                ok = np.isin(kh, common)
                out.append((kh[ok], kp[ok], kc[ok], sub[ok]))
            return n_raw, out
        return dict(V=V, gather=gather, range_sums=range_sums, neigh=neigh, links_nbr=links_nbr, set_links=set_links, runs=runs,
                    runs_to_blocks=runs_to_blocks, sparse=sparse, refine_filter=refine_filter)

    def sketch_table(self, a, w, masks):
        "what stays on the device in the CUDA backend: here simply the host table"
        return _HostTable(self.sketch(a, w, masks))

    def lookup(self, keys):
        out = np.full(len(keys), 0xFFFFFFFF, dtype=np.uint32)
        if len(self._h_sorted):
            i = np.searchsorted(self._h_sorted, keys)
            i[i >= len(self._h_sorted)] = 0
            hit = self._h_sorted[i] == keys
            out[hit] = self._h_order[i[hit]]
        return out
