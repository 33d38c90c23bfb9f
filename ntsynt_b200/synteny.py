"""Graph stage of the B200 path: minimizer graph -> collinear paths -> synteny blocks, with ntSynt's
window-refinement rounds.  Host side of kernel family (iv).

This is NOT a translation of the reference's python-igraph code.  The reference keeps the whole
graph as igraph objects keyed by decimal strings; here the bulk lives in flat arrays produced by the
device join (ntsynt_b200/csrc/nts_graph.cu):

  * vertices are numbered by their rank in the ORIENTING assembly's filtered minimizer list, so an
    edge supported by all G assemblies is always (i, i+1): the weight-filtered graph of the reference
    (max degree 2) is a `nbr[V, 2]` array, chains are runs of consecutive ids, and only the few
    irregular places (bubbles, refinement splices, erosion) are handled one by one on the host in the
    reference's order;
  * per-assembly positions / contigs are `POS[G, V]`, `CTG[G, V]` columns; orientation tallies, indel
    splits and the >= 4-minimizer filter are vectorised segment operations over concatenated paths.

Reference semantics followed (file:line are in /root/reference):
  bin/ntsynt_synteny.py        :66-106 find_synteny_blocks      :117-157 beds + masks
                               :194-226 find_mx_in_blocks        :256-280 filter_minimizers_synteny_blocks
                               :282-290 update_list_mx_info      :292-362 flag overlaps / erode / refine
                               :364-426 indels, >=4 filter       :428-472 merge_collinear_blocks
                               :476-530 refine_block_coordinates :566-590 run_graph_simplification
                               :593-647 main_synteny
  subprojects/ntJoin/bin/ntjoin_utils.py :70-80 incident-weight guard :83-141 build_graph
                               :152-165 filter_minimizers        :167-193 read_minimizers
  subprojects/ntJoin/bin/ntjoin.py :78-87 filter_graph_global    :89-102 determine_source_vertex
                               :114-151 find_paths
  bin/synteny_block.py         :48-65 orientation rule           :72-85 row format  :102-109 sort order
  bin/assembly_block.py        :17-27 block coordinates
Only `-n` = number of assemblies (the pipeline's only setting, bin/ntsynt_run_pipeline.smk:95-103)
is supported; every assembly weight is 1 (bin/ntsynt_synteny.py:32).
"""
import datetime
import re
import sys
from collections import defaultdict

import numpy as np

FA_TSV_RE = re.compile(r"^(\S+)\.k\d+\.w\d+.tsv")


_NO_U64 = np.zeros(0, dtype=np.uint64)


def _log(*a):
    print(datetime.datetime.today(), ":", *a, file=sys.stdout, flush=True)


class Block:
    """one synteny block: a path of vertices with one contig and one orientation per assembly.
    The path is a list of segments (lo, hi, dir): base-vertex ids lo..hi joined by (i, i+1) links,
    traversed upwards (dir = +1) or downwards (dir = -1); a lone vertex is (v, v, 1)."""
    __slots__ = ("segs", "ctg", "ori", "first_id", "last_id", "first_pos", "last_pos", "n", "broken_reason")

    def __init__(self, segs, ctg, ori, first_id, last_id, first_pos, last_pos, n):
        self.segs = segs
        self.ctg = ctg                # [G] contig index per assembly
        self.ori = ori                # [G] '+', '-'
        self.first_id, self.last_id = first_id, last_id
        self.first_pos = first_pos    # [G] position of the first minimizer
        self.last_pos = last_pos      # [G] position of the last minimizer
        self.n = n                    # number of minimizers
        self.broken_reason = None

    def start(self, a):               # bin/assembly_block.py:17-19
        return min(int(self.first_pos[a]), int(self.last_pos[a]))

    def end(self, a, k):              # bin/assembly_block.py:21-23
        return max(int(self.first_pos[a]), int(self.last_pos[a])) + k


def seg_first(seg):
    return seg[0] if seg[2] > 0 else seg[1]


def seg_last(seg):
    return seg[1] if seg[2] > 0 else seg[0]


def seg_ids(seg):
    lo, hi, d = seg
    return np.arange(lo, hi + 1, dtype=np.int64) if d > 0 else np.arange(hi, lo - 1, -1, dtype=np.int64)


class IntervalIndex:
    "half-open interval overlap queries over one (assembly, contig): exists [s,e) with s < b and a < e"

    def __init__(self, starts, ends):
        order = np.argsort(starts, kind="stable")
        self.starts = np.asarray(starts, dtype=np.int64)[order]
        self.maxend = np.maximum.accumulate(np.asarray(ends, dtype=np.int64)[order])

    def overlaps(self, a, b):
        "vectorised: a, b int64 arrays"
        n = np.searchsorted(self.starts, b, side="left")
        ok = n > 0
        res = np.zeros(len(a), dtype=bool)
        res[ok] = self.maxend[n[ok] - 1] > a[ok]
        return res


class _LazyRow:
    "one assembly's row of a LazyMatrix (read-only view)"

    def __init__(self, mat, a):
        self.mat, self.a = mat, a

    def __getitem__(self, idx):
        return self.mat[self.a, idx]


class LazyMatrix:
    """A [G, V] column of the device vertex table (positions or contigs), read where the host looks through
    `fetch(ids) -> [G, n]` (nts_graph_gather).  What later rounds write lives on the host: vertices added after
    round 0 in dense rows indexed by id - V0, the few round-0 entries that update_list_mx_info overwrites in a
    per-assembly dict (a byte mask over the ids says where to look).  Supports the index forms the engine uses:
    M[:, ids], M[a, ids], M[a] (a row view), M[a, ids] = vals; ids may be an int (that axis is then dropped)."""

    def __init__(self, fetch, G, V0, dtype):
        self.fetch, self.G, self.V0, self.dtype = fetch, G, V0, np.dtype(dtype)
        self.over = [dict() for _ in range(G)]          # overwritten round-0 entries
        self.ov_mask = np.zeros(max(V0, 1), dtype=bool)
        self.extra = np.zeros((G, 1024), dtype=self.dtype)   # vertices >= V0
        self._memo = {}                         # small read cache for single-vertex queries (erosion walks)

    def touched_base(self):
        "sorted ids < V0 whose entry differs (or may differ) from the device's in some assembly"
        keys = set()
        for d in self.over:
            keys.update(d)
        return np.array(sorted(keys), dtype=np.int64)

    def _get(self, rows, ids):
        ids = np.asarray(ids, dtype=np.int64)
        out = np.empty((self.G, len(ids)), dtype=self.dtype)
        base = ids < self.V0
        if base.all():
            bi = ids
            if len(ids):
                out[:] = self.fetch(ids)
        else:
            bi = ids[base]
            if len(bi):
                out[:, base] = self.fetch(bi)
            x = ids[~base] - self.V0
            self._reserve(int(x.max()) + 1)
            out[:, ~base] = self.extra[:, x]                # (zero until written, like the dense form's spare rows)
        if len(bi) and any(self.over):
            hit = np.flatnonzero(self.ov_mask[bi])
            if len(hit):
                at = np.flatnonzero(base)[hit] if len(bi) != len(ids) else hit
                for a in rows:
                    d = self.over[a]
                    if d:
                        for i, v in zip(at.tolist(), bi[hit].tolist()):
                            got = d.get(v)
                            if got is not None:
                                out[a, i] = got
        return out

    def _reserve(self, need):
        if need > self.extra.shape[1]:
            grown = np.zeros((self.G, max(need, 2 * self.extra.shape[1])), dtype=self.dtype)
            grown[:, :self.extra.shape[1]] = self.extra
            self.extra = grown

    def prefetch(self, ids):
        "cache the columns of `ids` for later single-vertex reads"
        ids = np.unique(np.asarray(ids, dtype=np.int64))
        ids = ids[ids >= 0]
        if len(ids):
            cols = self._get(range(self.G), ids)
            self._memo.update(zip(ids.tolist(), cols.T))

    def __getitem__(self, key):
        if not isinstance(key, tuple):
            return _LazyRow(self, int(key))
        r, ids = key
        scalar = np.ndim(ids) == 0
        if scalar:
            got = self._memo.get(int(ids))
            col = got if got is not None else self._get(range(self.G), [int(ids)])[:, 0]
            return col.copy() if isinstance(r, slice) else col[int(r)]
        if isinstance(r, slice):
            return self._get(range(self.G), ids)
        return self._get([int(r)], ids)[int(r)]

    def __setitem__(self, key, vals):
        a, ids = key
        a = int(a)
        ids = np.atleast_1d(np.asarray(ids, dtype=np.int64))
        vals = np.broadcast_to(np.asarray(vals, dtype=self.dtype), ids.shape)
        new = ids >= self.V0
        if new.any():
            x = ids[new] - self.V0
            self._reserve(int(x.max()) + 1)
            self.extra[a, x] = vals[new]
        if not new.all():
            b = ids[~new]
            self.over[a].update(zip(b.tolist(), vals[~new].tolist()))
            self.ov_mask[b] = True
        if self._memo:
            for v in ids.tolist():
                self._memo.pop(v, None)


class LazyVector:
    "the h1 column, same idea: device gather for round-0 vertices, host dict for the vertices added later"

    def __init__(self, fetch, V0):
        self.fetch, self.V0 = fetch, V0
        self.extra = {}

    def __getitem__(self, ids):
        scalar = np.ndim(ids) == 0
        idx = np.atleast_1d(np.asarray(ids, dtype=np.int64))
        out = np.zeros(len(idx), dtype=np.uint64)
        base = idx < self.V0
        if base.any():
            out[base] = self.fetch(idx[base])
        for i in np.flatnonzero(~base).tolist():
            out[i] = self.extra[int(idx[i])]
        return out[0] if scalar else out

    def __setitem__(self, ids, vals):
        self.extra.update(zip(np.atleast_1d(np.asarray(ids, dtype=np.int64)).tolist(),
                              np.atleast_1d(np.asarray(vals, dtype=np.uint64)).tolist()))


class EdgeBirth:
    """(u, v) -> (round, sequence number) of the edges added by the refinement rounds, in build_graph's insertion
    order (an edge's id order in the reference's igraph object).  Kept as sorted key arrays per round; looked up one
    edge at a time, and only for the few old edges the simplification has to order."""

    def __init__(self):
        self.rounds = []            # (round_no, sorted packed keys, sequence numbers in that order)

    def add_round(self, round_no, eu, ev):
        if len(eu):
            key = (np.minimum(eu, ev).astype(np.int64) << np.int64(32)) | np.maximum(eu, ev).astype(np.int64)
            o = np.argsort(key, kind="stable")
            self.rounds.append((round_no, key[o], o))

    def get(self, key):
        u, v = key
        pk = (min(u, v) << 32) | max(u, v)
        for round_no, keys, seq in self.rounds:
            i = int(np.searchsorted(keys, pk))
            if i < len(keys) and int(keys[i]) == pk:
                return (round_no, int(seq[i]))
        return None

    def __contains__(self, key):
        return self.get(key) is not None

    def __getitem__(self, key):
        got = self.get(key)
        if got is None:
            raise KeyError(key)
        return got


class SyntenyEngine:
    """backend must provide:
         names[a]            TSV-style assembly names, ALREADY in the reference's processing order
                             (reverse-sorted, bin/ntsynt_synteny.py:34)
         contig_names[a][c], contig_lengths[a][c]
         sketch(a, w, masks) -> (h1 u64, pos u32, ctg u32) in (contig, position) order;
                             masks = per-contig (starts, ends) extra N intervals or None
         join(tables, order_asm) -> dict with H[V], POS/CTG/RANK/INV[G,V], link/degree[V] and the per-pair
                             arrays incmask/decmask/spread[V] of (i, i+1)            (device kernel iv)
         lookup(h1 array)    -> vertex id per key (0xFFFFFFFF if none)               (device join table)
    """

    def __init__(self, backend, k, w, w_rounds, bp, collinear_merge, z, m=90, simplify=True, prefix="out",
                 dev=False, write_files=True, quiet=False, interarrivals=False):
        self.be = backend
        self.interarrivals = interarrivals
        self.G = len(backend.names)
        if self.G < 2:
            raise ValueError("at least two assemblies are required")
        self.k, self.w, self.w_rounds = int(k), int(w), [int(x) for x in w_rounds]
        self.bp, self.z, self.m = int(bp), int(z), m
        cm = str(collinear_merge)
        if (mt := re.search(r"^(\d+)w$", cm)):
            self.collinear_merge = int(mt.group(1)) * self.w
        elif (mt := re.search(r"^(\d+)$", cm)):
            self.collinear_merge = int(mt.group(1))
        else:
            raise ValueError("--collinear-merge must be provided with an integer value or string in the form '<num>w'")
        self.simplify = simplify
        self.prefix = prefix
        self.dev = dev
        self.write_files = write_files
        self.quiet = quiet
        self.names = list(backend.names)
        # the assembly whose positions orient the paths: last of the (reverse-sorted) list
        # (ntjoin.py:93-94 takes .pop() of the max-weight assemblies; all weights are 1)
        self.orient = self.G - 1
        # output: rows within a block sorted by assembly name; blocks sorted on the smallest name
        self.name_order = sorted(range(self.G), key=lambda a: self.names[a])
        self.smallest = self.name_order[0]
        self.labels = [(mt.group(1) if (mt := re.search(FA_TSV_RE, nm)) else nm) for nm in self.names]
        self.outputs = {}
        self.stats = {}
        self.native = True       # irregular walks in C++ (csrc/nts_hostgraph.cu); False = the Python statements of the same rules

    def log(self, *a):
        if not self.quiet:
            _log(*a)

    def _tick(self, name):
        "accumulate wall time since the previous tick under stats['t_<name>'] (cheap phase timer)"
        import time
        now = time.perf_counter()
        last = getattr(self, "_tick_last", None)
        if last is not None and name:
            self.stats["t_" + name] = self.stats.get("t_" + name, 0.0) + (now - last)
        self._tick_last = now

    # ------------------------------------------------------------------ vertex storage
    def _init_vertices(self, j):
        self._prebuilt = "host" in j or "gather" in j
        self._pair_masks = j.get("pair_masks")
        self._h_extra = {}
        self._h_extra_cache = None
        self._pair_cache = None
        self._pair_orig, self._pair_delta = {}, {}      # corrections to the device prefix sums (see _refresh_pairs)
        self._br_touched = []                           # arrays of pairs (i, i+1) whose `conn` changed since the runs were last made
        self.sparse = set()                             # vertices that may hold a non-(i,i+1) edge
        self._ctg0 = {}                                 # (assembly, base vertex) -> round-0 contig, for the few overwritten entries
        self._dev = j if "gather" in j else None        # device-resident form: columns are read through gathers
        self._rk_cache, self._inv_cache = {}, {}
        if self._dev is not None:
            # the O(V) columns stay on the device (nts_graph_gather / _range_sums / _neigh / _runs_to_blocks); the host
            # holds the weight-filtered graph (nbr, conn), the three sparse lists and what later rounds write
            V = int(j["V"])
            self.V0 = self.V = V
            cap = V + 65536 + V // 16
            G = self.G
            self.H = LazyVector(lambda ids: j["gather"]("h1", ids), V)
            self.POS = LazyMatrix(lambda ids: j["gather"]("pos", ids), G, V, np.int64)
            self.CTG = LazyMatrix(lambda ids: j["gather"]("ctg", ids), G, V, np.int32)
            self.nbr, conn = j["links_nbr"](cap)
            self.conn = conn[:max(V - 1, 0)].view(bool)
            self.alive = np.zeros(cap, dtype=bool); self.alive[:V] = True
            self.RANK = self.INV = self.CI = self.CD = None
            self.incmask = self.decmask = self.spread = None
            breaks, deg3, big = j["sparse"](self.bp)
            self._breaks, self._deg3, self.big = breaks, deg3, big
            self._cum_dirty = False
            return
        if self._prebuilt:
            # lean form (device backend): the O(V) columns arrive in their final dtype and layout, written by the
            # device into pinned buffers with room for the vertices later rounds add
            V = int(j["V"])
            self.V0 = self.V = V
            cap = V + 65536 + V // 16
            self.H, self.POS, self.CTG, self.nbr, conn = j["host"](cap)
            self.conn = conn[:max(V - 1, 0)].view(bool)
            self.alive = np.zeros(cap, dtype=bool); self.alive[:V] = True
            self.RANK, self.INV = j["RANK"], j["INV"]
            self.incmask = self.decmask = self.spread = None             # never needed: see _refresh_pairs
            self.CI, self.CD = j["CI"], j["CD"]
            breaks, deg3, big = j["sparse"](self.bp)
            self._breaks, self._deg3, self.big = breaks, deg3, big
            self._cum_dirty = False
            return
        H = j["H"]
        V = len(H)
        self.V0 = V
        self.V = V
        cap = V + 65536 + V // 16
        self.H = np.empty(cap, dtype=np.uint64); self.H[:V] = H
        self.POS = np.zeros((self.G, cap), dtype=np.int64); self.POS[:, :V] = j["POS"]
        self.CTG = np.zeros((self.G, cap), dtype=np.int32); self.CTG[:, :V] = j["CTG"]
        self.RANK, self.INV = j["RANK"], j["INV"]                    # round-0 vertices only (uint32)
        self.alive = np.zeros(cap, dtype=bool); self.alive[:V] = True
        self.nbr = np.full((cap, 2), -1, dtype=np.int32)
        self.conn = np.zeros(max(V - 1, 0), dtype=bool)              # edge (i, i+1) present, base vertices
        self._breaks = self._deg3 = None
        # per-pair arrays of (i, i+1) and their prefix sums per assembly
        self.incmask = np.asarray(j["incmask"], dtype=np.uint32)
        self.decmask = np.asarray(j["decmask"], dtype=np.uint32)
        self.spread = np.asarray(j["spread"], dtype=np.uint32)
        self._cum_dirty = True
        if "CI" in j:                                   # device prefix sums (nts_graph_download_cums)
            self.CI, self.CD = j["CI"], j["CD"]
            self.big = np.flatnonzero(self.spread > self.bp)
            self._cum_dirty = False

    def _cums(self):
        "prefix sums (per assembly) of the direction bits of the pairs (i, i+1); CI[a, i] = sum over pairs < i"
        if self._cum_dirty:
            V = self.V0
            self.CI = np.zeros((self.G, V + 1), dtype=np.int32)
            self.CD = np.zeros((self.G, V + 1), dtype=np.int32)
            for a in range(self.G):
                np.cumsum((self.incmask >> np.uint32(a)) & np.uint32(1), out=self.CI[a, 1:])
                np.cumsum((self.decmask >> np.uint32(a)) & np.uint32(1), out=self.CD[a, 1:])
            self.big = np.flatnonzero(self.spread > self.bp)
            self._cum_dirty = False
        return self.CI, self.CD

    def _pair_values(self, idx):
        "direction masks and |dpos| spread of the pairs (i, i+1), i in idx, from the current positions (graph_links_kernel)"
        d = self.POS[:, idx + 1] - self.POS[:, idx]                    # [G, n]
        bits = (np.int64(1) << np.arange(self.G, dtype=np.int64))[:, None]
        inc = ((d > 0) * bits).sum(axis=0)
        dec = ((d < 0) * bits).sum(axis=0)
        ad = np.abs(d)
        return inc, dec, ad.max(axis=0) - ad.min(axis=0)

    def _pairs_around(self, vids):
        "indices of the pairs (v-1, v) and (v, v+1) of base vertices vids"
        v = np.asarray(vids, dtype=np.int64)
        p = np.unique(np.concatenate([v - 1, v]))
        return p[(p >= 0) & (p < self.V0 - 1)]

    def _refresh_pairs(self, pairs, old):
        """positions of base vertices were overwritten (update_list_mx_info): redo the values of the pairs around
        them.  `old` = _pair_values(pairs) taken BEFORE the overwrite.  The device prefix sums stay as they are; the
        few changed pairs are kept as corrections (pair index -> per-assembly delta against the ORIGINAL value)
        that _range_sums adds for the ranges containing them."""
        if self._cum_dirty:
            self._cums()
        if not len(pairs):
            return
        new = self._pair_values(pairs)
        G = self.G
        for i, oi, od, osp, ni, nd, nsp in zip(pairs.tolist(), old[0].tolist(), old[1].tolist(), old[2].tolist(),
                                               new[0].tolist(), new[1].tolist(), new[2].tolist()):
            orig_i, orig_d = self._pair_orig.setdefault(i, (oi, od))
            di = np.array([((ni >> a) & 1) - ((orig_i >> a) & 1) for a in range(G)], dtype=np.int64)
            dd = np.array([((nd >> a) & 1) - ((orig_d >> a) & 1) for a in range(G)], dtype=np.int64)
            if di.any() or dd.any():
                self._pair_delta[i] = (di, dd)
            else:
                self._pair_delta.pop(i, None)
            self._pair_cache = None
            if (nsp > self.bp) != (osp > self.bp):
                if nsp > self.bp:
                    self.big = np.insert(self.big, np.searchsorted(self.big, i), i)
                else:
                    self.big = np.delete(self.big, np.searchsorted(self.big, i))

    def _range_sums(self, lo, hi):
        "per-assembly counts of increasing / decreasing pairs (i, i+1) with lo <= i < hi, for arrays lo, hi"
        lo = np.minimum(lo, self.V0)          # vertices added after round 0 only form single-vertex segments
        hi = np.minimum(hi, self.V0)
        if self._dev is not None:
            up, down = self._dev["range_sums"](lo, hi)
        else:
            CI, CD = self._cums()
            up = (CI[:, hi] - CI[:, lo]).astype(np.int64)
            down = (CD[:, hi] - CD[:, lo]).astype(np.int64)
        if self._pair_delta:
            if self._pair_cache is None or self._pair_cache[0] != len(self._pair_delta):
                keys = np.array(sorted(self._pair_delta), dtype=np.int64)
                di = np.array([self._pair_delta[int(k)][0] for k in keys], dtype=np.int64).T      # [G, n]
                dd = np.array([self._pair_delta[int(k)][1] for k in keys], dtype=np.int64).T
                z = np.zeros((self.G, 1), dtype=np.int64)
                self._pair_cache = (len(self._pair_delta), keys, np.concatenate([z, np.cumsum(di, axis=1)], axis=1),
                                    np.concatenate([z, np.cumsum(dd, axis=1)], axis=1))
            _, keys, pdi, pdd = self._pair_cache
            a0, a1 = np.searchsorted(keys, lo), np.searchsorted(keys, hi)
            up += pdi[:, a1] - pdi[:, a0]
            down += pdd[:, a1] - pdd[:, a0]
        return up, down

    def _ctg_round0(self, a, v):
        got = self._ctg0.get((a, int(v))) if self._ctg0 else None
        return self.CTG[a, v] if got is None else got

    def _grow(self, need):
        cap = len(self.alive)
        if self.V + need <= cap:
            return
        new = max(cap * 2, self.V + need + 1024)
        if self._dev is None:
            self.H = np.concatenate([self.H, np.empty(new - cap, dtype=np.uint64)])
            self.POS = np.concatenate([self.POS, np.zeros((self.G, new - cap), dtype=np.int64)], axis=1)
            self.CTG = np.concatenate([self.CTG, np.zeros((self.G, new - cap), dtype=np.int32)], axis=1)
        self.alive = np.concatenate([self.alive, np.zeros(new - cap, dtype=bool)])
        self.nbr = np.concatenate([self.nbr, np.full((new - cap, 2), -1, dtype=np.int32)])

    def _lookup(self, keys):
        "vertex id per h1 (or -1) for a uint64 array: device join table, then the later additions"
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        out = np.full(len(keys), -1, dtype=np.int64)
        if len(keys) and self.V0:
            got = np.asarray(self.be.lookup(keys)).astype(np.int64)
            hit = got != 0xFFFFFFFF
            out[hit] = got[hit]
        if self._h_extra:
            if self._h_extra_cache is None or self._h_extra_cache[0] != len(self._h_extra):
                ek = np.fromiter(self._h_extra.keys(), dtype=np.uint64, count=len(self._h_extra))
                ev = np.fromiter(self._h_extra.values(), dtype=np.int64, count=len(self._h_extra))
                o = np.argsort(ek)
                self._h_extra_cache = (len(self._h_extra), ek[o], ev[o])
            _, ek, ev = self._h_extra_cache
            miss = np.flatnonzero(out < 0)
            if len(miss):
                j = np.searchsorted(ek, keys[miss])
                j[j >= len(ek)] = 0
                hit = ek[j] == keys[miss]
                out[miss[hit]] = ev[j[hit]]
        return out

    # ------------------------------------------------------------------ degree-2 graph on arrays
    def _add_edges(self, us, vs):
        "vectorised _add_edge for a batch of edges (a vertex may take two of them)"
        us = np.asarray(us, dtype=np.int64); vs = np.asarray(vs, dtype=np.int64)
        if not len(us):
            return
        X = np.concatenate([us, vs]); Y = np.concatenate([vs, us])
        order = np.argsort(X, kind="stable")
        X, Y = X[order], Y[order]
        first = np.ones(len(X), dtype=bool); first[1:] = X[1:] != X[:-1]
        start = np.maximum.accumulate(np.where(first, np.arange(len(X)), 0))
        r = np.arange(len(X)) - start                                  # 0, 1, .. within the vertex's group
        free0 = self.nbr[X, 0] < 0
        free1 = self.nbr[X, 1] < 0
        slot = np.where((r == 0) & free0, 0, 1)
        ok = np.where(slot == 0, free0, free1) & (r < 2) & ~((r == 1) & ~free0)
        if not ok.all():
            raise RuntimeError("internal error: vertex of degree > 2 in the weight-filtered graph")
        self.nbr[X, slot] = Y.astype(np.int32)
        lo, hi = np.minimum(us, vs), np.maximum(us, vs)
        base = ((hi - lo) == 1) & (hi < self.V0)
        self.conn[lo[base]] = True
        self._br_touched.append(lo[base])
        self.sparse.update(us[~base].tolist()); self.sparse.update(vs[~base].tolist())

    def _remove_edges(self, us, vs):
        us = np.asarray(us, dtype=np.int64); vs = np.asarray(vs, dtype=np.int64)
        if not len(us):
            return
        for x, y in ((us, vs), (vs, us)):
            for s_ in (0, 1):
                hit = self.nbr[x, s_] == y
                self.nbr[x[hit], s_] = -1
        lo = np.minimum(us, vs)
        base = (np.abs(us - vs) == 1) & (np.maximum(us, vs) < self.V0)
        self.conn[lo[base]] = False
        self._br_touched.append(lo[base])

    def _remove_vertices(self, ids):
        ids = np.unique(np.asarray(ids, dtype=np.int64))
        if not len(ids):
            return
        for s_ in (0, 1):
            nb = self.nbr[ids, s_]
            ok = nb >= 0
            self._remove_edges(ids[ok], nb[ok])
        self.alive[ids] = False

    def _remove_segments(self, segs):
        "delete every vertex of the given segments"
        ids = [seg_ids(sg) for sg in segs]
        if ids:
            self._remove_vertices(np.concatenate(ids))

    # ------------------------------------------------------------------ round-0 adjacency (implicit in ranks)
    def _rank_of(self, a, u):
        if self._dev is not None:
            got = self._rk_cache.get((a, u))
            if got is None:
                got = int(self._dev["gather"]("rank", np.array([u], dtype=np.int64))[a, 0])
            return got
        return int(self.RANK[a, u])

    def _at_rank(self, a, r):
        if self._dev is not None:
            got = self._inv_cache.get((a, r))
            if got is None:
                got = int(self._dev["gather"]("inv", np.array([r], dtype=np.int64))[a, 0])
            return got
        return int(self.INV[a, r])

    def _prefetch_edge_keys(self, edges):
        """device-resident form: _edge_key0 of these round-0 edges reads ranks, rank neighbours and contigs one value at
        a time; fetch them in four gathers"""
        if self._dev is None or not edges:
            return
        G, V0 = self.G, self.V0
        ends = np.unique(np.array([x for e in edges for x in e if x < V0], dtype=np.int64))
        if not len(ends):
            return
        rk = self._dev["gather"]("rank", ends)                            # [G, n]
        for a in range(G):
            self._rk_cache.update(zip(((a, v) for v in ends.tolist()), rk[a].tolist()))
        nxt = np.unique(rk.astype(np.int64) + 1)
        nxt = nxt[nxt < V0]
        iv = self._dev["gather"]("inv", nxt)                              # vertex at rank r + 1, per assembly
        for a in range(G):
            self._inv_cache.update(zip(((a, r) for r in nxt.tolist()), iv[a].tolist()))
        xs = np.unique(iv.astype(np.int64))
        rx = self._dev["gather"]("rank", xs)
        for a in range(G):
            self._rk_cache.update(zip(((a, v) for v in xs.tolist()), rx[a].tolist()))
        self.CTG.prefetch(np.concatenate([ends, xs]))

    def _adjacent(self, a, u, v):
        if u >= self.V0 or v >= self.V0:
            return False
        ru, rv = self._rank_of(a, u), self._rank_of(a, v)
        return abs(int(ru) - int(rv)) == 1 and self._ctg_round0(a, u) == self._ctg_round0(a, v)

    def _edge_key0(self, u, v):
        "position of edge {u,v} in build_graph's formatted_edges order for round 0 (ntjoin_utils.py:97-115)"
        def first_new(src, dst):
            for a in range(self.G):
                if self._adjacent(a, src, dst):
                    return a
            return None
        a0 = first_new(u, v)
        r = min(self._rank_of(a0, u), self._rank_of(a0, v))
        src = u if self._rank_of(a0, u) < self._rank_of(a0, v) else v
        tau = (a0, r)
        sigma = None
        for a in range(self.G):                      # first time src is the left element of a NEW pair
            rs = self._rank_of(a, src)
            if rs + 1 < self.V0:
                x = self._at_rank(a, rs + 1)
                if self._ctg_round0(a, x) == self._ctg_round0(a, src) and first_new(src, x) == a:
                    sigma = (a, rs)
                    break
        return (0, sigma, tau)

    # ------------------------------------------------------------------ simplification (ntsynt_synteny.py:566-590)
    def _simplify_round0(self, link, degree):
        """run_graph_simplification on the round-0 graph.  Only vertices of degree 3 can take part; their
        neighbourhoods are pulled out of the rank arrays in one vectorised step, the (few thousand) candidate
        edges are then visited in build_graph's edge order with plain Python containers."""
        G, V0 = self.G, self.V0
        cand = self._deg3 if self._deg3 is not None else np.flatnonzero(np.asarray(degree) == 3)
        bumped, removed = {}, []
        if not len(cand):
            return bumped, removed
        ctg0 = self.CTG                                 # round 0: nothing has been overwritten yet
        if self._dev is not None:
            # neighbourhoods of the candidates come from the device (nts_graph_neigh); same visiting order and rules in
            # C++ (csrc/nts_hostgraph.cu: nts_host_simplify_neigh)
            import ctypes as C
            from ._lib import check, lib, ptr
            cand64 = np.ascontiguousarray(cand, dtype=np.int64)
            left, right, rk = self._dev["neigh"](cand64)
            cap = 4 * len(cand64) + 4
            bs, bt, rm = (np.empty(cap, dtype=np.int64) for _ in range(3))
            n_out = C.c_int64()
            check(lib.nts_host_simplify_neigh(ptr(cand64, C.c_int64), len(cand64), ptr(left, C.c_int64), ptr(right, C.c_int64),
                                              ptr(rk, C.c_int64), G, ptr(bs, C.c_int64), ptr(bt, C.c_int64),
                                              ptr(rm, C.c_int64), cap, C.byref(n_out)))
            n = n_out.value
            return dict(zip(zip(bs[:n].tolist(), bt[:n].tolist()), [G] * n)), rm[:n].tolist()
        if self.native:
            # same visiting order and rules, in C++ (csrc/nts_hostgraph.cu: nts_host_simplify)
            import ctypes as C
            from ._lib import check, lib, ptr
            cand64 = np.ascontiguousarray(cand, dtype=np.int64)
            rank = np.ascontiguousarray(self.RANK, dtype=np.uint32)
            inv = np.ascontiguousarray(self.INV, dtype=np.uint32)
            if not (ctg0.dtype == np.int32 and ctg0.strides[1] == 4):
                ctg0 = np.ascontiguousarray(ctg0, dtype=np.int32)
            cap = 4 * len(cand64) + 4
            bs, bt, rm = (np.empty(cap, dtype=np.int64) for _ in range(3))
            n_out = C.c_int64()
            check(lib.nts_host_simplify(ptr(cand64, C.c_int64), len(cand64), ptr(rank, C.c_uint32), ptr(inv, C.c_uint32),
                                        ctg0.ctypes.data_as(C.POINTER(C.c_int32)), ctg0.strides[0] // 4, V0, G,
                                        ptr(bs, C.c_int64), ptr(bt, C.c_int64), ptr(rm, C.c_int64), cap, C.byref(n_out)))
            n = n_out.value
            return dict(zip(zip(bs[:n].tolist(), bt[:n].tolist()), [G] * n)), rm[:n].tolist()
        left, right, ranks = [], [], []
        for a in range(G):
            r = self.RANK[a, cand].astype(np.int64)
            lf = np.full(len(cand), -1, dtype=np.int64)
            rt = np.full(len(cand), -1, dtype=np.int64)
            ok = r > 0
            x = self.INV[a, r[ok] - 1].astype(np.int64)
            lf[ok] = np.where(ctg0[a, x] == ctg0[a, cand[ok]], x, -1)
            ok = r + 1 < V0
            x = self.INV[a, r[ok] + 1].astype(np.int64)
            rt[ok] = np.where(ctg0[a, x] == ctg0[a, cand[ok]], x, -1)
            left.append(lf.tolist()); right.append(rt.tolist()); ranks.append(r.tolist())
        cl = cand.tolist()
        pos_of = {u: i for i, u in enumerate(cl)}
        nb = {}                                   # u -> {x: number of supporting assemblies}
        for i, u in enumerate(cl):
            d = {}
            for a in range(G):
                for x in (left[a][i], right[a][i]):
                    if x >= 0:
                        d[x] = d.get(x, 0) + 1
            nb[u] = d

        def adjacent(a, src_i, dst):
            return left[a][src_i] == dst or right[a][src_i] == dst

        def edge_key(u, v):                       # ntjoin_utils.py:97-115, see _edge_key0
            iu = pos_of[u]
            a0 = next(a for a in range(G) if adjacent(a, iu, v))
            ru, rv = ranks[a0][iu], ranks[a0][pos_of[v]]
            src, isrc = (u, iu) if ru < rv else (v, pos_of[v])
            tau = (a0, min(ru, rv))
            sigma = None
            for a in range(G):
                x = right[a][isrc]
                if x >= 0 and not any(adjacent(b, isrc, x) for b in range(a)):
                    sigma = (a, ranks[a][isrc])
                    break
            return (0, sigma, tau)

        edges = set()
        for u in cl:
            for x in nb[u]:
                if x in pos_of:
                    edges.add((u, x) if u < x else (x, u))
        order = sorted(edges, key=lambda e: edge_key(*e))

        # number of full-weight edges at each candidate; a bump (visible to later iterations) adds one at both ends
        fullc = {u: sum(1 for c_ in d.values() if c_ == G) for u, d in nb.items()}
        for s_, t_ in order:
            if fullc[s_] == 1 and fullc[t_] == 1:   # node_partially_anchored on both ends
                common = [x for x in nb[s_] if x != t_ and x in nb[t_]]
                if len(common) == 1:            # the edge itself + exactly one 2-step path
                    removed.append(common[0])
                    if (s_, t_) not in bumped and nb[s_][t_] != G:
                        fullc[s_] += 1; fullc[t_] += 1
                    bumped[(s_, t_)] = G
        return bumped, removed

    # ------------------------------------------------------------------ paths of a max-degree-2 graph
    def _find_paths(self, device_pure=False, as_arrays=False):
        """ntjoin.py:114-151 on the weight-filtered graph: every component that is a simple path with
        two distinct ends gives one path, oriented from the end with the smaller position in the
        orienting assembly.  A path is a list of segments (lo, hi, dir).
        device_pure (device-resident form): the plain (i, i+1) runs whose vertices still hold their round-0
        positions are not turned into paths here; they are returned as a second value (starts, ends) for
        nts_graph_runs_to_blocks.
        as_arrays: the paths come back as (path_off, seg_lo, seg_hi, seg_dir) arrays instead of lists of tuples."""
        V0 = self.V0
        opos = self.POS[self.orient]
        paths = []
        dev = self._dev is not None
        if V0 and dev:
            # chain extraction on the device: push the pairs the host edited, get the runs of >= 2 vertices back
            if self._br_touched:
                t = np.unique(np.concatenate(self._br_touched))
                if len(t):
                    self._dev["set_links"](t, self.conn[t].astype(np.uint8))
            self._br_touched = []
            starts, ends = self._dev["runs"]()
            self._tick("p_runs")
        elif V0:
            if self._breaks is None:
                self._breaks = np.flatnonzero(~self.conn)
            elif self._br_touched:
                # pairs whose link changed since the list was made: their truth is conn[i]
                t = np.unique(np.concatenate(self._br_touched))
                keep = self._breaks[~np.isin(self._breaks, t)]
                self._breaks = np.union1d(keep, t[~self.conn[t]])
            self._br_touched = []
            starts = np.concatenate([[0], self._breaks + 1])
            ends = np.concatenate([starts[1:] - 1, [V0 - 1]])
        else:
            starts = ends = np.zeros(0, dtype=np.int64)

        def run_of(v):
            "index of the listed run holding base vertices v, or -1 (single-vertex runs are not listed on the device path)"
            r = np.searchsorted(starts, v, side="right") - 1
            ok = r >= 0
            ok[ok] = ends[r[ok]] >= v[ok]
            return np.where(ok, r, -1)
        # vertices that really hold a sparse edge (non-consecutive neighbour, or any edge of a later vertex)
        rs = np.zeros(0, dtype=np.int64)                   # the real sparse vertices, ascending
        if self.sparse:
            sv = np.fromiter(self.sparse, dtype=np.int64, count=len(self.sparse))
            nb = self.nbr[sv]
            has = ((nb >= 0) & ((np.abs(nb - sv[:, None]) != 1) | (np.maximum(nb, sv[:, None]) >= V0))).any(axis=1)
            rs = np.sort(sv[has])
            self.sparse = set(rs.tolist())
        real_sparse = len(rs) > 0
        is_sp_run = np.zeros(len(starts), dtype=bool)
        base_sp = rs[rs < V0]
        if len(base_sp) and len(starts):
            r = run_of(base_sp)
            is_sp_run[r[r >= 0]] = True
        plain = (ends > starts) & ~is_sp_run
        dev_runs = None
        if device_pure:
            # runs holding a vertex whose position / contig was overwritten after round 0 stay on the host
            touched = np.union1d(self.POS.touched_base(), self.CTG.touched_base())
            stale = np.zeros(len(starts), dtype=bool)
            if len(touched) and len(starts):
                r = run_of(touched)
                stale[r[r >= 0]] = True
            dv = np.flatnonzero(plain & ~stale)
            dev_runs = (starts[dv], ends[dv])
            plain &= stale
        pure = np.flatnonzero(plain)
        arr_lo = arr_hi = np.zeros(0, dtype=np.int64)
        arr_dir = np.zeros(0, dtype=np.int8)
        arr_off = np.zeros(1, dtype=np.int64)
        if len(pure):
            pa, pb = opos[starts[pure]], opos[ends[pure]]
            if as_arrays:
                ok = pa != pb
                arr_lo, arr_hi = starts[pure][ok].astype(np.int64), ends[pure][ok].astype(np.int64)
                arr_dir = np.where(pa[ok] < pb[ok], 1, -1).astype(np.int8)
                arr_off = np.arange(len(arr_lo) + 1, dtype=np.int64)
            else:
                for a, b, x, y in zip(starts[pure].tolist(), ends[pure].tolist(), pa.tolist(), pb.tolist()):
                    if x < y:
                        paths.append([(a, b, 1)])
                    elif y < x:
                        paths.append([(a, b, -1)])
        if dev:
            self._tick("p_pure")
        if real_sparse and (self.native or dev):
            # the walk over runs joined by sparse edges, in C++ (csrc/nts_hostgraph.cu: nts_host_walk_paths)
            import ctypes as C
            from ._lib import check, lib, ptr
            sv = rs
            starts64 = np.ascontiguousarray(starts, dtype=np.int64)
            ends64 = np.ascontiguousarray(ends, dtype=np.int64)
            nbr32 = self.nbr if self.nbr.flags.c_contiguous else np.ascontiguousarray(self.nbr)
            cap = 2 * len(sv) + 2
            slo, shi, poff = (np.empty(cap + 1, dtype=np.int64) for _ in range(3))
            sdir = np.empty(cap + 1, dtype=np.int8)
            n_p, n_s = C.c_int64(), C.c_int64()
            if dev:
                # positions only where a path can end: the sparse vertices and both ends of the runs that hold them
                ra, rb = sv.copy(), sv.copy()
                bm = sv < V0
                if bm.any() and len(starts):
                    r = run_of(sv[bm])
                    hit = r >= 0
                    ia = np.flatnonzero(bm)[hit]
                    ra[ia], rb[ia] = starts[r[hit]], ends[r[hit]]
                oid = np.unique(np.concatenate([sv, ra, rb]))
                oval = np.ascontiguousarray(opos[oid], dtype=np.int64)
                check(lib.nts_host_walk_paths_sparse(nbr32.ctypes.data_as(C.POINTER(C.c_int32)), V0, ptr(starts64, C.c_int64),
                                                     ptr(ends64, C.c_int64), len(starts64), ptr(sv, C.c_int64), len(sv),
                                                     ptr(oid, C.c_int64), ptr(oval, C.c_int64), len(oid), ptr(slo, C.c_int64),
                                                     ptr(shi, C.c_int64), sdir.ctypes.data_as(C.POINTER(C.c_int8)),
                                                     ptr(poff, C.c_int64), cap, C.byref(n_p), C.byref(n_s)))
            else:
                opos64 = np.ascontiguousarray(opos, dtype=np.int64)
                check(lib.nts_host_walk_paths(nbr32.ctypes.data_as(C.POINTER(C.c_int32)), V0, ptr(starts64, C.c_int64),
                                              ptr(ends64, C.c_int64), len(starts64), ptr(sv, C.c_int64), len(sv),
                                              ptr(opos64, C.c_int64), ptr(slo, C.c_int64), ptr(shi, C.c_int64),
                                              sdir.ctypes.data_as(C.POINTER(C.c_int8)), ptr(poff, C.c_int64), cap,
                                              C.byref(n_p), C.byref(n_s)))
            if as_arrays:
                ns, np_ = n_s.value, n_p.value
                arr_off = np.concatenate([arr_off, arr_off[-1] + poff[1:np_ + 1]])
                arr_lo = np.concatenate([arr_lo, slo[:ns]]); arr_hi = np.concatenate([arr_hi, shi[:ns]])
                arr_dir = np.concatenate([arr_dir, sdir[:ns]])
            else:
                segs_all = list(zip(slo[:n_s.value].tolist(), shi[:n_s.value].tolist(), sdir[:n_s.value].tolist()))
                off = poff[:n_p.value + 1].tolist()
                paths.extend(segs_all[off[i]:off[i + 1]] for i in range(n_p.value))
        elif real_sparse:
            seen_runs = set()
            sv = rs
            ra, rb = sv.copy(), sv.copy()
            bm = sv < V0
            if bm.any():
                ri = np.searchsorted(starts, sv[bm], side="right") - 1
                ra[bm], rb[bm] = starts[ri], ends[ri]
            ends_all = np.unique(np.concatenate([ra, rb]))
            # run of every vertex the walk can stand on: the sparse vertices and the ends of their runs
            bounds = dict(zip(sv.tolist(), zip(ra.tolist(), rb.tolist())))
            bounds.update(zip(ra.tolist(), zip(ra.tolist(), rb.tolist())))
            bounds.update(zip(rb.tolist(), zip(ra.tolist(), rb.tolist())))

            def run_bounds(v):
                got = bounds.get(v)
                if got is not None:
                    return got
                if v >= V0:
                    return v, v
                r = int(np.searchsorted(starts, v, side="right") - 1)
                return int(starts[r]), int(ends[r])
            nb_all = self.nbr[ends_all]
            cand_ends = ends_all[(nb_all >= 0).sum(axis=1) == 1].tolist()
            keys_arr = np.unique(np.concatenate([sv, ends_all]))
            nbr_of = dict(zip(keys_arr.tolist(), self.nbr[keys_arr].tolist()))
            for e0 in cand_ends:
                if run_bounds(e0)[0] in seen_runs:
                    continue
                segs, prev, cur, ok = [], -1, e0, True
                while True:
                    a, b = run_bounds(cur)
                    if a in seen_runs:
                        ok = False
                        break
                    seen_runs.add(a)
                    if cur == a:
                        segs.append((a, b, 1)); last = b
                    else:
                        segs.append((a, b, -1)); last = a
                    inside = -1
                    if b > a:
                        inside = last - 1 if last == b else last + 1
                    nxt = -1
                    for y in (nbr_of.get(last) or self.nbr[last].tolist()):
                        if y < 0 or y == inside:
                            continue
                        if a == b and y == prev:
                            continue
                        nxt = y
                    if nxt < 0:
                        break
                    prev, cur = last, nxt
                if not ok:
                    continue
                first, lastv = seg_first(segs[0]), seg_last(segs[-1])
                if first == lastv:
                    continue
                pa, pb = opos[first], opos[lastv]
                if pa < pb:
                    paths.append(segs)
                elif pb < pa:
                    paths.append([(lo, hi, -d) for lo, hi, d in reversed(segs)])
        if as_arrays:
            paths = (arr_off, arr_lo, arr_hi, arr_dir)
        return (paths, dev_runs) if device_pure else paths

    # ------------------------------------------------------------------ blocks from paths
    def _blocks_from_paths(self, paths):
        """find_synteny_blocks (ntsynt_synteny.py:66-106) for every path, vectorised over the flattened
        segments and junctions of all paths.  Orientation tallies come from prefix sums of the per-pair
        direction masks (segments) plus the few junction steps.  Vertices of unoriented blocks are deleted."""
        if not paths:
            return []
        G = self.G
        nseg = np.array([len(p) for p in paths], dtype=np.int64)
        flat = [sg for p in paths for sg in p]
        lo = np.array([sg[0] for sg in flat], dtype=np.int64)
        hi = np.array([sg[1] for sg in flat], dtype=np.int64)
        up_dir = np.array([sg[2] > 0 for sg in flat])
        pid = np.repeat(np.arange(len(paths)), nseg)
        seg0 = np.concatenate([[0], np.cumsum(nseg)])                 # first segment index of each path
        first = np.where(up_dir, lo, hi)
        last = np.where(up_dir, hi, lo)
        # junctions: between segment j-1 and j of the same path
        jn = np.flatnonzero(pid[1:] == pid[:-1]) + 1                  # index of the segment AFTER the junction
        start_seg = seg0[:-1].copy()
        jinc = jdec = None
        if len(jn):
            u, v = last[jn - 1], first[jn]
            chg = (self.CTG[:, u] != self.CTG[:, v]).any(axis=0)
            # only the LAST run of constant contigs of a path becomes a block (past_start_flag is never set,
            # ntsynt_synteny.py:71,77): the block starts at the last junction with a contig change
            if chg.any():
                np.maximum.at(start_seg, pid[jn[chg]], jn[chg])
            dp = self.POS[:, v] - self.POS[:, u]
            jkeep = jn > start_seg[pid[jn]]                           # junctions inside the block
            jinc = np.zeros((G, len(paths)), dtype=np.int64)
            jdec = np.zeros((G, len(paths)), dtype=np.int64)
            for a in range(G):
                jinc[a] = np.bincount(pid[jn[jkeep]], weights=(dp[a, jkeep] > 0), minlength=len(paths))
                jdec[a] = np.bincount(pid[jn[jkeep]], weights=(dp[a, jkeep] < 0), minlength=len(paths))
        keep_seg = np.arange(len(flat)) >= start_seg[pid]
        up, down = self._range_sums(lo, hi)
        sinc = np.where(up_dir[None, :], up, down) * keep_seg[None, :]
        sdec = np.where(up_dir[None, :], down, up) * keep_seg[None, :]
        inc = np.zeros((G, len(paths)), dtype=np.int64)
        dec = np.zeros((G, len(paths)), dtype=np.int64)
        for a in range(G):
            inc[a] = np.bincount(pid, weights=sinc[a], minlength=len(paths))
            dec[a] = np.bincount(pid, weights=sdec[a], minlength=len(paths))
        if jinc is not None:
            inc += jinc; dec += jdec
        n_all = np.bincount(pid, weights=(hi - lo + 1) * keep_seg, minlength=len(paths)).astype(np.int64)
        f_id = first[start_seg]
        l_id = last[seg0[1:] - 1]
        plus = (inc == n_all - 1) | (n_all == 1)
        minus = ~plus & (dec == n_all - 1)
        simple = (plus | minus).all(axis=0).tolist()
        ori_simple = np.where(plus, "+", "-").T.tolist()
        ctg_f = self.CTG[:, f_id].T.tolist()
        pos_f = self.POS[:, f_id].T.tolist()
        pos_l = self.POS[:, l_id].T.tolist()
        n_list, f_list, l_list = n_all.tolist(), f_id.tolist(), l_id.tolist()
        s_list, e_list = (start_seg - seg0[:-1]).tolist(), nseg.tolist()
        blocks, to_remove = [], []
        for x, segs in enumerate(paths):
            if s_list[x]:
                segs = segs[s_list[x]:]
            n = n_list[x]
            if simple[x]:
                ori = ori_simple[x]
            else:
                ori = []
                for a in range(G):
                    if plus[a, x]:
                        ori.append("+")
                    elif minus[a, x]:
                        ori.append("-")
                    else:
                        positive = int(inc[a, x]) / float(n - 1) * 100
                        negative = 100 - positive
                        ori.append("+" if positive >= self.m else ("-" if negative >= self.m else "?"))
                if "?" in ori:
                    to_remove.extend(segs)
                    continue
            blocks.append(Block(segs, ctg_f[x], ori, f_list[x], l_list[x], pos_f[x], pos_l[x], n))
        if to_remove:
            self._remove_segments(to_remove)
        return blocks

    def _split_indels(self, blocks):
        "check_for_indels + break_synteny_block (ntsynt_synteny.py:364-409)"
        if not blocks:
            return blocks
        self._cums()
        big = self.big
        # vectorised detection: blocks with no large-spread pair inside a segment and no large-spread junction
        nseg = np.array([len(b.segs) for b in blocks], dtype=np.int64)
        flat = [sg for b in blocks for sg in b.segs]
        lo = np.array([sg[0] for sg in flat], dtype=np.int64)
        hi = np.array([sg[1] for sg in flat], dtype=np.int64)
        bid = np.repeat(np.arange(len(blocks)), nseg)
        dirty = np.zeros(len(blocks), dtype=bool)
        if len(big):
            inside = np.searchsorted(big, lo) != np.searchsorted(big, hi)
            dirty[bid[inside]] = True
        jn = np.flatnonzero(bid[1:] == bid[:-1]) + 1
        if len(jn):
            up_dir = np.array([sg[2] > 0 for sg in flat])
            first = np.where(up_dir, lo, hi)
            last = np.where(up_dir, hi, lo)
            dd = np.abs(self.POS[:, first[jn]] - self.POS[:, last[jn - 1]])
            wide = (dd.max(axis=0) - dd.min(axis=0)) > self.bp
            dirty[bid[jn[wide]]] = True
        if not dirty.any():
            return blocks
        if self._dev is not None:
            # the loop below reads positions one vertex at a time: fetch the segment ends of the dirty blocks at once
            ends_ = [x for bx in np.flatnonzero(dirty).tolist() for sg in blocks[bx].segs for x in (sg[0], sg[1])]
            cuts_ = [big[np.searchsorted(big, sg[0]):np.searchsorted(big, sg[1])] for bx in np.flatnonzero(dirty).tolist()
                     for sg in blocks[bx].segs if sg[1] > sg[0]]
            cuts_ = np.concatenate(cuts_) if cuts_ else np.zeros(0, dtype=np.int64)
            self.POS.prefetch(np.concatenate([np.asarray(ends_, dtype=np.int64), cuts_, cuts_ + 1]))
        out = []
        rm_u, rm_v = [], []
        for bx, b in enumerate(blocks):
            if not dirty[bx]:
                out.append(b)
                continue
            pieces, cur = [], []
            for jx, (slo, shi, d) in enumerate(b.segs):
                if jx > 0:
                    u, v = seg_last(b.segs[jx - 1]), seg_first(b.segs[jx])
                    dd = np.abs(self.POS[:, v] - self.POS[:, u])
                    if int(dd.max() - dd.min()) > self.bp:
                        rm_u.append(u); rm_v.append(v)
                        pieces.append(cur); cur = []
                cuts = big[np.searchsorted(big, slo):np.searchsorted(big, shi)] if shi > slo else ()
                if len(cuts):
                    rm_u.extend(int(c) for c in cuts); rm_v.extend(int(c) + 1 for c in cuts)
                    if d > 0:
                        s0 = slo
                        for c in cuts:
                            cur.append((s0, int(c), 1)); pieces.append(cur); cur = []
                            s0 = int(c) + 1
                        cur.append((s0, shi, 1))
                    else:
                        s0 = shi
                        for c in cuts[::-1]:
                            cur.append((int(c) + 1, s0, -1)); pieces.append(cur); cur = []
                            s0 = int(c)
                        cur.append((slo, s0, -1))
                else:
                    cur.append((slo, shi, d))
            pieces.append(cur)
            for segs in pieces:
                f, l = seg_first(segs[0]), seg_last(segs[-1])
                n = sum(h_ - l_ + 1 for l_, h_, _ in segs)
                out.append(Block(segs, b.ctg, list(b.ori), f, l, self.POS[:, f].tolist(), self.POS[:, l].tolist(), n))
        if rm_u:
            self._remove_edges(np.array(rm_u, dtype=np.int64), np.array(rm_v, dtype=np.int64))
        return out

    def _filter_small(self, blocks, min_mx):
        "filter_synteny_blocks (ntsynt_synteny.py:411-426)"
        keep, rm = [], []
        for b in blocks:
            if b.n >= min_mx:
                keep.append(b)
            else:
                rm.extend(b.segs)
        if rm:
            self._remove_segments(rm)
        return keep

    def _host_blocks_native(self, path_off, lo, hi, sdir):
        """_blocks_from_paths + _split_indels + _filter_small(4) for paths given as segment arrays, in C++
        (csrc/nts_hostgraph.cu: nts_host_paths_to_blocks); the graph edits it asks for are applied here"""
        import ctypes as C
        from ._lib import check, lib, ptr
        G, n_paths, n_seg = self.G, len(path_off) - 1, len(lo)
        if not n_paths:
            return []
        if self._dev is None:
            self._cums()
        up, down = self._range_sums(lo, hi)
        up, down = np.ascontiguousarray(up, dtype=np.int64), np.ascontiguousarray(down, dtype=np.int64)
        big = np.ascontiguousarray(self.big, dtype=np.int64)
        i0 = np.searchsorted(big, lo)
        cnt = np.maximum(np.searchsorted(big, hi) - i0, 0)
        tot = int(cnt.sum())
        cut_off = np.zeros(n_seg + 1, dtype=np.int64)
        np.cumsum(cnt, out=cut_off[1:])
        cuts = np.ascontiguousarray(big[np.repeat(i0, cnt) + (np.arange(tot) - np.repeat(cut_off[:-1], cnt))]) if tot else big[:0]
        # columns of the position table: segment starts, segment ends, cut pairs c, c + 1
        cols = np.concatenate([lo, hi, cuts, cuts + 1])
        pos = np.ascontiguousarray(self.POS[:, cols], dtype=np.int64)
        ctg = np.ascontiguousarray(self.CTG[:, cols], dtype=np.int32)
        cap = n_seg + tot + 2
        i64 = lambda n: np.empty(n, dtype=np.int64)          # noqa: E731
        b_off, b_n, b_first, b_last = i64(cap + 1), i64(cap), i64(cap), i64(cap)
        b_ori, b_ctg = np.empty((cap, G), dtype=np.int8), np.empty((cap, G), dtype=np.int32)
        b_fpos, b_lpos = np.empty((cap, G), dtype=np.int64), np.empty((cap, G), dtype=np.int64)
        o_lo, o_hi, o_dir = i64(cap), i64(cap), np.empty(cap, dtype=np.int8)
        r_lo, r_hi, e_u, e_v = i64(cap), i64(cap), i64(cap), i64(cap)
        counts = (C.c_int64 * 4)()
        p8 = lambda x: x.ctypes.data_as(C.POINTER(C.c_int8))      # noqa: E731
        path_off = np.ascontiguousarray(path_off, dtype=np.int64)
        lo, hi = np.ascontiguousarray(lo, dtype=np.int64), np.ascontiguousarray(hi, dtype=np.int64)
        sdir = np.ascontiguousarray(sdir, dtype=np.int8)
        self._tick("bh_prep")
        check(lib.nts_host_paths_to_blocks(n_paths, ptr(path_off, C.c_int64), ptr(lo, C.c_int64), ptr(hi, C.c_int64), p8(sdir), G,
                                           ptr(up, C.c_int64), ptr(down, C.c_int64), ptr(cuts if tot else np.zeros(1, np.int64), C.c_int64),
                                           ptr(cut_off, C.c_int64), ptr(pos, C.c_int64), ctg.ctypes.data_as(C.POINTER(C.c_int32)),
                                           int(self.bp), float(self.m), 4, cap, ptr(b_off, C.c_int64), ptr(b_n, C.c_int64),
                                           ptr(b_first, C.c_int64), ptr(b_last, C.c_int64), p8(b_ori),
                                           b_ctg.ctypes.data_as(C.POINTER(C.c_int32)), ptr(b_fpos, C.c_int64), ptr(b_lpos, C.c_int64),
                                           ptr(o_lo, C.c_int64), ptr(o_hi, C.c_int64), p8(o_dir), ptr(r_lo, C.c_int64),
                                           ptr(r_hi, C.c_int64), ptr(e_u, C.c_int64), ptr(e_v, C.c_int64), counts))
        nb, no, nr, ne = (int(x) for x in counts)
        self._tick("bh_c")
        segs = list(zip(o_lo[:no].tolist(), o_hi[:no].tolist(), o_dir[:no].tolist()))
        off = b_off[:nb + 1].tolist()
        ori = b_ori[:nb].view("S1").astype("U1").tolist()             # '+' / '-' per assembly
        ctg_l, fp, lp = b_ctg[:nb].tolist(), b_fpos[:nb].tolist(), b_lpos[:nb].tolist()
        blocks = [Block(segs[off[i]:off[i + 1]], ctg_l[i], ori[i], f, l, fp[i], lp[i], n)
                  for i, (f, l, n) in enumerate(zip(b_first[:nb].tolist(), b_last[:nb].tolist(), b_n[:nb].tolist()))]
        self._tick("bh_wrap")
        self.stats.setdefault("host_blocks", []).append([n_paths, n_seg, nb, no, nr, ne])
        if ne:
            self._remove_edges(e_u[:ne].copy(), e_v[:ne].copy())
        if nr:
            rl, rh = r_lo[:nr], r_hi[:nr]
            n = rh - rl + 1
            self._remove_vertices(np.repeat(rl, n) + (np.arange(int(n.sum())) - np.repeat(np.cumsum(n) - n, n)))
        return blocks

    # ------------------------------------------------------------------ paths -> blocks (one round)
    def _extract_blocks(self):
        """find_paths -> find_synteny_blocks -> check_for_indels -> filter_synteny_blocks(4) for the current graph.
        Device-resident form: the plain (i, i+1) runs go through nts_graph_runs_to_blocks (kernel iv-c) and come back
        as a compact block table; only the components with a non-(i, i+1) edge, and the runs holding a vertex whose
        position was overwritten after round 0, are walked on the host."""
        if self._dev is None:
            if self.native:
                poff, lo, hi, sdir = self._find_paths(as_arrays=True)
                self._tick("paths")
                self.stats["paths"] = len(poff) - 1
                blocks = self._host_blocks_native(poff, lo, hi, sdir)
                self._tick("blocks")
                return blocks
            paths = self._find_paths()
            self._tick("paths")
            self.stats["paths"] = len(paths)
            blocks = self._blocks_from_paths(paths)
            self._tick("blocks")
            blocks = self._split_indels(blocks)
            self._tick("indels")
            return self._filter_small(blocks, 4)
        paths, (ds, de) = self._find_paths(device_pure=True, as_arrays=self.native)
        self._tick("p_walk")
        res = self._dev["runs_to_blocks"](ds, de, self.bp, float(self.m), 4)
        self._tick("b_dev")
        if self.native:
            blocks = self._host_blocks_native(*paths)
            n_host = len(paths[0]) - 1
        else:
            blocks = self._blocks_from_paths(paths)
            blocks = self._split_indels(blocks)
            blocks = self._filter_small(blocks, 4)
            n_host = len(paths)
        self._tick("b_host")
        lo, hi = res["b_lo"].astype(np.int64), res["b_hi"].astype(np.int64)
        self.stats["paths"] = n_host + len(lo) + len(res["r_lo"])
        if len(lo):
            o = np.argsort(lo)
            lo, hi, up = lo[o], hi[o], res["b_dir"][o] > 0
            plus = res["b_plus"][o]
            f_id, l_id = np.where(up, lo, hi), np.where(up, hi, lo)
            both = np.concatenate([f_id, l_id])
            pos = self.POS[:, both]
            ctg_f = self.CTG[:, f_id].T.tolist()
            pos_f, pos_l = pos[:, :len(lo)].T.tolist(), pos[:, len(lo):].T.tolist()
            ori = np.where((plus.astype(np.int64)[:, None] >> np.arange(self.G)) & 1, "+", "-").tolist()
            for i, (l_, h_, u_, f_, e_) in enumerate(zip(lo.tolist(), hi.tolist(), up.tolist(), f_id.tolist(), l_id.tolist())):
                blocks.append(Block([(l_, h_, 1 if u_ else -1)], ctg_f[i], ori[i], f_, e_, pos_f[i], pos_l[i], h_ - l_ + 1))
        self._tick("b_make")
        # what the kernel deleted: unoriented / short pieces, and the edges cut as indels
        if len(res["r_lo"]):
            rl, rh = res["r_lo"].astype(np.int64), res["r_hi"].astype(np.int64)
            n = rh - rl + 1
            self._remove_vertices(np.repeat(rl, n) + (np.arange(int(n.sum())) - np.repeat(np.cumsum(n) - n, n)))
        if len(res["cuts"]):
            c = res["cuts"].astype(np.int64)
            self._remove_edges(c, c + 1)
        self.stats["dev_runs"] = self.stats.get("dev_runs", []) + [[int(len(ds)), int(len(lo)), int(len(res["r_lo"])), int(len(res["cuts"])), n_host]]
        self._tick("b_apply")
        return blocks

    # ------------------------------------------------------------------ output
    def _sort_blocks(self, blocks):
        "SyntenyBlock.__lt__ (synteny_block.py:102-109): contig NAME string, then start, of the smallest assembly"
        a = self.smallest
        cn = self.be.contig_names[a]
        return sorted(blocks, key=lambda b: (cn[int(b.ctg[a])], b.start(a)))

    def _long_enough(self, b):
        return all(b.end(a, self.k) - b.start(a) >= self.z for a in range(self.G))

    def _block_coords(self, blocks):
        "start[nb, G], end[nb, G] of the blocks (bin/assembly_block.py:17-23), ctg[nb, G]"
        if not blocks:
            z = np.zeros((0, self.G), dtype=np.int64)
            return z, z, z
        fp = np.array([b.first_pos for b in blocks], dtype=np.int64)
        lp = np.array([b.last_pos for b in blocks], dtype=np.int64)
        ctg = np.array([b.ctg for b in blocks], dtype=np.int64)
        return np.minimum(fp, lp), np.maximum(fp, lp) + self.k, ctg

    def _emit(self, key, blocks, verbose=False):
        st, en, ctg = self._block_coords(blocks)
        keep = np.flatnonzero(((en - st) >= self.z).all(axis=1)) if len(blocks) else []
        text = []
        if len(keep):
            order = self.name_order
            st_l, en_l, ctg_l = st[keep][:, order].tolist(), en[keep][:, order].tolist(), ctg[keep][:, order].tolist()
            labels = [self.labels[a] for a in order]
            cnames = [self.be.contig_names[a] for a in order]
            for num, bi in enumerate(keep.tolist()):
                b = blocks[bi]
                tail = f"\t{b.n}\t{b.broken_reason}\n" if verbose else f"\t{b.n}\n"
                s_, e_, c_ = st_l[num], en_l[num], ctg_l[num]
                for j, a in enumerate(order):
                    text.append(f"{num}\t{labels[j]}\t{cnames[j][c_[j]]}\t{s_[j]}\t{e_[j]}\t{b.ori[a]}{tail}")
        text = "".join(text)
        self.outputs[key] = text
        if self.write_files:
            suffix = {"initial": ".synteny_blocks.tsv", "pre_merge": ".pre-collinear-merge.synteny_blocks.tsv",
                      "final": ".synteny_blocks.tsv"}[key]
            with open(f"{self.prefix}{suffix}", "w", encoding="utf-8") as fh:
                fh.write(text)
        return text

    def _print_interarrivals(self, blocks):
        """print_interarrivals (bin/ntsynt_synteny.py:557-564): <prefix>.interarrivals.tsv, one distance per line between
        consecutive minimizers of every block of the initial graph, block by block, assembly by assembly.  (The
        reference's block order follows igraph vertex ids, which come from iterating a Python set of strings -- it is
        not reproducible between runs; the file is a bag of distances.)"""
        parts = []
        if blocks:
            ids = [np.concatenate([seg_ids(sg) for sg in b.segs]) if len(b.segs) > 1 else seg_ids(b.segs[0]) for b in blocks]
            n = np.array([len(x) for x in ids], dtype=np.int64)
            P = np.asarray(self.POS[:, np.concatenate(ids)], dtype=np.int64)
            D = np.abs(np.diff(P, axis=1))
            off = np.cumsum(n) - n
            parts = [D[:, o:o + c - 1].ravel() for o, c in zip(off.tolist(), n.tolist()) if c > 1]
        flat = np.concatenate(parts) if parts else np.zeros(0, dtype=np.int64)
        text = "".join(f"{v}\n" for v in flat.tolist())
        self.outputs["interarrivals"] = text
        if self.write_files:
            with open(f"{self.prefix}.interarrivals.tsv", "w", encoding="utf-8") as fh:
                fh.write(text)
        return text

    def _check_non_overlapping(self, blocks):
        """check_non_overlapping (bin/ntsynt_synteny.py:234-253), the --dev check of the final blocks: walking the
        blocks in output order, a block whose extent on some assembly shares >= z bases with an earlier block on the
        same contig gets a warning on stderr (one per block and assembly).  Blocks shorter than z on any assembly are
        neither checked nor remembered.  Returns the warnings as (assembly name, contig name, start, end)."""
        warnings = []
        if not blocks:
            return warnings
        st, en, ctg = self._block_coords(blocks)
        ok = ((en - st) >= self.z).all(axis=1)
        hits = []
        for a in range(self.G):
            seen = {}                                   # contig -> ([starts], [ends]) of the earlier blocks
            for i in np.flatnonzero(ok).tolist():
                c, s_, e_ = int(ctg[i, a]), int(st[i, a]), int(en[i, a])
                ss, ee = seen.setdefault(c, ([], []))
                if ss:
                    ov = np.minimum(np.asarray(ee), e_) - np.maximum(np.asarray(ss), s_)
                    if (ov >= max(self.z, 1)).any():
                        hits.append((i, a, self.names[a], self.be.contig_names[a][c], s_, e_))
                ss.append(s_); ee.append(e_)
        for _, _, nm, cn, s_, e_ in sorted(hits):       # the reference walks block-major
            print("WARNING: detected overlapping segments for this block:", nm, cn, s_, e_, "\n", file=sys.stderr, flush=True)
            warnings.append((nm, cn, s_, e_))
        self.outputs["overlap_warnings"] = warnings
        return warnings

    # ------------------------------------------------------------------ merge (ntsynt_synteny.py:428-472)
    def _merge_collinear(self, blocks):
        out = []
        G, k = self.G, self.k
        st, en, _ = self._block_coords(blocks)
        st_l, en_l = st.tolist(), en.tolist()
        cur = blocks[0]
        cur_st, cur_en = st_l[0], en_l[0]
        cur_ori, cur_ctg = list(cur.ori), [int(x) for x in cur.ctg]
        for i in range(1, len(blocks)):
            b = blocks[i]
            b_ori, b_ctg = list(b.ori), [int(x) for x in b.ctg]
            same_ori = cur_ori == b_ori
            same_ctg = cur_ctg == b_ctg
            bs, be_ = st_l[i], en_l[i]
            # get_difference_between_blocks: distance from the end of the first block to the start of the second,
            # measured the other way round where both are on the minus strand
            diffs = [(cur_st[a] - be_[a]) if (cur_ori[a] == "-" and b_ori[a] == "-") else (bs[a] - cur_en[a]) for a in range(G)]
            mx, mn = max(diffs), min(diffs)
            spread = mx - mn
            if (not same_ori) or (not same_ctg) or spread > self.bp - k or mx >= self.collinear_merge:
                if not same_ctg:
                    b.broken_reason = "id_change"
                elif not same_ori:
                    b.broken_reason = "ori_change"
                elif mn < 0:
                    b.broken_reason = "inconsistent_order"
                elif spread > self.bp - k:
                    b.broken_reason = "indel"
                elif mx >= self.collinear_merge:
                    b.broken_reason = "merge"
                out.append(cur)
                cur, cur_st, cur_en, cur_ori, cur_ctg = b, bs, be_, b_ori, b_ctg
            else:
                # extend: coordinates come from the first minimizer of `cur` and the last of `b`
                cur.last_pos = b.last_pos
                cur.last_id = b.last_id
                cur.n += b.n
                cur.segs = None
                cur_st = [min(int(f), int(l)) for f, l in zip(cur.first_pos, cur.last_pos)]
                cur_en = [max(int(f), int(l)) + k for f, l in zip(cur.first_pos, cur.last_pos)]
        out.append(cur)
        return out

    # ------------------------------------------------------------------ refinement helpers
    def _masks_for(self, blocks, prev_w):
        "get_synteny_bed_lists + mask_assemblies_with_synteny_extents (ntsynt_synteny.py:117-157)"
        thr = max(2 * prev_w, prev_w + self.k + 1)
        shrink = prev_w + self.k
        st, en, ctg = self._block_coords(blocks)
        masks = []
        for a in range(self.G):
            lens = np.asarray(self.be.contig_lengths[a], dtype=np.int64)
            s, e, c = st[:, a], en[:, a], ctg[:, a]
            big = (e - s) > thr
            s2 = np.maximum(s[big] + shrink, 0)
            c2 = c[big]
            e2 = np.minimum(e[big] - shrink, lens[c2]) if len(c2) else s2
            ok = s2 < e2             # an interval the negative slop empties is dropped (unpinned)
            s2, e2, c2 = s2[ok], e2[ok], c2[ok]
            o = np.lexsort((e2, s2, c2))
            s2, e2, c2 = s2[o], e2[o], c2[o]
            # union per contig (bedtools maskfasta semantics): an interval starts a new run unless it begins at or
            # before the furthest end seen so far in its contig.  All contigs at once on one axis (contig << 40 | pos):
            # a contig's intervals lie above everything of the contigs before it, so one running maximum serves all
            lst = [(_NO_U64, _NO_U64)] * len(self.be.contig_names[a])
            if len(s2):
                base = c2 << np.int64(40)
                run_max = np.maximum.accumulate(e2 + base)
                heads = np.flatnonzero(np.r_[True, (s2 + base)[1:] > run_max[:-1]])
                ends_ = (run_max[np.r_[heads[1:] - 1, len(s2) - 1]] - base[heads]).astype(np.uint64)
                starts_ = s2[heads].astype(np.uint64)
                ch = c2[heads]
                cb = np.flatnonzero(np.r_[True, ch[1:] != ch[:-1]])           # first run of each contig
                for c_, i0, i1 in zip(ch[cb].tolist(), cb.tolist(), np.r_[cb[1:], len(ch)].tolist()):
                    lst[c_] = (starts_[i0:i1], ends_[i0:i1])
            masks.append(lst)
        return masks

    @staticmethod
    def _dedup(h1, pos, ctg):
        "read_minimizers (ntjoin_utils.py:167-193): drop every h1 seen more than once in the file"
        if not len(h1):
            return h1, pos, ctg
        _, inv, cnt = np.unique(h1, return_inverse=True, return_counts=True)
        keep = cnt[inv] == 1
        return h1[keep], pos[keep], ctg[keep]

    def _refine_round(self, blocks, new_w, prev_w, last_round, round_no=1):
        G = self.G
        # --- new minimizers from the masked assemblies (generate_additional_minimizers :532-541)
        masks = self._masks_for(blocks, prev_w)
        # u_r of SURVEY 8d: the fraction of every assembly the masked round still has to sketch
        self.stats.setdefault("unmasked", []).append(
            [round(1.0 - sum(int((e - s).sum()) for s, e in masks[a]) / max(sum(self.be.contig_lengths[a]), 1), 6) for a in range(G)])
        self._tick("r_masks")
        use_dev = self._dev is not None and "refine_filter" in self._dev and hasattr(self.be, "sketch_table")
        term_ids = np.array([x for b in blocks for x in (b.first_id, b.last_id)], dtype=np.int64)
        if use_dev:
            # the new tables never leave the device: duplicate removal, the block filter and the G-way intersection are
            # kernels (nts_graph_refine_filter); the host gets the few survivors
            tables = [self.be.sketch_table(a, new_w, masks[a]) for a in range(G)]
            self._tick("r_sketch")
            seg_lo = np.array([sg[0] for b in blocks for sg in b.segs], dtype=np.int64)
            seg_hi = np.array([sg[1] for b in blocks for sg in b.segs], dtype=np.int64)
            o = np.argsort(seg_lo)
            SH = np.int64(40)
            bst, ben, bctg = self._block_coords(blocks)
            ivs, ive, ivo = [], [], [0]
            for a in range(G):
                s_, e_, c_ = bst[:, a], ben[:, a] - self.k, bctg[:, a]
                ok = (e_ - s_) >= 2
                base_ = c_[ok] << SH
                ii = IntervalIndex(base_ + s_[ok] + 1, base_ + e_[ok])
                ivs.append(ii.starts); ive.append(ii.maxend); ivo.append(ivo[-1] + len(ii.starts))
            xk = np.fromiter(self._h_extra.keys(), dtype=np.uint64, count=len(self._h_extra))
            xv = np.fromiter(self._h_extra.values(), dtype=np.int64, count=len(self._h_extra))
            xo = np.argsort(xk)
            self._tick("r_blockinfo")
            n_raw, lists = self._dev["refine_filter"](tables, seg_lo[o], seg_hi[o], np.unique(term_ids), xk[xo], xv[xo],
                                                      np.concatenate(ivs) if ivs else np.zeros(0, dtype=np.int64),
                                                      np.concatenate(ive) if ive else np.zeros(0, dtype=np.int64), ivo)
            for t in tables:
                t.close()
            self.stats.setdefault("new_raw", []).append(n_raw)
            common = np.sort(lists[0][0]) if len(lists[0][0]) else np.zeros(0, dtype=np.uint64)
            self._tick("r_filter")
        else:
            new = []
            for a in range(G):
                h1, pos, ctg = self.be.sketch(a, new_w, masks[a])
                new.append(self._dedup(h1, pos.astype(np.int64), ctg.astype(np.int64)))
            self.stats.setdefault("new_raw", []).append([int(len(x[0])) for x in new])
            self._tick("r_sketch")
            # --- terminal / internal minimizers and block intervals (find_mx_in_blocks :205-226)
            term_ids = np.array([x for b in blocks for x in (b.first_id, b.last_id)], dtype=np.int64)
            terminal_h = set(int(x) for x in self.H[term_ids]) if len(term_ids) else set()
            seg_lo = np.array([sg[0] for b in blocks for sg in b.segs], dtype=np.int64)
            seg_hi = np.array([sg[1] for b in blocks for sg in b.segs], dtype=np.int64)
            is_internal = np.zeros(self.V, dtype=bool)
            long_ = np.flatnonzero(seg_hi - seg_lo > 64)
            for lo, hi in zip(seg_lo[long_].tolist(), seg_hi[long_].tolist()):          # the few long runs: slice fills
                is_internal[lo:hi + 1] = True
            short_ = np.flatnonzero(seg_hi - seg_lo <= 64)                                 # the many short ones: one scatter
            if len(short_):
                n_ = seg_hi[short_] - seg_lo[short_] + 1
                off_ = np.repeat(np.cumsum(n_) - n_, n_)
                is_internal[np.repeat(seg_lo[short_], n_) + (np.arange(int(n_.sum())) - off_)] = True
            if len(term_ids):
                is_internal[term_ids] = False
            # open intervals (start, last minimizer) of the blocks per assembly, on one axis: coordinate = contig << 40 | pos
            # (intervals of different contigs cannot meet there, so one index answers the queries of every contig)
            SH = np.int64(40)
            bst, ben, bctg = self._block_coords(blocks)
            intervals = []
            for a in range(G):
                s_, e_, c_ = bst[:, a], ben[:, a] - self.k, bctg[:, a]
                ok = (e_ - s_) >= 2
                base_ = c_[ok] << SH
                intervals.append(IntervalIndex(base_ + s_[ok] + 1, base_ + e_[ok]) if ok.any() else None)
            self._tick("r_blockinfo")
            # --- filter_minimizers_synteny_blocks (:256-280), vectorised over all contigs of an assembly
            # all new minimizers of all assemblies through ONE device lookup
            sizes = [len(x[0]) for x in new]
            all_h1 = np.concatenate([x[0] for x in new]) if sum(sizes) else np.zeros(0, dtype=np.uint64)
            all_vid = self._lookup(all_h1) if len(all_h1) else np.zeros(0, dtype=np.int64)
            offs = np.concatenate([[0], np.cumsum(sizes)])
            kept = []       # per assembly: (h1, pos, ctg, sublist_id)
            for a in range(G):
                h1, pos, ctg = new[a]
                n = len(h1)
                if n == 0:
                    kept.append((h1, pos, ctg, np.zeros(0, dtype=np.int64)))
                    continue
                vid_new = all_vid[offs[a]:offs[a + 1]]
                internal_hit = np.zeros(n, dtype=bool)
                known = vid_new >= 0
                internal_hit[known] = is_internal[vid_new[known]]
                ii = intervals[a]
                key = (ctg.astype(np.int64) << SH) + pos
                inside = ii.overlaps(key, key + 1) if ii is not None else np.zeros(n, dtype=bool)
                keep = ~internal_hit & ~inside
                kh, kp, kc, kk = h1[keep], pos[keep], ctg[keep], key[keep]
                # cut between consecutive kept minimizers of one contig whose span overlaps a block interval
                cut = np.ones(len(kh), dtype=bool)
                if len(kh) > 1:
                    same = kc[1:] == kc[:-1]
                    ov = np.zeros(len(kh) - 1, dtype=bool)
                    if ii is not None:
                        sel = np.flatnonzero(same)
                        ov[sel] = ii.overlaps(kk[:-1][sel], kk[1:][sel])
                    cut[1:] = ~same | ov
                kept.append((kh, kp, kc, np.cumsum(cut) - 1))
            self._tick("r_filter")
            # --- G-way intersection (ntjoin_utils.filter_minimizers :152-165)
            # every assembly's keys are distinct by now (read_minimizers dropped the repeated ones), so a key is common
            # to all G lists iff it occurs G times in their concatenation: one sort instead of G sorts + G-1 merges
            allk = np.concatenate([x[0] for x in kept])
            if len(allk):
                uk, cnt_k = np.unique(allk, return_counts=True)
                common = uk[cnt_k == G]
            else:
                common = allk
            lists = []
            for a in range(G):
                kh, kp, kc, sub = kept[a]
                if len(kh) and len(common):
                    j = np.searchsorted(common, kh)
                    j[j >= len(common)] = 0
                    ok = common[j] == kh
                else:
                    ok = np.zeros(len(kh), dtype=bool)
                lists.append((kh[ok], kp[ok], kc[ok], sub[ok]))
        self.stats.setdefault("new_common", []).append(int(len(common)))
        # --- update_list_mx_info (:282-290) + vertex ids for every surviving minimizer
        ids_per_asm = []
        if len(common):
            cid = self._lookup(common)
            missing = np.nonzero(cid < 0)[0]
            self._grow(len(missing))
            if len(missing):
                nv = np.arange(self.V, self.V + len(missing), dtype=np.int64)
                self.V += len(missing)
                self.H[nv] = common[missing]
                self._h_extra.update(zip(common[missing].tolist(), nv.tolist()))
                self.alive[nv] = False
                self.nbr[nv] = -1
                cid[missing] = nv
            # pass 1: which base vertices get a new position / contig (assembly a only writes row a)
            touched, changed_per_asm = [], []
            cur_P, cur_C = self.POS[:, cid], self.CTG[:, cid]                     # one gather each, in `common` order
            ar_ = np.arange(len(common))
            for a in range(G):
                kh, kp, kc, _ = lists[a]
                # every list holds exactly the common keys, each once: its rank order IS the index into `common`
                # (an argsort and a scatter instead of a binary search per key)
                if len(kh) != len(common):
                    raise RuntimeError("internal error: a filtered minimizer list is not a permutation of the common keys")
                j_ = np.empty(len(kh), dtype=np.int64)
                j_[np.argsort(kh)] = ar_
                vid = cid[j_]
                cur_c = cur_C[a, j_]
                changed = (cur_P[a, j_] != kp) | (cur_c != kc)
                sel = changed & (vid < self.V0)
                tb = vid[sel]
                for v_, c_ in zip(tb.tolist(), cur_c[sel].tolist()):           # keep the round-0 contig of what gets overwritten
                    self._ctg0.setdefault((a, v_), c_)
                touched.append(tb)
                ids_per_asm.append(vid)
                changed_per_asm.append(changed | (vid >= self.V0))
            touched_base = np.unique(np.concatenate(touched)) if touched else np.zeros(0, dtype=np.int64)
            pairs = old = None
            if len(touched_base):
                self.stats["base_overwrites"] = self.stats.get("base_overwrites", 0) + len(touched_base)
                pairs = self._pairs_around(touched_base)
                old = self._pair_values(pairs)
            # pass 2: overwrite
            for a in range(G):
                _, kp, kc, _ = lists[a]
                ch = changed_per_asm[a]                   # (the other entries already hold these values)
                self.POS[a, ids_per_asm[a][ch]] = kp[ch]
                self.CTG[a, ids_per_asm[a][ch]] = kc[ch]
            if pairs is not None:
                self._refresh_pairs(pairs, old)
        else:
            ids_per_asm = [np.zeros(0, dtype=np.int64) for _ in range(G)]
        self._tick("r_update")
        # --- build_graph in extend mode (ntjoin_utils.py:83-141): consecutive pairs of every sub-list, in
        #     (assembly, position) order; an edge's id order is its first appearance, its weight the number of
        #     assemblies (pairs) that support it
        S, T = [], []
        for a in range(G):
            vid, sub = ids_per_asm[a], lists[a][3]
            if len(vid) > 1:
                m_ = sub[:-1] == sub[1:]
                S.append(vid[:-1][m_]); T.append(vid[1:][m_])
        S = np.concatenate(S) if S else np.zeros(0, dtype=np.int64)
        T = np.concatenate(T) if T else np.zeros(0, dtype=np.int64)
        lo_, hi_ = np.minimum(S, T), np.maximum(S, T)
        # distinct pairs with their first appearance and multiplicity: an (unstable) argsort of the keys, group
        # boundaries, and the smallest original index of every group
        ekey = (lo_ << np.int64(32)) | hi_
        if len(ekey):
            so_ = np.argsort(ekey)
            ks_ = ekey[so_]
            gb = np.flatnonzero(np.r_[True, ks_[1:] != ks_[:-1]])
            first_ix = np.minimum.reduceat(so_, gb)
            cnt_ = np.diff(np.r_[gb, len(ks_)])
        else:
            first_ix = cnt_ = np.zeros(0, dtype=np.int64)
        eo = np.argsort(first_ix)
        first_ix, cnt_ = first_ix[eo], cnt_[eo]
        eu, ev = lo_[first_ix], hi_[first_ix]                           # distinct new pairs in insertion order
        # add_vertices: new (or previously deleted) vertices, except the blocks' terminal minimizers
        allv = np.sort(cid) if len(common) else np.zeros(0, dtype=np.int64)     # every list names the same vertices
        if len(allv):
            term_arr = np.unique(self.H[term_ids]) if len(term_ids) else np.zeros(0, dtype=np.uint64)
            revive = allv[~self.alive[allv] & ~np.isin(self.H[allv], term_arr)]
            self.alive[revive] = True
            self.nbr[revive] = -1
        # edges already in the graph are skipped (either orientation)
        if len(eu):
            has = (self.nbr[eu, 0] == ev) | (self.nbr[eu, 1] == ev)
            eu, ev, cnt_ = eu[~has], ev[~has], cnt_[~has]
            if not (self.alive[eu].all() and self.alive[ev].all()):
                raise RuntimeError("internal error: edge to a vertex that is not in the graph")
        cnt_ = cnt_.astype(np.int64)
        self._edge_birth.add_round(round_no, eu, ev)
        # incident-weight guard (check_added_edges_incident_weights :70-80), per touched vertex
        fl = np.zeros(len(eu), dtype=bool)
        cand = np.zeros(0, dtype=np.int64)
        at_cand = np.zeros(0, dtype=np.int64)
        if len(eu):
            ends = np.concatenate([eu, ev])
            tv, inv_ = np.unique(ends, return_inverse=True)
            new_w = np.bincount(inv_, weights=np.concatenate([cnt_, cnt_]), minlength=len(tv))
            new_deg = np.bincount(inv_, minlength=len(tv))
            old_deg = (self.nbr[tv] >= 0).sum(axis=1)
            heavy = (G * old_deg + new_w) > 2 * G
            fl = heavy[inv_[:len(eu)]] | heavy[inv_[len(eu):]]
            # simplification candidates: touched vertices with exactly three neighbours in the extended graph
            cand_mask = (old_deg + new_deg) == 3
            cand = tv[cand_mask]                       # sorted
            if len(cand):
                at_cand = np.flatnonzero(cand_mask[inv_[:len(eu)]] | cand_mask[inv_[len(eu):]])
        # --- simplification runs on the graph WITH the flagged edges; its weight bumps survive only
        #     when nothing was flagged (same object), its vertex deletions never do -- so with a
        #     flagged edge there is nothing to compute
        if self.simplify and len(cand) and not fl.any():
            cnt_[self._simplify_extended(eu, ev, cnt_, cand, at_cand)] = G
        self._tick("r_graph")
        # --- weight filter (+ flagged pairs on the last round)
        full = ~fl & (cnt_ >= G)
        if full.any():
            self._add_edges(eu[full], ev[full])
        low = ~fl & (cnt_ < G)
        if last_round and low.any():
            self._erode(np.stack([eu[low], ev[low]], axis=1))
        return None

    def _simplify_extended(self, eu, ev, wt, cand, at_cand):
        """run_graph_simplification on the extended graph; returns the indices of the NEW edges whose weight it sets to G
        (bumps of old edges change nothing: they are at full weight already).
        eu / ev / wt: the new edges in insertion order and their weights; cand: the touched vertices with exactly three
        neighbours (sorted); at_cand: indices of the new edges with an end in `cand`.

        A candidate edge joins two candidates; it is bumped when both ends have exactly one full-weight edge at that
        moment (a bump adds one at both ends) and share exactly one other neighbour.  The shared-neighbour test does
        not depend on the visiting order, so it is taken for all candidate edges at once on [n, 3] neighbour tables;
        only the edges that pass it are visited in edge-id order."""
        G = self.G
        n = len(cand)
        # neighbour table: the old neighbours (weight G) and the other ends of the new edges, three per candidate
        rows = self.nbr[cand]
        ci_old, slot_old = np.nonzero(rows >= 0)
        a_, b_ = eu[at_cand], ev[at_cand]
        pa, pb = np.searchsorted(cand, a_), np.searchsorted(cand, b_)
        pa[pa >= n] = 0; pb[pb >= n] = 0
        ia, ib = cand[pa] == a_, cand[pb] == b_
        ci = np.concatenate([ci_old, pa[ia], pb[ib]])
        nb = np.concatenate([rows[ci_old, slot_old].astype(np.int64), b_[ia], a_[ib]])
        ww = np.concatenate([np.full(len(ci_old), G, dtype=np.int64), wt[at_cand][ia], wt[at_cand][ib]])
        eid = np.concatenate([np.full(len(ci_old), -1, dtype=np.int64), at_cand[ia], at_cand[ib]])     # new edge index, -1 = old
        if len(ci) != 3 * n:
            raise RuntimeError("internal error: a simplification candidate without exactly three neighbours")
        o = np.argsort(ci, kind="stable")
        N3, W3, E3 = nb[o].reshape(n, 3), ww[o].reshape(n, 3), eid[o].reshape(n, 3)
        # candidate edges: (candidate, neighbour) entries whose neighbour is a candidate too, once per edge
        pos = np.searchsorted(cand, N3)
        pos[pos >= n] = 0
        take = (cand[pos] == N3) & (cand[:, None] < N3)
        si, sl = np.nonzero(take)
        if not len(si):
            return []
        ti = pos[si, sl]
        t_v = N3[si, sl]
        A, B = N3[si], N3[ti]
        shared = ((A[:, :, None] == B[:, None, :]).any(axis=2) & (A != t_v[:, None])).sum(axis=1)
        ok = shared == 1
        if not ok.any():
            return []
        si, sl, ti, t_v = si[ok], sl[ok], ti[ok], t_v[ok]
        w_e, e_e = W3[si, sl], E3[si, sl]
        # edge-id order: old edges first (their relative order only matters among themselves), then the new ones in
        # insertion order
        old_ix = np.flatnonzero(e_e < 0).tolist()
        new_ix = np.flatnonzero(e_e >= 0)
        new_ix = new_ix[np.argsort(e_e[new_ix], kind="stable")].tolist()
        if old_ix:
            keys = dict(zip(old_ix, zip(cand[si[old_ix]].tolist(), t_v[old_ix].tolist())))
            self._prefetch_edge_keys([k_ for k_ in keys.values() if k_ not in self._edge_birth])
            old_ix.sort(key=lambda i: self._old_edge_key(keys[i]))
        fullc = (W3 == G).sum(axis=1).tolist()      # full-weight edges at a candidate; a bump adds one at both ends
        si_l, ti_l, w_l, e_l = si.tolist(), ti.tolist(), w_e.tolist(), e_e.tolist()
        out = []
        for i in old_ix + new_ix:
            a, b = si_l[i], ti_l[i]
            if fullc[a] == 1 and fullc[b] == 1:
                if w_l[i] != G:
                    fullc[a] += 1; fullc[b] += 1
                if e_l[i] >= 0:
                    out.append(e_l[i])
        return out

    def _old_edge_key(self, key):
        if key in self._edge_birth:
            return self._edge_birth[key]
        return self._edge_key0(*key)

    def _erode(self, low):
        """refine_graph + erode_edges (ntsynt_synteny.py:305-362) for the low-weight edges removed
        on the last round, in edge order"""
        to_remove = set()

        def name(x):
            return str(int(self.H[x]))

        pos_memo = {}

        def pos_of(x):
            got = pos_memo.get(x)
            if got is None:
                got = pos_memo[x] = [int(v) for v in self.POS[:, x]]
            return got

        def overlap(s, t):
            k = self.k
            return any(abs(a - b) < k for a, b in zip(pos_of(s), pos_of(t)))

        # the graph is not modified inside the loop (removals are applied at the end), so the degree test of every
        # flagged pair can be taken up front
        la = np.array(low, dtype=np.int64).reshape(-1, 2)
        both1 = ((self.nbr[la[:, 0]] >= 0).sum(axis=1) == 1) & ((self.nbr[la[:, 1]] >= 0).sum(axis=1) == 1)
        sel = la[both1]
        if len(sel):
            # the walk of a pair starts only if its two vertices are less than k apart in some assembly: that first
            # test for all pairs at once, the walk for the few that pass it
            P = np.asarray(self.POS[:, sel.ravel()], dtype=np.int64).reshape(self.G, -1, 2)
            sel = sel[(np.abs(P[:, :, 0] - P[:, :, 1]) < self.k).any(axis=0)]
        hv = self.H[sel.ravel()].reshape(-1, 2) if len(sel) else np.zeros((0, 2), dtype=np.uint64)
        if self._dev is not None and len(sel):
            # positions around the flagged pairs in one gather (the walk below reads them one vertex at a time)
            ring = sel.ravel()
            for _ in range(3):
                nb = self.nbr[ring].ravel()
                ring = np.unique(np.concatenate([ring, nb[nb >= 0]]))
            self.POS.prefetch(ring)
            cols = np.asarray(self.POS[:, ring], dtype=np.int64)
            pos_memo.update(zip(ring.tolist(), cols.T.tolist()))
        for (u, v), (hu, hv_) in zip(sel.tolist(), hv.tolist()):
            # igraph reports (source, target) = (min id, max id); the reference then orders by NAME string
            s, t = (u, v)
            if str(int(hu)) > str(int(hv_)):
                s, t = t, s
            erode_target = True
            cs, ct = s, t
            visited = {cs, ct}
            while overlap(cs, ct):
                ev = ct if erode_target else cs
                for y in self.nbr[ev]:
                    if y >= 0:
                        to_remove.add((min(ev, int(y)), max(ev, int(y))))
                cand = [int(y) for y in self.nbr[ev] if y >= 0 and int(y) not in visited]
                if not cand:
                    break
                assert len(cand) == 1
                if erode_target:
                    ct = cand[0]; erode_target = False; visited.add(ct)
                else:
                    cs = cand[0]; erode_target = True; visited.add(cs)
        if to_remove:
            us, vs = zip(*to_remove)
            self._remove_edges(np.array(us), np.array(vs))

    # ------------------------------------------------------------------ driver (main_synteny :593-647)
    def run(self):
        """the whole graph stage.  The cyclic collector is paused for its duration: the stage makes a few hundred
        thousand small objects and no reference cycles, and one full collection in a process with large libraries
        loaded costs tens of milliseconds"""
        import gc
        was_enabled = gc.isenabled()
        gc.disable()
        try:
            return self._run()
        finally:
            if was_enabled:
                gc.enable()

    def _run(self):
        G = self.G
        if len(self.w_rounds) != len(set(self.w_rounds)):
            print("Error: duplicate values found in w_rounds!", file=sys.stderr, flush=True)
            raise SystemExit(1)
        self.log("Sketching and joining minimizers, w =", self.w)
        self._tick(None)
        tables = [self.be.sketch(a, self.w, None) for a in range(G)]
        self._tick("sketch0")
        j = self.be.join(tables, self.orient)
        self._tick("join")
        link, degree = j.get("link"), j.get("degree")
        self.stats["vertices"] = int(j["V"]) if "V" in j else int(len(j["H"]))
        if getattr(self, "dot_path", None):
            self.log("Printing graph", self.dot_path)
            self.be.write_dot(self.dot_path, j)
        self._init_vertices(j)
        self._tick("init")
        self._edge_birth = EdgeBirth()
        V = self.V0
        # --- simplification, weight filter
        bumped, removed = ({}, [])
        if self.simplify:
            self.log("Running graph simplificaton")
            bumped, removed = self._simplify_round0(link, degree)
        self.stats["simplified_vertices"] = len(set(removed))
        self._tick("simplify0")
        self.log("Filtering the graph")
        if V > 1 and not self._prebuilt:
            self.conn[:] = np.asarray(link[:V - 1], dtype=bool)
            ar = np.arange(1, V, dtype=np.int32)
            self.nbr[:V - 1, 1] = np.where(self.conn, ar, -1)        # slot 1: right neighbour
            self.nbr[1:V, 0] = np.where(self.conn, ar - 1, -1)       # slot 0: left neighbour
        if removed:
            self._remove_vertices(np.array(removed, dtype=np.int64))
        if bumped:
            be_ = np.array(list(bumped.keys()), dtype=np.int64)
            be_ = be_[self.alive[be_[:, 0]] & self.alive[be_[:, 1]]]
            has = (self.nbr[be_[:, 0], 0] == be_[:, 1]) | (self.nbr[be_[:, 0], 1] == be_[:, 1])
            self._add_edges(be_[~has, 0], be_[~has, 1])
        # --- paths, blocks
        self.log("Finding paths")
        self._tick("filter0")
        self.log("Finding synteny blocks")
        blocks = self._extract_blocks()
        if self.interarrivals:
            self._print_interarrivals(blocks)
        ordered = self._sort_blocks(blocks)
        self._tick("filter_sort")
        if not ordered:
            print("Error - no paths found. Try adjusting the specified k/w parameters.")
            raise SystemExit(1)
        if self.write_files or not self.w_rounds:      # with refinement rounds the final table replaces this one
            self._emit("initial", ordered)
        self.log("Done initial synteny blocks")
        # --- refinement rounds (refine_block_coordinates :476-530)
        prev_w = self.w
        for ri, new_w in enumerate(self.w_rounds):
            self.log("Extending synteny blocks with w =", new_w)
            last = new_w == self.w_rounds[-1]
            self._tick("emit")
            self._refine_round(blocks, new_w, prev_w, last, ri + 1)
            self._tick("refine")
            blocks = self._extract_blocks()
            ordered = self._sort_blocks(blocks)
            self._tick("filter_sort")
            if last or self.write_files:               # every round overwrites the same table: only the last one stays
                self._emit("pre_merge", ordered)
            if last:
                merged = self._merge_collinear(ordered) if ordered else []
                if merged:
                    st_, en_, _ = self._block_coords(merged)
                    merged = [merged[i] for i in np.flatnonzero(((en_ - st_) >= self.z).all(axis=1)).tolist()]
                merged = self._merge_collinear(merged) if merged else []
                if self.dev:
                    self._check_non_overlapping(merged)
                self._emit("final", merged, verbose=True)
            prev_w = new_w
        self._tick("emit")
        self.log("Done extended synteny blocks")
        return self.outputs.get("final", self.outputs.get("initial"))
