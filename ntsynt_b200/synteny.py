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


def _log(*a):
    print(datetime.datetime.today(), ":", *a, file=sys.stdout, flush=True)


class Block:
    "one synteny block: a path of vertices with one contig and one orientation per assembly"
    __slots__ = ("vids", "ctg", "ori", "first_pos", "last_pos", "n", "broken_reason")

    def __init__(self, vids, ctg, ori, first_pos, last_pos, n):
        self.vids = vids              # int64 vertex ids in path order (None after a collinear merge)
        self.ctg = ctg                # [G] contig index per assembly
        self.ori = ori                # [G] '+', '-'
        self.first_pos = first_pos    # [G] position of the first minimizer
        self.last_pos = last_pos      # [G] position of the last minimizer
        self.n = n                    # number of minimizers
        self.broken_reason = None

    def start(self, a):               # bin/assembly_block.py:17-19
        return min(int(self.first_pos[a]), int(self.last_pos[a]))

    def end(self, a, k):              # bin/assembly_block.py:21-23
        return max(int(self.first_pos[a]), int(self.last_pos[a])) + k


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


class SyntenyEngine:
    """backend must provide:
         names[a]            TSV-style assembly names, ALREADY in the reference's processing order
                             (reverse-sorted, bin/ntsynt_synteny.py:34)
         contig_names[a][c], contig_lengths[a][c]
         sketch(a, w, masks) -> (h1 u64, pos u32, ctg u32) in (contig, position) order;
                             masks = per-contig (starts, ends) extra N intervals or None
         join(tables, order_asm) -> (H, POS, CTG, RANK, link, degree)   (device kernel iv)
    """

    def __init__(self, backend, k, w, w_rounds, bp, collinear_merge, z, m=90, simplify=True, prefix="out",
                 dev=False, write_files=True, quiet=False):
        self.be = backend
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

    def log(self, *a):
        if not self.quiet:
            _log(*a)

    # ------------------------------------------------------------------ vertex storage
    def _init_vertices(self, H, POS, CTG, RANK):
        V = len(H)
        self.V0 = V
        self.V = V
        cap = V + 1024
        self.H = np.empty(cap, dtype=np.uint64); self.H[:V] = H
        self.POS = np.zeros((self.G, cap), dtype=np.int64); self.POS[:, :V] = POS
        self.CTG = np.zeros((self.G, cap), dtype=np.int64); self.CTG[:, :V] = CTG
        self.RANK = np.asarray(RANK, dtype=np.int64)                 # round-0 vertices only
        self.INV = np.empty_like(self.RANK)
        ar = np.arange(V, dtype=np.int64)
        for a in range(self.G):
            self.INV[a, self.RANK[a]] = ar
        self.alive = np.zeros(cap, dtype=bool); self.alive[:V] = True
        self.nbr = np.full((cap, 2), -1, dtype=np.int64)
        # h1 -> id lookup: sorted view of the round-0 keys + dict for later additions
        self._h_order = np.argsort(self.H[:V], kind="stable")
        self._h_sorted = self.H[:V][self._h_order]
        self._h_extra = {}

    def _grow(self, need):
        cap = len(self.H)
        if self.V + need <= cap:
            return
        new = max(cap * 2, self.V + need + 1024)
        self.H = np.concatenate([self.H, np.empty(new - cap, dtype=np.uint64)])
        self.POS = np.concatenate([self.POS, np.zeros((self.G, new - cap), dtype=np.int64)], axis=1)
        self.CTG = np.concatenate([self.CTG, np.zeros((self.G, new - cap), dtype=np.int64)], axis=1)
        self.alive = np.concatenate([self.alive, np.zeros(new - cap, dtype=bool)])
        self.nbr = np.concatenate([self.nbr, np.full((new - cap, 2), -1, dtype=np.int64)])

    def _lookup(self, keys):
        "vertex id per h1 (or -1) for a uint64 array"
        keys = np.asarray(keys, dtype=np.uint64)
        out = np.full(len(keys), -1, dtype=np.int64)
        if len(self._h_sorted) and len(keys):
            i = np.searchsorted(self._h_sorted, keys)
            i[i >= len(self._h_sorted)] = 0
            hit = self._h_sorted[i] == keys
            out[hit] = self._h_order[i[hit]]
        if self._h_extra:
            for j in np.nonzero(out < 0)[0]:
                out[j] = self._h_extra.get(int(keys[j]), -1)
        return out

    # ------------------------------------------------------------------ degree-2 graph on arrays
    def _has_edge(self, u, v):
        return self.nbr[u, 0] == v or self.nbr[u, 1] == v

    def _add_edge(self, u, v):
        for x, y in ((u, v), (v, u)):
            if self.nbr[x, 0] < 0:
                self.nbr[x, 0] = y
            elif self.nbr[x, 1] < 0:
                self.nbr[x, 1] = y
            else:
                raise RuntimeError("internal error: vertex of degree > 2 in the weight-filtered graph")

    def _remove_edges(self, us, vs):
        us = np.asarray(us, dtype=np.int64); vs = np.asarray(vs, dtype=np.int64)
        for x, y in ((us, vs), (vs, us)):
            for s in (0, 1):
                hit = self.nbr[x, s] == y
                self.nbr[x[hit], s] = -1

    def _remove_vertices(self, ids):
        ids = np.unique(np.asarray(ids, dtype=np.int64))
        if not len(ids):
            return
        for s in (0, 1):
            nb = self.nbr[ids, s]
            ok = nb >= 0
            self._remove_edges(ids[ok], nb[ok])
        self.alive[ids] = False

    def _degree(self, ids=None):
        n = self.nbr if ids is None else self.nbr[ids]
        return (n >= 0).sum(axis=1)

    # ------------------------------------------------------------------ round-0 adjacency (implicit in ranks)
    def _adjacent(self, a, u, v):
        if u >= self.V0 or v >= self.V0:
            return False
        ru, rv = self.RANK[a, u], self.RANK[a, v]
        return abs(int(ru) - int(rv)) == 1 and self._ctg0[a, u] == self._ctg0[a, v]

    def _neighbors0(self, u):
        "distinct round-0 neighbours of u with their weights (number of supporting assemblies)"
        res = {}
        for a in range(self.G):
            r = int(self.RANK[a, u])
            for rr in (r - 1, r + 1):
                if 0 <= rr < self.V0:
                    x = int(self.INV[a, rr])
                    if self._ctg0[a, x] == self._ctg0[a, u]:
                        res[x] = res.get(x, 0) + 1
        return res

    def _edge_key0(self, u, v):
        "position of edge {u,v} in build_graph's formatted_edges order for round 0 (ntjoin_utils.py:97-115)"
        def first_new(src, dst):
            for a in range(self.G):
                if self._adjacent(a, src, dst):
                    return a
            return None
        a0 = first_new(u, v)
        r = min(int(self.RANK[a0, u]), int(self.RANK[a0, v]))
        src = u if int(self.RANK[a0, u]) < int(self.RANK[a0, v]) else v
        tau = (a0, r)
        sigma = None
        for a in range(self.G):                      # first time src is the left element of a NEW pair
            rs = int(self.RANK[a, src])
            if rs + 1 < self.V0:
                x = int(self.INV[a, rs + 1])
                if self._ctg0[a, x] == self._ctg0[a, src] and first_new(src, x) == a:
                    sigma = (a, rs)
                    break
        return (0, sigma, tau)

    # ------------------------------------------------------------------ simplification (ntsynt_synteny.py:566-590)
    def _simplify_round0(self, link, degree):
        G = self.G
        cand_v = np.nonzero(degree == 3)[0]
        bumped = {}
        removed = []
        if len(cand_v):
            cset = set(int(x) for x in cand_v)
            nb_cache = {}

            def nbrs(u):
                if u not in nb_cache:
                    nb_cache[u] = self._neighbors0(u)
                return nb_cache[u]

            edges = set()
            for u in cset:
                for x in nbrs(u):
                    if x in cset:
                        edges.add((min(u, x), max(u, x)))
            order = sorted(edges, key=lambda e: self._edge_key0(*e))

            def weight(u, x):
                return bumped.get((min(u, x), max(u, x)), nbrs(u)[x])

            def anchored(u):
                return sum(1 for x in nbrs(u) if weight(u, x) == G) == 1

            for s, t in order:
                if anchored(s) and anchored(t):
                    common = [x for x in nbrs(s) if x != t and x in nbrs(t)]
                    if len(common) == 1:            # the edge itself + exactly one 2-step path
                        removed.append(common[0])
                        bumped[(s, t)] = G
        return bumped, removed

    # ------------------------------------------------------------------ paths of a max-degree-2 graph
    def _find_paths(self):
        """ntjoin.py:114-151 on the weight-filtered graph: every component that is a simple path
        with two distinct ends gives one path, oriented from the end with the smaller position in
        the orienting assembly.  Returns a list of int64 id arrays."""
        V = self.V
        nbr = self.nbr[:V]
        deg = (nbr >= 0).sum(axis=1)
        ids = np.arange(V, dtype=np.int64)
        right = (nbr[:, 0] == ids + 1) | (nbr[:, 1] == ids + 1)          # simple link i -> i+1
        # runs of consecutive ids joined by simple links
        start = np.ones(V, dtype=bool)
        start[1:] = ~right[:-1]
        run_start = np.nonzero(start)[0]
        run_end = np.append(run_start[1:] - 1, V - 1) if V else run_start
        run_len = run_end - run_start + 1
        # sparse (non i,i+1) edges
        left = np.zeros(V, dtype=bool)
        left[1:] = right[:-1]
        n_simple = right.astype(np.int64) + left.astype(np.int64)
        has_sparse = deg > n_simple
        opos = self.POS[self.orient]
        paths = []
        # pure runs: no sparse edge at either end
        pure = (~has_sparse[run_start]) & (~has_sparse[run_end]) & (run_len >= 2)
        for a, b in zip(run_start[pure], run_end[pure]):
            pa, pb = opos[a], opos[b]
            if pa < pb:
                paths.append(np.arange(a, b + 1, dtype=np.int64))
            elif pb < pa:
                paths.append(np.arange(b, a - 1, -1, dtype=np.int64))
        # components with sparse edges: walk run by run from every degree-1 end
        if has_sparse.any():
            run_of = np.cumsum(start) - 1
            visited_runs = set()
            sp_runs = np.unique(np.concatenate([run_of[np.nonzero(has_sparse)[0]]]))
            ends = []
            for rid in sp_runs:
                for x in {int(run_start[rid]), int(run_end[rid])}:
                    if deg[x] == 1:
                        ends.append(x)
            # also ends of runs reachable only through sparse edges are found by walking
            for e0 in ends:
                if int(run_of[e0]) in visited_runs:
                    continue
                seq = []
                prev, cur = -1, e0
                ok = True
                while True:
                    rid = int(run_of[cur])
                    if rid in visited_runs:
                        ok = False       # cycle guard (cannot happen from a degree-1 start)
                        break
                    visited_runs.add(rid)
                    a, b = int(run_start[rid]), int(run_end[rid])
                    if cur == a:
                        seq.append(np.arange(a, b + 1, dtype=np.int64)); last = b
                    else:
                        seq.append(np.arange(b, a - 1, -1, dtype=np.int64)); last = a
                    # leave the run through the sparse neighbour of `last`
                    inside = (last - 1 if last == b and b > a else (last + 1 if last == a and b > a else -1))
                    nxt = -1
                    for s in (0, 1):
                        y = int(nbr[last, s])
                        if y >= 0 and y != inside and y != (prev if a == b else -2):
                            nxt = y
                    if a == b and nxt < 0:
                        # singleton run: both slots may be sparse; pick the one that is not `prev`
                        cand = [int(nbr[last, s]) for s in (0, 1) if nbr[last, s] >= 0 and nbr[last, s] != prev]
                        nxt = cand[0] if cand else -1
                    if nxt < 0:
                        break
                    prev, cur = last, nxt
                if not ok:
                    continue
                p = np.concatenate(seq)
                if len(p) < 2:
                    continue
                pa, pb = opos[p[0]], opos[p[-1]]
                if pa < pb:
                    paths.append(p)
                elif pb < pa:
                    paths.append(p[::-1].copy())
        return paths

    # ------------------------------------------------------------------ blocks from paths (vectorised)
    def _blocks_from_paths(self, paths):
        """find_synteny_blocks (ntsynt_synteny.py:66-106) for every path.  Returns blocks; vertices of
        unoriented blocks are deleted from the graph."""
        if not paths:
            return []
        G = self.G
        lens = np.array([len(p) for p in paths], dtype=np.int64)
        pv = np.concatenate(paths)
        off = np.concatenate([[0], np.cumsum(lens)])
        N = len(pv)
        pid = np.repeat(np.arange(len(paths)), lens)
        ctg = self.CTG[:, pv]                         # [G, N]
        pos = self.POS[:, pv]
        # contig change between consecutive path vertices (any assembly)
        same_path = np.zeros(N, dtype=bool)
        same_path[1:] = pid[1:] == pid[:-1]
        chg = np.zeros(N, dtype=bool)
        chg[1:] = (ctg[:, 1:] != ctg[:, :-1]).any(axis=0) & same_path[1:]
        # only the LAST run of constant contigs of each path becomes a block (past_start_flag is
        # never set, ntsynt_synteny.py:71,77): block start = last change index, else path start
        idx = np.arange(N, dtype=np.int64)
        last_chg = np.full(len(paths), -1, dtype=np.int64)
        if chg.any():
            np.maximum.at(last_chg, pid[chg], idx[chg])
        bstart = np.where(last_chg >= 0, last_chg, off[:-1])
        bend = off[1:]
        blocks, to_remove = [], []
        # orientation tallies per assembly over consecutive pairs inside the block
        inside = np.zeros(N, dtype=bool)             # pair (j-1, j) inside the block
        inside[1:] = same_path[1:] & (idx[1:] > bstart[pid[1:]])
        inc = np.zeros((G, len(paths)), dtype=np.int64)
        dec = np.zeros((G, len(paths)), dtype=np.int64)
        if N > 1:
            d = pos[:, 1:] - pos[:, :-1]
            m_in = inside[1:]
            for a in range(G):
                np.add.at(inc[a], pid[1:][m_in & (d[a] > 0)], 1)
                np.add.at(dec[a], pid[1:][m_in & (d[a] < 0)], 1)
        for p in range(len(paths)):
            s, e = int(bstart[p]), int(bend[p])
            n = e - s
            ori = []
            for a in range(G):
                if n == 1 or inc[a, p] == n - 1:
                    ori.append("+")
                elif dec[a, p] == n - 1:
                    ori.append("-")
                else:
                    positive = int(inc[a, p]) / float(n - 1) * 100
                    negative = 100 - positive
                    ori.append("+" if positive >= self.m else ("-" if negative >= self.m else "?"))
            vids = pv[s:e]
            if "?" in ori:
                to_remove.append(vids)
                continue
            blocks.append(Block(vids, ctg[:, s].copy(), ori, pos[:, s].copy(), pos[:, e - 1].copy(), n))
        if to_remove:
            self._remove_vertices(np.concatenate(to_remove))
        return blocks

    def _split_indels(self, blocks):
        "check_for_indels + break_synteny_block (ntsynt_synteny.py:364-409)"
        out = []
        rm_u, rm_v = [], []
        for b in blocks:
            if b.n < 2:
                out.append(b)
                continue
            pos = self.POS[:, b.vids]
            d = np.abs(pos[:, 1:] - pos[:, :-1])
            spread = d.max(axis=0) - d.min(axis=0)
            brk = np.nonzero(spread > self.bp)[0]
            if not len(brk):
                out.append(b)
                continue
            rm_u.append(b.vids[brk]); rm_v.append(b.vids[brk + 1])
            cuts = [0] + [int(x) + 1 for x in brk] + [b.n]
            for s, e in zip(cuts[:-1], cuts[1:]):
                vids = b.vids[s:e]
                out.append(Block(vids, b.ctg, list(b.ori), pos[:, s].copy(), pos[:, e - 1].copy(), e - s))
        if rm_u:
            self._remove_edges(np.concatenate(rm_u), np.concatenate(rm_v))
        return out

    def _filter_small(self, blocks, min_mx):
        "filter_synteny_blocks (ntsynt_synteny.py:411-426)"
        keep, rm = [], []
        for b in blocks:
            if b.n >= min_mx:
                keep.append(b)
            else:
                rm.append(b.vids)
        if rm:
            self._remove_vertices(np.concatenate(rm))
        return keep

    # ------------------------------------------------------------------ output
    def _sort_blocks(self, blocks):
        "SyntenyBlock.__lt__ (synteny_block.py:102-109): contig NAME string, then start, of the smallest assembly"
        a = self.smallest
        cn = self.be.contig_names[a]
        return sorted(blocks, key=lambda b: (cn[int(b.ctg[a])], b.start(a)))

    def _long_enough(self, b):
        return all(b.end(a, self.k) - b.start(a) >= self.z for a in range(self.G))

    def _block_rows(self, b, num, verbose=False):
        rows = []
        for a in self.name_order:
            row = (f"{num}\t{self.labels[a]}\t{self.be.contig_names[a][int(b.ctg[a])]}\t{b.start(a)}"
                   f"\t{b.end(a, self.k)}\t{b.ori[a]}\t{b.n}")
            if verbose:
                row = f"{row.strip()}\t{b.broken_reason}"
            rows.append(row + "\n")
        return "".join(rows)

    def _emit(self, key, blocks, verbose=False):
        text, num = [], 0
        for b in blocks:
            if not self._long_enough(b):
                continue
            text.append(self._block_rows(b, num, verbose))
            num += 1
        text = "".join(text)
        self.outputs[key] = text
        if self.write_files:
            suffix = {"initial": ".synteny_blocks.tsv", "pre_merge": ".pre-collinear-merge.synteny_blocks.tsv",
                      "final": ".synteny_blocks.tsv"}[key]
            with open(f"{self.prefix}{suffix}", "w", encoding="utf-8") as fh:
                fh.write(text)
        return text

    # ------------------------------------------------------------------ merge (ntsynt_synteny.py:428-472)
    def _gap(self, b1, b2, a):
        if b1.ori[a] == "-" and b2.ori[a] == "-":
            return b1.start(a) - b2.end(a, self.k)
        return b2.start(a) - b1.end(a, self.k)

    def _merge_collinear(self, blocks):
        out = []
        cur = blocks[0]
        for b in blocks[1:]:
            same_ori = all(cur.ori[a] == b.ori[a] for a in range(self.G))
            same_ctg = all(int(cur.ctg[a]) == int(b.ctg[a]) for a in range(self.G))
            diffs = [self._gap(cur, b, a) for a in range(self.G)]
            spread = max(diffs) - min(diffs)
            if (not same_ori) or (not same_ctg) or spread > self.bp - self.k or max(diffs) >= self.collinear_merge:
                if not same_ctg:
                    b.broken_reason = "id_change"
                elif not same_ori:
                    b.broken_reason = "ori_change"
                elif any(x < 0 for x in diffs):
                    b.broken_reason = "inconsistent_order"
                elif spread > self.bp - self.k:
                    b.broken_reason = "indel"
                elif max(diffs) >= self.collinear_merge:
                    b.broken_reason = "merge"
                out.append(cur)
                cur = b
            else:
                # extend: coordinates come from the first minimizer of `cur` and the last of `b`
                cur.last_pos = b.last_pos
                cur.n += b.n
                cur.vids = None
        out.append(cur)
        return out

    # ------------------------------------------------------------------ refinement helpers
    def _masks_for(self, blocks, prev_w):
        "get_synteny_bed_lists + mask_assemblies_with_synteny_extents (ntsynt_synteny.py:117-157)"
        thr = max(2 * prev_w, prev_w + self.k + 1)
        shrink = prev_w + self.k
        masks = []
        for a in range(self.G):
            per = defaultdict(list)
            for b in blocks:
                s, e = b.start(a), b.end(a, self.k)
                if e - s > thr:
                    c = int(b.ctg[a])
                    s2 = max(s + shrink, 0)
                    e2 = min(e - shrink, int(self.be.contig_lengths[a][c]))
                    if s2 < e2:      # an interval the negative slop empties is dropped (unpinned)
                        per[c].append((s2, e2))
            lst = []
            for c in range(len(self.be.contig_names[a])):
                iv = sorted(per.get(c, []))
                # union (bedtools maskfasta semantics): merge overlapping intervals
                ms, me = [], []
                for s, e in iv:
                    if ms and s <= me[-1]:
                        me[-1] = max(me[-1], e)
                    else:
                        ms.append(s); me.append(e)
                lst.append((np.array(ms, dtype=np.uint64), np.array(me, dtype=np.uint64)))
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
        new = []
        for a in range(G):
            h1, pos, ctg = self.be.sketch(a, new_w, masks[a])
            new.append(self._dedup(h1, pos.astype(np.int64), ctg.astype(np.int64)))
        # --- terminal / internal minimizers and block intervals (find_mx_in_blocks :205-226)
        term_ids = np.array([x for b in blocks for x in (int(b.vids[0]), int(b.vids[-1]))], dtype=np.int64)
        terminal_h = set(int(x) for x in self.H[term_ids]) if len(term_ids) else set()
        internal_ids = [b.vids[1:-1] for b in blocks if b.n > 2]
        internal_h = np.sort(self.H[np.concatenate(internal_ids)]) if internal_ids else np.zeros(0, dtype=np.uint64)
        intervals = [defaultdict(list) for _ in range(G)]
        for b in blocks:
            for a in range(G):
                s, e = b.start(a), max(int(b.first_pos[a]), int(b.last_pos[a]))
                if e - s < 2:
                    continue
                intervals[a][int(b.ctg[a])].append((s + 1, e))
        # --- filter_minimizers_synteny_blocks (:256-280), vectorised per contig
        kept = []       # per assembly: (h1, pos, ctg, sublist_id)
        for a in range(G):
            h1, pos, ctg = new[a]
            n = len(h1)
            if n == 0:
                kept.append((h1, pos, ctg, np.zeros(0, dtype=np.int64)))
                continue
            is_internal = np.zeros(n, dtype=bool)
            if len(internal_h):
                j = np.searchsorted(internal_h, h1)
                j[j >= len(internal_h)] = 0
                is_internal = internal_h[j] == h1
            inside = np.zeros(n, dtype=bool)
            idx_by_ctg = {}
            for c in np.unique(ctg):
                c = int(c)
                if c in intervals[a]:
                    st, en = zip(*intervals[a][c])
                    ii = IntervalIndex(st, en)
                    idx_by_ctg[c] = ii
                    sel = np.nonzero(ctg == c)[0]
                    inside[sel] = ii.overlaps(pos[sel], pos[sel] + 1)
            keep = ~is_internal & ~inside
            kh, kp, kc = h1[keep], pos[keep], ctg[keep]
            # cut between consecutive kept minimizers of one contig whose span overlaps a block interval
            cut = np.ones(len(kh), dtype=bool)
            if len(kh) > 1:
                same = kc[1:] == kc[:-1]
                ov = np.zeros(len(kh) - 1, dtype=bool)
                for c, ii in idx_by_ctg.items():
                    sel = np.nonzero(same & (kc[1:] == c))[0]
                    if len(sel):
                        ov[sel] = ii.overlaps(kp[:-1][sel], kp[1:][sel])
                cut[1:] = ~same | ov
            kept.append((kh, kp, kc, np.cumsum(cut) - 1))
        # --- G-way intersection (ntjoin_utils.filter_minimizers :152-165)
        common = None
        for a in range(G):
            hs = np.unique(kept[a][0])
            common = hs if common is None else np.intersect1d(common, hs, assume_unique=True)
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
            for j in missing:
                vid = self.V
                self.V += 1
                self.H[vid] = common[j]
                self._h_extra[int(common[j])] = vid
                self.alive[vid] = False
                self.nbr[vid] = -1
                cid[j] = vid
            for a in range(G):
                kh, kp, kc, _ = lists[a]
                j = np.searchsorted(common, kh)
                vid = cid[j]
                self.POS[a, vid] = kp
                self.CTG[a, vid] = kc
                ids_per_asm.append(vid)
        else:
            ids_per_asm = [np.zeros(0, dtype=np.int64) for _ in range(G)]
        # --- build_graph in extend mode (ntjoin_utils.py:83-141)
        new_edges = {}          # (min,max) -> [support count, (s,t) as first inserted]
        new_order = []
        for a in range(G):
            vid, sub = ids_per_asm[a], lists[a][3]
            for i in range(len(vid) - 1):
                if sub[i] != sub[i + 1]:
                    continue
                s, t = int(vid[i]), int(vid[i + 1])
                key = (s, t) if s < t else (t, s)
                if key in new_edges:
                    new_edges[key][0] += 1
                else:
                    new_edges[key] = [1, (s, t)]
                    new_order.append(key)
            for x in vid:
                x = int(x)
                if int(self.H[x]) not in terminal_h and not self.alive[x]:
                    self.alive[x] = True        # add_vertices: new (or previously deleted) vertex
                    self.nbr[x] = -1
        # edges already in the graph are skipped (either orientation)
        fresh = [key for key in new_order if not self._has_edge(*key)]
        for key in fresh:
            for x in key:
                if not self.alive[x]:
                    raise RuntimeError("internal error: edge to a vertex that is not in the graph")
        inc_new = defaultdict(list)
        for seq, key in enumerate(fresh):
            inc_new[key[0]].append(key); inc_new[key[1]].append(key)
            self._edge_birth[key] = (round_no, seq)
        wt = {key: new_edges[key][0] for key in fresh}

        def old_nbrs(x):
            return [int(y) for y in self.nbr[x] if y >= 0]

        # incident-weight guard (check_added_edges_incident_weights :70-80)
        def incident_weight(x):
            return G * len(old_nbrs(x)) + sum(wt[e] for e in inc_new[x])
        flagged = [key for key in fresh if incident_weight(key[0]) > 2 * G or incident_weight(key[1]) > 2 * G]
        flagged_set = set(flagged)
        # --- simplification runs on the graph WITH the flagged edges; its weight bumps survive only
        #     when nothing was flagged (same object), its vertex deletions never do
        if self.simplify:
            bumps = self._simplify_extended(fresh, wt, inc_new, old_nbrs)
            if not flagged:
                for key in bumps:
                    if key in wt:
                        wt[key] = G
        # --- weight filter (+ flagged pairs on the last round)
        surviving = [key for key in fresh if key not in flagged_set]
        low = [key for key in surviving if wt[key] < G]
        for key in surviving:
            if wt[key] >= G:
                self._add_edge(*key)
        if last_round and low:
            self._erode(low)
        return None

    def _simplify_extended(self, fresh, wt, inc_new, old_nbrs):
        "run_graph_simplification on the extended graph; returns the edges whose weight it sets to G"
        G = self.G
        touched = set(inc_new.keys())

        def nbrs(x):
            res = {y: G for y in old_nbrs(x)}
            for e in inc_new.get(x, ()):
                y = e[1] if e[0] == x else e[0]
                res[y] = wt[e]
            return res

        cand_v = {x for x in touched if len(nbrs(x)) == 3}
        if not cand_v:
            return []
        bumped = {}

        def weight(u, x):
            key = (u, x) if u < x else (x, u)
            return bumped.get(key, nbrs(u)[x])

        def anchored(u):
            return sum(1 for x in nbrs(u) if weight(u, x) == G) == 1

        # candidate edges in edge-id order: old edges first (their relative order only matters among
        # themselves), then the new edges in insertion order
        old_c, new_c = [], []
        fresh_pos = {key: i for i, key in enumerate(fresh)}
        for u in cand_v:
            for x in nbrs(u):
                if x in cand_v:
                    key = (u, x) if u < x else (x, u)
                    if key in fresh_pos:
                        new_c.append(key)
                    else:
                        old_c.append(key)
        old_c = sorted(set(old_c), key=self._old_edge_key)
        new_c = sorted(set(new_c), key=lambda e: fresh_pos[e])
        out = []
        for s, t in old_c + new_c:
            if anchored(s) and anchored(t):
                common = [x for x in nbrs(s) if x != t and x in nbrs(t)]
                if len(common) == 1:
                    bumped[(s, t)] = G
                    out.append((s, t))
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

        def overlap(s, t):
            return bool((np.abs(self.POS[:, s] - self.POS[:, t]) < self.k).any())

        deg = lambda x: int((self.nbr[x] >= 0).sum())   # noqa: E731
        for (u, v) in low:
            # igraph reports (source, target) = (min id, max id); the reference then orders by NAME string
            s, t = (u, v)
            if name(s) > name(t):
                s, t = t, s
            if deg(s) != 1 or deg(t) != 1:
                continue
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
        G = self.G
        if len(self.w_rounds) != len(set(self.w_rounds)):
            print("Error: duplicate values found in w_rounds!", file=sys.stderr, flush=True)
            raise SystemExit(1)
        self.log("Sketching and joining minimizers, w =", self.w)
        tables = [self.be.sketch(a, self.w, None) for a in range(G)]
        H, POS, CTG, RANK, link, degree = self.be.join(tables, self.orient)
        self.stats["vertices"] = int(len(H))
        self._init_vertices(H, POS, CTG, RANK)
        self._ctg0 = self.CTG[:, :self.V0].copy()
        self._edge_birth = {}
        V = self.V0
        # --- simplification, weight filter
        bumped, removed = ({}, [])
        if self.simplify:
            self.log("Running graph simplificaton")
            bumped, removed = self._simplify_round0(np.asarray(link), np.asarray(degree))
        self.stats["simplified_vertices"] = len(set(removed))
        self.log("Filtering the graph")
        lk = np.asarray(link[:max(V - 1, 0)], dtype=bool) if V else np.zeros(0, dtype=bool)
        ids = np.arange(max(V - 1, 0), dtype=np.int64)
        self.nbr[ids[lk], 1] = ids[lk] + 1          # slot 1: right neighbour
        self.nbr[ids[lk] + 1, 0] = ids[lk]          # slot 0: left neighbour
        if removed:
            self._remove_vertices(np.array(removed, dtype=np.int64))
        for (s, t) in bumped:
            if self.alive[s] and self.alive[t] and not self._has_edge(s, t):
                self._add_edge(s, t)
        # --- paths, blocks
        self.log("Finding paths")
        paths = self._find_paths()
        self.stats["paths"] = len(paths)
        self.log("Finding synteny blocks")
        blocks = self._blocks_from_paths(paths)
        blocks = self._split_indels(blocks)
        blocks = self._filter_small(blocks, 4)
        ordered = self._sort_blocks(blocks)
        if not ordered:
            print("Error - no paths found. Try adjusting the specified k/w parameters.")
            raise SystemExit(1)
        self._emit("initial", ordered)
        self.log("Done initial synteny blocks")
        # --- refinement rounds (refine_block_coordinates :476-530)
        prev_w = self.w
        for ri, new_w in enumerate(self.w_rounds):
            self.log("Extending synteny blocks with w =", new_w)
            last = new_w == self.w_rounds[-1]
            self._refine_round(blocks, new_w, prev_w, last, ri + 1)
            paths = self._find_paths()
            blocks = self._blocks_from_paths(paths)
            blocks = self._split_indels(blocks)
            blocks = self._filter_small(blocks, 4)
            ordered = self._sort_blocks(blocks)
            self._emit("pre_merge", ordered)
            if last:
                merged = self._merge_collinear(ordered) if ordered else []
                merged = [b for b in merged if self._long_enough(b)]
                merged = self._merge_collinear(merged) if merged else []
                self._emit("final", merged, verbose=True)
            prev_w = new_w
        self.log("Done extended synteny blocks")
        return self.outputs.get("final", self.outputs.get("initial"))
