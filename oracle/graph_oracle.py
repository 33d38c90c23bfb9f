"""CPU restatement of ntSynt's graph stage -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

A plain, dictionary-and-string restatement of what bin/ntsynt_run.py does, in memory (no files,
no subprocesses), used (a) as the checker for the CUDA path on the GPU box, where /root/reference
does not exist, and (b) as the graph leg of bench.py's cpu_baseline.  It is deliberately the
"slow, obvious" formulation: minimizers are decimal strings, the graph is an igraph-like object
(oracle/shims/igraph.py), every step is sequential.  Pinned in the build container against the
reference's own code run under shims (oracle/ref_harness.py) and against the golden block files
(tests/test_graph_oracle.py).

Each function names the reference lines it follows (paths relative to /root/reference).
"""
import os
import re
import sys
from collections import defaultdict

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "shims"))
import igraph as ig  # noqa: E402  (the stand-in under oracle/shims)
import ncls  # noqa: E402

from . import sketch_oracle as so  # noqa: E402

FA_RE = re.compile(r"^(\S+)\.k\d+\.w\d+.tsv")


# ---------------------------------------------------------------- minimizer lists
def sketch_lists(records, k, w, common):
    "indexlr output of one genome as [(contig, [(mx_string, pos), ...])] (SURVEY A.4)"
    out = []
    for name, seq in records:
        h1, pos = so.minimize(seq, k, w, common)
        out.append((name, [(str(int(h)), int(p)) for h, p in zip(h1, pos)]))
    return out


def read_minimizers(lines):
    "subprojects/ntJoin/bin/ntjoin_utils.py:167-193"
    info, lists, dups = {}, [], set()
    for contig, toks in lines:
        if not toks:
            continue
        lists.append([mx for mx, _ in toks])
        for mx, pos in toks:
            if mx in info:
                dups.add(mx)
            else:
                info[mx] = (contig, pos)
    info = {mx: v for mx, v in info.items() if mx not in dups}
    return info, [[mx for mx in lst if mx not in dups] for lst in lists]


def filter_minimizers(list_mxs):
    "ntjoin_utils.py:152-165"
    sets = [{mx for lst in list_mxs[a] for mx in lst} for a in list_mxs]
    common = set.intersection(*sets)
    return {a: [[mx for mx in lst if mx in common] for lst in list_mxs[a]] for a in list_mxs}


# ---------------------------------------------------------------- graph construction
def _incident_weight(graph, v):
    return sum(graph.es[e]["weight"] for e in graph.incident(v))


def build_graph(list_mxs, weights, graph=None, black_list=None):
    "ntjoin_utils.py:83-141 (incl. check_added_edges_incident_weights :70-80)"
    if graph is None:
        graph, prev = ig.Graph(), {}
    else:
        prev = {e.index: {"support": e["support"], "weight": e["weight"]} for e in graph.es}
    vertices, edges = set(), defaultdict(dict)
    for asm in list_mxs:
        for lst in list_mxs[asm]:
            for a, b in zip(lst, lst[1:]):
                if a in edges and b in edges[a]:
                    edges[a][b].append(asm)
                elif b in edges and a in edges[b]:
                    edges[b][a].append(asm)
                else:
                    edges[a][b] = [asm]
                if black_list is None or a not in black_list:
                    vertices.add(a)
            if lst and (black_list is None or lst[-1] not in black_list):
                vertices.add(lst[-1])
    pairs = [(s, t) for s in edges for t in edges[s]]
    if prev:
        have = set(graph.vs["name"])
        vertices = {v for v in vertices if v not in have}
    graph.add_vertices(sorted(vertices))
    if prev:
        have_e = {(graph.vs[e.source]["name"], graph.vs[e.target]["name"]) for e in graph.es}
        pairs = [(s, t) for s, t in pairs if (s, t) not in have_e and (t, s) not in have_e]
    graph.add_edges(pairs)
    attrs = {graph.get_eid(s, t): {"support": edges[s][t], "weight": sum(weights[f] for f in edges[s][t])}
             for s, t in pairs}
    attrs.update(prev)
    graph.es["support"] = [attrs[e]["support"] for e in sorted(attrs)]
    graph.es["weight"] = [attrs[e]["weight"] for e in sorted(attrs)]
    if prev:
        limit = sum(weights.values()) * 2
        bad = [graph.get_eid(s, t) for s, t in pairs
               if _incident_weight(graph, s) > limit or _incident_weight(graph, t) > limit]
        if bad:
            graph = graph.copy()
            graph.delete_edges(bad)
    return graph


class GraphOracle:
    """In-memory ntSynt graph stage.  genomes: list of (tsv_name, records) with
    records = [(contig, seq_bytes)]; order as given by the user (re-sorted like the reference)."""

    def __init__(self, genomes, k, w, w_rounds, bp, collinear_merge, z, common, m=90, simplify=True,
                 restart_on_gap=False, n=0):
        self.k, self.w, self.w_rounds, self.bp, self.z, self.m = k, w, list(w_rounds), bp, z, m
        cm = str(collinear_merge)
        self.collinear_merge = int(cm[:-1]) * w if cm.endswith("w") else int(cm)
        self.common = common
        self.simplify = simplify
        self.genomes = dict(genomes)
        self.files = sorted(self.genomes, reverse=True)          # bin/ntsynt_synteny.py:34
        self.weights = {f: 1 for f in self.files}
        self.n = n or len(self.files)                            # -n, minimum edge weight (bin/ntsynt_synteny.py:46-47)
        self.info, self.lists = {}, {}
        self.graph = None
        self.outputs = {}

    # ------------------------------------------------------------ blocks
    class Blk:
        def __init__(self, asms):
            self.ctg = {a: None for a in asms}
            self.mxs = {a: [] for a in asms}     # [(mx, pos)]
            self.ori = {a: None for a in asms}
            self.reason = None

        def start(self, a):
            return min(self.mxs[a][0][1], self.mxs[a][-1][1])

        def end(self, a, k):
            return max(self.mxs[a][0][1], self.mxs[a][-1][1]) + k

        def count(self):
            return len(next(iter(self.mxs.values())))

    def _orient(self, blk):
        "bin/synteny_block.py:48-65"
        for a, lst in blk.mxs.items():
            p = [x[1] for x in lst]
            if all(x < y for x, y in zip(p, p[1:])):
                blk.ori[a] = "+"
            elif all(x > y for x, y in zip(p, p[1:])):
                blk.ori[a] = "-"
            else:
                up = [x < y for x, y in zip(p, p[1:])].count(True) / float(len(p) - 1) * 100
                blk.ori[a] = "+" if up >= self.m else ("-" if 100 - up >= self.m else "?")
        return all(o in "+-" for o in blk.ori.values())

    def find_synteny_blocks(self, path):
        "bin/ntsynt_synteny.py:66-106 (past_start_flag is never set: only the last run survives)"
        out, drop = [], []
        cur = self.Blk(self.info.keys())
        for mx in path:
            if all(self.info[a][mx][0] == cur.ctg[a] for a in self.info):
                for a in self.info:
                    cur.mxs[a].append((mx, self.info[a][mx][1]))
            else:
                cur = self.Blk(self.info.keys())
                for a in self.info:
                    cur.ctg[a], pos = self.info[a][mx]
                    cur.mxs[a].append((mx, pos))
        if self._orient(cur):
            out.append(cur)
        else:
            drop.extend(self.graph.vs.find(mx).index for mx, _ in next(iter(cur.mxs.values())))
        if drop:
            g = self.graph.copy()
            g.delete_vertices(drop)
            self.graph = g
        return out

    def check_for_indels(self, blocks):
        "bin/ntsynt_synteny.py:364-409"
        out, rm = [], []
        for b in blocks:
            cuts = []
            asms = sorted(b.mxs)
            for i in range(b.count() - 1):
                d = [abs(b.mxs[a][i][1] - b.mxs[a][i + 1][1]) for a in asms]
                if max(d) - min(d) > self.bp:
                    cuts.append(i + 1)
                    rm.append(self.graph.get_eid(b.mxs[asms[0]][i][0], b.mxs[asms[0]][i + 1][0]))
            if not cuts:
                out.append(b)
                continue
            bounds = [0] + cuts + [b.count()]
            for s, e in zip(bounds, bounds[1:]):
                nb = self.Blk(b.mxs.keys())
                for a in b.mxs:
                    nb.ctg[a], nb.ori[a], nb.mxs[a] = b.ctg[a], b.ori[a], b.mxs[a][s:e]
                out.append(nb)
        g = self.graph.copy()
        g.delete_edges(rm)
        self.graph = g
        return out

    def filter_blocks(self, blocks, min_mx):
        "bin/ntsynt_synteny.py:411-426"
        keep, drop = [], []
        for b in blocks:
            if b.count() >= min_mx:
                keep.append(b)
            else:
                drop.extend(self.graph.vs.find(mx).index for mx, _ in next(iter(b.mxs.values())))
        g = self.graph.copy()
        g.delete_vertices(drop)
        self.graph = g
        return keep

    def _sorted(self, blocks):
        "bin/synteny_block.py:102-109"
        a = sorted(self.files)[0]
        return sorted(blocks, key=lambda b: (b.ctg[a], b.start(a)))

    def _long(self, b):
        return all(b.end(a, self.k) - b.start(a) >= self.z for a in b.mxs)

    def _text(self, blocks, verbose=False):
        "bin/synteny_block.py:72-85"
        rows, num = [], 0
        for b in blocks:
            if not self._long(b):
                continue
            for a in sorted(b.mxs):
                label = m.group(1) if (m := re.search(FA_RE, a)) else a
                row = f"{num}\t{label}\t{b.ctg[a]}\t{b.start(a)}\t{b.end(a, self.k)}\t{b.ori[a]}\t{len(b.mxs[a])}"
                if verbose:
                    row += f"\t{b.reason}"
                rows.append(row + "\n")
            num += 1
        return "".join(rows)

    # ------------------------------------------------------------ graph steps
    def simplify_graph(self, graph):
        "bin/ntsynt_synteny.py:548-590"
        top = sum(self.weights.values())

        def anchored(v):
            return [graph.es[e]["weight"] for e in graph.incident(v)].count(top) == 1

        drop = []
        for e in graph.es:
            s, t = e.source, e.target
            if graph.degree(s) == 3 and graph.degree(t) == 3 and anchored(s) and anchored(t):
                paths = graph.get_all_simple_paths(s, t, cutoff=2)
                if len(paths) == 2:
                    for p in paths:
                        if len(p) == 3:
                            drop.append(p[1])
                            e["weight"] = top
        g = graph.copy()
        g.delete_vertices(drop)
        return g

    def weight_filter(self, graph, flag=False):
        "subprojects/ntJoin/bin/ntjoin.py:78-87; bin/ntsynt_synteny.py:292-303"
        if not flag and self.n <= min(self.weights.values()):
            return graph
        low = [e.index for e in graph.es if e["weight"] < self.n]
        pairs = [(graph.es[i].source, graph.es[i].target) for i in low]
        g = graph.copy()
        g.delete_edges(low)
        return (g, pairs) if flag else g

    def find_paths(self):
        """subprojects/ntJoin/bin/ntjoin.py:68-76,89-151.  With n = number of assemblies every component is linear already;
        below that, edges lighter than a rising threshold are taken off the branch vertices of a component until it is"""
        ref = self.files[-1]       # .pop() of the equally weighted assemblies
        out = []
        for comp in self.graph.components():
            whole = self.graph.subgraph(comp)
            floor, top = self.n, sum(self.weights.values())
            while any(v.degree() > 2 for v in whole.vs) and floor <= top:
                light = [e for v in whole.vs if v.degree() > 2 for e in whole.incident(v.index) if whole.es[e]["weight"] < floor]
                g = whole.copy()
                g.delete_edges(light)
                whole = g
                floor += 1
            for part in whole.components():
                sub = whole.subgraph(part)
                ends = [v.index for v in sub.vs if v.degree() == 1]
                if len(ends) != 2:
                    continue
                pos = [self.info[ref][sub.vs[v]["name"]][1] for v in ends]
                src = [v for v, p in zip(ends, pos) if p == min(pos)].pop()
                dst = [v for v, p in zip(ends, pos) if p == max(pos)].pop()
                path = sub.get_shortest_paths(src, dst)[0]
                if len(path) == sub.vcount() and len(path) - 1 == sub.ecount() and len(set(path)) == len(path):
                    out.append([sub.vs[v]["name"] for v in path])
        return out

    def _blocks(self):
        blocks = [b for p in self.find_paths() for b in self.find_synteny_blocks(p)]
        blocks = self.check_for_indels(blocks)
        return self.filter_blocks(blocks, 4)

    # ------------------------------------------------------------ refinement
    def _masked_records(self, asm, blocks, w):
        "bin/ntsynt_synteny.py:117-157; empty/inverted intervals after the negative slop are dropped"
        per = defaultdict(list)
        for b in blocks:
            s, e = b.start(asm), b.end(asm, self.k)
            if e - s > max(2 * w, w + self.k + 1):
                per[b.ctg[asm]].append((s, e))
        out = []
        for name, seq in self.genomes[asm]:
            buf = None
            for s, e in per.get(name, ()):
                s2, e2 = max(s + w + self.k, 0), min(e - (w + self.k), len(seq))
                if s2 < e2:
                    if buf is None:
                        buf = bytearray(seq)
                    buf[s2:e2] = b"N" * (e2 - s2)
            out.append((name, bytes(buf) if buf is not None else seq))
        return out

    def refine(self, blocks, new_w, prev_w, last):
        "bin/ntsynt_synteny.py:476-541"
        new_info, new_lists = {}, {}
        for asm in self.files:
            recs = self._masked_records(asm, blocks, prev_w)
            new_info[asm], new_lists[asm] = read_minimizers(sketch_lists(recs, self.k, new_w, self.common))
        terminal, internal, ivs = set(), set(), defaultdict(dict)
        for b in blocks:                                                    # :205-226
            for a in b.mxs:
                first, last_mx = b.mxs[a][0], b.mxs[a][-1]
                terminal.update((first[0], last_mx[0]))
                lo, hi = min(first[1], last_mx[1]), max(first[1], last_mx[1])
                if hi - lo >= 2:
                    ivs[a].setdefault(b.ctg[a], []).append((lo + 1, hi))
                internal.update(mx for mx, _ in b.mxs[a][1:-1])
        trees = {a: {c: ncls.NCLS(*zip(*[(s, e, 1) for s, e in v])) for c, v in ivs[a].items()} for a in ivs}
        filt = {}
        for a in new_lists:                                                 # :256-280
            res = []
            for lst in new_lists[a]:
                cur = []
                for mx in lst:
                    ctg, pos = new_info[a][mx]
                    tree = trees.get(a, {}).get(ctg)
                    if cur and tree is not None:
                        prev = new_info[a][cur[-1]][1]
                        if tree.has_overlap(min(prev, pos), max(prev, pos)):
                            res.append(cur)
                            cur = []
                    if mx not in internal and (tree is None or not tree.has_overlap(pos, pos + 1)):
                        cur.append(mx)
                res.append(cur)
            filt[a] = res
        filt = filter_minimizers(filt)
        valid = {mx for a in filt for lst in filt[a] for mx in lst}         # :282-290
        for a, d in new_info.items():
            for mx in d:
                if mx in valid:
                    self.info[a][mx] = d[mx]
        graph = build_graph(filt, self.weights, graph=self.graph, black_list=terminal)
        if self.simplify:
            self.graph = self.simplify_graph(self.graph)                    # result overwritten below (quirk)
        if last:
            self.graph, pairs = self.weight_filter(graph, flag=True)
            self._erode(pairs)
        else:
            self.graph = self.weight_filter(graph)
        return self._blocks()

    def _erode(self, pairs):
        "bin/ntsynt_synteny.py:305-362"
        g = self.graph
        name = lambda v: g.vs[v]["name"]          # noqa: E731

        def close(a, b):
            return any(abs(d[a][1] - d[b][1]) < self.k for d in self.info.values())

        kill = set()
        for s, t in pairs:
            if name(s) > name(t):
                s, t = t, s
            if g.degree(s) != 1 or g.degree(t) != 1:
                continue
            cs, ct, at_target, seen = s, t, True, {s, t}
            while close(name(cs), name(ct)):
                v = ct if at_target else cs
                kill.update(g.incident(v))
                nxt = [u for u in g.neighbors(v) if u not in seen]
                if not nxt:
                    break
                assert len(nxt) == 1
                if at_target:
                    ct = nxt[0]
                else:
                    cs = nxt[0]
                seen.add(nxt[0])
                at_target = not at_target
        if kill:
            g2 = g.copy()
            g2.delete_edges(sorted(kill))
            self.graph = g2

    def merge_collinear(self, blocks):
        "bin/ntsynt_synteny.py:428-472"
        def gap(b1, b2, a):
            if b1.ori[a] == "-" and b2.ori[a] == "-":
                return b1.start(a) - b2.end(a, self.k)
            return b2.start(a) - b1.end(a, self.k)
        out, cur = [], blocks[0]
        for b in blocks[1:]:
            same_o = all(cur.ori[a] == b.ori[a] for a in cur.mxs)
            same_c = all(cur.ctg[a] == b.ctg[a] for a in cur.mxs)
            d = [gap(cur, b, a) for a in cur.mxs]
            if not same_o or not same_c or max(d) - min(d) > self.bp - self.k or max(d) >= self.collinear_merge:
                if not same_c:
                    b.reason = "id_change"
                elif not same_o:
                    b.reason = "ori_change"
                elif any(x < 0 for x in d):
                    b.reason = "inconsistent_order"
                elif max(d) - min(d) > self.bp - self.k:
                    b.reason = "indel"
                else:
                    b.reason = "merge"
                out.append(cur)
                cur = b
            else:
                for a in b.mxs:
                    cur.mxs[a].extend(b.mxs[a])
        out.append(cur)
        return out

    # ------------------------------------------------------------ driver
    def run(self, round0=None):
        """bin/ntsynt_synteny.py:593-647.  round0: optional {tsv_name: sketch lines} made elsewhere (the
        reference reads the round-0 sketches from the indexlr TSVs)"""
        for f in self.files:
            lines = round0[f] if round0 is not None else sketch_lists(self.genomes[f], self.k, self.w, self.common)
            self.info[f], self.lists[f] = read_minimizers(lines)
        self.graph = build_graph(filter_minimizers(self.lists), self.weights)
        self.round0_edges = [(self.graph.vs[e.source]["name"], self.graph.vs[e.target]["name"], e["weight"])
                             for e in self.graph.es]
        if self.simplify:
            self.graph = self.simplify_graph(self.graph)
        self.graph = self.weight_filter(self.graph)
        blocks = self._blocks()
        ordered = self._sorted(blocks)
        if not ordered:
            raise SystemExit("Error - no paths found. Try adjusting the specified k/w parameters.")
        self.outputs["initial"] = self._text(ordered)
        prev_w = self.w
        for new_w in self.w_rounds:
            last = new_w == self.w_rounds[-1]
            blocks = self.refine(blocks, new_w, prev_w, last)
            ordered = self._sorted(blocks)
            self.outputs["pre_merge"] = self._text(ordered)
            if last:
                merged = self.merge_collinear(ordered)
                merged = [b for b in merged if self._long(b)]
                merged = self.merge_collinear(merged)
                self.outputs["final"] = self._text(merged, verbose=True)
            prev_w = new_w
        return self.outputs.get("final", self.outputs["initial"])
