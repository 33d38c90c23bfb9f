"""Minimal pure-Python stand-in for the slice of python-igraph that ntSynt's graph
stage touches.  TEST INFRASTRUCTURE ONLY: it lets the
reference's own bin/ntsynt_run.py run unmodified in this container (python-igraph is
not installed) so that golden fixtures can be generated, and it backs
oracle/graph_oracle.py.  Semantics kept: vertices / edges keep insertion order across
deletions (ids are compacted), an undirected edge reports source=min(id), target=max(id),
Edge objects hash/compare by (graph, index), new edges carry None attributes.
"""
from collections import deque


class Vertex:
    __slots__ = ("graph", "index")

    def __init__(self, graph, index):
        self.graph, self.index = graph, index

    def __getitem__(self, key):
        if key != "name":
            raise KeyError(key)
        return self.graph._names[self.index]

    def degree(self):
        return len(self.graph._inc()[self.index])

    def incident(self):
        return [Edge(self.graph, e) for e in self.graph._inc()[self.index]]

    def neighbors(self):
        g = self.graph
        return [Vertex(g, g._other(e, self.index)) for e in g._inc()[self.index]]

    def __hash__(self):
        return hash((id(self.graph), self.index))

    def __eq__(self, other):
        return isinstance(other, Vertex) and self.graph is other.graph and self.index == other.index


class Edge:
    __slots__ = ("graph", "index")

    def __init__(self, graph, index):
        self.graph, self.index = graph, index

    @property
    def source(self):
        return self.graph._src[self.index]

    @property
    def target(self):
        return self.graph._tgt[self.index]

    @property
    def tuple(self):
        return (self.source, self.target)

    def __getitem__(self, key):
        return self.graph._eattr[key][self.index]

    def __setitem__(self, key, value):
        g = self.graph
        if key not in g._eattr:
            g._eattr[key] = [None] * len(g._src)
        g._eattr[key][self.index] = value

    def __hash__(self):
        return hash((id(self.graph), self.index))

    def __eq__(self, other):
        return isinstance(other, Edge) and self.graph is other.graph and self.index == other.index


class VertexSeq:
    def __init__(self, graph, ids=None):
        self.graph, self.ids = graph, ids

    def _all(self):
        return range(len(self.graph._names)) if self.ids is None else self.ids

    def __call__(self):
        return self

    def __iter__(self):
        g = self.graph
        return (Vertex(g, i) for i in self._all())

    def __len__(self):
        return len(self._all())

    def __getitem__(self, key):
        if isinstance(key, str):
            if key != "name":
                raise KeyError(key)
            return [self.graph._names[i] for i in self._all()]
        if isinstance(key, int):
            return Vertex(self.graph, self._all()[key])
        return VertexSeq(self.graph, [self._all()[i] for i in key])

    def find(self, name=None, **kwargs):
        if name is None:
            name = kwargs["name"]
        if isinstance(name, int):
            return Vertex(self.graph, name)
        try:
            return Vertex(self.graph, self.graph._name_index()[name])
        except KeyError:
            raise ValueError(f"no such vertex: {name!r}") from None


class EdgeSeq:
    def __init__(self, graph, ids=None):
        self.graph, self.ids = graph, ids

    def _all(self):
        return range(len(self.graph._src)) if self.ids is None else self.ids

    def __call__(self):
        return self

    def __iter__(self):
        g = self.graph
        return (Edge(g, i) for i in self._all())

    def __len__(self):
        return len(self._all())

    def __getitem__(self, key):
        if isinstance(key, str):
            col = self.graph._eattr[key]
            return [col[i] for i in self._all()]
        if isinstance(key, int):
            return Edge(self.graph, self._all()[key])
        return EdgeSeq(self.graph, [self._all()[i] for i in key])

    def __setitem__(self, key, values):
        g = self.graph
        ids = list(self._all())
        if not isinstance(values, (list, tuple)):
            values = [values] * len(ids)
        if len(values) != len(ids):
            raise ValueError("attribute list length must match the number of edges")
        if key not in g._eattr:
            g._eattr[key] = [None] * len(g._src)
        col = g._eattr[key]
        for i, v in zip(ids, values):
            col[i] = v


class VertexClustering:
    def __init__(self, clusters):
        self._clusters = clusters

    def __iter__(self):
        return iter(self._clusters)

    def __len__(self):
        return len(self._clusters)

    def __getitem__(self, i):
        return self._clusters[i]


class Graph:
    def __init__(self):
        self._names = []
        self._src, self._tgt = [], []
        self._eattr = {}
        self._inc_cache = None
        self._name_cache = None

    # -- internals
    def _dirty(self):
        self._inc_cache = None
        self._name_cache = None

    def _inc(self):
        if self._inc_cache is None:
            inc = [[] for _ in self._names]
            for e, (s, t) in enumerate(zip(self._src, self._tgt)):
                inc[s].append(e)
                if t != s:
                    inc[t].append(e)
            self._inc_cache = inc
        return self._inc_cache

    def _name_index(self):
        if self._name_cache is None:
            self._name_cache = {}
            for i, n in enumerate(self._names):
                self._name_cache.setdefault(n, i)
        return self._name_cache

    def _other(self, e, v):
        return self._tgt[e] if self._src[e] == v else self._src[e]

    def _vid(self, v):
        if isinstance(v, Vertex):
            return v.index
        if isinstance(v, str):
            try:
                return self._name_index()[v]
            except KeyError:
                raise ValueError(f"no such vertex: {v!r}") from None
        return int(v)

    # -- public surface
    @property
    def vs(self):
        return VertexSeq(self)

    @property
    def es(self):
        return EdgeSeq(self)

    def vcount(self):
        return len(self._names)

    def ecount(self):
        return len(self._src)

    def copy(self):
        g = Graph()
        g._names = list(self._names)
        g._src, g._tgt = list(self._src), list(self._tgt)
        g._eattr = {k: list(v) for k, v in self._eattr.items()}
        return g

    def add_vertices(self, names):
        if isinstance(names, int):
            names = [None] * names
        self._names.extend(names)
        self._dirty()

    def add_edges(self, pairs):
        for s, t in pairs:
            a, b = self._vid(s), self._vid(t)
            self._src.append(min(a, b))
            self._tgt.append(max(a, b))
        for col in self._eattr.values():
            col.extend([None] * (len(self._src) - len(col)))
        self._inc_cache = None

    def get_eid(self, v1, v2):
        a, b = self._vid(v1), self._vid(v2)
        for e in self._inc()[a]:
            if self._other(e, a) == b:
                return e
        raise ValueError("no such edge")

    def incident(self, v):
        return list(self._inc()[self._vid(v)])

    def degree(self, v):
        return len(self._inc()[self._vid(v)])

    def neighbors(self, v):
        v = self._vid(v)
        return [self._other(e, v) for e in self._inc()[v]]

    def delete_edges(self, edges):
        drop = {e.index if isinstance(e, Edge) else int(e) for e in edges}
        if not drop:
            return
        keep = [i for i in range(len(self._src)) if i not in drop]
        self._src = [self._src[i] for i in keep]
        self._tgt = [self._tgt[i] for i in keep]
        self._eattr = {k: [v[i] for i in keep] for k, v in self._eattr.items()}
        self._inc_cache = None

    def delete_vertices(self, vertices):
        drop = {self._vid(v) for v in vertices}
        if not drop:
            return
        remap, names = {}, []
        for i, n in enumerate(self._names):
            if i not in drop:
                remap[i] = len(names)
                names.append(n)
        keep = [i for i in range(len(self._src)) if self._src[i] not in drop and self._tgt[i] not in drop]
        src = [remap[self._src[i]] for i in keep]
        tgt = [remap[self._tgt[i]] for i in keep]
        self._names = names
        self._src = [min(a, b) for a, b in zip(src, tgt)]
        self._tgt = [max(a, b) for a, b in zip(src, tgt)]
        self._eattr = {k: [v[i] for i in keep] for k, v in self._eattr.items()}
        self._dirty()

    def components(self):
        inc = self._inc()
        seen = [False] * len(self._names)
        clusters = []
        for v0 in range(len(self._names)):
            if seen[v0]:
                continue
            seen[v0] = True
            comp, dq = [v0], deque([v0])
            while dq:
                v = dq.popleft()
                for e in inc[v]:
                    u = self._other(e, v)
                    if not seen[u]:
                        seen[u] = True
                        comp.append(u)
                        dq.append(u)
            comp.sort()
            clusters.append(comp)
        return VertexClustering(clusters)

    connected_components = components

    def subgraph(self, vertices):
        ids = sorted({self._vid(v) for v in vertices})
        remap = {v: i for i, v in enumerate(ids)}
        g = Graph()
        g._names = [self._names[v] for v in ids]
        eids = sorted({e for v in ids for e in self._inc()[v]
                       if self._src[e] in remap and self._tgt[e] in remap})
        g._src = [remap[self._src[e]] for e in eids]
        g._tgt = [remap[self._tgt[e]] for e in eids]
        g._eattr = {k: [col[e] for e in eids] for k, col in self._eattr.items()}
        return g

    induced_subgraph = subgraph

    def get_all_simple_paths(self, v, to=None, cutoff=-1, mode="all"):
        s = self._vid(v)
        t = None if to is None else self._vid(to)
        inc = self._inc()
        out = []

        def rec(path):
            last = path[-1]
            if cutoff >= 0 and len(path) - 1 >= cutoff:
                return
            for e in inc[last]:
                u = self._other(e, last)
                if u in path:
                    continue
                new = path + [u]
                if t is None or u == t:
                    out.append(new)
                if u != t:
                    rec(new)

        rec([s])
        return out

    def get_shortest_paths(self, v, to=None, **_kw):
        s = self._vid(v)
        targets = [self._vid(to)] if not isinstance(to, (list, tuple)) else [self._vid(x) for x in to]
        inc = self._inc()
        prev = {s: None}
        dq = deque([s])
        while dq:
            x = dq.popleft()
            for e in inc[x]:
                u = self._other(e, x)
                if u not in prev:
                    prev[u] = x
                    dq.append(u)
        res = []
        for t in targets:
            if t not in prev:
                res.append([])
                continue
            p = []
            while t is not None:
                p.append(t)
                t = prev[t]
            res.append(p[::-1])
        return res
