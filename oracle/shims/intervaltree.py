"""Stand-in for the `intervaltree` package (TEST INFRASTRUCTURE ONLY), covering
bin/ntsynt_synteny.py:228-253,375-380."""
from collections import namedtuple


class Interval(namedtuple("Interval", ["begin", "end", "data"])):
    __slots__ = ()

    def __new__(cls, begin, end, data=None):
        return super().__new__(cls, begin, end, data)


class IntervalTree:
    def __init__(self):
        self._ivs = set()

    def __setitem__(self, key, data):
        if not isinstance(key, slice):
            raise TypeError("use tree[a:b] = data")
        if key.start >= key.stop:
            raise ValueError("IntervalTree: Null Interval objects not allowed in IntervalTree")
        self._ivs.add(Interval(key.start, key.stop, data))

    def __getitem__(self, key):
        if isinstance(key, slice):
            return {iv for iv in self._ivs if iv.begin < key.stop and key.start < iv.end}
        return {iv for iv in self._ivs if iv.begin <= key < iv.end}

    def slice(self, point):
        hit = [iv for iv in self._ivs if iv.begin < point < iv.end]
        for iv in hit:
            self._ivs.remove(iv)
            self._ivs.add(Interval(iv.begin, point, iv.data))
            self._ivs.add(Interval(point, iv.end, iv.data))

    def __iter__(self):
        return iter(self._ivs)

    def __len__(self):
        return len(self._ivs)

    def __bool__(self):
        return bool(self._ivs)
