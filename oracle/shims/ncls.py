"""Stand-in for the `ncls` package (TEST INFRASTRUCTURE ONLY): half-open interval
overlap test as used by bin/ntsynt_synteny.py:224-225,271,275."""
import bisect


class NCLS:
    def __init__(self, starts, ends, ids=None):
        order = sorted(range(len(starts)), key=lambda i: starts[i])
        self._starts = [int(starts[i]) for i in order]
        self._maxend = []
        m = None
        for i in order:
            e = int(ends[i])
            m = e if m is None or e > m else m
            self._maxend.append(m)

    def has_overlap(self, start, end):
        "True iff some [s, e) has s < end and start < e"
        n = bisect.bisect_left(self._starts, end)   # intervals with s < end
        return n > 0 and self._maxend[n - 1] > start
