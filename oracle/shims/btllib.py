"""Stand-in for btllib's Python module (TEST INFRASTRUCTURE ONLY).
bin/ntsynt_synteny.py:17 imports it; it is used only with --filter Filter (:605-607):
KmerBloomFilter(path).contains(seq).  The file is the harness's raw oracle filter
(b"ORCBF1\\n" + uint64 nbytes + bytes); membership = bit (canonical ntHash2 h0 of the k-mer) mod m."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
import numpy as np  # noqa: E402
from oracle import sketch_oracle as so  # noqa: E402


class KmerBloomFilter:
    def __init__(self, path, *args, **kwargs):
        if not isinstance(path, str) or args or kwargs:
            raise NotImplementedError("only KmerBloomFilter(path) is provided by the oracle shim")
        with open(path, "rb") as fh:
            assert fh.readline() == b"ORCBF1\n"
            n = int(np.frombuffer(fh.read(8), dtype=np.uint64)[0])
            self.bits = np.frombuffer(fh.read(n), dtype=np.uint8).copy()
        self.m = self.bits.size * 8

    def contains(self, seq):
        h0 = so.kmer_hash(seq.encode() if isinstance(seq, str) else bytes(seq))
        if h0 is None:
            return False
        i = h0 % self.m
        return bool((int(self.bits[i >> 3]) >> (i & 7)) & 1)
