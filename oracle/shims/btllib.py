"""Import-only stand-in for btllib's Python module (TEST INFRASTRUCTURE ONLY).
bin/ntsynt_synteny.py:17 imports it; it is used only with --filter Filter (:605-607)."""


class KmerBloomFilter:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("btllib is not available; --filter Filter is outside the oracle harness")
