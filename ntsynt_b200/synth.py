"""Synthetic genome workloads (SURVEY.md 8d recipe) -- host side: segment tables for the device
generator nts_genome_synthesize.  An ancestor with human-like contig proportions is never stored;
each genome is a list of segments copying ancestor intervals (forward / reverse complement), random
insertions and N runs, produced by inversions, translocations and indels.  Substitutions are applied
per base on the device.  Everything is keyed by integer seeds, so any rank regenerates any genome.
"""
import ctypes as C

import numpy as np

from ._lib import SynthSeg, check, lib
from . import device

from .synth_layout import (HUMAN_MBP, SEG_DTYPE, Layout, ancestor_layout, ancestor_nruns, contig_names,  # noqa: F401
                           genome_segments)

assert SEG_DTYPE.itemsize == C.sizeof(SynthSeg)


def _empty_packed(names):
    from .fasta import PackedGenome
    n = len(names)
    z = np.zeros(0, dtype=np.uint64)
    return PackedGenome(list(names), np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=np.uint64), z, np.zeros(n + 1, dtype=np.uint64),
                        z, z)


class Workload(Layout):
    "G synthetic genomes derived from one ancestor, materialised on the device"

    def materialize(self, ctx, g, contigs=None):
        """genome g resident in HBM (device.DeviceGenome).  contigs: only these contig indices are generated, the others
        are kept as empty records so that contig numbering stays global (contig-sharded multi-GPU runs)"""
        lengths, segs = self.segments(g)
        if contigs is not None:
            own = np.zeros(len(lengths), dtype=bool)
            own[np.asarray(list(contigs), dtype=np.int64)] = True
            lengths = np.where(own, lengths, 0).astype(np.uint64)
            segs = np.ascontiguousarray(segs[own[segs["dst_contig"]]])
            if not len(segs):                       # a rank that owns nothing of this genome still needs a valid object
                return device.DeviceGenome(ctx, _empty_packed(self.names))
        h = C.c_void_p()
        check(lib.nts_genome_synthesize(ctx._h, len(lengths), lengths.ctypes.data_as(C.POINTER(C.c_uint64)),
                                        segs.ctypes.data_as(C.POINTER(SynthSeg)), len(segs), self.seed,
                                        self.seed * 1000003 + g + 1, self.d / 200.0, self.n_repeat_fam,
                                        self.repeat_slot_prob, C.byref(h)))
        return device.DeviceGenome._from_handle(ctx, h, self.names, lengths)
