"""In-process ntSynt pipeline on one GPU: FASTA -> common Bloom filter -> sketches -> graph ->
synteny blocks.  Collapses the three processes of bin/ntsynt_run_pipeline.smk:55-103
(ntsynt_make_common_bf, indexlr x G, ntsynt_run.py) into one, with genomes, filters and
minimizer tables resident in HBM; intermediate files are written only on request.
"""
import os
import time

import numpy as np

from . import device, fasta
from .synteny import SyntenyEngine


def tsv_name(fasta_path, k, w):
    "the sketch file name the smk gives a genome: <basename>.k<k>.w<w>.tsv (smk:44-46,78)"
    return f"{os.path.basename(fasta_path)}.k{k}.w{w}.tsv"


def processing_order(names):
    "the reference processes assemblies in reverse-sorted TSV-name order (bin/ntsynt_synteny.py:34)"
    return sorted(range(len(names)), key=lambda i: names[i], reverse=True)


def build_common_bf(ctx, genomes, paths, k, fpr=0.025, nbytes=None, log=None):
    """src/ntsynt_make_common_bf.cpp:107-160 on the device.  `paths` are the strings the genomes
    were given as (they are sorted as plain strings, cpp:107; the size comes from the first).
    Returns the common filter (AND of per-genome bit arrays == the cascade with 1 hash fn)."""
    order = sorted(range(len(genomes)), key=lambda i: paths[i])
    if nbytes is None:
        nbytes = device.BloomFilter.size_for(genomes[order[0]].total_bases, fpr)
        if log:
            log(f"Genome size (bp): {genomes[order[0]].total_bases}")
    if log:
        log(f"BF size (bytes): {nbytes}")
    common = ctx.bloom(nbytes)
    if not log:
        # one call: genome 0 written as a whole filter, every further genome ANDed in by the apply pass (nts_bf_build_common)
        level = ctx.bloom(nbytes) if len(order) > 1 else None
        common.build_common(level, [genomes[i] for i in order], k)
        if level is not None:
            level.close()
        return common
    common.insert_genome(genomes[order[0]], k)
    if log:
        log(f"Bloom filter FPR: {common.fpr()}")
    if len(order) > 1:
        level = ctx.bloom(nbytes)
        for n, i in enumerate(order[1:]):
            if n:
                level.clear()
            level.insert_genome(genomes[i], k)
            common.iand(level)
            if log:
                log(f"Bloom filter FPR: {common.fpr()}")
        level.close()
    return common


def ingest_and_build(ctx, fastas, k, fpr=0.025, common=True, timing=None, lazy=False):
    """FASTA files -> device genomes + common Bloom filter as a pipeline: the files are read concurrently
    (fasta.read_fastas: inflate + scan + multi-threaded pack per file); genome i is uploaded on the copy stream and
    inserted into the filter as soon as it is parsed, in the sorted-path order of src/ntsynt_make_common_bf.cpp:107-160,
    while the later files are still being read.  Returns (packed, genomes, bf) in the order of `fastas`.
    lazy: the last AND pass is left out and (packed, genomes, bf, last) comes back -- the common filter is bf AND last
    (None with fewer than two genomes), which the sketches take as a pair (nts_sketch2)."""
    G = len(fastas)
    bf_order = sorted(range(G), key=lambda i: str(fastas[i]))
    packed, genomes = [None] * G, [None] * G
    bf = level = None
    t0 = time.perf_counter()
    wait = 0.0
    it = fasta.read_fastas([fastas[i] for i in bf_order])
    for j in range(G):
        tw = time.perf_counter()
        _, pk = next(it)
        wait += time.perf_counter() - tw
        i = bf_order[j]
        packed[i] = pk
        genomes[i] = ctx.upload(pk, async_copy=True)
        if not common:
            continue
        if j == 0:
            bf = ctx.bloom(device.BloomFilter.size_for(genomes[i].total_bases, fpr))
            bf.set_genome(genomes[i], k)
        else:
            if level is None:
                level = ctx.bloom(bf.nbytes)
            level.set_genome(genomes[i], k)
            if j + 1 < G or not lazy:
                bf.iand(level)
    if level is not None and not lazy:
        level.close()
        level = None
    ctx.sync()
    if timing is not None:
        timing["ingest_build_s"] = time.perf_counter() - t0
        timing["waited_for_parser_s"] = wait
    return (packed, genomes, bf, level) if lazy else (packed, genomes, bf)


class CudaBackend:
    "SyntenyEngine backend on the CUDA library (the only backend the package ships)"

    def __init__(self, ctx, genomes, names, contig_names, contig_lengths, k, common=None, repeat=None, round0=None,
                 filter_mode=None, common2=None):
        """filter_mode (bin/ntsynt_synteny.py:172-187,601-609): "Indexlr" hands the repeat filter to the sketch kernel
        (indexlr -r: a k-mer in it is never a minimizer), "Filter" drops the minimizers of every sketch whose k-mer is
        in it (read_minimizers(tsv, repeat_bf), ntjoin_utils.py:182)."""
        if filter_mode not in (None, "Filter", "Indexlr"):
            raise ValueError(f"unknown repeat filter mode {filter_mode!r}")
        if filter_mode and repeat is None:
            raise ValueError("If --filter is specified, must supply repeat Bloom filter with --repeat")
        self.filter_mode = filter_mode
        self.ctx, self.genomes = ctx, genomes
        self.names = list(names)
        self.contig_names = contig_names
        self.contig_lengths = contig_lengths
        self.k = k
        self.common, self.repeat = common, repeat
        self.common2 = common2    # the common filter is (common AND common2) when the last AND pass was left out
        self.round0 = round0      # optional pre-made round-0 tables (bin/ntsynt_run.py: sketches read from TSVs)
        self.graph = None
        self.timing = {"sketch_ms": 0.0, "join_ms": 0.0}

    def sketch(self, a, w, masks):
        t0 = time.perf_counter()
        if masks is None and self.round0 is not None:
            mx = self._drop_repeats(a, self.round0[a])      # tables read from TSV files (bin/ntsynt_run.py)
        else:
            mx = self._sketch(a, w, masks)
        self.timing["sketch_ms"] += (time.perf_counter() - t0) * 1e3
        if masks is None:
            return mx                      # round 0: stays on the device for the join
        out = mx.to_numpy()
        mx.close()
        return out

    def sketch_table(self, a, w, masks):
        "a masked refinement sketch that stays on the device (consumed by MinimizerGraph.refine_filter)"
        t0 = time.perf_counter()
        mx = self._sketch(a, w, masks)
        self.timing["sketch_ms"] += (time.perf_counter() - t0) * 1e3
        return mx

    def _sketch(self, a, w, masks):
        rep = self.repeat if self.filter_mode == "Indexlr" else None
        return self._drop_repeats(a, self.ctx.sketch(self.genomes[a], self.k, w, common=self.common, repeat=rep, masks=masks,
                                                     common2=self.common2))

    def _drop_repeats(self, a, mx):
        if self.filter_mode != "Filter":
            return mx
        out = mx.drop_in_filter(self.genomes[a], self.repeat, self.k)
        mx.close()
        return out

    def join(self, tables, order_asm):
        t0 = time.perf_counter()
        if getattr(self, "graph", None) is not None:
            self.graph.close()
        self.graph = device.MinimizerGraph(self.ctx, tables, order_asm)      # kept alive for lookup()
        res = self.graph.join_result()
        for t in tables:
            t.close()
        self.timing["join_ms"] += (time.perf_counter() - t0) * 1e3
        return res

    def lookup(self, keys):
        return self.graph.lookup(keys)

    def write_dot(self, path, j):
        from . import io
        H, POS, CTG = self.graph.vertices()[:3]
        io.write_mx_dot(path, self.names, self.contig_names, H, POS, CTG, self.graph.edges())

    def close(self):
        if getattr(self, "graph", None) is not None:
            self.graph.close()
            self.graph = None


def run_ntsynt(fastas, k=24, w=1000, w_rounds=(100, 10), indel=10000, merge="10000", block_size=500, fpr=0.025,
               prefix="ntSynt", simplify=True, common=True, device_index=0, write_files=True, quiet=True,
               packed=None, ctx=None, intermediates=False):
    """The whole path for `fastas` (paths; .gz accepted).  Returns (final_tsv_text, engine).
    `packed`: optional pre-parsed fasta.PackedGenome list (same order as `fastas`)."""
    own_ctx = ctx is None
    ctx = ctx or device.Context(device_index)
    bases = [os.path.basename(f)[:-3] if f.endswith(".gz") else os.path.basename(f) for f in fastas]
    names = [tsv_name(b, k, w) for b in bases]
    order = processing_order(names)
    timing = {}
    bf2 = None
    if packed is None:
        packed, genomes, bf, bf2 = ingest_and_build(ctx, fastas, k, fpr, common, timing, lazy=True)
        if intermediates and bf2 is not None:         # <prefix>.common.bf is one filter
            bf.iand(bf2)
            bf2.close()
            bf2 = None
    else:
        genomes = [ctx.upload(p) for p in packed]
        # the filter is sized from the lexicographically first PATH STRING as given (src/ntsynt_make_common_bf.cpp:107,116),
        # not from the first basename
        bf = build_common_bf(ctx, genomes, [str(f) for f in fastas], k, fpr) if common else None
    be = CudaBackend(ctx, [genomes[i] for i in order], [names[i] for i in order],
                     [packed[i].names for i in order], [[int(x) for x in packed[i].lengths] for i in order], k,
                     common=bf, common2=bf2)
    eng = SyntenyEngine(be, k, w, list(w_rounds), indel, merge, block_size, simplify=simplify, prefix=prefix,
                        write_files=write_files, quiet=quiet)
    if intermediates:
        # the files the smk rules leave behind (smk:44-103): .fai, <prefix>.common.bf, sketch TSVs, .mx.dot
        from . import io
        for b, p in zip(bases, packed):
            if p.fai:
                fasta.write_fai(p, b + ".fai")
        if bf is not None:
            io.save_bf(f"{prefix}.common.bf", bf, k)
        for b, p, g in zip(bases, packed, genomes):
            t = ctx.sketch(g, k, w, common=bf)
            with open(tsv_name(b, k, w), "w", encoding="utf-8") as fh:
                io.write_sketch_tsv(fh, p, t.to_numpy(), k)
            t.close()
        eng.dot_path = f"{prefix}.mx.dot"
    out = eng.run()
    eng.backend_timing = be.timing
    eng.ingest_timing = timing
    eng.total_bases = sum(p.total_bases for p in packed)
    be.close()
    if bf is not None:
        bf.close()
    if bf2 is not None:
        bf2.close()
    for g in genomes:
        g.close()
    if own_ctx:
        ctx.close()
    return out, eng
