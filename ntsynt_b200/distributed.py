"""Multi-GPU plumbing: one process per GPU, genomes assigned to ranks, per-GPU Bloom
filters merged by one NCCL all-reduce(sum) over packed counters, minimizer tables all-gathered, graph
stage on rank 0.  The host side channel (rendezvous, the 128-byte NCCL id, table sizes, barriers) is
whatever the launcher offers -- bench.py uses torch.distributed with the gloo backend.
The reference has no counterpart: it is a single process (src/ntsynt_make_common_bf.cpp:136-160).
"""
import ctypes as C

import numpy as np

from ._lib import check, lib, ptr
from . import device


def assign_genomes(n_genomes, world):
    "round-robin ownership: genome g lives on rank g % world; returns [list of genome ids per rank]"
    return [[g for g in range(n_genomes) if g % world == r] for r in range(world)]


def owner_of(genome, world):
    return genome % world


def assign_contigs(lengths, world):
    """contig-sharded ownership (for fewer genomes than GPUs): longest-first greedy packing of the
    contigs into `world` bins; returns [sorted contig ids per rank].  A k-mer lies inside one contig, so Bloom inserts
    need no halo; a minimizer window never crosses a contig, so neither does the sketch."""
    load = [0] * world
    bins = [[] for _ in range(world)]
    for c in sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i)):
        r = min(range(world), key=lambda i: (load[i], i))
        bins[r].append(c)
        load[r] += int(lengths[c])
    return [sorted(b) for b in bins]


def field_bits(world):
    "counter width used by the all-reduce merge: smallest of 2/4/8 bits that can hold `world`"
    return 2 if world <= 3 else 4 if world <= 15 else 8


def merge_wire_bytes(filter_bytes, world):
    "bytes each rank hands to ncclAllReduce for one filter merge"
    return filter_bytes * field_bits(world)


class Comm:
    "NCCL communicator bound to a Context (nts_nccl_*)"

    def __init__(self, ctx, rank, world, unique_id):
        self.ctx, self.rank, self.world = ctx, rank, world
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(unique_id))
        h = C.c_void_p()
        check(lib.nts_nccl_init(ctx._h, buf, int(rank), int(world), C.byref(h)))
        self._h = h

    @staticmethod
    def new_unique_id():
        buf = (C.c_uint8 * 128)()
        check(lib.nts_nccl_unique_id(buf))
        return bytes(buf)

    def allreduce_and(self, bf):
        "bf := AND over all ranks' filters (replaces the cascade of cpp:136-160 across GPUs)"
        check(lib.nts_bf_allreduce_and(self._h, bf._h))

    def allreduce_or(self, bf):
        check(lib.nts_bf_allreduce_or(self._h, bf._h))

    def barrier(self):
        "stream-ordered barrier across the ranks (one-word all-reduce on the context's stream; the host does not wait)"
        check(lib.nts_nccl_barrier(self._h))

    def allgather_tables(self, table, counts, genome_for_rank):
        "every rank's minimizer table, as MinimizerTable objects on this rank"
        cnt = np.asarray(counts, dtype=np.uint64)
        out = (C.c_void_p * self.world)()
        check(lib.nts_mxs_allgather(self._h, table._h, ptr(cnt, C.c_uint64), out))
        return [device.MinimizerTable(self.ctx, C.c_void_p(out[r]), genome_for_rank[r]) for r in range(self.world)]

    def close(self):
        if self._h:
            lib.nts_nccl_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # pylint: disable=broad-except
            pass


class PeerMerge:
    """Bloom-filter merge over NVLink peer memory (nts_p2p_*): reduce-scatter + all-gather kernels that read
    the peers' arrays directly.  `gather` / `barrier` are the launcher's host side channel:
    gather(obj) -> list of every rank's obj in rank order; barrier() -> None."""

    def __init__(self, bf, rank, world, gather, barrier):
        self.bf, self.rank, self.world, self.barrier = bf, rank, world, barrier
        mine = (C.c_uint8 * 64)()
        check(lib.nts_bf_ipc_handle(bf._h, mine))
        handles = b"".join(gather(bytes(mine)))
        buf = (C.c_uint8 * len(handles)).from_buffer_copy(handles)
        h = C.c_void_p()
        rc = lib.nts_p2p_open(bf._h, buf, int(rank), int(world), C.byref(h))
        # every rank must agree: one rank without peer access would leave the others waiting in merge()
        self.ok = all(gather(rc == 0))
        self._h = h if rc == 0 else None
        if not self.ok and self._h:
            lib.nts_p2p_close(self._h)
            self._h = None
        barrier()

    def merge(self, op="and", comm=None):
        """filter := AND (or OR) over all ranks; every rank must call it.  The three inter-GPU orderings (every filter
        complete / every slice reduced / nobody still reads my slices) are stream-ordered NCCL barriers when a Comm is
        given, host barriers of the launcher's side channel otherwise."""
        if comm is None:
            self.bf.ctx.sync()
        sync = comm.barrier if comm is not None else self.barrier
        sync()                               # every filter is complete
        check(lib.nts_p2p_reduce_scatter(self._h, 0 if op == "and" else 1))
        sync()                               # every slice is reduced
        check(lib.nts_p2p_all_gather(self._h))
        sync()                               # nobody still reads my slices

    def close(self):
        if self._h:
            self.barrier()
            lib.nts_p2p_close(self._h)
            self._h = None


class ShardedMerge:
    """common = AND over genomes of (OR over ranks of the per-rank partial filters) over NVLink peer memory
    (nts_p2p_reduce_and_of_or + nts_p2p_all_gather): the merge of a contig-sharded run.  `parts[g]` is this rank's
    filter of its contigs of genome g, `common` receives the result on every rank."""

    def __init__(self, parts, common, rank, world, gather, barrier):
        self.parts, self.common, self.rank, self.world, self.barrier = parts, common, rank, world, barrier
        self._sets, self._out = [], None
        ok = True
        for bf in list(parts) + [common]:
            mine = (C.c_uint8 * 64)()
            check(lib.nts_bf_ipc_handle(bf._h, mine))
            handles = b"".join(gather(bytes(mine)))
            buf = (C.c_uint8 * len(handles)).from_buffer_copy(handles)
            h = C.c_void_p()
            rc = lib.nts_p2p_open(bf._h, buf, int(rank), int(world), C.byref(h))
            ok = ok and rc == 0
            self._sets.append(h if rc == 0 else None)
        self.ok = all(gather(ok))
        if self.ok:
            self._out = self._sets.pop()
        else:
            self.close()
        barrier()

    def merge(self, comm=None):
        if comm is None:
            self.common.ctx.sync()
        sync = comm.barrier if comm is not None else self.barrier
        sync()                               # every partial filter is complete
        arr = (C.c_void_p * len(self._sets))(*self._sets)
        check(lib.nts_p2p_reduce_and_of_or(arr, len(self._sets), self._out))
        sync()                               # every slice of the common filter is reduced
        check(lib.nts_p2p_all_gather(self._out))
        sync()                               # nobody still reads my slice / my partial filters

    def close(self):
        for h in self._sets + [self._out]:
            if h:
                lib.nts_p2p_close(h)
        self._sets, self._out = [], None


class OwnedBuild:
    """Common filter of a multi-GPU run WITHOUT building per-GPU filters and merging them (nts_bin_* / nts_bf_apply_owned):
    every rank bins the k-mers it holds of genome g (pass 1 of the partitioned insert, one plan agreed by all ranks), a
    stream-ordered barrier, then the rank that owns a slice of the filter applies the buckets of EVERY rank to that slice,
    reading them over NVLink peer memory -- 4 bytes per k-mer on the wire instead of G x (P - 1) / P partial filters --,
    ORs within the genome and ANDs across genomes on its slice; one all-gather distributes the finished filter.
    `shards_of_genome[g]` = what this rank holds of genome g (a contig shard; or, with one genome per GPU, the whole
    genome on its owner and None elsewhere).  Bucket scratch is double-buffered, so one barrier per genome suffices."""

    SLOTS = (2, 3)          # scratch slots of their own (slot 0 belongs to the single-GPU insert)

    def __init__(self, common, level, rank, world, k, plan_valid, gather, barrier, comm):
        self.common, self.level, self.rank, self.world, self.k = common, level, rank, world, int(k)
        self.comm, self.barrier = comm, barrier
        ctx = common.ctx
        self.ctx = ctx
        self._peers = []
        self._p2p = None
        ok = True
        try:
            for slot in self.SLOTS:
                check(lib.nts_bin_prepare(common._h, slot, int(plan_valid)))
        except Exception:              # pylint: disable=broad-except
            ok = False
        ok = all(gather(ok))
        if ok:
            for slot in self.SLOTS:
                hi, hc = (C.c_uint8 * 64)(), (C.c_uint8 * 64)()
                check(lib.nts_bin_ipc_handles(ctx._h, slot, hi, hc))
                got = gather((bytes(hi), bytes(hc)))
                items = b"".join(x[0] for x in got)
                curs = b"".join(x[1] for x in got)
                h = C.c_void_p()
                rc = lib.nts_binpeer_open(ctx._h, (C.c_uint8 * len(items)).from_buffer_copy(items),
                                          (C.c_uint8 * len(curs)).from_buffer_copy(curs), int(rank), int(world), slot, C.byref(h))
                ok = ok and rc == 0
                self._peers.append(h if rc == 0 else None)
            mine = (C.c_uint8 * 64)()
            check(lib.nts_bf_ipc_handle(common._h, mine))
            handles = b"".join(gather(bytes(mine)))
            h = C.c_void_p()
            rc = lib.nts_p2p_open(common._h, (C.c_uint8 * len(handles)).from_buffer_copy(handles), int(rank), int(world), C.byref(h))
            ok = ok and rc == 0
            self._p2p = h if rc == 0 else None
        self.ok = all(gather(ok))
        if self.ok:
            off, n = C.c_uint64(), C.c_uint64()
            check(lib.nts_p2p_slice(self._p2p, int(rank), C.byref(off), C.byref(n)))
            self.off16, self.n16 = int(off.value), int(n.value)
        else:
            self.close()
        barrier()

    def build(self, shards_of_genome, genomes_are_sources=False):
        """common := AND_g OR_rank bits; returns the number of k-mers (over all ranks: caller sums) that overflowed their
        bucket -- if it is not zero on every rank the result is incomplete and the caller must fall back to a merge.
        genomes_are_sources: one genome per GPU (shards_of_genome = [my genome]); every source is a genome of its own, so
        the AND is taken across the sources instead of across the calls."""
        lv, cm, off, n = self.level._h, self.common._h, self.off16, self.n16
        first = True
        for g, shard in enumerate(shards_of_genome):
            slot = g & 1
            check(lib.nts_bin_genome(cm, shard._h, self.k, self.SLOTS[slot]))
            self.comm.barrier()                                  # every rank's buckets of this genome are complete
            if not genomes_are_sources:
                check(lib.nts_bf_range_op(lv, None, off, n, 2))                       # level[slice] = 0
            for i in range(self.world):
                s = (self.rank + i) % self.world                 # rotated: every source serves one reader at a time
                if genomes_are_sources:
                    check(lib.nts_bf_range_op(lv, None, off, n, 2))
                check(lib.nts_bf_apply_owned(lv, self._peers[slot], s, off, n))
                if genomes_are_sources:
                    check(lib.nts_bf_range_op(cm, lv, off, n, 2 if first else 0))
                    first = False
            if not genomes_are_sources:
                check(lib.nts_bf_range_op(cm, lv, off, n, 2 if first else 0))         # common[slice] = / &= level[slice]
                first = False
        self.comm.barrier()                                      # every slice is final (and nobody still reads my buckets)
        check(lib.nts_p2p_all_gather(self._p2p))
        self.comm.barrier()
        over = 0
        for slot in self.SLOTS:
            n_over = C.c_uint64()
            check(lib.nts_bin_overflow(self.ctx._h, slot, C.byref(n_over)))
            over += int(n_over.value)
        return over

    def close(self):
        for h in self._peers:
            if h:
                lib.nts_binpeer_close(h)
        self._peers = []
        if self._p2p:
            lib.nts_p2p_close(self._p2p)
            self._p2p = None


def gather_sharded_table(comm, table, n_contigs, owner_of_contig, gather_objects, genome=None):
    """every rank sketched its own contigs of one genome (empty records elsewhere, global contig numbering): all-gather
    the tables and put the contigs back in order.  Returns the whole genome's table on this rank."""
    off = table.contig_offsets(n_contigs)
    info = gather_objects((len(table), off.tolist()))
    tabs = comm.allgather_tables(table, [n for n, _ in info], [None] * comm.world)
    parts, src, cnt = [], [], []
    for c in range(n_contigs):
        r = owner_of_contig[c]
        o = info[r][1]
        parts.append(tabs[r]); src.append(o[c]); cnt.append(o[c + 1] - o[c])
    out = device.MinimizerTable.concat(comm.ctx, parts, src, cnt, genome)
    for t in tabs:
        t.close()
    return out


class ShardedSketcher:
    """Refinement-round sketches of a contig-sharded run: rank 0 (which runs the graph stage) announces (genome, w,
    masks) over the host side channel, EVERY rank sketches its own contigs of that genome under the masks, and the
    tables are all-gathered and put back in contig order -- the masked rounds are sharded like round 0 instead of
    being re-sketched by rank 0 alone.  The other ranks sit in serve() until rank 0 says done()."""

    def __init__(self, comm, ctx, shards, k, common, n_contigs, owner_of_contig, bcast_object, gather_objects):
        self.comm, self.ctx, self.shards, self.k, self.common = comm, ctx, shards, k, common
        self.n_contigs, self.owner_of_contig = n_contigs, owner_of_contig
        self.bcast, self.gather_objects = bcast_object, gather_objects

    def _do(self, g, w, masks):
        t = self.ctx.sketch(self.shards[g], self.k, w, common=self.common, masks=masks)
        whole = gather_sharded_table(self.comm, t, self.n_contigs, self.owner_of_contig, self.gather_objects)
        t.close()
        return whole

    def sketch(self, g, w, masks):
        "rank 0: the whole genome's masked table"
        self.bcast(("sketch", g, w, masks))
        return self._do(g, w, masks)

    def done(self):
        self.bcast(("done",))

    def serve(self):
        "ranks > 0: take part in every sketch rank 0 asks for"
        while True:
            cmd = self.bcast(None)
            if cmd[0] == "done":
                return
            self._do(cmd[1], cmd[2], cmd[3]).close()


class GatheredBackend:
    """SyntenyEngine backend for rank 0 of a one-genome-per-GPU run: round-0 tables were sketched on
    their owner ranks and all-gathered; refinement sketches (tiny, <1 % unmasked) run locally on the
    copies of the genomes rank 0 holds."""

    def __init__(self, ctx, genomes, names, contig_names, contig_lengths, k, common, round0_tables, masked_sketch=None):
        "masked_sketch(a, w, masks) -> device table: where the refinement sketches come from (default: local copies)"
        self.masked_sketch = masked_sketch
        self.ctx, self.genomes = ctx, genomes
        self.names = list(names)
        self.contig_names, self.contig_lengths = contig_names, contig_lengths
        self.k, self.common = k, common
        self.round0 = round0_tables

    def sketch(self, a, w, masks):
        if masks is None:
            return self.round0[a]
        mx = self.sketch_table(a, w, masks)
        out = mx.to_numpy()
        mx.close()
        return out

    def sketch_table(self, a, w, masks):
        if self.masked_sketch is not None:
            return self.masked_sketch(a, w, masks)
        return self.ctx.sketch(self.genomes[a], self.k, w, common=self.common, masks=masks)

    def join(self, tables, order_asm):
        if getattr(self, "graph", None) is not None:
            self.graph.close()
        self.graph = device.MinimizerGraph(self.ctx, tables, order_asm)      # kept alive for lookup()
        return self.graph.join_result()

    def lookup(self, keys):
        return self.graph.lookup(keys)

    def close(self):
        if getattr(self, "graph", None) is not None:
            self.graph.close()
            self.graph = None
