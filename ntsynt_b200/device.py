"""Thin object layer over the C-ABI: Context, DeviceGenome, BloomFilter, MinimizerTable.

Mirrors the reference's operators for this path:
  BloomFilter.insert_genome  <- bf->insert(record.seq)              src/ntsynt_make_common_bf.cpp:128-131
  BloomFilter.iand           <- the contains/insert cascade          src/ntsynt_make_common_bf.cpp:136-160
  Context.sketch             <- indexlr --long --pos [-s bf] [-r bf] bin/ntsynt_run_pipeline.smk:83-85
"""
import ctypes as C
import os

import numpy as np

from ._lib import NtsError, check, lib, ptr  # noqa: F401


def device_count():
    n = C.c_int()
    check(lib.nts_device_count(C.byref(n)))
    return n.value


class Context:
    def __init__(self, device=0):
        h = C.c_void_p()
        check(lib.nts_ctx_create(int(device), C.byref(h)))
        self._h = h
        self.device = device

    def close(self):
        if self._h:
            PinnedPool.release(id(self))
            lib.nts_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # pylint: disable=broad-except
            pass

    def sync(self):
        check(lib.nts_ctx_sync(self._h))

    def timer_start(self):
        check(lib.nts_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_float()
        check(lib.nts_timer_stop(self._h, C.byref(ms)))
        return ms.value

    @property
    def launches(self):
        return int(lib.nts_launch_count(self._h))

    @property
    def sketch_escalated(self):
        "dense sub-tiles the sparse sketch kernel handed to the dense selector so far (statistics)"
        return int(lib.nts_sketch_escalated(self._h))

    @property
    def sketch_queried_all(self):
        "sketches that looked every slot up (filters passing < 10 % of the k-mers; statistics)"
        return int(lib.nts_sketch_queried_all(self._h))

    @property
    def part_inserts(self):
        "Bloom inserts that took the partitioned path so far (statistics)"
        return int(lib.nts_part_inserts(self._h))

    @property
    def part_overflow_items(self):
        "items those inserts applied through their overflow lists (heavy-hitter k-mers)"
        return int(lib.nts_part_overflow_items(self._h))

    def mem_info(self):
        f, t = C.c_uint64(), C.c_uint64()
        check(lib.nts_mem_info(self._h, C.byref(f), C.byref(t)))
        return f.value, t.value

    # -- genome
    def upload(self, packed, async_copy=False):
        """async_copy: queue the H2D copy on the context's copy stream and return at once (the packed words should be
        page-locked and are kept alive by the returned object); consumers order themselves after the copy"""
        return DeviceGenome(self, packed, async_copy)

    # -- Bloom filter
    def bloom(self, nbytes):
        return BloomFilter(self, nbytes)

    # -- sketch (indexlr)
    def sketch(self, genome, k, w, common=None, repeat=None, masks=None, common2=None):
        """masks: optional list (per contig) of (start, end) arrays of extra N intervals.
        common2: second part of the common filter (common AND common2; BloomFilter.build_common(lazy=True))."""
        mo = ms = me = None
        if masks is not None:
            off = [0]
            starts, ends = [], []
            for c in range(genome.n_contigs):
                iv = masks[c] if c < len(masks) and masks[c] is not None else ((), ())
                s = np.asarray(iv[0], dtype=np.uint64)
                e = np.asarray(iv[1], dtype=np.uint64)
                starts.append(s)
                ends.append(e)
                off.append(off[-1] + len(s))
            mo = np.asarray(off, dtype=np.uint64)
            ms = np.concatenate(starts) if starts else np.zeros(0, dtype=np.uint64)
            me = np.concatenate(ends) if ends else np.zeros(0, dtype=np.uint64)
            if ms.size == 0:
                ms = np.zeros(1, dtype=np.uint64)
                me = np.zeros(1, dtype=np.uint64)
        h = C.c_void_p()
        check(lib.nts_sketch2(self._h, genome._h, common._h if common else None, common2._h if common2 else None,
                              repeat._h if repeat else None, int(k), int(w), ptr(mo, C.c_uint64), ptr(ms, C.c_uint64),
                              ptr(me, C.c_uint64), C.byref(h)))
        return MinimizerTable(self, h, genome)

    def hash_contig(self, genome, contig, k):
        n = int(genome.lengths[contig])
        nk = max(n - k + 1, 0)
        h0 = np.zeros(max(nk, 1), dtype=np.uint64)
        valid = np.zeros(max(nk, 1), dtype=np.uint8)
        check(lib.nts_hash_contig(self._h, genome._h, int(contig), int(k), ptr(h0, C.c_uint64), ptr(valid, C.c_uint8)))
        return h0[:nk], valid[:nk]


class DeviceGenome:
    def __init__(self, ctx, packed, async_copy=False):
        self.ctx = ctx
        self._keep = packed.words if async_copy else None      # the copy may still be reading it
        self.names = list(packed.names)
        self.lengths = np.asarray(packed.lengths, dtype=np.uint64)
        self.n_contigs = len(self.names)
        h = C.c_void_p()
        words = packed.words if packed.words.size else np.zeros(1, dtype=np.uint64)
        ns = packed.nrun_start if packed.nrun_start.size else np.zeros(1, dtype=np.uint64)
        nl = packed.nrun_len if packed.nrun_len.size else np.zeros(1, dtype=np.uint64)
        check((lib.nts_genome_upload_async if async_copy else lib.nts_genome_upload)(ctx._h, self.n_contigs, ptr(self.lengths, C.c_uint64),
                                    ptr(packed.word_off, C.c_uint64), ptr(words, C.c_uint64), int(packed.words.size),
                                    ptr(packed.nrun_off, C.c_uint64), ptr(ns, C.c_uint64), ptr(nl, C.c_uint64),
                                    C.byref(h)))
        self._h = h

    @classmethod
    def _from_handle(cls, ctx, h, names, lengths):
        self = cls.__new__(cls)
        self.ctx, self._h = ctx, h
        self.names = list(names)
        self.lengths = np.asarray(lengths, dtype=np.uint64)
        self.n_contigs = len(self.names)
        return self

    @property
    def total_bases(self):
        return int(lib.nts_genome_size(self._h))

    def close(self):
        if self._h:
            lib.nts_genome_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # pylint: disable=broad-except
            pass


class BloomFilter:
    def __init__(self, ctx, nbytes):
        self.ctx = ctx
        h = C.c_void_p()
        check(lib.nts_bf_create(ctx._h, int(nbytes), C.byref(h)))
        self._h = h
        self.nbytes = int(nbytes)

    @staticmethod
    def size_for(genome_size, fpr):
        "approximate_bf_size, src/ntsynt_make_common_bf.cpp:28-40 (+ btllib 8-byte round-up)"
        return int(lib.nts_bf_bytes(int(genome_size), float(fpr)))

    def clear(self):
        check(lib.nts_bf_clear(self._h))

    def insert_genome(self, genome, k):
        check(lib.nts_bf_insert_genome(self._h, genome._h, int(k)))

    def set_genome(self, genome, k):
        "self = bits(genome): clear + insert_genome without the zero-fill pass"
        check(lib.nts_bf_set_genome(self._h, genome._h, int(k)))

    def build_common(self, level, genomes, k, lazy=False):
        """self = AND over the genomes of their k-mer bit arrays (src/ntsynt_make_common_bf.cpp:107-160); `level` is a
        scratch filter of the same size (None for one genome); both are zeroed inside.
        lazy: the last AND pass is left out; returns True when `level` then holds the last genome's bits and the common
        filter is the pair (self, level) -- what Context.sketch takes as common / common2 -- else self is complete."""
        arr = (C.c_void_p * len(genomes))(*[g._h for g in genomes])
        if lazy:
            apart = C.c_int(0)
            check(lib.nts_bf_build_common_lazy(self._h, level._h if level is not None else None, arr, len(genomes), int(k),
                                               C.byref(apart)))
            return bool(apart.value)
        check(lib.nts_bf_build_common(self._h, level._h if level is not None else None, arr, len(genomes), int(k)))
        return self

    def build_from_and(self, filters):
        "self = AND of the given filters (same size)"
        self.clear()
        self.ior(filters[0])
        for f in filters[1:]:
            self.iand(f)
        return self

    def insert_repeats(self, scratch, genome, k):
        check(lib.nts_bf_insert_repeats(self._h, scratch._h, genome._h, int(k)))

    def iand(self, other):
        check(lib.nts_bf_and(self._h, other._h))
        return self

    def ior(self, other):
        check(lib.nts_bf_or(self._h, other._h))
        return self

    def popcount(self):
        n = C.c_uint64()
        check(lib.nts_bf_popcount(self._h, C.byref(n)))
        return n.value

    def fpr(self):
        "occupancy (1 hash function): what btllib's get_fpr() prints"
        return self.popcount() / (self.nbytes * 8.0)

    def to_numpy(self):
        out = np.empty(self.nbytes, dtype=np.uint8)
        check(lib.nts_bf_download(self._h, ptr(out, C.c_uint8)))
        return out

    def from_numpy(self, arr):
        arr = np.ascontiguousarray(arr, dtype=np.uint8)
        if arr.size != self.nbytes:
            raise ValueError("size mismatch")
        check(lib.nts_bf_upload(self._h, ptr(arr, C.c_uint8)))
        return self

    def close(self):
        if self._h:
            lib.nts_bf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # pylint: disable=broad-except
            pass


class MinimizerTable:
    "device-resident (h1, pos, contig) triples in indexlr emission order"

    def __init__(self, ctx, handle, genome):
        self.ctx, self._h, self.genome = ctx, handle, genome

    @classmethod
    def from_numpy(cls, ctx, h1, pos, contig, genome=None):
        "device table from host arrays in (contig, position) order (bin/ntsynt_run.py reads indexlr TSVs)"
        h1 = np.ascontiguousarray(h1, dtype=np.uint64)
        pos = np.ascontiguousarray(pos, dtype=np.uint32)
        contig = np.ascontiguousarray(contig, dtype=np.uint32)
        h = C.c_void_p()
        n = len(h1)
        z64, z32 = np.zeros(1, dtype=np.uint64), np.zeros(1, dtype=np.uint32)
        check(lib.nts_mxs_upload(ctx._h, n, ptr(h1 if n else z64, C.c_uint64), ptr(pos if n else z32, C.c_uint32),
                                 ptr(contig if n else z32, C.c_uint32), C.byref(h)))
        return cls(ctx, h, genome)

    def __len__(self):
        return int(lib.nts_mxs_count(self._h))

    def drop_in_filter(self, genome, bf, k):
        "a new table without the minimizers whose k-mer is in `bf` (the graph stage's --filter Filter)"
        h = C.c_void_p()
        check(lib.nts_mxs_drop_in_bf(self.ctx._h, self._h, genome._h, bf._h, int(k), C.byref(h)))
        return MinimizerTable(self.ctx, h, genome)

    def contig_offsets(self, n_contigs):
        "off[n_contigs + 1]: rows [off[c], off[c+1]) belong to contig c"
        off = np.zeros(n_contigs + 1, dtype=np.uint64)
        check(lib.nts_mxs_contig_offsets(self._h, int(n_contigs), ptr(off, C.c_uint64)))
        return off

    @classmethod
    def concat(cls, ctx, parts, src_off, cnt, genome=None):
        "new table = rows [src_off[i], src_off[i] + cnt[i]) of parts[i], in order"
        arr = (C.c_void_p * max(len(parts), 1))(*[t._h for t in parts])
        so = np.ascontiguousarray(src_off, dtype=np.uint64)
        cn = np.ascontiguousarray(cnt, dtype=np.uint64)
        h = C.c_void_p()
        z = np.zeros(1, dtype=np.uint64)
        check(lib.nts_mxs_concat(ctx._h, arr, ptr(so if len(so) else z, C.c_uint64), ptr(cn if len(cn) else z, C.c_uint64),
                                 len(parts), C.byref(h)))
        return cls(ctx, h, genome)

    def to_numpy(self):
        n = len(self)
        h1 = np.empty(max(n, 1), dtype=np.uint64)
        pos = np.empty(max(n, 1), dtype=np.uint32)
        ctg = np.empty(max(n, 1), dtype=np.uint32)
        check(lib.nts_mxs_download(self._h, ptr(h1, C.c_uint64), ptr(pos, C.c_uint32), ptr(ctg, C.c_uint32)))
        return h1[:n], pos[:n], ctg[:n]

    def close(self):
        if self._h:
            lib.nts_mxs_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # pylint: disable=broad-except
            pass


class PinnedPool:
    """page-locked host arrays reused across calls (D2H into pageable memory runs at a few GB/s).  One set of buffers
    per Context (two GPUs or two threads never share one); a buffer that has to grow is retired, not freed, so numpy
    views handed out earlier stay valid memory until the context is closed.  A view's CONTENTS are valid until the next
    MinimizerGraph of the same context downloads into the same buffer (the engine of a finished run keeps no live
    dependency on them: its results are the output texts)."""
    _bufs = {}
    _retired = {}

    @classmethod
    def get(cls, key, shape, dtype, owner=None):
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        key = (owner, key)
        ent = cls._bufs.get(key)
        if ent is None or ent[1] < n:
            if ent is not None:
                cls._retired.setdefault(owner, []).append(ent[0])
            p = C.c_void_p()
            check(lib.nts_host_alloc(max(n, 1) + 64, C.byref(p)))
            ent = (p, n + 64)
            cls._bufs[key] = ent
        raw = (C.c_uint8 * max(n, 1)).from_address(ent[0].value)
        return np.frombuffer(raw, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    @classmethod
    def release(cls, owner):
        "free every buffer of a context (Context.close)"
        for key in [k for k in cls._bufs if k[0] == owner]:
            lib.nts_host_free(cls._bufs.pop(key)[0])
        for p in cls._retired.pop(owner, []):
            lib.nts_host_free(p)


class MinimizerGraph:
    """Device join of G minimizer tables (kernel iv): vertex table in the orienting assembly's
    list order, full-weight links and vertex degrees.  See nts_graph_build in the header."""

    def __init__(self, ctx, tables, order_asm):
        self.ctx = ctx
        self.n_asm = len(tables)
        arr = (C.c_void_p * self.n_asm)(*[t._h for t in tables])
        h = C.c_void_p()
        check(lib.nts_graph_build(ctx._h, arr, self.n_asm, int(order_asm), C.byref(h)))
        self._h = h
        self.order_asm = order_asm

    def __len__(self):
        return int(lib.nts_graph_vertices(self._h))

    def vertices(self):
        "h1[V] u64, pos[G,V] u32, contig[G,V] u32, rank[G,V] u32, link[V] u8, degree[V] u8"
        V, G = len(self), self.n_asm
        n = max(V, 1)
        h1 = PinnedPool.get("h1", (n,), np.uint64, owner=id(self.ctx))
        pos = PinnedPool.get("pos", (G, n), np.uint32, owner=id(self.ctx))
        ctg = PinnedPool.get("ctg", (G, n), np.uint32, owner=id(self.ctx))
        rank = PinnedPool.get("rank", (G, n), np.uint32, owner=id(self.ctx))
        link = PinnedPool.get("link", (n,), np.uint8, owner=id(self.ctx))
        deg = PinnedPool.get("deg", (n,), np.uint8, owner=id(self.ctx))
        if V:
            check(lib.nts_graph_download_vertices(self._h, ptr(h1, C.c_uint64), ptr(pos, C.c_uint32),
                                                  ptr(ctg, C.c_uint32), ptr(rank, C.c_uint32), ptr(link, C.c_uint8),
                                                  ptr(deg, C.c_uint8)))
        return h1[:V], pos[:, :V], ctg[:, :V], rank[:, :V], link[:V], deg[:V]

    def links(self):
        "inv[G,V] u32, incmask[V], decmask[V], spread[V] u32: per-pair arrays of (i, i+1)"
        V, G = len(self), self.n_asm
        n = max(V, 1)
        inv = PinnedPool.get("inv", (G, n), np.uint32, owner=id(self.ctx))
        inc = PinnedPool.get("inc", (n,), np.uint32, owner=id(self.ctx))
        dec = PinnedPool.get("dec", (n,), np.uint32, owner=id(self.ctx))
        spread = PinnedPool.get("spread", (n,), np.uint32, owner=id(self.ctx))
        if V:
            check(lib.nts_graph_download_links(self._h, ptr(inv, C.c_uint32), ptr(inc, C.c_uint32), ptr(dec, C.c_uint32),
                                               ptr(spread, C.c_uint32)))
        return inv[:, :V], inc[:V], dec[:V], spread[:V]

    def lookup(self, keys):
        "vertex id per h1 (0xFFFFFFFF if it is not a vertex), through the device join table"
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        out = np.full(max(len(keys), 1), 0xFFFFFFFF, dtype=np.uint32)
        if len(keys):
            check(lib.nts_graph_lookup(self._h, ptr(keys, C.c_uint64), len(keys), ptr(out, C.c_uint32)))
        return out[:len(keys)]

    def cums(self):
        "CI, CD int32 [G, V+1]: device prefix sums of the per-pair direction bits"
        V, G = len(self), self.n_asm
        ci = PinnedPool.get("ci", (G, V + 1), np.uint32, owner=id(self.ctx))
        cd = PinnedPool.get("cd", (G, V + 1), np.uint32, owner=id(self.ctx))
        ci[:, 0] = 0; cd[:, 0] = 0
        if V:
            check(lib.nts_graph_download_cums(self._h, ptr(ci, C.c_uint32), ptr(cd, C.c_uint32)))
        return ci.view(np.int32), cd.view(np.int32)

    def host_arrays(self, cap):
        """H[cap] u64, POS[G,cap] i64, CTG[G,cap] i32, nbr[cap,2] i32, conn[cap] u8: the graph stage's working
        columns in their final dtypes, written by the device into pinned buffers with room for later vertices"""
        V, G = len(self), self.n_asm
        cap = max(int(cap), V, 1)
        H = PinnedPool.get("H", (cap,), np.uint64, owner=id(self.ctx))
        POS = PinnedPool.get("POS64", (G, cap), np.int64, owner=id(self.ctx))
        CTG = PinnedPool.get("CTG32", (G, cap), np.int32, owner=id(self.ctx))
        nbr = PinnedPool.get("nbr", (cap, 2), np.int32, owner=id(self.ctx))
        conn = PinnedPool.get("conn", (cap,), np.uint8, owner=id(self.ctx))
        if V:
            check(lib.nts_graph_download_host_arrays(self._h, cap, ptr(H, C.c_uint64), ptr(POS, C.c_longlong),
                                                     ptr(CTG, C.c_int32), ptr(nbr, C.c_int32), ptr(conn, C.c_uint8)))
        return H, POS, CTG, nbr, conn

    def sparse_lists(self, bp, breaks=True):
        """(breaks, deg3, big) sorted int64 arrays -- see nts_graph_sparse_lists in the header; breaks=False leaves the
        (long) list of pairs without a link on the device (the lean form gets its runs from nts_graph_runs)"""
        cnt = (C.c_uint64 * 3)()
        bp = min(int(bp), 0xFFFFFFFF)
        check(lib.nts_graph_sparse_lists(self._h, bp, None, None, None, cnt))
        if not breaks:
            cnt[0] = 0
        arrs = [np.empty(max(int(c), 1), dtype=np.uint32) for c in cnt]
        check(lib.nts_graph_sparse_lists(self._h, bp, ptr(arrs[0], C.c_uint32) if breaks else None, ptr(arrs[1], C.c_uint32),
                                         ptr(arrs[2], C.c_uint32), cnt))
        if not breaks:
            cnt[0] = 0
        return tuple(a[:int(c)].astype(np.int64) for a, c in zip(arrs, cnt))

    def rank_inv(self):
        "RANK[G,V], INV[G,V] u32 (round-0 neighbourhoods of the simplification candidates)"
        V, G = len(self), self.n_asm
        n = max(V, 1)
        rank = PinnedPool.get("rank", (G, n), np.uint32, owner=id(self.ctx))
        inv = PinnedPool.get("inv", (G, n), np.uint32, owner=id(self.ctx))
        if V:
            check(lib.nts_graph_download_vertices(self._h, None, None, None, ptr(rank, C.c_uint32), None, None))
            check(lib.nts_graph_download_links(self._h, ptr(inv, C.c_uint32), None, None, None))
        return rank[:, :V], inv[:, :V]

    def pair_masks(self):
        "incmask, decmask, spread [V] u32 of the pairs (i, i+1) (fetched only when a refinement round overwrites positions)"
        V = len(self)
        n = max(V, 1)
        inc = PinnedPool.get("inc", (n,), np.uint32, owner=id(self.ctx))
        dec = PinnedPool.get("dec", (n,), np.uint32, owner=id(self.ctx))
        spread = PinnedPool.get("spread", (n,), np.uint32, owner=id(self.ctx))
        if V:
            check(lib.nts_graph_download_links(self._h, None, ptr(inc, C.c_uint32), ptr(dec, C.c_uint32), ptr(spread, C.c_uint32)))
        return inc[:V], dec[:V], spread[:V]


    # ---- lean form: the O(V) columns stay on the device, the host asks for what it walks
    _GATHER = {"h1": (0, np.uint64, False), "pos": (1, np.int64, True), "ctg": (2, np.int32, True),
               "rank": (3, np.uint32, True), "inv": (4, np.uint32, True)}

    def gather(self, what, ids):
        "column `what` at vertex ids: h1 -> u64[n]; pos / ctg / rank / inv -> [G, n]"
        code, dt, per_asm = self._GATHER[what]
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        n = len(ids)
        out = np.zeros((self.n_asm, n) if per_asm else (n,), dtype=dt)
        if n:
            check(lib.nts_graph_gather(self._h, code, ptr(ids, C.c_int64), n, out.ctypes.data_as(C.c_void_p)))
        return out

    def range_sums(self, lo, hi):
        "up, down i64 [G, n]: increasing / decreasing pairs (j, j+1), lo <= j < hi, per assembly (device prefix sums)"
        lo = np.ascontiguousarray(lo, dtype=np.int64)
        hi = np.ascontiguousarray(hi, dtype=np.int64)
        n = len(lo)
        up = np.zeros((self.n_asm, n), dtype=np.int64)
        down = np.zeros((self.n_asm, n), dtype=np.int64)
        if n:
            check(lib.nts_graph_range_sums(self._h, ptr(lo, C.c_int64), ptr(hi, C.c_int64), n, ptr(up, C.c_int64),
                                           ptr(down, C.c_int64)))
        return up, down

    def neigh(self, cand):
        "left, right, rank i64 [n, G] of the candidates (neighbour on the same contig line per assembly, or -1)"
        cand = np.ascontiguousarray(cand, dtype=np.int64)
        n = len(cand)
        out = [np.zeros((n, self.n_asm), dtype=np.int64) for _ in range(3)]
        if n:
            check(lib.nts_graph_neigh(self._h, ptr(cand, C.c_int64), n, *[ptr(x, C.c_int64) for x in out]))
        return tuple(out)

    def links_nbr(self, cap):
        "nbr[cap, 2] i32, conn[cap] u8 of the weight-filtered graph (pinned, room for later vertices)"
        V = len(self)
        cap = max(int(cap), V, 1)
        nbr = PinnedPool.get("nbr", (cap, 2), np.int32, owner=id(self.ctx))
        conn = PinnedPool.get("conn", (cap,), np.uint8, owner=id(self.ctx))
        if V:
            check(lib.nts_graph_download_links_nbr(self._h, cap, ptr(nbr, C.c_int32), ptr(conn, C.c_uint8)))
        return nbr, conn

    def set_links(self, idx, val):
        "push the host's edits of pairs (i, i+1) back to the device link bitmap"
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        val = np.ascontiguousarray(val, dtype=np.uint8)
        if len(idx):
            check(lib.nts_graph_set_links(self._h, ptr(idx, C.c_int64), ptr(val, C.c_uint8), len(idx)))

    def runs(self):
        "(starts, ends) i64, ascending: maximal runs of >= 2 base vertices joined by full-weight edges (device)"
        n = C.c_uint64()
        check(lib.nts_graph_runs(self._h, None, None, C.byref(n)))
        st = np.zeros(max(n.value, 1), dtype=np.int64)
        en = np.zeros(max(n.value, 1), dtype=np.int64)
        if n.value:
            check(lib.nts_graph_runs(self._h, ptr(st, C.c_int64), ptr(en, C.c_int64), C.byref(n)))
        return st[:n.value], en[:n.value]

    def runs_to_blocks(self, starts, ends, bp, m_pct, min_mx):
        """paths, blocks, indel cuts and the >= min_mx filter for plain (i, i+1) runs on the device (kernel iv-c);
        dict of b_lo, b_hi, b_plus, b_dir (surviving blocks), r_lo, r_hi (deleted intervals), cuts"""
        starts = np.ascontiguousarray(starts, dtype=np.int64)
        ends = np.ascontiguousarray(ends, dtype=np.int64)
        n = len(starts)
        z32 = np.zeros(0, dtype=np.uint32)
        if not n:
            return dict(b_lo=z32, b_hi=z32, b_plus=z32, b_dir=np.zeros(0, dtype=np.int8), r_lo=z32, r_hi=z32, cuts=z32)
        bp = min(int(bp), 0xFFFFFFFF)
        nb = C.c_uint64()
        check(lib.nts_graph_big_count(self._h, bp, C.byref(nb)))
        cap = n + nb.value + 1
        a = [np.zeros(cap, dtype=np.uint32) for _ in range(6)]
        d = np.zeros(cap, dtype=np.int8)
        cnt = (C.c_uint64 * 3)()
        check(lib.nts_graph_runs_to_blocks(self._h, ptr(starts, C.c_int64), ptr(ends, C.c_int64), n, bp, float(m_pct),
                                           int(min_mx), ptr(a[0], C.c_uint32), ptr(a[1], C.c_uint32), ptr(a[2], C.c_uint32),
                                           d.ctypes.data_as(C.POINTER(C.c_int8)), ptr(a[3], C.c_uint32), ptr(a[4], C.c_uint32),
                                           ptr(a[5], C.c_uint32), cnt, cap))
        nb_, nr, nc = (int(x) for x in cnt)
        return dict(b_lo=a[0][:nb_], b_hi=a[1][:nb_], b_plus=a[2][:nb_], b_dir=d[:nb_], r_lo=a[3][:nr], r_hi=a[4][:nr],
                    cuts=a[5][:nc])

    def refine_filter(self, tables, seg_lo, seg_hi, term, x_key, x_vid, iv_start, iv_maxend, iv_off):
        """one refinement round's minimizer filtering on the device (nts_graph_refine_filter); returns
        (n_raw[G], [(h1 u64, pos i64, ctg i64, sub i64) per assembly])"""
        G = self.n_asm
        arr = (C.c_void_p * G)(*[t._h for t in tables])
        u32 = lambda x: np.ascontiguousarray(x, dtype=np.uint32)          # noqa: E731
        seg_lo, seg_hi, term, x_vid = u32(seg_lo), u32(seg_hi), u32(term), u32(x_vid)
        x_key = np.ascontiguousarray(x_key, dtype=np.uint64)
        iv_start = np.ascontiguousarray(iv_start, dtype=np.int64)
        iv_maxend = np.ascontiguousarray(iv_maxend, dtype=np.int64)
        iv_off = np.ascontiguousarray(iv_off, dtype=np.uint64)
        n_raw = np.zeros(G, dtype=np.uint64)
        off = np.zeros(G + 1, dtype=np.uint64)
        z32, z64, zi = np.zeros(1, dtype=np.uint32), np.zeros(1, dtype=np.uint64), np.zeros(1, dtype=np.int64)
        cap = 1 << 16
        while True:
            o_h1 = np.zeros(cap, dtype=np.uint64)
            o_pos, o_ctg, o_sub = (np.zeros(cap, dtype=np.uint32) for _ in range(3))
            check(lib.nts_graph_refine_filter(
                self._h, arr, ptr(seg_lo if len(seg_lo) else z32, C.c_uint32), ptr(seg_hi if len(seg_hi) else z32, C.c_uint32),
                len(seg_lo), ptr(term if len(term) else z32, C.c_uint32), len(term), ptr(x_key if len(x_key) else z64, C.c_uint64),
                ptr(x_vid if len(x_vid) else z32, C.c_uint32), len(x_key), ptr(iv_start if len(iv_start) else zi, C.c_int64),
                ptr(iv_maxend if len(iv_maxend) else zi, C.c_int64), ptr(iv_off, C.c_uint64), ptr(n_raw, C.c_uint64),
                ptr(off, C.c_uint64), ptr(o_h1, C.c_uint64), ptr(o_pos, C.c_uint32), ptr(o_ctg, C.c_uint32), ptr(o_sub, C.c_uint32),
                cap))
            if int(off[G]) <= cap:
                break
            cap = int(off[G])
        out = []
        for a in range(G):
            i0, i1 = int(off[a]), int(off[a + 1])
            out.append((o_h1[i0:i1].copy(), o_pos[i0:i1].astype(np.int64), o_ctg[i0:i1].astype(np.int64), o_sub[i0:i1].astype(np.int64)))
        return [int(x) for x in n_raw], out

    def join_result(self, full=False, lean=None):
        """everything SyntenyEngine needs from the join, as a dict.  Default: device-resident form -- positions,
        contigs, hashes, ranks and the direction prefix sums stay in HBM and are read through gather / range_sums /
        neigh / runs / runs_to_blocks; only the weight-filtered graph (links_nbr) and three sparse lists come to the
        host.  lean="host" gives the round-1 form (host-ready O(V) columns); full=True also returns the raw columns
        (tests, .mx.dot writer)."""
        V = len(self)
        if lean is None:
            lean = os.environ.get("NTS_GRAPH_FORM", "device")
        if lean != "host" and not full:
            return dict(V=V, gather=self.gather, range_sums=self.range_sums, neigh=self.neigh, links_nbr=self.links_nbr,
                        set_links=self.set_links, runs=self.runs, runs_to_blocks=self.runs_to_blocks,
                        refine_filter=self.refine_filter,
                        sparse=lambda bp: self.sparse_lists(bp, breaks=False))
        RANK, INV = self.rank_inv()
        CI, CD = self.cums()
        res = dict(V=V, RANK=RANK, INV=INV, CI=CI, CD=CD, host=self.host_arrays, sparse=self.sparse_lists,
                   pair_masks=self.pair_masks)
        if full:
            H, POS, CTG, _, link, deg = self.vertices()
            _, inc, dec, spread = self.links()
            res.update(H=H, POS=POS, CTG=CTG, link=link, degree=deg, incmask=inc, decmask=dec, spread=spread)
        return res

    def edges(self):
        "(u, v, support) of the distinct adjacency edges in build_graph's first-insertion order"
        n = C.c_uint64()
        check(lib.nts_graph_edges(self._h, C.byref(n)))
        E = n.value
        u = np.empty(max(E, 1), dtype=np.uint32)
        v = np.empty(max(E, 1), dtype=np.uint32)
        s = np.empty(max(E, 1), dtype=np.uint32)
        if E:
            check(lib.nts_graph_download_edges(self._h, ptr(u, C.c_uint32), ptr(v, C.c_uint32), ptr(s, C.c_uint32)))
        return u[:E], v[:E], s[:E]

    def close(self):
        if self._h:
            lib.nts_graph_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # pylint: disable=broad-except
            pass


def _genome_nruns(self):
    "(nrun_off[n_contigs+1], nrun_start, nrun_len) of a device genome"
    off = np.zeros(self.n_contigs + 1, dtype=np.uint64)
    n = C.c_uint64()
    check(lib.nts_genome_nruns(self._h, ptr(off, C.c_uint64), None, None, 0, C.byref(n)))
    cap = max(int(n.value), 1)
    st = np.zeros(cap, dtype=np.uint64)
    ln = np.zeros(cap, dtype=np.uint64)
    check(lib.nts_genome_nruns(self._h, ptr(off, C.c_uint64), ptr(st, C.c_uint64), ptr(ln, C.c_uint64), cap,
                               C.byref(n)))
    return off, st[:n.value], ln[:n.value]


def _genome_contig_words(self, contig):
    nw = int(lib.nts_packed_words(int(self.lengths[contig])))
    words = np.zeros(max(nw, 1), dtype=np.uint64)
    check(lib.nts_genome_download_contig(self._h, int(contig), ptr(words, C.c_uint64)))
    return words[:nw]


def _genome_contig_ascii(self, contig, start=0, length=None):
    "ASCII of (part of) a contig, N runs restored -- for host-side consumers (CPU baseline, --seq output)"
    L = int(self.lengths[contig])
    length = L - start if length is None else min(length, L - start)
    words = self.contig_words(contig)
    buf = C.create_string_buffer(max(length, 1))
    check(lib.nts_unpack_ascii(ptr(words, C.c_uint64), int(start), int(length), buf))
    seq = bytearray(buf.raw[:length])
    off, st, ln = self.nruns()
    for i in range(int(off[contig]), int(off[contig + 1])):
        a, b = max(int(st[i]), start), min(int(st[i] + ln[i]), start + length)
        if a < b:
            seq[a - start:b - start] = b"N" * (b - a)
    return bytes(seq)


def _genome_to_packed(self):
    "host copy (fasta.PackedGenome) of a device genome"
    from .fasta import PackedGenome
    parts, woff, off = [], [], 0
    for c in range(self.n_contigs):
        w = self.contig_words(c)
        parts.append(w); woff.append(off); off += len(w)
    no, ns, nl = self.nruns()
    return PackedGenome(self.names, self.lengths, woff, np.concatenate(parts) if parts else np.zeros(0, np.uint64),
                        no, ns, nl)


DeviceGenome.nruns = _genome_nruns
DeviceGenome.contig_words = _genome_contig_words
DeviceGenome.contig_ascii = _genome_contig_ascii
DeviceGenome.to_packed = _genome_to_packed


# ---- profiling / transfer counters / pinned memory (bench.py)
def _ctx_prof_enable(self, on=True):
    check(lib.nts_prof_enable(self._h, 1 if on else 0))


def _ctx_prof_reset(self):
    check(lib.nts_prof_reset(self._h))


def _ctx_prof(self):
    "{family: (ms, units, launches)} accumulated since the last reset"
    out = {}
    for i in range(lib.nts_prof_count()):
        ms, units, n = C.c_double(), C.c_double(), C.c_uint64()
        check(lib.nts_prof_get(self._h, i, C.byref(ms), C.byref(units), C.byref(n)))
        out[lib.nts_prof_name(i).decode()] = (ms.value, units.value, n.value)
    return out


def _ctx_xfer(self):
    a, b = C.c_uint64(), C.c_uint64()
    check(lib.nts_xfer_bytes(self._h, C.byref(a), C.byref(b)))
    return a.value, b.value


Context.prof_enable = _ctx_prof_enable
Context.prof_reset = _ctx_prof_reset
Context.prof = _ctx_prof
Context.xfer_bytes = _ctx_xfer


class PinnedU64:
    "page-locked uint64 host array (genome words staged for H2D copies)"

    def __init__(self, n):
        p = C.c_void_p()
        check(lib.nts_host_alloc(int(n) * 8, C.byref(p)))
        self._p = p
        self.array = np.ctypeslib.as_array((C.c_uint64 * max(int(n), 1)).from_address(p.value))[:n]

    def close(self):
        if self._p:
            self.array = None
            lib.nts_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # pylint: disable=broad-except
            pass
