"""Bit-exact parity of the partitioned Bloom inserts against the direct RED.OR kernel and the C oracle
(oracle/ntsynt_oracle.c, which follows src/ntsynt_make_common_bf.cpp:122-160):

  * the production pair bf_bin_kernel + bf_apply_kernel (csrc/nts_bin.cuh) -- the path every real genome takes and
    bench.py times -- at many buckets, with a short last region, and with bucket capacities shrunk so that the
    overflow-to-direct-atomics branch fires (knobs NTS_BF_REGION_SHIFT, NTS_BF_CAP_SCALE);
  * the atomics-free three-pass variant (csrc/nts_part.cuh, NTS_BF_IMPL=3; knobs NTS_BF_P1MAX, NTS_BF_P2,
    NTS_BF_CAP_SCALE, NTS_BF_OVF_CAP).

The knobs only change bucket counts and capacities; every setting must give the same bits."""
import os
from contextlib import contextmanager

import numpy as np
import pytest

from ntsynt_b200 import device, fasta, synth
from oracle import sketch_oracle as so

pytestmark = pytest.mark.gpu
K = 24


@contextmanager
def env(**kw):
    old = {k: os.environ.get(k) for k in kw}
    os.environ.update({k: str(v) for k, v in kw.items()})
    try:
        yield
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _records(gen):
    return [(gen.names[c], gen.contig_ascii(c)) for c in range(gen.n_contigs)]


@pytest.fixture(scope="module")
def small(cuda_ctx):
    "3 x 6 Mbp genomes with N runs and repeats + direct-kernel and oracle bits"
    wl = synth.Workload(3, 6_000_000, 1.0)
    gens = [wl.materialize(cuda_ctx, g) for g in range(3)]
    nbytes = device.BloomFilter.size_for(gens[0].total_bases, 0.025)
    per = []
    n0 = cuda_ctx.part_inserts
    with env(NTS_BF_PARTITION=0):
        for g in gens:
            bf = cuda_ctx.bloom(nbytes)
            bf.insert_genome(g, K)
            per.append(bf.to_numpy().copy())
            bf.close()
    assert cuda_ctx.part_inserts == n0
    want0 = so.genome_bits(_records(gens[0]), K, nbytes)
    assert np.array_equal(per[0], want0)                      # direct kernel == C oracle
    return gens, nbytes, per


# (6 Mbp genome: m = 2.4e8 bits; a region's byte flags must fit one CTA, so P1MAX * P2 >= ~1100)
PLANS = [dict(NTS_BF_P1MAX=1024, NTS_BF_P2=1024),             # R = 256 bits
         dict(NTS_BF_P1MAX=100, NTS_BF_P2=16),                # R = 148 Kbit: one CTA per SM in the apply pass
         dict(NTS_BF_P1MAX=1000, NTS_BF_P2=4),
         dict(NTS_BF_P1MAX=2, NTS_BF_P2=1024),
         dict(NTS_BF_P1MAX=1024, NTS_BF_P2=2),
         dict(NTS_BF_P1MAX=333, NTS_BF_P2=8)]


@pytest.mark.parametrize("plan", PLANS)
@pytest.mark.parametrize("match", [0, 1])
def test_three_pass_insert_modes_equal_direct_kernel(cuda_ctx, small, plan, match):
    gens, nbytes, per = small
    with env(NTS_BF_PARTITION=1, NTS_BF_IMPL=3, NTS_BF_MATCH=match, **plan):
        n0 = cuda_ctx.part_inserts
        bf = cuda_ctx.bloom(nbytes)
        bf.from_numpy(np.full(nbytes, 0xFF, dtype=np.uint8))
        bf.set_genome(gens[0], K)                              # SET: overwrites whatever was there
        assert np.array_equal(bf.to_numpy(), per[0])
        bf.insert_genome(gens[1], K)                           # OR into an existing filter
        assert np.array_equal(bf.to_numpy(), per[0] | per[1])
        bf.clear(); bf.insert_genome(gens[2], K)
        assert np.array_equal(bf.to_numpy(), per[2])
        lvl = cuda_ctx.bloom(nbytes)
        lvl.from_numpy(np.full(nbytes, 0xAA, dtype=np.uint8))
        for n in (1, 2, 3):                                    # SET, then AND into the other filter, swapping
            bf.build_common(lvl if n > 1 else None, gens[:n], K)
            want = per[0].copy()
            for x in per[1:n]:
                want &= x
            assert np.array_equal(bf.to_numpy(), want), n
        assert cuda_ctx.part_inserts - n0 == 3 + 6
        bf.close(); lvl.close()


@pytest.mark.parametrize("scale", [0.9, 0.5, 0.05])
def test_three_pass_bucket_overflow_goes_through_the_overflow_list(cuda_ctx, small, scale):
    "capacities below the expected load: the surplus of every bucket is applied from the overflow list"
    gens, nbytes, per = small
    with env(NTS_BF_PARTITION=1, NTS_BF_IMPL=3, NTS_BF_P1MAX=64, NTS_BF_P2=64, NTS_BF_CAP_SCALE=scale, NTS_BF_OVF_CAP=50_000_000):
        o0 = cuda_ctx.part_overflow_items
        bf, lvl = cuda_ctx.bloom(nbytes), cuda_ctx.bloom(nbytes)
        bf.build_common(lvl, gens, K)
        assert np.array_equal(bf.to_numpy(), per[0] & per[1] & per[2])
        bf.insert_genome(gens[1], K)
        assert np.array_equal(bf.to_numpy(), per[1])                  # (a & b & c) | b == b
        assert cuda_ctx.part_overflow_items - o0 > (0.05 if scale > 0.6 else 0.4) * gens[0].total_bases
        bf.close(); lvl.close()


def test_three_pass_exhausted_overflow_list_falls_back_to_the_direct_kernel(cuda_ctx, small):
    gens, nbytes, per = small
    with env(NTS_BF_PARTITION=1, NTS_BF_IMPL=3, NTS_BF_P1MAX=64, NTS_BF_P2=64, NTS_BF_CAP_SCALE=0.5, NTS_BF_OVF_CAP=1000):
        bf, lvl = cuda_ctx.bloom(nbytes), cuda_ctx.bloom(nbytes)
        bf.build_common(lvl, gens, K)
        assert np.array_equal(bf.to_numpy(), per[0] & per[1] & per[2])
        bf.set_genome(gens[2], K)
        assert np.array_equal(bf.to_numpy(), per[2])
        bf.insert_genome(gens[0], K)
        assert np.array_equal(bf.to_numpy(), per[2] | per[0])
        bf.close(); lvl.close()


@pytest.mark.parametrize("impl", [2, 1, 3])
def test_heavy_hitter_kmers(cuda_ctx, impl):
    "one k-mer repeated 2 M times (poly-A) plus a tandem repeat: far more copies than any bucket holds"
    rng = np.random.default_rng(5)
    rnd = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 3_000_000).tobytes()
    recs = [("polyA", b"A" * 2_000_000), ("tandem", b"ACGTTGCAAT" * 150_000), ("rnd", rnd)]
    g = cuda_ctx.upload(fasta.pack_records(recs))
    nbytes = device.BloomFilter.size_for(g.total_bases, 0.025)
    want = so.genome_bits(recs, K, nbytes)
    with env(NTS_BF_PARTITION=1, NTS_BF_IMPL=impl, NTS_BF_BIN=impl, NTS_BF_P1MAX=128, NTS_BF_P2=128, NTS_BF_REGION_SHIFT=18):
        bf = cuda_ctx.bloom(nbytes)
        o0, n0 = cuda_ctx.part_overflow_items, cuda_ctx.part_inserts
        bf.set_genome(g, K)
        assert np.array_equal(bf.to_numpy(), want)
        assert cuda_ctx.part_inserts - n0 == 1
        if impl == 3:
            assert cuda_ctx.part_overflow_items - o0 > 1_500_000
        bf.close()


# ---------------------------------------------------------------------------------------- the production pair
@pytest.mark.parametrize("rank", [2, 1])
@pytest.mark.parametrize("shift", [14, 18, 21, 23, 27])
@pytest.mark.parametrize("scale", [1.0, 0.6, 0.05])
def test_pair_many_buckets_short_last_region_and_overflow_branch(cuda_ctx, small, shift, scale, rank):
    """bf_bin_kernel + bf_apply_kernel with 2^shift-bit regions (m = 2.4e8 bits: 1024 buckets at shift 18 -- 14 asks
    for more than the 1024 the kernel supports and is widened --, 29 with a short last region at 23, 2 at 27) and
    with capacities below the load (items past a bucket's capacity are applied with direct atomics).
    rank = 2: bf_rank_bin_kernel (csrc/nts_rank.cuh, ranking by private counters; 5-bit high digit above 512 buckets,
    4-bit below), rank = 1: bf_bin_kernel (shared-memory atomicAdd ranking)"""
    gens, nbytes, per = small
    with env(NTS_BF_PARTITION=1, NTS_BF_REGION_SHIFT=shift, NTS_BF_CAP_SCALE=scale, NTS_BF_BIN=rank):
        n0 = cuda_ctx.part_inserts
        bf, lvl = cuda_ctx.bloom(nbytes), cuda_ctx.bloom(nbytes)
        bf.from_numpy(np.full(nbytes, 0xFF, dtype=np.uint8))
        bf.set_genome(gens[0], K)
        assert np.array_equal(bf.to_numpy(), per[0])
        bf.insert_genome(gens[1], K)
        assert np.array_equal(bf.to_numpy(), per[0] | per[1])
        lvl.from_numpy(np.full(nbytes, 0x55, dtype=np.uint8))
        for n in (1, 2, 3):
            bf.build_common(lvl if n > 1 else None, gens[:n], K)
            want = per[0].copy()
            for x in per[1:n]:
                want &= x
            assert np.array_equal(bf.to_numpy(), want), n
        assert cuda_ctx.part_inserts - n0 == 2 + 6
        bf.close(); lvl.close()


def test_default_plan_at_150_mbp_equals_direct_kernel_and_oracle(cuda_ctx):
    """what a real genome gets (no knobs) at a filter of 5.9e9 bits (> 2^32: 23 regions of 32 MB): the pair's SET and
    AND against the direct kernel on the device, genome 0 against the C oracle on the host; then the same with 706
    regions and shrunk capacities, and with the three-pass variant"""
    wl = synth.Workload(2, 150_000_000, 1.0)
    gens = [wl.materialize(cuda_ctx, g) for g in range(2)]
    nbytes = device.BloomFilter.size_for(gens[0].total_bases, 0.025)
    assert nbytes * 8 > 1 << 32
    n0 = cuda_ctx.part_inserts
    common, lvl = cuda_ctx.bloom(nbytes), cuda_ctx.bloom(nbytes)
    common.build_common(lvl, gens, K)
    assert cuda_ctx.part_inserts - n0 == 2                    # both inserts took the partitioned path
    first = cuda_ctx.bloom(nbytes)
    first.set_genome(gens[0], K)
    with env(NTS_BF_PARTITION=0):
        d0, d1 = cuda_ctx.bloom(nbytes), cuda_ctx.bloom(nbytes)
        d0.insert_genome(gens[0], K); d1.insert_genome(gens[1], K)
    a = first.to_numpy()
    assert np.array_equal(a, d0.to_numpy())
    d0.iand(d1)
    assert np.array_equal(common.to_numpy(), d0.to_numpy())
    # C oracle (OpenMP over records; a few seconds)
    recs = _records(gens[0])
    assert np.array_equal(a, so.genome_bits(recs, K, nbytes))
    want = d0.to_numpy()
    for knobs in (dict(NTS_BF_REGION_SHIFT=23), dict(NTS_BF_REGION_SHIFT=23, NTS_BF_CAP_SCALE=0.7), dict(NTS_BF_IMPL=3),
                  dict(NTS_BF_BIN=1), dict(NTS_BF_BIN=1, NTS_BF_REGION_SHIFT=24)):
        with env(**knobs):
            common.build_common(lvl, gens, K)
        assert np.array_equal(common.to_numpy(), want), knobs
    for x in (common, lvl, first, d0, d1):
        x.close()
