"""Multi-GPU parity (needs >= 2 CUDA devices; skipped otherwise): the NCCL counter merge must give
exactly the single-GPU AND, and the one-genome-per-GPU run must print the single-GPU block table."""
import os
import subprocess
import sys
import textwrap

import pytest

from ntsynt_b200 import device

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

WORKER = textwrap.dedent("""
    import os, sys, hashlib
    import numpy as np
    sys.path.insert(0, {root!r})
    import bench
    from ntsynt_b200 import device, distributed, pipeline, synth
    d = bench.Dist()
    ctx = device.Context(d.local_rank)
    wl = synth.Workload(2, 20_000_000, 1.0, n_contigs=6)
    gens = [wl.materialize(ctx, g) for g in range(2)]
    k = 24
    nbytes = device.BloomFilter.size_for(gens[0].total_bases, 0.025)
    ref = pipeline.build_common_bf(ctx, gens, [wl.file_name(g) for g in range(2)], k, nbytes=nbytes)
    mine = ctx.bloom(nbytes)
    mine.insert_genome(gens[d.rank], k)
    ident = distributed.Comm.new_unique_id() if d.rank == 0 else b""
    comm = distributed.Comm(ctx, d.rank, d.world, d.bcast_bytes(ident, 128))
    both = ctx.bloom(nbytes).from_numpy(mine.to_numpy())
    comm.allreduce_and(mine)
    assert np.array_equal(mine.to_numpy(), ref.to_numpy()), "NCCL counter merge != AND"
    comm.allreduce_or(both)
    a = ctx.bloom(nbytes); a.insert_genome(gens[0], k)
    b = ctx.bloom(nbytes); b.insert_genome(gens[1], k)
    a.ior(b)
    assert np.array_equal(both.to_numpy(), a.to_numpy()), "NCCL counter merge != OR"
    # peer-memory merge (CUDA IPC + NVLink loads) gives the same bits as the NCCL counter merge
    pm_bf = ctx.bloom(nbytes)
    pm_bf.insert_genome(gens[d.rank], k)
    pm = distributed.PeerMerge(pm_bf, d.rank, d.world, d.gather_objects, d.barrier)
    pm.merge("and")
    assert np.array_equal(pm_bf.to_numpy(), ref.to_numpy()), "peer-memory merge != AND"
    pm_bf.clear(); pm_bf.insert_genome(gens[d.rank], k)
    pm.merge("or")
    assert np.array_equal(pm_bf.to_numpy(), a.to_numpy()), "peer-memory merge != OR"
    # the same merge ordered by stream barriers (one-word NCCL all-reduces) instead of host barriers,
    # back to back without any host synchronisation in between
    for _ in range(3):
        pm_bf.clear(); pm_bf.insert_genome(gens[d.rank], k)
        pm.merge("and", comm=comm)
    assert np.array_equal(pm_bf.to_numpy(), ref.to_numpy()), "peer-memory merge (stream barriers) != AND"
    pm.close()
    # one genome per GPU WITHOUT per-GPU filters: every rank bins its genome, the owner of a slice applies every rank's
    # buckets (each source is a genome of its own: AND across the sources)
    ob_c, ob_l = ctx.bloom(nbytes), ctx.bloom(nbytes)
    os.environ["NTS_BF_REGION_SHIFT"] = "22"
    ob = distributed.OwnedBuild(ob_c, ob_l, d.rank, d.world, k, int(d.max(int(gens[d.rank].total_bases))), d.gather_objects,
                                d.barrier, comm)
    del os.environ["NTS_BF_REGION_SHIFT"]
    assert ob.ok
    for rep in range(2):
        ob_c.from_numpy(np.full(nbytes, 0x77, dtype=np.uint8))
        assert ob.build([gens[d.rank]], genomes_are_sources=True) == 0
        assert np.array_equal(ob_c.to_numpy(), ref.to_numpy()), "owned build (one genome per GPU) != AND"
    ob.close()
    t = ctx.sketch(gens[d.rank], k, 1000, common=mine)
    counts = d.gather_objects(len(t))
    tabs = comm.allgather_tables(t, counts, gens)
    for r in range(2):
        want = ctx.sketch(gens[r], k, 1000, common=ref).to_numpy()
        got = tabs[r].to_numpy()
        assert all(np.array_equal(x, y) for x, y in zip(want, got)), "all-gathered table differs"
    comm.close(); d.barrier(); d.close()
    print("rank", d.rank, "ok")
""")


def test_nccl_merge_and_allgather(tmp_path):
    if device.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29563", str(script)],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:]
    assert res.stdout.count("ok") >= 2


SHARD_WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    sys.path.insert(0, {root!r})
    import bench
    from ntsynt_b200 import device, distributed, pipeline, synth
    d = bench.Dist()
    ctx = device.Context(d.local_rank)
    G, k, w = 3, 24, 1000
    wl = synth.Workload(G, 20_000_000, 1.3, n_contigs=6)
    whole = [wl.materialize(ctx, g) for g in range(G)]
    nbytes = device.BloomFilter.size_for(whole[0].total_bases, 0.025)
    ref = pipeline.build_common_bf(ctx, whole, [wl.file_name(g) for g in range(G)], k, nbytes=nbytes)
    own = distributed.assign_contigs(wl.anc_lengths, d.world)
    assert sorted(c for b in own for c in b) == list(range(6))
    owner = [next(r for r in range(d.world) if c in own[r]) for c in range(6)]
    shards = [wl.materialize(ctx, g, contigs=own[d.rank]) for g in range(G)]
    for g in range(G):                      # a shard holds exactly its contigs, bit for bit, and nothing else
        for c in range(6):
            if c in own[d.rank]:
                assert np.array_equal(shards[g].contig_words(c), whole[g].contig_words(c))
            else:
                assert int(shards[g].lengths[c]) == 0
    parts = [ctx.bloom(nbytes) for _ in range(G)]
    common = ctx.bloom(nbytes)
    common.from_numpy(np.full(nbytes, 0x5A, dtype=np.uint8))
    ident = distributed.Comm.new_unique_id() if d.rank == 0 else b""
    comm = distributed.Comm(ctx, d.rank, d.world, d.bcast_bytes(ident, 128))
    sm = distributed.ShardedMerge(parts, common, d.rank, d.world, d.gather_objects, d.barrier)
    assert sm.ok
    for rep in range(2):
        for g in range(G):
            parts[g].set_genome(shards[g], k)
        sm.merge(comm=comm if rep else None)
        assert np.array_equal(common.to_numpy(), ref.to_numpy()), "AND_g OR_rank over peer memory != single-GPU filter"
    sm.close()
    # hash-range owned build: every rank only bins; the owner of a slice applies every rank's buckets over peer memory
    lvl = ctx.bloom(nbytes)
    plan = int(d.max(max(int(x.total_bases) for x in shards)))
    os.environ["NTS_BF_REGION_SHIFT"] = "22"          # 2^22-bit regions: 40 of them, slices cut through regions
    ob = distributed.OwnedBuild(common, lvl, d.rank, d.world, k, plan, d.gather_objects, d.barrier, comm)
    del os.environ["NTS_BF_REGION_SHIFT"]
    assert ob.ok
    for rep in range(2):
        common.from_numpy(np.full(nbytes, 0xA5, dtype=np.uint8))
        lvl.from_numpy(np.full(nbytes, 0x3C, dtype=np.uint8))
        over = ob.build(shards)
        assert over == 0
        assert np.array_equal(common.to_numpy(), ref.to_numpy()), "owned build != single-GPU filter"
    ob.close()
    # the north-star form: one counter all-reduce per genome, AND locally
    for g in range(G):
        parts[g].set_genome(shards[g], k)
        comm.allreduce_or(parts[g])
    nccl = ctx.bloom(nbytes).build_from_and(parts)
    assert np.array_equal(nccl.to_numpy(), ref.to_numpy()), "NCCL OR per genome + AND != single-GPU filter"
    for g in range(G):
        t = ctx.sketch(shards[g], k, w, common=common)
        full = distributed.gather_sharded_table(comm, t, 6, owner, d.gather_objects)
        want = ctx.sketch(whole[g], k, w, common=ref).to_numpy()
        assert all(np.array_equal(x, y) for x, y in zip(want, full.to_numpy())), "contig-sharded sketch differs"
    comm.close(); d.barrier(); d.close()
    print("rank", d.rank, "ok")
""")


def _torchrun(script, port, n=2, timeout=900):
    return subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
                           "--master-addr", "127.0.0.1", "--master-port", str(port), *script],
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout)


def test_contig_sharded_merge_and_sketch(tmp_path):
    "SURVEY 8e P2: contigs of every genome spread over the ranks; AND over genomes of OR over ranks; sketch by owner"
    if device.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(SHARD_WORKER.format(root=ROOT))
    res = _torchrun([str(script)], 29571)
    assert res.returncode == 0, (res.stdout + res.stderr)[-3000:]
    assert res.stdout.count("ok") >= 2


@pytest.mark.parametrize("flags", [["--genomes", "3", "--divergence", "1.3"], ["--genomes", "2", "--divergence", "1"]])
def test_bench_block_table_is_the_same_on_one_and_two_gpus(flags):
    "bench.py prints a sha1 of the final block table: contig-sharded (G = 3) and one-genome-per-GPU (G = 2) runs == 1 GPU"
    if device.device_count() < 2:
        pytest.skip("needs two GPUs")
    import json
    common = [os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--no-cpu", "--genome-mbp", "40", *flags]
    one = subprocess.run([sys.executable, *common, "--gpus", "1"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                         timeout=900)
    assert one.returncode == 0, one.stderr[-3000:]
    two = _torchrun([*common, "--gpus", "2"], 29573)
    assert two.returncode == 0, two.stderr[-3000:]
    a = json.loads(one.stdout.strip().splitlines()[-1])
    b = json.loads(two.stdout.strip().splitlines()[-1])
    assert a["config"]["blocks"] > 0
    assert a["config"]["blocks_sha1"] == b["config"]["blocks_sha1"]
    assert a["config"]["vertices"] == b["config"]["vertices"]
