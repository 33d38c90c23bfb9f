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
