"""world_size-2 gloo test (CPU) of the multi-GPU host logic: genome ownership, the 128-byte id
broadcast plumbing of bench.py, and the counter-merge arithmetic (expand -> all-reduce(sum) ->
threshold) that nts_bf_allreduce_and performs on the device, restated with numpy here."""
import os
import subprocess
import sys
import textwrap

import numpy as np

from ntsynt_b200 import distributed

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ownership_plan():
    assert distributed.assign_genomes(8, 8) == [[g] for g in range(8)]
    assert distributed.assign_genomes(5, 2) == [[0, 2, 4], [1, 3]]
    assert all(distributed.owner_of(g, 4) == g % 4 for g in range(9))
    assert [distributed.field_bits(w) for w in (2, 3, 4, 8, 15, 16)] == [2, 2, 4, 4, 4, 8]
    assert distributed.merge_wire_bytes(100, 8) == 400


WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    sys.path.insert(0, {root!r})
    import bench
    from ntsynt_b200 import distributed
    d = bench.Dist()
    assert d.world == 2
    # id broadcast: rank 0's bytes reach rank 1
    ident = bytes(range(128)) if d.rank == 0 else b""
    got = d.bcast_bytes(ident, 128)
    assert got == bytes(range(128))
    # counter merge == AND / OR of the per-rank bit arrays
    import torch, torch.distributed as td
    rng = np.random.default_rng(100 + d.rank)
    bits = rng.integers(0, 256, 4096, dtype=np.uint8)
    field = distributed.field_bits(d.world)
    unpacked = np.unpackbits(bits, bitorder="little").astype(np.int32)          # one counter per bit
    t = torch.from_numpy(unpacked.copy())
    td.all_reduce(t, op=td.ReduceOp.SUM)
    cnt = t.numpy()
    assert cnt.max() < (1 << field)
    and_bits = np.packbits((cnt == d.world).astype(np.uint8), bitorder="little")
    or_bits = np.packbits((cnt > 0).astype(np.uint8), bitorder="little")
    all_bits = d.gather_objects(bits.tobytes())
    a, b = (np.frombuffer(x, dtype=np.uint8) for x in all_bits)
    assert np.array_equal(and_bits, a & b) and np.array_equal(or_bits, a | b)
    assert d.max(d.rank + 1.0) == 2.0 and d.sum(1.0) == 2.0
    # rank 0's command reaches rank 1 (the side channel of distributed.ShardedSketcher)
    cmd = d.bcast_object(("sketch", 2, 250, [np.arange(3)]) if d.rank == 0 else None)
    assert cmd[0] == "sketch" and cmd[1] == 2 and cmd[2] == 250 and cmd[3][0].tolist() == [0, 1, 2]
    d.barrier(); d.close()
    print("rank", d.rank, "ok")
""")


def test_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29561")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29561", str(script)],
                         env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:]
    assert res.stdout.count("ok") >= 2


def test_assign_contigs_covers_everything_and_balances():
    "contig-sharded ownership (SURVEY 8e P2): every contig has exactly one owner; human-like sizes balance within ~15 %"
    from ntsynt_b200 import distributed, synth
    lens = synth.ancestor_layout(3_000_000_000)
    for world in (1, 2, 3, 4, 8):
        own = distributed.assign_contigs(lens, world)
        assert sorted(c for b in own for c in b) == list(range(len(lens)))
        load = [sum(int(lens[c]) for c in b) for b in own]
        assert max(load) <= 1.15 * sum(load) / world
    assert distributed.assign_contigs([5, 1, 1], 8)[0] == [0]          # more ranks than contigs: some own nothing
