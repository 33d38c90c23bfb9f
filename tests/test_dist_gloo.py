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
    # ShardedSketcher protocol and gather_sharded_table's contig re-ordering, with stand-ins for the device tables:
    # rank 0 asks for two sketches, rank 1 serves them, both see the whole genome's rows in contig order
    from ntsynt_b200 import device
    n_contigs, owner = 5, [0, 1, 1, 0, 1]
    class Tab:
        def __init__(self, rows): self.rows = rows                    # list of (contig, value)
        def __len__(self): return len(self.rows)
        def contig_offsets(self, n):
            cnt = np.bincount([c for c, _ in self.rows], minlength=n)
            return np.concatenate([[0], np.cumsum(cnt)])
        def close(self): pass
    class Comm:
        world, ctx = d.world, None
        def allgather_tables(self, t, sizes, _):
            got = d.gather_objects(t.rows)
            assert [len(x) for x in got] == sizes
            return [Tab(x) for x in got]
    class Ctx:
        def sketch(self, shard, k, w, common=None, masks=None):
            # rows of my contigs only; the value encodes (genome, w, mask) so a mixed-up command would show
            return Tab([(c, (shard, w, int(masks[c]), i)) for c in range(n_contigs) if owner[c] == d.rank for i in range(c + 1)])
    def concat(ctx, parts, src, cnt, genome):
        return Tab([row for p_, s_, n_ in zip(parts, src, cnt) for row in p_.rows[s_:s_ + n_]])
    device.MinimizerTable.concat = staticmethod(concat)
    seen = []
    svc = distributed.ShardedSketcher(Comm(), Ctx(), ["g0", "g1"], 24, None, n_contigs, owner, d.bcast_object, d.gather_objects)
    real_do = svc._do
    def spy(g, w, masks):
        out = real_do(g, w, masks); seen.append((g, w, out.rows)); return out
    svc._do = spy
    if d.rank == 0:
        m = np.arange(10, 15)
        a = svc.sketch(1, 250, m)
        b = svc.sketch(0, 100, m + 5)
        svc.done()
    else:
        svc.serve()
    assert [(g, w) for g, w, _ in seen] == [(1, 250), (0, 100)]
    for (g, w, rows), base in zip(seen, (10, 15)):
        assert rows == [(c, ("g%d" % g, w, base + c, i)) for c in range(n_contigs) for i in range(c + 1)]
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
