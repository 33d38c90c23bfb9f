"""where the end-to-end minus resident gap goes: H2D rate, async overlap with the Bloom build"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from ntsynt_b200 import device, synth
ctx = device.Context(0)
wl = synth.Workload(2, 3_000_000_000, 1.0)
gens = [wl.materialize(ctx, g) for g in range(2)]
packed = []
for g in gens:
    pk = g.to_packed()
    pin = device.PinnedU64(len(pk.words)); pin.array[:] = pk.words; pk.words = pin.array
    packed.append((pk, pin))
nbytes = device.BloomFilter.size_for(gens[0].total_bases, 0.025)
common, level = ctx.bloom(nbytes), ctx.bloom(nbytes)
def t(fn, n=3):
    best = 1e9
    for _ in range(n):
        ctx.sync(); t0 = time.perf_counter(); fn(); ctx.sync(); best = min(best, time.perf_counter() - t0)
    return best * 1e3
def up_sync():
    for pk, _ in packed:
        ctx.upload(pk).close()
def up_async():
    gs = [ctx.upload(pk, async_copy=True) for pk, _ in packed]
    for g in gs: g.close()
print("upload sync  2 genomes ms", t(up_sync))
print("upload async 2 genomes ms", t(up_async))
print("build resident ms", t(lambda: common.build_common(level, gens, 24)))
def fresh(async_copy):
    gs = [ctx.upload(pk, async_copy=async_copy) for pk, _ in packed]
    t0 = time.perf_counter()
    common.build_common(level, gs, 24)
    for g in gs: g.close()
print("upload sync + build ms", t(lambda: fresh(False)))
print("upload async + build ms", t(lambda: fresh(True)))
