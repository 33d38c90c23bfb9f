#!/usr/bin/env python3
"""Randomised comparison of the three statements of the graph stage on small seeded genomes (CPU only):

    python scripts/fuzz_graph_stage.py engine  <first seed> <last seed>    product engine (device-resident, lean and dense
                                                                           forms, fed by the test backend) vs oracle/graph_oracle.py
    python scripts/fuzz_graph_stage.py oracle  <first seed> <last seed>    oracle/graph_oracle.py vs the reference's OWN
                                                                           bin/ntsynt_run.py under oracle/ref_harness.py (build container only)
    python scripts/fuzz_graph_stage.py oracle-n <first seed> <last seed>   the same with -n (minimum edge weight) drawn below the
                                                                           number of assemblies: the branch-resolution loop of
                                                                           ntjoin.py:68-76,114-123 (restated in the oracle only)
    python scripts/fuzz_graph_stage.py direct  <first seed> <last seed>    product engine vs the reference's own code, with
                                                                           --filter Filter | Indexlr | none drawn as well

Every seed draws the number of genomes (2-6), contigs, divergence, indels, inversions, translocations, duplications, N
runs, soft-masking, k, w, the refinement rounds, --indel, --collinear-merge, -z and --simplify-graph.  TEST INFRASTRUCTURE:
nothing here is on the product path."""
import os
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth_small  # noqa: E402
from oracle import sketch_oracle as so  # noqa: E402
from oracle.graph_oracle import GraphOracle  # noqa: E402


def draw(seed):
    "the case of one seed: FASTA records per genome and the run's parameters"
    r = np.random.default_rng(seed)
    G = int(r.integers(2, 7))
    contig_lens = tuple(int(x) for x in r.integers(30000, 160000, int(r.integers(1, 5))))
    gens = synth_small.make_genomes(seed, G, contig_lens=contig_lens, sub=float(r.choice([0.001, 0.004, 0.01, 0.02])),
                                    indel=float(r.choice([0.0, 0.0005, 0.002])), n_inv=int(r.integers(0, 9)),
                                    n_trans=int(r.integers(0, 6)), n_dup=int(r.integers(0, 8)), n_nruns=int(r.integers(0, 6)),
                                    lowercase=bool(r.integers(0, 2)))
    k = int(r.choice([12, 16, 20, 24]))
    w = int(r.choice([20, 40, 80]))
    w_rounds = [x for x in [[], [10], [20, 5], [30, 15, 6]][int(r.integers(0, 4))] if x < w]
    par = dict(k=k, w=w, w_rounds=w_rounds, indel=int(r.choice([100, 300, 2000, 50000])),
               merge=str(r.choice(["400", "2w", "0", "100000"])), z=int(r.choice([0, 50, 200, 1000])), simplify=bool(r.integers(0, 4)))
    return gens, par


def write_case(gens, tmp):
    paths = []
    for i, recs in enumerate(gens):
        p = os.path.join(tmp, f"g{chr(65 + i)}.fa")
        synth_small.write_fasta(p, recs)
        paths.append(p)
    return paths


def run_oracle(paths, par, bits, n=0):
    go = GraphOracle([(os.path.basename(p) + f".k{par['k']}.w{par['w']}.tsv", so.read_fasta(p)) for p in paths], par["k"], par["w"],
                     par["w_rounds"], par["indel"], par["merge"], par["z"], bits, simplify=par["simplify"], n=n)
    try:
        go.run()
    except SystemExit:                 # "no paths found"
        return {}
    return go.outputs


def engine_outputs(paths, par, forms=("dev", True, False), repeat_bits=None, filter_mode=None):
    "the product engine in its three vertex-storage forms; returns ([outputs per form], common filter bits)"
    from backends import OracleBackend
    from ntsynt_b200.synteny import SyntenyEngine
    k, w = par["k"], par["w"]
    tsv = [f"{os.path.basename(p)}.k{k}.w{w}.tsv" for p in paths]
    order = sorted(range(len(paths)), key=lambda i: tsv[i], reverse=True)
    outs = []
    for lean in forms:
        be = OracleBackend([paths[i] for i in order], [tsv[i] for i in order], k, lean=lean, repeat_bits=repeat_bits,
                           filter_mode=filter_mode)
        eng = SyntenyEngine(be, k, w, par["w_rounds"], par["indel"], par["merge"], par["z"], write_files=False, quiet=True,
                            simplify=par["simplify"])
        try:
            eng.run()
        except SystemExit:
            eng.outputs = {}
        outs.append(eng.outputs)
    return outs, be.bits


def reference_outputs(paths, par, tmp, filter_mode=None, n=0):
    from oracle import ref_harness
    res = ref_harness.run_reference(paths, os.path.join(tmp, "wd"), "fz", k=par["k"], w=par["w"], w_rounds=par["w_rounds"],
                                    indel=par["indel"], merge=par["merge"], block_size=par["z"], simplify=par["simplify"],
                                    filter_mode=filter_mode, repeat_fpr=0.1 if filter_mode else None, n=n)
    if res["returncode"] != 0 and "no paths found" not in res["log"]:
        raise RuntimeError("the reference run failed:\n" + res["log"][-2000:])
    out = {key: (open(f).read() if os.path.exists(f) else None) for key, f in (("final", res["blocks"]), ("pre_merge", res["pre_merge"]))}
    out["repeat_bits"] = res["repeat_bits"]
    return out


def check_seed(mode, seed):
    "None if the seed does not apply to the mode, else True / False"
    gens, par = draw(seed)
    keys = ("final", "pre_merge") if par["w_rounds"] else ("initial",)
    tmp = tempfile.mkdtemp()
    try:
        paths = write_case(gens, tmp)
        if mode == "engine":
            outs, bits = engine_outputs(paths, par)
            want = run_oracle(paths, par, bits)
            return all(o.get(kk) == want.get(kk) for o in outs for kk in keys)
        if not par["w_rounds"]:
            return None                # bin/ntsynt_run.py always runs refinement rounds
        if mode == "direct":
            filter_mode = [None, "Filter", "Indexlr"][seed % 3]
            ref = reference_outputs(paths, par, tmp, filter_mode)
            outs, _ = engine_outputs(paths, par, forms=(["dev", False][seed % 2],), repeat_bits=ref["repeat_bits"],
                                     filter_mode=filter_mode)
            return all(outs[0].get(kk) == ref[kk] for kk in keys)
        n = 0
        if mode == "oracle-n":
            if len(gens) < 3:
                return None
            n = 1 + seed % (len(gens) - 1)
        ref = reference_outputs(paths, par, tmp, n=n)
        bits = so.common_bf([(os.path.basename(p), so.read_fasta(p)) for p in paths], par["k"], 0.025)
        want = run_oracle(paths, par, bits, n=n)
        return all(ref[kk] == want.get(kk) for kk in keys)
    finally:
        shutil.rmtree(tmp)


def main():
    mode, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    assert mode in ("engine", "oracle", "oracle-n", "direct")
    t0, bad, n = time.time(), 0, 0
    for seed in range(lo, hi):
        ok = check_seed(mode, seed)
        if ok is None:
            continue
        n += 1
        if not ok:
            bad += 1
            print("MISMATCH seed", seed, draw(seed)[1], flush=True)
    print(f"{mode}: {n} cases, {bad} mismatches, {time.time() - t0:.0f} s")
    return 1 if bad else 0


if __name__ == "__main__":
    raise SystemExit(main())
