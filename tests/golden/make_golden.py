#!/usr/bin/env python3
"""Generate the committed golden fixtures under tests/golden/ -- run in the BUILD CONTAINER only
(needs /root/reference).  Two kinds of fixture:

1. hash_kats.json: (k-mer, k, h1) known answers sampled from the reference's own golden indexlr
   outputs tests/expected_result/*.k{24,20}.w1000.tsv (real btllib output) -- pins ntHash2 + h1.
2. mini/: three small synthetic genomes (tests/synth_small.py, fixed seeds) and the outputs of the
   REFERENCE's own graph stage (bin/ntsynt_run.py run unmodified under oracle/shims, sketches made
   by the golden-pinned oracle): sketch TSVs, .mx.dot edge list, pre-merge and final block files,
   for a 2-genome and a 3-genome run.
"""
import gzip
import json
import os
import re
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth_small  # noqa: E402
from oracle import ref_harness  # noqa: E402

REF = ref_harness.REF
MINI = os.path.join(HERE, "mini")
MINI_PARAMS = dict(k=24, w=100, w_rounds=(40, 10), indel=500, merge="300", block_size=200)


def hash_kats():
    out = []
    exp = os.path.join(REF, "tests", "expected_result")
    for f in sorted(os.listdir(exp)):
        m = re.search(r"\.k(\d+)\.w\d+\.tsv$", f)
        if not m:
            continue
        k = int(m.group(1))
        n = 0
        for line in open(os.path.join(exp, f), encoding="utf-8"):
            _, rest = line.rstrip("\n").split("\t")
            for tok in rest.split(" "):
                n += 1
                if n % 150 == 0:
                    h1, _, seq = tok.split(":")
                    out.append([seq, k, h1])
    with open(os.path.join(HERE, "hash_kats.json"), "w", encoding="utf-8") as fh:
        json.dump({"source": "bcgsc/ntSynt v1.0.4 tests/expected_result/*.k*.w1000.tsv (every 150th triple)",
                   "kats": out}, fh)
    return len(out)


def dot_edges(path):
    "sorted [u, v, weight] list from a .mx.dot"
    edges = []
    for line in open(path, encoding="utf-8"):
        m = re.match(r'^"(\d+)" --"(\d+)" \[weight=(\d+) ', line)
        if m:
            u, v = sorted((m.group(1), m.group(2)))
            edges.append([u, v, int(m.group(3))])
    return sorted(edges)


def mini():
    os.makedirs(MINI, exist_ok=True)
    gens = synth_small.make_genomes(20261017, 3, contig_lens=(150000, 100000, 60000), sub=0.01, n_inv=3, n_trans=2,
                                    n_dup=2, n_nruns=3, lowercase=True)
    names = ["miniA.fa", "miniB.fa", "miniC.fa"]
    tmp = tempfile.mkdtemp(prefix="mkgold_")
    for n, recs in zip(names, gens):
        synth_small.write_fasta(os.path.join(tmp, n), recs)
        with open(os.path.join(tmp, n), "rb") as fin, gzip.GzipFile(os.path.join(MINI, n + ".gz"), "wb", mtime=0) as fout:
            shutil.copyfileobj(fin, fout)
    for tag, sel in (("AB", names[:2]), ("ABC", names)):
        wd = os.path.join(tmp, tag)
        res = ref_harness.run_reference([os.path.join(tmp, n) for n in sel], wd, f"mini-{tag}", **MINI_PARAMS)
        assert res["returncode"] == 0, res["log"][-2000:]
        out = os.path.join(MINI, tag)
        os.makedirs(out, exist_ok=True)
        shutil.copyfile(res["blocks"], os.path.join(out, "synteny_blocks.tsv"))
        shutil.copyfile(res["pre_merge"], os.path.join(out, "pre-collinear-merge.synteny_blocks.tsv"))
        for t in res["tsvs"]:
            with open(t, "rb") as fin, gzip.GzipFile(os.path.join(out, os.path.basename(t) + ".gz"), "wb", mtime=0) as fo:
                shutil.copyfileobj(fin, fo)
        with gzip.GzipFile(os.path.join(out, "mx_dot_edges.json.gz"), "wb", mtime=0) as fo:
            fo.write(json.dumps(dot_edges(res["dot"])).encode())
        for n in sel:
            shutil.copyfile(os.path.join(wd, n + ".fai"), os.path.join(out, n + ".fai"))
    with open(os.path.join(MINI, "params.json"), "w", encoding="utf-8") as fh:
        json.dump({**MINI_PARAMS, "w_rounds": list(MINI_PARAMS["w_rounds"]), "fpr": 0.025,
                   "made_by": "tests/golden/make_golden.py: reference bin/ntsynt_run.py under oracle/shims"}, fh)
    shutil.rmtree(tmp)


def presets():
    """tests/golden/presets/: the d >= 1 presets of bin/ntSynt:89-99 (G = 3 with -d 1.3, G = 5 with -d 12) on seeded
    genomes; block files written by the reference's own bin/ntsynt_run.py under oracle/shims"""
    import preset_cases as pc
    os.makedirs(pc.PRESET_DIR, exist_ok=True)
    meta = {}
    for tag, case in pc.CASES.items():
        gens = pc.genomes(tag)
        tmp = tempfile.mkdtemp(prefix="mkpre_")
        for n, recs in zip(pc.names(tag), gens):
            synth_small.write_fasta(os.path.join(tmp, n), recs)
        p = case["params"]
        res = ref_harness.run_reference([os.path.join(tmp, n) for n in pc.names(tag)], os.path.join(tmp, "wd"), tag,
                                        k=p["k"], w=p["w"], w_rounds=p["w_rounds"], indel=p["indel"], merge=p["merge"],
                                        block_size=p["block_size"])
        assert res["returncode"] == 0, res["log"][-2000:]
        out = os.path.join(pc.PRESET_DIR, tag)
        os.makedirs(out, exist_ok=True)
        shutil.copyfile(res["blocks"], os.path.join(out, "synteny_blocks.tsv"))
        shutil.copyfile(res["pre_merge"], os.path.join(out, "pre-collinear-merge.synteny_blocks.tsv"))
        meta[tag] = {**{k: v for k, v in case.items() if k != "lens"}, "lens": list(case["lens"]),
                     "genomes_sha1": pc.digest(gens),
                     "made_by": "tests/golden/make_golden.py presets(): reference bin/ntsynt_run.py under oracle/shims"}
        shutil.rmtree(tmp)
    with open(os.path.join(pc.PRESET_DIR, "params.json"), "w", encoding="utf-8") as fh:
        json.dump(meta, fh, indent=1)


if __name__ == "__main__":
    print("hash KATs:", hash_kats())
    mini()
    print("mini fixtures written to", MINI)
    presets()
    print("preset fixtures written")
