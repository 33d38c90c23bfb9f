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


FILTER_FPR = 0.2      # a small repeat filter (many false positives) so that both modes change the result visibly


def filters():
    """tests/golden/mini/ABC_{filter,indexlr}/: the reference's graph stage with --filter Filter / --filter Indexlr
    (bin/ntsynt_synteny.py:172-187,601-609) and --interarrivals (:557-564) on the mini genomes; the repeat filter is
    bin/ntsynt_make_repeat_bfs.py's (oracle restatement).  interarrivals are stored sorted: the reference's order
    follows a Python set of strings."""
    names = ["miniA.fa", "miniB.fa", "miniC.fa"]
    tmp = tempfile.mkdtemp(prefix="mkfilt_")
    for n in names:
        with gzip.open(os.path.join(MINI, n + ".gz"), "rb") as fin, open(os.path.join(tmp, n), "wb") as fout:
            shutil.copyfileobj(fin, fout)
    for tag, mode in (("ABC_filter", "Filter"), ("ABC_indexlr", "Indexlr"), ("ABC", None)):
        wd = os.path.join(tmp, tag)
        res = ref_harness.run_reference([os.path.join(tmp, n) for n in names], wd, f"mini-{tag}", filter_mode=mode,
                                        repeat_fpr=FILTER_FPR, interarrivals=True, **MINI_PARAMS)
        assert res["returncode"] == 0, res["log"][-2000:]
        out = os.path.join(MINI, tag)
        os.makedirs(out, exist_ok=True)
        vals = sorted(int(x) for x in open(res["interarrivals"], encoding="utf-8").read().split())
        with gzip.GzipFile(os.path.join(out, "interarrivals.sorted.txt.gz"), "wb", mtime=0) as fo:
            fo.write(("\n".join(map(str, vals)) + "\n").encode())
        if mode is None:
            assert open(res["blocks"]).read() == open(os.path.join(out, "synteny_blocks.tsv")).read()
            continue
        shutil.copyfile(res["blocks"], os.path.join(out, "synteny_blocks.tsv"))
        shutil.copyfile(res["pre_merge"], os.path.join(out, "pre-collinear-merge.synteny_blocks.tsv"))
        with gzip.GzipFile(os.path.join(MINI, "repeat_bits.bin.gz"), "wb", mtime=0) as fo:      # same filter in both modes
            fo.write(res["repeat_bits"].tobytes())
    shutil.rmtree(tmp)


def overlap_cases():
    """tests/golden/overlap_cases.json: seeded block tables and the warnings the reference's own
    NtSyntSynteny.check_non_overlapping (bin/ntsynt_synteny.py:234-253) prints for them"""
    import contextlib
    import io
    import random
    import types
    sys.path[:0] = [os.path.join(ROOT, "oracle", "shims"), os.path.join(REF, "bin"), os.path.join(REF, "subprojects", "ntJoin", "bin")]
    import ntsynt_synteny as ref

    class AB:
        def __init__(self, c, s, e):
            self.c, self.s, self.e = c, s, e

        def get_block_contig_start_end(self):
            return self.c, self.s, self.e

        def get_block_length(self):
            return self.e - self.s
    cases = []
    for seed in range(6):
        rng = random.Random(seed)
        G, z = rng.choice([2, 3]), rng.choice([50, 200])
        blocks = []
        for _ in range(rng.randrange(5, 40)):
            row = []
            for a in range(G):
                s = rng.randrange(0, 3000)
                row.append([f"ctg{rng.randrange(2)}", s, s + rng.randrange(10, 900)])
            blocks.append(row)
        objs = [types.SimpleNamespace(assembly_blocks={f"asm{a}": AB(*row[a]) for a in range(G)}) for row in blocks]
        fake = types.SimpleNamespace(args=types.SimpleNamespace(z=z))
        fake.get_overlapping_region = lambda s, e, iv, _f=fake: ref.NtSyntSynteny.get_overlapping_region(_f, s, e, iv)
        err = io.StringIO()
        with contextlib.redirect_stderr(err):
            ref.NtSyntSynteny.check_non_overlapping(fake, objs)
        warn = [ln.split("block: ", 1)[1].split() for ln in err.getvalue().splitlines() if ln.startswith("WARNING")]
        cases.append({"G": G, "z": z, "blocks": blocks, "warnings": [[w[0], w[1], int(w[2]), int(w[3])] for w in warn]})
    with open(os.path.join(HERE, "overlap_cases.json"), "w", encoding="utf-8") as fh:
        json.dump({"made_by": "tests/golden/make_golden.py overlap_cases(): reference check_non_overlapping", "cases": cases}, fh)
    return sum(len(c["warnings"]) for c in cases)


if __name__ == "__main__":
    if "--filters" in sys.argv:           # only the round-2 additions (the older fixtures are left untouched)
        filters()
        print("overlap warnings recorded:", overlap_cases())
        sys.exit(0)
    print("hash KATs:", hash_kats())
    mini()
    print("mini fixtures written to", MINI)
    presets()
    print("preset fixtures written")
