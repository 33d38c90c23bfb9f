#!/usr/bin/env python3
"""Run the REFERENCE's own graph stage (bin/ntsynt_run.py, unmodified, imported from
/root/reference) on oracle-made sketches -- TEST INFRASTRUCTURE ONLY, usable only in the
build container (needs /root/reference).  Used to (1) validate oracle/graph_oracle.py and
the product's graph stage, (2) generate the golden fixtures under tests/golden/ (see
tests/golden/make_golden.py).

Third-party modules the reference imports (igraph, ncls, intervaltree, pybedtools, btllib)
and the executables it spawns (indexlr, seqtk) are replaced by oracle/shims/.
"""
import argparse
import gzip
import os
import shutil
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import sketch_oracle as so  # noqa: E402

REF = os.environ.get("NTSYNT_REFERENCE", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REF, "bin", "ntsynt_run.py"))


def write_fai(fasta_path, fai_path):
    "samtools faidx restatement (bin/ntsynt_run_pipeline.smk:48-53): name,len,offset,linebases,linewidth"
    rows = []
    with open(fasta_path, "rb") as fh:
        off = 0
        name = None
        for line in fh:
            if line.startswith(b">"):
                if name is not None:
                    rows.append((name, length, seq_off, lb, lw))
                name = line[1:].split()[0].decode()
                length, seq_off, lb, lw = 0, off + len(line), 0, 0
            else:
                stripped = line.rstrip(b"\r\n")
                if lb == 0:
                    lb, lw = len(stripped), len(line)
                length += len(stripped)
            off += len(line)
        if name is not None:
            rows.append((name, length, seq_off, lb, lw))
    with open(fai_path, "w", encoding="utf-8") as out:
        for r in rows:
            out.write("\t".join(map(str, r)) + "\n")


def save_bf(path, bits):
    with open(path, "wb") as fh:
        fh.write(b"ORCBF1\n")
        fh.write(np.uint64(bits.size).tobytes())
        fh.write(bits.tobytes())


def stage_fasta(src, workdir):
    "copy (and gunzip) a FASTA into workdir; returns the basename"
    base = os.path.basename(src)
    if base.endswith(".gz"):
        base = base[:-3]
        with gzip.open(src, "rb") as fin, open(os.path.join(workdir, base), "wb") as fout:
            shutil.copyfileobj(fin, fout)
    else:
        dst = os.path.join(workdir, base)
        if os.path.abspath(src) != os.path.abspath(dst):
            shutil.copyfile(src, dst)
    return base


def run_reference(fastas, workdir, prefix, k=24, w=1000, w_rounds=(100, 10), indel=10000, merge="10000",
                  block_size=500, fpr=0.025, simplify=True, common=True, restart_on_gap=False, quiet=True,
                  filter_mode=None, repeat_fpr=None, interarrivals=False, dev=False, n=0):
    """Full ntSynt run = oracle BF + oracle sketches + the reference's ntsynt_run.py.
    `fastas`: paths (may be .gz).  Mirrors bin/ntsynt_run_pipeline.smk:44-103.
    Returns dict of output paths."""
    if not reference_available():
        raise RuntimeError(f"reference not found at {REF}")
    os.makedirs(workdir, exist_ok=True)
    bases = [stage_fasta(f, workdir) for f in fastas]
    genomes = [(b, so.read_fasta(os.path.join(workdir, b))) for b in bases]
    for b in bases:
        write_fai(os.path.join(workdir, b), os.path.join(workdir, b + ".fai"))
    bf_path = None
    bits = None
    if common:
        # make_common_bf sorts the paths as given on its command line (cpp:107); the smk
        # passes the user's paths, tests run in the FASTA directory -> basenames.
        bits = so.common_bf(genomes, k, fpr)
        bf_path = f"{prefix}.common.bf"
        save_bf(os.path.join(workdir, bf_path), bits)
    rep_path = rep_bits = None
    if filter_mode:
        # rule make_repeat_bf (smk:65-72): ntsynt_make_repeat_bfs.py --genome <refs as given> -p <prefix>.repeat --fpr <fpr> -k <k>
        # -> size from the FIRST genome as given (every base counts), bits = k-mers seen twice within a genome
        import math
        size_bits = math.ceil(-sum(len(s) for _, s in genomes[0][1]) / math.log(1 - (repeat_fpr or fpr)))
        rep_bytes = max(int(math.ceil(int(size_bits / 8) / 8.0)) * 8, 8)
        rep_bits = so.repeat_bf(genomes, k, rep_bytes)
        rep_path = f"{prefix}.repeat.bf"
        save_bf(os.path.join(workdir, rep_path), rep_bits)
    tsvs = []
    for b, recs in genomes:
        tsv = f"{b}.k{k}.w{w}.tsv"
        # rule indexlr (smk:74-85) passes -r only when the pipeline's `repeat` switch is on; we tie it to --filter Indexlr
        so.write_sketch_tsv(os.path.join(workdir, tsv), recs, k, w, bits, rep_bits if filter_mode == "Indexlr" else None)
        tsvs.append(tsv)
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(HERE, "shims"), os.path.join(REF, "bin"),
                                        os.path.join(REF, "subprojects", "ntJoin", "bin")])
    env["PATH"] = os.path.join(HERE, "shims", "bin") + os.pathsep + env["PATH"]
    env["ORC_RESTART_ON_GAP"] = "1" if restart_on_gap else "0"
    env["PYTHONHASHSEED"] = env.get("PYTHONHASHSEED", "0")
    cmd = [sys.executable, os.path.join(REF, "bin", "ntsynt_run.py"), *tsvs, "-k", str(k), "-w", str(w),
           "--w-rounds", *map(str, w_rounds), "-p", prefix, "--bp", str(indel), "--collinear-merge", str(merge),
           "-z", str(block_size)]
    if common:
        cmd += ["--common", bf_path]
    if simplify:
        cmd += ["--simplify-graph"]
    cmd += ["--btllib_t", "1", "--fastas", *bases]
    if filter_mode:
        cmd += ["--filter", filter_mode, "--repeat", rep_path]
    if interarrivals:
        cmd += ["--interarrivals"]
    if n:
        cmd += ["-n", str(n)]
    if dev:
        cmd += ["--dev"]
    res = subprocess.run(cmd, cwd=workdir, env=env, stdout=subprocess.PIPE if quiet else None,
                         stderr=subprocess.STDOUT if quiet else None, text=True)
    out = {
        "returncode": res.returncode,
        "log": res.stdout if quiet else "",
        "tsvs": [os.path.join(workdir, t) for t in tsvs],
        "blocks": os.path.join(workdir, f"{prefix}.synteny_blocks.tsv"),
        "pre_merge": os.path.join(workdir, f"{prefix}.pre-collinear-merge.synteny_blocks.tsv"),
        "dot": os.path.join(workdir, f"{prefix}.mx.dot"),
        "bf": os.path.join(workdir, bf_path) if bf_path else None,
        "repeat_bf": os.path.join(workdir, rep_path) if rep_path else None,
        "repeat_bits": rep_bits,
        "interarrivals": os.path.join(workdir, f"{prefix}.interarrivals.tsv"),
    }
    return out


def main():
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("fastas", nargs="+")
    ap.add_argument("--workdir", required=True)
    ap.add_argument("-p", "--prefix", default="ntSynt")
    ap.add_argument("-k", type=int, default=24)
    ap.add_argument("-w", type=int, default=1000)
    ap.add_argument("--w_rounds", nargs="+", type=int, default=[100, 10])
    ap.add_argument("--indel", type=int, default=10000)
    ap.add_argument("--merge", default="10000")
    ap.add_argument("--block_size", type=int, default=500)
    a = ap.parse_args()
    out = run_reference(a.fastas, a.workdir, a.prefix, a.k, a.w, a.w_rounds, a.indel, a.merge, a.block_size,
                        quiet=False)
    sys.exit(out["returncode"])


if __name__ == "__main__":
    main()
