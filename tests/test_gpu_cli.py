"""Drop-in executables under bin/ (SURVEY.md 8b), run as subprocesses on the GPU box with the reference's
own command lines (bin/ntsynt_run_pipeline.smk:55-103, ntjoin_utils.py:197-198)."""
import gzip
import json
import os
import shutil
import subprocess
import sys

import pytest

from conftest import MINI, mini_expected, mini_fastas

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "bin")
pytestmark = pytest.mark.gpu


def run(cmd, cwd):
    res = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:]
    return res.stdout


def stage(tmp_path, tag):
    names = []
    for p in mini_fastas(tag):
        n = os.path.basename(p)[:-3]
        with gzip.open(p, "rb") as fi, open(tmp_path / n, "wb") as fo:
            shutil.copyfileobj(fi, fo)
        names.append(n)
    return names


def test_smk_rule_commands_reproduce_the_fixture(tmp_path, mini_params):
    "make_common_bf -> indexlr (both spellings) -> ntsynt_run.py, exactly as the smk rules spell them"
    p = mini_params
    names = stage(tmp_path, "ABC")
    k, w = p["k"], p["w"]
    out = run([sys.executable, os.path.join(BIN, "ntsynt_make_common_bf"), "--genome", *names, "-p", "mini.common", "--fpr",
               str(p["fpr"]), "-k", str(k), "-t", "4"], tmp_path)
    assert "BF size (bytes):" in out and "Final Bloom filter FPR:" in out
    for i, n in enumerate(names):
        tsv = f"{n}.k{k}.w{w}.tsv"
        if i == 0:      # smk spelling, stdout redirect
            txt = run([sys.executable, os.path.join(BIN, "indexlr"), "-k", str(k), "-w", str(w), "--long", "--seq", "--pos",
                       "-t", "5", "-s", "mini.common.bf", n], tmp_path)
            (tmp_path / tsv).write_text(txt)
        else:           # ntjoin_utils.run_indexlr spelling
            run([sys.executable, os.path.join(BIN, "indexlr"), n, "--seq", "--long", "--pos", f"-k{k}", f"-w{w}", "-t4",
                 "-s", "mini.common.bf", "-o", tsv], tmp_path)
        assert (tmp_path / tsv).read_text() == mini_expected("ABC", tsv + ".gz")
    run([sys.executable, os.path.join(BIN, "ntsynt_run.py"), *[f"{n}.k{k}.w{w}.tsv" for n in names], "-k", str(k), "-w", str(w),
         "--w-rounds", *map(str, p["w_rounds"]), "-p", "mini-ABC", "--bp", str(p["indel"]), "--collinear-merge", p["merge"],
         "-z", str(p["block_size"]), "--common", "mini.common.bf", "--simplify-graph", "--btllib_t", "4", "--fastas", *names],
        tmp_path)
    assert (tmp_path / "mini-ABC.synteny_blocks.tsv").read_text() == mini_expected("ABC", "synteny_blocks.tsv")
    assert (tmp_path / "mini-ABC.pre-collinear-merge.synteny_blocks.tsv").read_text() == \
        mini_expected("ABC", "pre-collinear-merge.synteny_blocks.tsv")
    for n in names:
        assert (tmp_path / (n + ".fai")).read_text() == open(os.path.join(MINI, "ABC", n + ".fai")).read()
    # .mx.dot: same edge multiset as the reference's round-0 graph
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden import dot_edges
    with gzip.open(os.path.join(MINI, "ABC", "mx_dot_edges.json.gz"), "rt") as fh:
        assert dot_edges(str(tmp_path / "mini-ABC.mx.dot")) == json.load(fh)


def test_ntsynt_cli_end_to_end(tmp_path, mini_params):
    "bin/ntSynt with explicit parameters (tests/ntsynt_tests.py:8-23 style) + --dev intermediates"
    p = mini_params
    names = stage(tmp_path, "AB")
    run([sys.executable, os.path.join(BIN, "ntSynt"), "--force", *names, f"-k{p['k']}", "-w", str(p["w"]), "-d", "0.5",
         "--prefix", "mini-AB", "--indel", str(p["indel"]), "--merge", p["merge"], "--block_size", str(p["block_size"]),
         "--w_rounds", *map(str, p["w_rounds"]), "--dev"], tmp_path)
    assert (tmp_path / "mini-AB.synteny_blocks.tsv").read_text() == mini_expected("AB", "synteny_blocks.tsv")
    tsv = f"{names[0]}.k{p['k']}.w{p['w']}.tsv"
    assert (tmp_path / tsv).read_text() == mini_expected("AB", tsv + ".gz")
    assert (tmp_path / "mini-AB.common.bf").exists() and (tmp_path / "mini-AB.mx.dot").exists()
    # fastas_list spelling
    (tmp_path / "list.tsv").write_text("\n".join(names) + "\n")
    run([sys.executable, os.path.join(BIN, "ntSynt"), "--fastas_list", "list.tsv", f"-k{p['k']}", "-w", str(p["w"]), "-d", "0.5",
         "--prefix", "mini-fof", "--indel", str(p["indel"]), "--merge", p["merge"], "--block_size", str(p["block_size"]),
         "--w_rounds", *map(str, p["w_rounds"])], tmp_path)
    assert (tmp_path / "mini-fof.synteny_blocks.tsv").read_text() == mini_expected("AB", "synteny_blocks.tsv")


def test_ntsynt_cli_reads_gz_inputs_through_the_ingest_pipeline(tmp_path, mini_params):
    "bin/ntSynt on .fa.gz paths in another directory: files parsed concurrently, genome i inserted while the next is read"
    p = mini_params
    src = tmp_path / "data"
    src.mkdir()
    paths = []
    for f in mini_fastas("ABC"):
        shutil.copyfile(f, src / os.path.basename(f))
        paths.append(os.path.join("data", os.path.basename(f)))
    out = run([sys.executable, os.path.join(BIN, "ntSynt"), *paths, f"-k{p['k']}", "-w", str(p["w"]), "-d", "0.5",
               "--prefix", "gz-ABC", "--indel", str(p["indel"]), "--merge", p["merge"], "--block_size", str(p["block_size"]),
               "--w_rounds", *map(str, p["w_rounds"]), "--benchmark"], tmp_path)
    assert "ingest + Bloom filter (pipelined)" in out
    assert (tmp_path / "gz-ABC.synteny_blocks.tsv").read_text() == mini_expected("ABC", "synteny_blocks.tsv")


def test_cli_argument_errors(tmp_path):
    res = subprocess.run([sys.executable, os.path.join(BIN, "ntSynt"), "a.fa", "-d", "1"], cwd=tmp_path, stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 2 and "at least two" in res.stdout
    res = subprocess.run([sys.executable, os.path.join(BIN, "ntsynt_make_common_bf"), "--genome", "x.fa"], cwd=tmp_path,
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 1


@pytest.mark.parametrize("tag,mode", [("ABC_filter", "Filter"), ("ABC_indexlr", "Indexlr")])
def test_repeat_filter_chain_reproduces_the_reference_fixture(tmp_path, mini_params, tag, mode):
    """rule make_repeat_bf -> indexlr -s -r -> ntsynt_run.py --filter <mode> --repeat --interarrivals --dev
    (bin/ntsynt_run_pipeline.smk:65-103; bin/ntsynt_synteny.py:172-187,557-564,601-609): block files and the
    interarrival distances were written by the reference's own graph stage (tests/golden/make_golden.py filters())."""
    import numpy as np
    from ntsynt_b200 import io
    p = mini_params
    names = stage(tmp_path, "ABC")
    k, w = p["k"], p["w"]
    run([sys.executable, os.path.join(BIN, "ntsynt_make_common_bf"), "--genome", *names, "-p", "mini.common", "--fpr",
         str(p["fpr"]), "-k", str(k), "-t", "4"], tmp_path)
    out = run([sys.executable, os.path.join(BIN, "ntsynt_make_repeat_bfs.py"), "--genome", *names, "-p", "mini.repeat", "--fpr", "0.2",
               "-k", str(k), "-t", "4"], tmp_path)
    assert "Calculated Bloom filter size:" in out
    bits, kk = io.load_bf_bytes(str(tmp_path / "mini.repeat.bf"))
    with gzip.open(os.path.join(MINI, "repeat_bits.bin.gz"), "rb") as fh:
        assert kk == k and np.array_equal(bits, np.frombuffer(fh.read(), dtype=np.uint8))      # == bin/ntsynt_make_repeat_bfs.py restated
    for n in names:
        rep = ["-r", "mini.repeat.bf"] if mode == "Indexlr" else []
        run([sys.executable, os.path.join(BIN, "indexlr"), n, "--seq", "--long", "--pos", f"-k{k}", f"-w{w}", "-t4",
             "-s", "mini.common.bf", *rep, "-o", f"{n}.k{k}.w{w}.tsv"], tmp_path)
    log = run([sys.executable, os.path.join(BIN, "ntsynt_run.py"), *[f"{n}.k{k}.w{w}.tsv" for n in names], "-k", str(k), "-w", str(w),
               "--w-rounds", *map(str, p["w_rounds"]), "-p", "mini-F", "--bp", str(p["indel"]), "--collinear-merge", p["merge"],
               "-z", str(p["block_size"]), "--common", "mini.common.bf", "--simplify-graph", "--btllib_t", "4", "--fastas", *names,
               "--filter", mode, "--repeat", "mini.repeat.bf", "--interarrivals", "--dev"], tmp_path)
    assert "WARNING: detected overlapping" not in log
    assert (tmp_path / "mini-F.synteny_blocks.tsv").read_text() == mini_expected(tag, "synteny_blocks.tsv")
    assert (tmp_path / "mini-F.pre-collinear-merge.synteny_blocks.tsv").read_text() == \
        mini_expected(tag, "pre-collinear-merge.synteny_blocks.tsv")
    got = sorted(int(x) for x in (tmp_path / "mini-F.interarrivals.tsv").read_text().split())
    assert got == [int(x) for x in mini_expected(tag, "interarrivals.sorted.txt.gz").split()]


def test_filter_without_repeat_is_the_reference_error(tmp_path):
    res = subprocess.run([sys.executable, os.path.join(BIN, "ntsynt_run.py"), "a.fa.k24.w100.tsv", "b.fa.k24.w100.tsv", "-k", "24",
                          "-w", "100", "--fastas", "a.fa", "b.fa", "--filter", "Filter"], cwd=tmp_path, stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True)
    assert res.returncode != 0 and "must supply repeat Bloom filter with --repeat" in res.stdout
