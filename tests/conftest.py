import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
MINI = os.path.join(GOLDEN, "mini")
REF_DEMO = os.path.join(GOLDEN, "_ref_demo")            # staged copy of the reference's demo data (git-ignored)
REFERENCE = os.environ.get("NTSYNT_REFERENCE", "/root/reference")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than ~20 s on CPU")


def _demo_dir():
    "directory holding celegans-*.fa.gz + expected_result/, or None"
    if os.path.isdir(os.path.join(REF_DEMO, "expected_result")):
        return REF_DEMO
    if os.path.isdir(os.path.join(REFERENCE, "tests", "expected_result")):
        return os.path.join(REFERENCE, "tests")
    return None


@pytest.fixture(scope="session")
def demo_dir():
    d = _demo_dir()
    if d is None:
        pytest.skip("reference demo data not staged (tests/golden/stage_ref_demo.py)")
    return d


@pytest.fixture(scope="session")
def mini_params():
    with open(os.path.join(MINI, "params.json"), encoding="utf-8") as fh:
        return json.load(fh)


def mini_fastas(tag):
    names = ["miniA.fa", "miniB.fa"] + (["miniC.fa"] if tag == "ABC" else [])
    return [os.path.join(MINI, n + ".gz") for n in names]


def mini_expected(tag, name):
    p = os.path.join(MINI, tag, name)
    if p.endswith(".gz"):
        with gzip.open(p, "rt", encoding="utf-8") as fh:
            return fh.read()
    with open(p, encoding="utf-8") as fh:
        return fh.read()


def parse_sketch_tsv(text):
    "{contig: (h1 list, pos list)} from indexlr --long --pos --seq output"
    out = {}
    for line in text.splitlines():
        name, _, rest = line.partition("\t")
        toks = rest.split(" ") if rest else []
        out[name] = ([int(t.split(":")[0]) for t in toks], [int(t.split(":")[1]) for t in toks])
    return out


@pytest.fixture(scope="session")
def cuda_ctx():
    from ntsynt_b200 import device
    return device.Context(0)
