"""The d >= 1 presets of bin/ntSynt:89-99 (BASELINE configs 2-5 use them) on seeded synth_small genomes.
Shared by tests/golden/make_golden.py (which runs the REFERENCE's own bin/ntsynt_run.py on them and commits the
block files under tests/golden/presets/), by the CPU tests (graph oracle == those files; reference re-run when
/root/reference is present) and by the GPU tests (CUDA path == those files)."""
import hashlib
import json
import os

import synth_small

HERE = os.path.dirname(os.path.abspath(__file__))
PRESET_DIR = os.path.join(HERE, "golden", "presets")

CASES = {
    # -d 1.3  ->  --block_size 1000 --indel 50000 --merge 100000 --w_rounds 250 100   (bin/ntSynt:92-94)
    "d1.3_G3": dict(G=3, seed=1301, lens=(1_500_000, 900_000, 600_000), sub=0.0065,
                    params=dict(k=24, w=1000, w_rounds=[250, 100], indel=50000, merge="100000", block_size=1000)),
    # -d 12   ->  --block_size 10000 --indel 100000 --merge 1000000 --w_rounds 500 250   (bin/ntSynt:95-97)
    "d12_G5": dict(G=5, seed=1205, lens=(2_500_000, 1_500_000, 1_000_000), sub=0.03,
                   params=dict(k=24, w=1000, w_rounds=[500, 250], indel=100000, merge="1000000", block_size=10000)),
}


def names(tag):
    return [f"g{chr(65 + i)}.fa" for i in range(CASES[tag]["G"])]


def genomes(tag):
    c = CASES[tag]
    return synth_small.make_genomes(c["seed"], c["G"], contig_lens=c["lens"], sub=c["sub"], n_inv=4, n_trans=3, n_dup=3,
                                    n_nruns=4, lowercase=True)


def digest(gens):
    h = hashlib.sha1()
    for recs in gens:
        for name, seq in recs:
            h.update(name.encode()); h.update(seq)
    return h.hexdigest()


def expected(tag, which="synteny_blocks.tsv"):
    with open(os.path.join(PRESET_DIR, tag, which), encoding="utf-8") as fh:
        return fh.read()


def checked_genomes(tag):
    "the seeded genomes, verified against the digest recorded when the reference made the fixture"
    gens = genomes(tag)
    with open(os.path.join(PRESET_DIR, "params.json"), encoding="utf-8") as fh:
        want = json.load(fh)[tag]["genomes_sha1"]
    if digest(gens) != want:
        raise AssertionError(f"{tag}: synth_small no longer reproduces the genomes the fixture was made from")
    return gens
