"""Small CPU-side synthetic genomes for parity tests (numpy only; independent of the device
generator in ntsynt_b200/synth.py).  An ancestor with a few contigs; each genome gets substitutions,
indels, inversions, translocations, segmental duplications and N runs."""
import numpy as np

_COMP = bytes.maketrans(b"ACGTN", b"TGCAN")


def _revcomp(b):
    return bytes(b).translate(_COMP)[::-1]


def make_genomes(seed, n_genomes=2, contig_lens=(200000, 150000), sub=0.01, indel=0.0005, n_inv=3, n_trans=2,
                 n_dup=2, n_nruns=3, lowercase=False):
    rng = np.random.default_rng(seed)
    anc = [bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), L, p=[0.3, 0.2, 0.2, 0.3])) for L in contig_lens]
    genomes = []
    for g in range(n_genomes):
        contigs = [bytearray(c) for c in anc]
        for _ in range(n_inv):
            c = int(rng.integers(len(contigs)))
            L = len(contigs[c])
            n = int(rng.integers(500, max(L // 5, 600)))
            a = int(rng.integers(0, L - n))
            contigs[c][a:a + n] = _revcomp(contigs[c][a:a + n])
        for _ in range(n_trans):
            c1, c2 = int(rng.integers(len(contigs))), int(rng.integers(len(contigs)))
            L = len(contigs[c1])
            n = int(rng.integers(500, max(L // 8, 600)))
            a = int(rng.integers(0, L - n))
            seg = contigs[c1][a:a + n]
            del contigs[c1][a:a + n]
            p = int(rng.integers(0, len(contigs[c2])))
            contigs[c2][p:p] = seg
        for _ in range(n_dup):
            c = int(rng.integers(len(contigs)))
            L = len(contigs[c])
            n = int(rng.integers(300, 3000))
            a = int(rng.integers(0, L - n))
            p = int(rng.integers(0, L))
            contigs[c][p:p] = contigs[c][a:a + n]
        out = []
        for ci, c in enumerate(contigs):
            arr = np.frombuffer(bytes(c), dtype=np.uint8).copy()
            m = rng.random(len(arr)) < sub
            arr[m] = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), int(m.sum()))
            c = bytearray(arr.tobytes())
            n_ev = rng.poisson(len(c) * indel)
            for p in sorted(rng.integers(1, len(c) - 1, n_ev), reverse=True):
                n = int(rng.geometric(1 / 3.0))
                if rng.random() < 0.5:
                    del c[p:p + n]
                else:
                    c[p:p] = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), n))
            for _ in range(n_nruns):
                n = int(rng.integers(1, 3000))
                a = int(rng.integers(0, max(len(c) - n, 1)))
                c[a:a + n] = b"N" * n
            if lowercase:
                a = int(rng.integers(0, len(c) // 2))
                c[a:a + 5000] = bytes(c[a:a + 5000]).lower()
            out.append((f"ctg{ci + 1}", bytes(c)))
        genomes.append(out)
    return genomes


def write_fasta(path, records, width=70):
    with open(path, "wb") as fh:
        for name, seq in records:
            fh.write(b">" + name.encode() + b" synthetic\n")
            for i in range(0, len(seq), width):
                fh.write(seq[i:i + width] + b"\n")
