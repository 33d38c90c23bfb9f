"""Stand-in for pybedtools/bedtools (TEST INFRASTRUCTURE ONLY) covering the one chain
used at bin/ntsynt_synteny.py:144-147:
    BedTool(str, from_string=True).slop(g=fai, l=-(w+k), r=-(w+k)).sort().mask_fasta(fi=, fo=)
slop clamps to [0, chrom_len]; intervals that a negative slop leaves empty or inverted
(start >= end) are DROPPED -- real bedtools' behaviour there is unpinned."""


class BedTool:
    def __init__(self, data, from_string=False):
        self.ivs = []
        if from_string:
            for line in data.splitlines():
                f = line.split("\t")
                if len(f) >= 3:
                    self.ivs.append((f[0], int(f[1]), int(f[2])))
        else:
            self.ivs = list(data)

    def slop(self, g, l=0, r=0):
        sizes = {}
        with open(g, encoding="utf-8") as fh:
            for line in fh:
                f = line.rstrip("\n").split("\t")
                sizes[f[0]] = int(f[1])
        out = []
        for c, s, e in self.ivs:
            s2 = max(s - l, 0)
            e2 = min(e + r, sizes[c])
            if s2 < e2:
                out.append((c, s2, e2))
        return BedTool(out)

    def sort(self):
        return BedTool(sorted(self.ivs))

    def mask_fasta(self, fi, fo):
        by = {}
        for c, s, e in self.ivs:
            by.setdefault(c, []).append((s, e))
        with open(fi, "rb") as fin, open(fo, "wb") as fout:
            name, chunks = None, []

            def flush():
                if name is None:
                    return
                seq = bytearray(b"".join(chunks))
                for s, e in by.get(name.split()[0].decode() if name.split() else "", []):
                    seq[s:e] = b"N" * (min(e, len(seq)) - s)
                fout.write(b">" + name + b"\n")
                for i in range(0, len(seq), 60):
                    fout.write(seq[i:i + 60] + b"\n")

            for line in fin:
                if line.startswith(b">"):
                    flush()
                    name, chunks = line[1:].rstrip(b"\r\n"), []
                else:
                    chunks.append(line.strip())
            flush()
        return self
