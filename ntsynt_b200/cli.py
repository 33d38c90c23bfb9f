"""Command-line entry points that keep the reference's contracts:
   ntSynt                 bin/ntSynt:43-175 (flags, divergence presets); runs the whole path in one process
   ntsynt_make_common_bf  src/ntsynt_make_common_bf.cpp:46-72 (flags), :107-164 (behaviour, log lines)
   indexlr                bin/ntsynt_run_pipeline.smk:83-85 and ntjoin_utils.py:197-198 (both spellings)
   ntsynt_run.py          bin/ntsynt_run.py:10-50
   ntsynt_make_repeat_bfs.py   bin/ntsynt_make_repeat_bfs.py:10-69
   denovo_synteny_block_stats.py   analysis_scripts/denovo_synteny_block_stats.py
   sort_ntsynt_blocks.py  visualization_scripts/sort_ntsynt_blocks.py
"""
import argparse
import os
import re
import sys
import time

import numpy as np

NTSYNT_VERSION = "ntSynt v1.0.4 (ntsynt_b200)"
MAX_W = 9216      # the window selector keeps a window in shared memory (limits)


def check_w(w, what="-w"):
    if w > MAX_W:
        raise SystemExit(f"ntsynt_b200: {what} {w} is above the largest window this build supports ({MAX_W})")


# ------------------------------------------------------------------------------------------------ ntSynt
def main_ntsynt(argv=None):
    epilog = "\n".join(["Default parameter settings for divergence values:",
                        "< 1% divergence:\t--block_size 500 --indel 10000 --merge 10000 --w_rounds 100 10",
                        "1% - 10% divergence:\t--block_size 1000 --indel 50000 --merge 100000 --w_rounds 250 100",
                        "> 10% divergence:\t--block_size 10000 --indel 100000 --merge 1000000 --w_rounds 500 250",
                        "If any of these parameters are set manually, those values will override the above."])
    ap = argparse.ArgumentParser(description="ntSynt: Multi-genome synteny detection using minimizer graphs "
                                             "(B200-native path)", formatter_class=argparse.RawTextHelpFormatter,
                                 epilog=epilog)
    ap.add_argument("fastas", nargs="*", help="Input genome fasta files")
    ap.add_argument("--fastas_list", type=str, help="File listing input genome fasta files, one per line")
    ap.add_argument("-d", "--divergence", required=True, type=float,
                    help="Approx. maximum percent sequence divergence between input genomes")
    ap.add_argument("-p", "--prefix", help="Prefix for ntSynt output files [ntSynt.k<k>.w<w>]")
    ap.add_argument("-k", type=int, default=24, help="Minimizer k-mer size [24]")
    ap.add_argument("-w", type=int, default=1000, help="Minimizer window size [1000]")
    ap.add_argument("-t", type=int, default=12, help="Number of threads [12] (accepted for compatibility)")
    ap.add_argument("--fpr", type=float, default=0.025, help="False positive rate for Bloom filter creation [0.025]")
    ap.add_argument("-b", "--block_size", type=int, help="Minimum synteny block size (bp)")
    ap.add_argument("--merge", type=str, help="Maximum distance between collinear synteny blocks for merging (bp or Nw)")
    ap.add_argument("--w_rounds", nargs="+", type=int, help="List of decreasing window sizes for refinement")
    ap.add_argument("--indel", type=int, help="Threshold for indel detection (bp)")
    ap.add_argument("--no-common", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--no-simplify-graph", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("-n", "--dry-run", action="store_true", help="Print the parameters and exit")
    ap.add_argument("--benchmark", action="store_true", help="Print per-stage timings")
    ap.add_argument("-f", "--force", action="store_true", help="Accepted for compatibility (every step always runs)")
    ap.add_argument("--dev", action="store_true", help="Also write the intermediate files (.fai, sketch TSVs, .common.bf, .mx.dot)")
    ap.add_argument("--gpu", type=int, default=0, help="CUDA device index [0]")
    ap.add_argument("-v", "--version", action="version", version=NTSYNT_VERSION)
    args = ap.parse_args(argv)
    if not args.prefix:
        args.prefix = f"ntSynt.k{args.k}.w{args.w}"
    d = args.divergence
    if d < 1:
        dflt = (10000, "10000", [100, 10], 500)
    elif d <= 10:
        dflt = (50000, "100000", [250, 100], 1000)
    elif d <= 100:
        dflt = (100000, "1000000", [500, 250], 10000)
    else:
        ap.error("--divergence must be a value between 0 and 100")
    args.indel, args.merge = args.indel or dflt[0], args.merge or dflt[1]
    args.w_rounds, args.block_size = args.w_rounds or dflt[2], args.block_size or dflt[3]
    for w in args.w_rounds:
        if w > args.w:
            ap.error("All values specified for --w_rounds must be smaller than -w")
    check_w(args.w)
    if not args.fastas and not args.fastas_list:
        ap.error("Please supply the input genome fasta files as positional arguments, "
                 "or specify a file listing the files (one fasta per line) with --fastas_list")
    if args.fastas and args.fastas_list:
        ap.error("Please supply the input genome fasta files as positional arguments, "
                 "or specify a single file (one fasta per line) with --fastas_list, NOT both.")
    if args.fastas_list:
        with open(args.fastas_list, "r", encoding="utf-8") as fin:
            fastas = [line.strip() for line in fin if line.strip()]
    else:
        fastas = args.fastas
    if len(fastas) < 2:
        ap.error("Must supply at least two reference genomes to compare")
    print("\n".join(["Running ntSynt...", f"Specified percent divergence: {d}", "Parameter settings:",
                     f"\tfastas {fastas}", f"\t--divergence {d}", f"\t--block_size {args.block_size}",
                     f"\t--merge {args.merge}", f"\t--w_rounds {args.w_rounds}", f"\t--indel {args.indel}",
                     f"\t-p {args.prefix}", f"\t-k {args.k}", f"\t-w {args.w}", f"\t-t {args.t}",
                     f"\t--fpr {args.fpr}"]), flush=True)
    for f in fastas:
        if not os.path.isfile(f):
            raise FileNotFoundError(f"Input file {f} not found.")
    if args.dry_run:
        return 0
    from . import device, fasta, io, pipeline
    t0 = time.perf_counter()
    ctx = device.Context(args.gpu)
    # ingest is a pipeline (pipeline.ingest_and_build): files parsed concurrently, genome i uploaded and inserted into the
    # common filter while the later files are still being read
    out, eng = pipeline.run_ntsynt(fastas, k=args.k, w=args.w, w_rounds=args.w_rounds, indel=args.indel, merge=args.merge,
                                   block_size=args.block_size, fpr=args.fpr, prefix=args.prefix,
                                   simplify=not args.no_simplify_graph, common=not args.no_common, write_files=True,
                                   quiet=False, ctx=ctx, intermediates=args.dev)
    t2 = time.perf_counter()
    if args.benchmark:
        tm = eng.ingest_timing
        print(f"ingest + Bloom filter (pipelined) {tm.get('ingest_build_s', 0.0):.3f} s, of which waiting for the parser "
              f"{tm.get('waited_for_parser_s', 0.0):.3f} s; whole run {t2 - t0:.3f} s; {eng.total_bases / (t2 - t0):.3e} bp/s", flush=True)
    print("Done ntSynt!")
    return 0


# ------------------------------------------------------------------------------------------------ make_common_bf
def main_make_common_bf(argv=None):
    ap = argparse.ArgumentParser(prog="ntsynt_make_common_bf")
    ap.add_argument("--genome", nargs="+", required=True, help="Input genome file(s)")
    ap.add_argument("-k", required=True, type=int, help="k-mer size (bp)")
    ap.add_argument("--fpr", type=float, default=0.025, help="False positive rate for Bloom filter")
    ap.add_argument("-p", default="common_bf", help="Prefix for output Bloom filter")
    ap.add_argument("--bf", type=int, help="Bloom filter size in bytes (optional)")
    ap.add_argument("-t", type=int, default=12, help="Number of threads (accepted for compatibility)")
    ap.add_argument("--gpu", type=int, default=0)
    ap.add_argument("--bf-format", choices=["btllib", "native"], default="btllib",
                    help="layout of the .bf file: btllib's header (default; unpinned by any reference fixture) or our own container")
    try:
        args = ap.parse_args(argv)
    except SystemExit as exc:
        if exc.code not in (0, None):
            sys.exit(1)
        raise
    from . import device, fasta, io, pipeline
    print("Parameters:")
    print("\t\t--genome " + " ".join(args.genome) + " ")
    print(f"\t\t-t {args.t}\n\t\t-k {args.k}\n\t\t--fpr {args.fpr:g}\n\t\t-p {args.p}", flush=True)
    ctx = device.Context(args.gpu)
    nbytes = None
    if args.bf is not None:
        nbytes = int(np.ceil(args.bf / 8.0)) * 8
        print(f"\t\t--bf {args.bf}")
    else:
        print("Calculating BF size based on input genome size")
    genomes = [ctx.upload(fasta.read_fasta(g)) for g in args.genome]
    bf = pipeline.build_common_bf(ctx, genomes, list(args.genome), args.k, args.fpr, nbytes=nbytes, log=print)
    print(f"Final Bloom filter FPR: {bf.fpr()}")
    io.save_bf(args.p + ".bf", bf, args.k, fmt=args.bf_format)
    print("Done!", flush=True)
    return 0


# ------------------------------------------------------------------------------------------------ indexlr
def parse_indexlr_args(argv):
    "both spellings: `-k 24 -w 1000 ... fa` and `fa --seq --long --pos -k24 -w1000 -t4 -s bf -o out`"
    opt = dict(k=None, w=None, t=1, s=None, r=None, o=None, seq=False, pos=False, long=False, fa=None)
    i = 0
    while i < len(argv):
        a = argv[i]
        if a in ("--seq", "--pos", "--long"):
            opt[a[2:]] = True
        elif a in ("--help", "-h"):
            print("indexlr -k K -w W [--long] [--pos] [--seq] [-t T] [-s in.bf] [-r out.bf] [-o FILE] <fasta>")
            sys.exit(0)
        elif (m := re.match(r"^-([kwtsro])(.*)$", a)):
            flag, val = m.group(1), m.group(2)
            if val == "":
                i += 1
                if i >= len(argv):
                    raise SystemExit(f"indexlr: option -{flag} needs a value")
                val = argv[i]
            opt[flag] = int(val) if flag in "kwt" else val
        elif a.startswith("-") and a != "-":
            raise SystemExit(f"indexlr: unknown option {a}")
        else:
            opt["fa"] = a
        i += 1
    if opt["k"] is None or opt["w"] is None or opt["fa"] is None:
        raise SystemExit("indexlr: -k, -w and a FASTA file are required")
    check_w(opt["w"])
    return opt


def main_indexlr(argv=None):
    opt = parse_indexlr_args(sys.argv[1:] if argv is None else argv)
    from . import device, fasta, io
    ctx = device.Context(int(os.environ.get("NTSYNT_B200_GPU", "0")))
    packed = fasta.read_fasta(opt["fa"])
    genome = ctx.upload(packed)
    common = io.load_bf(ctx, opt["s"])[0] if opt["s"] and opt["s"] != "None" else None
    repeat = io.load_bf(ctx, opt["r"])[0] if opt["r"] and opt["r"] != "None" else None
    table = ctx.sketch(genome, opt["k"], opt["w"], common=common, repeat=repeat).to_numpy()
    out = open(opt["o"], "w", encoding="utf-8") if opt["o"] else sys.stdout
    io.write_sketch_tsv(out, packed, table, opt["k"], with_seq=opt["seq"], with_pos=opt["pos"])
    if opt["o"]:
        out.close()
    return 0


# ------------------------------------------------------------------------------------------------ ntsynt_run.py
def main_ntsynt_run(argv=None):
    ap = argparse.ArgumentParser(description="Run the dynamic minimizer graph stage of ntSynt (B200-native)")
    ap.add_argument("FILES", nargs="+", help="Minimizer TSV files of input assemblies")
    ap.add_argument("--fastas", nargs="+", required=True, type=str, help="Assembly fasta files")
    ap.add_argument("-n", default=0, type=int, help="Minimum edge weight [Number of input assemblies]")
    ap.add_argument("-p", default="out", type=str, help="Output prefix [out]")
    ap.add_argument("-k", required=True, type=int, help="k-mer size used for minimizer step")
    ap.add_argument("-w", required=True, type=int, help="Window size used for minimizers")
    ap.add_argument("-z", type=int, default=500, help="Minimum synteny block size (bp) [500]")
    ap.add_argument("--filter", choices=["Filter", "Indexlr"], type=str, help="Type of repeat filtering")
    ap.add_argument("--common", type=str, help="Input common BF for minimizer selection")
    ap.add_argument("--repeat", type=str, help="Repeat BF (must be included if --filter is specified)")
    ap.add_argument("--btllib_t", type=int, default=4, help="accepted for compatibility")
    ap.add_argument("--w-rounds", default=[100, 10], nargs="+", type=int, help="decreasing list of 'w' values")
    ap.add_argument("--bp", default=500, type=int, help="Maximum tolerated indel size [500]")
    ap.add_argument("--collinear-merge", default="1w", type=str, help="Maximum distance between collinear blocks")
    ap.add_argument("--simplify-graph", action="store_true", help="Run minimizer graph simplification")
    ap.add_argument("-m", default=90, type=int, help="percent of increasing/decreasing positions for orientation [90]")
    ap.add_argument("--dev", action="store_true")
    ap.add_argument("--interarrivals", action="store_true")
    ap.add_argument("-v", "--version", action="version", version=NTSYNT_VERSION)
    print(f"Running {NTSYNT_VERSION}", flush=True)
    args = ap.parse_args(argv)
    if args.n not in (0, len(args.FILES)):
        raise SystemExit("ntsynt_b200: only -n = number of assemblies (the pipeline's setting) is supported")
    check_w(max([args.w] + list(args.w_rounds)), "-w / --w-rounds")
    if args.filter and not args.repeat:                            # bin/ntsynt_synteny.py:601-602
        raise ValueError("If --filter is specified, must supply repeat Bloom filter with --repeat")
    from . import device, fasta, io, pipeline
    from .synteny import FA_TSV_RE, SyntenyEngine
    files = sorted(args.FILES, reverse=True)                       # bin/ntsynt_synteny.py:34
    by_base = {os.path.basename(f): f for f in args.fastas}        # :137
    ctx = device.Context(int(os.environ.get("NTSYNT_B200_GPU", "0")))
    packed, genomes, tables = [], [], []
    for tsv in files:
        mt = re.search(FA_TSV_RE, os.path.basename(tsv))
        if not mt:
            print("ERROR: Target assembly minimizer TSV file must follow the naming convention:")
            print("\ttarget_assembly.fa.k<k>.w<w>.tsv, where <k> and <w> are parameters used for minimizering")
            sys.exit(1)
        fa = by_base[mt.group(1)]
        pk = fasta.read_fasta(fa)
        if not os.path.exists(mt.group(1) + ".fai") and pk.fai:
            fasta.write_fai(pk, mt.group(1) + ".fai")
        names, h1, pos, ctg = io.read_sketch_tsv(tsv)
        remap = {n: i for i, n in enumerate(pk.names)}
        ctg = np.array([remap[names[c]] for c in ctg.tolist()], dtype=np.uint32) if len(ctg) else ctg
        packed.append(pk)
        genomes.append(ctx.upload(pk))
        tables.append(device.MinimizerTable.from_numpy(ctx, h1, pos, ctg, genomes[-1]))
    common = io.load_bf(ctx, args.common)[0] if args.common else None
    repeat = io.load_bf(ctx, args.repeat)[0] if args.repeat and args.filter else None
    be = pipeline.CudaBackend(ctx, genomes, [os.path.basename(f) for f in files], [p.names for p in packed],
                              [[int(x) for x in p.lengths] for p in packed], args.k, common=common, round0=tables,
                              repeat=repeat, filter_mode=args.filter)
    eng = SyntenyEngine(be, args.k, args.w, args.w_rounds, args.bp, args.collinear_merge, args.z, m=args.m,
                        simplify=args.simplify_graph, prefix=args.p, dev=args.dev, write_files=True, quiet=False,
                        interarrivals=args.interarrivals)
    if os.environ.get("NTSYNT_B200_NO_DOT") != "1":
        eng.dot_path = f"{args.p}.mx.dot"
    eng.run()
    print("DONE!", flush=True)
    return 0


# ------------------------------------------------------------------------------------------------ ntsynt_make_repeat_bfs.py
def parse_bf_size(text, ap):
    "bin/ntsynt_make_repeat_bfs.py:10-24: <num>[BkMG]"
    m = re.search(r"^(\d+)([BkMG])$", text)
    if not m:
        ap.print_help()
        ap.error(f"Invalid input value for --bf: {text}")
    num, unit = int(m.group(1)), m.group(2)
    return {"B": num, "k": int(num * 1e3), "M": int(num * 1e6), "G": int(num * 1e9)}[unit]


def main_make_repeat_bfs(argv=None):
    """Bloom filter of the k-mers seen twice or more within an input genome (bin/ntsynt_make_repeat_bfs.py:56-69):
    per genome a fresh filter; a k-mer whose bit is already set goes to the repeat filter, else its bit is set.
    With one hash function that is: bit i of the repeat filter is set iff two or more k-mer occurrences of one
    genome hash to i -- what nts_bf_insert_repeats computes (order-free, bit-exact vs the sequential statement)."""
    import math
    ap = argparse.ArgumentParser(description="Generating BF of k-mer 2+ multiplicities")
    ap.add_argument("--genome", help="Input genome file(s)", nargs="+")
    ap.add_argument("-k", help="K-mer size (bp)", required=True, type=int)
    ap.add_argument("--bf", help="Bloom filter size [accepted units: B (bytes), k (kilobytes), M (megabytes), G (gigabytes)]",
                    type=str)
    ap.add_argument("-t", help="Number of threads [4] (accepted for compatibility)", required=False, type=int, default=4)
    ap.add_argument("-p", help="Prefix for output BF", default="out.bf", type=str)
    ap.add_argument("--fpr", help="False positive rate for Bloom filter. Only used if --bf is not specified. [0.01]",
                    default=0.01, type=float)
    ap.add_argument("--gpu", type=int, default=0)
    ap.add_argument("--bf-format", choices=["btllib", "native"], default="btllib", help="layout of the .bf file")
    args = ap.parse_args(argv)
    if not args.genome:
        ap.error("--genome is required")
    from . import device, fasta, io
    ctx = device.Context(args.gpu)
    first = fasta.read_fasta(args.genome[0])
    if not args.bf:
        # approximate_bf_size (:26-34): every base of the first genome counts, N included
        size_bits = math.ceil((-1 * int(sum(int(x) for x in first.lengths))) / math.log(1 - args.fpr))
        bf_bytes = int(size_bits / 8)
        print(f"Calculated Bloom filter size: {bf_bytes} bytes")
    else:
        bf_bytes = parse_bf_size(args.bf, ap)
    nbytes = max(int(math.ceil(bf_bytes / 8.0)) * 8, 8)           # btllib rounds the array up to whole 64-bit words
    rep, scratch = ctx.bloom(nbytes), ctx.bloom(nbytes)
    for i, path in enumerate(args.genome):
        g = ctx.upload(first if i == 0 else fasta.read_fasta(path))
        scratch.clear()
        rep.insert_repeats(scratch, g, args.k)
        g.close()
    io.save_bf(f"{args.p}.bf", rep, args.k, fmt=args.bf_format)
    return 0


# ------------------------------------------------------------------------------------------------ block stats
def block_stats(tsv, fais):
    """analysis_scripts/denovo_synteny_block_stats.py: the ten summary numbers of a synteny-block TSV.
    Returns (header list, value list) exactly as that script prints them."""
    sizes = {}
    for fai in fais:
        if (mt := re.search(r"^(\S+).fai", fai)):
            with open(fai, "r", encoding="utf-8") as fh:
                sizes[os.path.basename(mt.group(1))] = sum(int(line.strip().split("\t")[1]) for line in fh)
    lens, ids, tally = {}, {}, {}
    with open(tsv, "r", encoding="utf-8") as fh:
        for line in fh:
            f = line.strip().split("\t")
            lens.setdefault(f[1], []).append(int(f[4]) - int(f[3]))
            ids.setdefault(f[1], []).append(f[0])
            tally.setdefault(f[0], set()).add(f[1])
    G = len(fais)
    full = {a: [l for l, b in zip(lens[a], ids[a]) if len(tally[b]) >= G] for a in lens}

    def ng50(blocks, size):
        acc = 0
        for b in sorted(blocks, reverse=True):
            acc += b
            if acc >= size * 0.5:
                return b
        return 0

    num_blocks = sum(len(v) for v in lens.values()) / G
    num_all = sum(len(v) for v in full.values()) / G
    total = sum(sum(v) for v in lens.values()) / G
    cov = sum(sum(v) / sizes[a] * 100 for a, v in lens.items()) / G
    min_size, min_asm = sorted(((s, a) for a, s in sizes.items()), key=lambda x: x[0])[0]
    cov_min = sum(lens[min_asm]) / min_size * 100
    cov_all = sum(sum(v) / sizes[a] * 100 for a, v in full.items()) / G
    avg = sum(np.mean(v) for v in lens.values()) / G
    med = sum(np.median(v) for v in lens.values()) / G
    a_ng50 = sum(ng50(v, sizes[a]) for a, v in lens.items()) / G
    a_n50 = sum(ng50(v, sum(v)) for v in lens.values()) / G
    head = ["Number_blocks", "Number_blocks_all_asm", "Average_coverage", "Average_coverage_all_asm",
            "Coverage_min_genome_size", "Average_length", "Median_length", "Total_length", "NG50_length", "N50_length"]
    return head, [int(num_blocks), int(num_all), cov, cov_all, cov_min, avg, med, total, int(a_ng50), int(a_n50)]


def main_block_stats(argv=None):
    ap = argparse.ArgumentParser(description="Compute de novo stats on synteny blocks")
    ap.add_argument("--tsv", help="ntSynt synteny block file", required=True)
    ap.add_argument("--fai", help="FAI files for the compared genomes", nargs="+", required=True)
    args = ap.parse_args(argv)
    head, vals = block_stats(args.tsv, args.fai)
    print(*head, sep="\t")
    print("\t".join(str(v) for v in vals))
    return 0


# ------------------------------------------------------------------------------------------------ sort blocks
def sort_blocks_text(text, order):
    """visualization_scripts/sort_ntsynt_blocks.py:16-41: the rows of every block re-ordered by the rank of their assembly
    in `order` (dict name -> rank); 8-column rows keep all columns, other rows their first six."""
    out, cur, cur_id = [], [], None

    def flush():
        for row in sorted(cur, key=lambda r: order[r[1]]):
            out.append("\t".join(row) + "\n")
    for line in text.splitlines():
        f = line.strip().split("\t")
        row = f if len(f) == 8 else f[:6]
        if cur_id is not None and row[0] != cur_id:
            flush()
            cur = []
        cur.append(row)
        cur_id = row[0]
    flush()
    return "".join(out)


def main_sort_blocks(argv=None):
    ap = argparse.ArgumentParser(description="Sort the assemblies in the ntSynt synteny blocks in specified order")
    ap.add_argument("--synteny_blocks", help="Input synteny blocks", required=True, type=str)
    ap.add_argument("--sort_order", help="Desired assembly sort order", nargs="+", required=True)
    ap.add_argument("--fais", help="The assembly sort order option lists the FAI files for the assemblies", action="store_true")
    args = ap.parse_args(argv)
    if args.fais:
        order = {re.search(r"^(\S+)\.fai$", os.path.basename(os.path.realpath(a))).group(1): i for i, a in enumerate(args.sort_order)}
    else:
        order = {a: i for i, a in enumerate(args.sort_order)}
    with open(args.synteny_blocks, "r", encoding="utf-8") as fh:
        sys.stdout.write(sort_blocks_text(fh.read(), order))
    return 0
