#!/usr/bin/env python3
"""bench.py -- genome bp/s through the sketch + Bloom-filter + graph path (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference path

A step = one pass of the whole hot path over one batch of synthetic genomes: Bloom-filter zero-fill,
per-genome insert, merge (the AND of the cascade: all levels but the last are ANDed, the last stays apart and the
sketches look candidates up in both parts -- nts_bf_build_common_lazy / nts_sketch2), round-0 sketch, minimizer
join + graph, every refinement round (masked re-sketch + graph), erosion, collinear merges, final block table
as TSV text.

N = 1  : BASELINE.json configs[1]: 2 synthetic ~3 Gbp human-like genomes, d = 1 %, k = 24, w = 1000,
         presets of bin/ntSynt:92-94 (block_size 1000, indel 50000, merge 100000, w_rounds 250 100).
N > 1  : configs[4] (default, G = N or a multiple): one 3 Gbp genome per GPU, per-GPU Bloom filters merged over NVLink
         -- by default with the peer-memory reduce-scatter / all-gather kernels (csrc/nts_p2p.cu), or with --merge nccl
         by one NCCL all-reduce(sum) over packed counters (the north-star form; bit-identical, ~4x the wire volume;
         both are timed alone in config.merge_alone_ms) -- then every rank sketches its genome and rank 0 runs the
         graph stage on the gathered tables.  Weak scaling: per-GPU work is fixed.
         configs[2] / configs[3] (--genomes 3 --divergence 1.3, --genomes 5 --divergence 12; any G that is not a
         multiple of N, or --shard contig): CONTIG-sharded ownership -- every rank inserts its contigs of every genome,
         common = AND over genomes of (OR over ranks) in one peer-memory kernel, every rank sketches its contigs, the
         tables are put back in contig order on rank 0; the masked refinement rounds are sketched by every rank on
         its own contigs too.  Strong scaling: the job is fixed.

`value` times the path with the packed genomes already resident in HBM; `e2e` times the same call
chain starting from packed genomes in PINNED HOST memory (H2D inside the timed region, in growing chunks so that
the first Bloom insert starts during the copy) and ending with the TSV text on the host.  Timing: CUDA events on the library's stream, bracketed by barriers,
max over ranks.  Inputs (1.5 GB of bases, 2 x 14.8 GB of filter) are far larger than the 126 MB L2,
so no explicit L2 flush is needed between iterations.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "genome bp/sec through sketch+BF+graph path"
K, W = 24, 1000


def presets(divergence):
    "bin/ntSynt:89-99"
    if divergence < 1:
        return dict(indel=10000, merge="10000", w_rounds=[100, 10], block_size=500)
    if divergence <= 10:
        return dict(indel=50000, merge="100000", w_rounds=[250, 100], block_size=1000)
    return dict(indel=100000, merge="1000000", w_rounds=[500, 250], block_size=10000)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- distributed plumbing
class Dist:
    "torch.distributed (gloo) is used only for rendezvous, barriers and max-over-ranks"

    def __init__(self):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.td = None
        if self.world > 1:
            import torch
            import torch.distributed as td
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            td.init_process_group(backend="gloo", rank=self.rank, world_size=self.world)
            self.td, self.torch = td, torch

    def barrier(self):
        if self.td:
            self.td.barrier()

    def max(self, x):
        if not self.td:
            return x
        t = self.torch.tensor([float(x)], dtype=self.torch.float64)
        self.td.all_reduce(t, op=self.td.ReduceOp.MAX)
        return float(t[0])

    def sum(self, x):
        if not self.td:
            return x
        t = self.torch.tensor([float(x)], dtype=self.torch.float64)
        self.td.all_reduce(t, op=self.td.ReduceOp.SUM)
        return float(t[0])

    def bcast_bytes(self, b, n):
        if not self.td:
            return b
        t = self.torch.zeros(n, dtype=self.torch.uint8)
        if self.rank == 0:
            t = self.torch.frombuffer(bytearray(b), dtype=self.torch.uint8).clone()
        self.td.broadcast(t, 0)
        return bytes(t.numpy().tobytes())

    def bcast_object(self, obj):
        "rank 0's object on every rank"
        if not self.td:
            return obj
        box = [obj]
        self.td.broadcast_object_list(box, src=0)
        return box[0]

    def gather_objects(self, obj):
        if not self.td:
            return [obj]
        out = [None] * self.world
        self.td.all_gather_object(out, obj)
        return out

    def close(self):
        if self.td:
            self.td.destroy_process_group()


def nccl_env():
    """NCCL's log goes to stderr (stdout carries the one JSON line): the communicator banner (ranks, NVLS, transports)
    stays visible to whoever reads the run's stderr"""
    os.environ.setdefault("NCCL_DEBUG", "INFO")
    os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")


def sha1_text(text):
    import hashlib
    return hashlib.sha1((text or "").encode()).hexdigest()


# ----------------------------------------------------------------------------- CPU arm
def cpu_sample_records(layout, G, sample_mbp_per_genome):
    """first slice of every contig of every genome, as ASCII records, written by the oracle's own statement of the
    workload generator (oracle/synth_oracle.c) -- the CPU arm never touches the CUDA library"""
    from oracle import sketch_oracle as so
    per = max(int(sample_mbp_per_genome * 1e6 / len(layout.names)), 20000)
    return [so.synth_records(layout, g, per_contig=per) for g in range(G)]


def cpu_sample_mbp(args, G):
    "per-genome CPU sample: --cpu-sample-mbp for 2 genomes, scaled so that the whole sample (and its run time) stays the same"
    return args.cpu_sample_mbp * 2.0 / max(G, 2)


def cpu_path(records_per_genome, file_names, divergence, threads, log=None):
    """The reference's CPU path restated (oracle/): make_common_bf (OpenMP over records, atomic byte-OR;
    src/ntsynt_make_common_bf.cpp:122-160), indexlr per genome (5 threads, 2 genomes at a time;
    bin/ntsynt_run_pipeline.smk:79-80, bin/ntSynt:154), then the graph stage (single-threaded Python,
    bin/ntsynt_synteny.py:33).  Returns (seconds, total bases, final TSV text)."""
    import ctypes as C
    import numpy as np
    from oracle import sketch_oracle as so
    from oracle.graph_oracle import GraphOracle
    L = so.lib()
    ps = presets(divergence)
    t0 = time.perf_counter()
    order = sorted(range(len(file_names)), key=lambda i: file_names[i])
    n0 = sum(len(s) for _, s in records_per_genome[order[0]])
    nbytes = so.bf_bytes(n0, 0.025)
    m = nbytes * 8

    def rec_arrays(recs):
        seqs = (C.c_char_p * len(recs))(*[s for _, s in recs])
        lens = (C.c_size_t * len(recs))(*[len(s) for _, s in recs])
        return seqs, lens
    bits = np.zeros(nbytes, dtype=np.uint8)
    seqs, lens = rec_arrays(records_per_genome[order[0]])
    L.orc_common_bf_level1(so._u8p(bits), m, seqs, lens, len(lens), K, threads)
    for i in order[1:]:
        nxt = np.zeros(nbytes, dtype=np.uint8)
        seqs, lens = rec_arrays(records_per_genome[i])
        L.orc_common_bf_cascade(so._u8p(bits), so._u8p(nxt), m, seqs, lens, len(lens), K, threads)
        bits = nxt
    if log:
        log(f"common Bloom filter ({nbytes} bytes) after {time.perf_counter() - t0:.1f} s")
    # round-0 indexlr: a worker per record, 5 threads per genome, 2 genomes at a time (smk:79-80, bin/ntSynt:154)
    tsv = [f"{fn}.k{K}.w{W}.tsv" for fn in file_names]
    round0 = {}

    def sketch_one(gi):
        recs = records_per_genome[gi]
        seqs, lens = rec_arrays(recs)
        caps = [max(64, 4 * (len(s) // W) + 64) for _, s in recs]
        h1 = [np.empty(c, dtype=np.uint64) for c in caps]
        ps_ = [np.empty(c, dtype=np.uint64) for c in caps]
        u64p = C.POINTER(C.c_uint64)
        h1p = (u64p * len(recs))(*[so._u64p(a) for a in h1])
        pp = (u64p * len(recs))(*[so._u64p(a) for a in ps_])
        capa = (C.c_size_t * len(recs))(*caps)
        cnt = (C.c_size_t * len(recs))()
        L.orc_sketch_records(seqs, lens, len(recs), K, W, so._u8p(bits), m, h1p, pp, capa, cnt, min(5, threads))
        round0[tsv[gi]] = [(recs[i][0], [(str(int(h)), int(p)) for h, p in zip(h1[i][:cnt[i]], ps_[i][:cnt[i]])])
                           for i in range(len(recs))]
    pending = list(range(len(file_names)))
    while pending:
        batch, pending = pending[:2], pending[2:]
        ths = [threading.Thread(target=sketch_one, args=(gi,)) for gi in batch]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
    if log:
        log(f"round-0 sketches after {time.perf_counter() - t0:.1f} s")
    go = GraphOracle(list(zip(tsv, records_per_genome)), K, W, ps["w_rounds"], ps["indel"], ps["merge"],
                     ps["block_size"], bits)
    try:
        text = go.run(round0=round0)
    except SystemExit:
        text = ""
    dt = time.perf_counter() - t0
    total = sum(len(s) for recs in records_per_genome for _, s in recs)
    return dt, total, text


def ingest_probe(layout, sample_mbp):
    """FASTA ingest, reported separately from the path: the first slice of every contig of one genome as
    60-column FASTA text in memory -> ntsynt_b200.fasta.parse_fasta_bytes (native scan + multi-threaded 2-bit pack)"""
    import numpy as np
    from ntsynt_b200 import fasta
    recs = cpu_sample_records(layout, 1, sample_mbp)[0]
    parts = []
    for name, seq in recs:
        n = len(seq) - len(seq) % 60
        a = np.frombuffer(seq, dtype=np.uint8)
        lines = np.empty((n // 60, 61), dtype=np.uint8)
        lines[:, :60] = a[:n].reshape(-1, 60)
        lines[:, 60] = 10
        parts += [b">" + name.encode() + b"\n", lines.tobytes(), seq[n:] + b"\n" if len(seq) > n else b""]
    text = b"".join(parts)
    best, bases = None, 0
    for _ in range(3):
        t0 = time.perf_counter()
        pk = fasta.parse_fasta_bytes(text)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        bases = pk.total_bases
    assert bases == sum(len(s) for _, s in recs)
    return {"value": bases / best, "unit": "bp/s", "threads": os.cpu_count() or 1,
            "sample": f"{bases} bp of genome 0 as 60-column FASTA text in memory ({len(text)} bytes), best of 3; "
                      f"not part of `value` / `e2e` (both sides of the comparison would pay it)"}


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args, dist):
    import numpy as np
    from ntsynt_b200 import device, pipeline, synth
    from ntsynt_b200.synteny import SyntenyEngine
    N = dist.world
    if N != args.gpus and dist.rank == 0 and N > 1:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE {N}", file=sys.stderr)
    ctx = device.Context(dist.local_rank)
    G = args.genomes or (2 if N == 1 else N)
    d = args.divergence
    ps = presets(d)
    wl = synth.Workload(G, int(args.genome_mbp * 1e6), d, seed=args.seed)
    file_names = [wl.file_name(g) for g in range(G)]
    names = [pipeline.tsv_name(f, K, W) for f in file_names]
    order = pipeline.processing_order(names)
    my_ids = list(range(G)) if N == 1 else [g for g in range(G) if g % N == dist.rank]
    if N > 1:
        shard = args.shard or ("genome" if G % N == 0 else "contig")
        if args.merge is None:
            args.merge = "owned" if shard == "contig" else "p2p"
        return run_ours_sharded(args, dist, ctx) if shard == "contig" else run_ours_multi(args, dist, ctx)
    gens = {g: wl.materialize(ctx, g) for g in my_ids}
    total_bp = sum(int(x.total_bases) for x in gens.values())
    size_sorted = sorted(range(G), key=lambda i: file_names[i])
    nbytes = device.BloomFilter.size_for(gens[size_sorted[0]].total_bases, 0.025)
    common, level = ctx.bloom(nbytes), ctx.bloom(nbytes)

    phase = {}

    def hot_path(gen_list):
        # src/ntsynt_make_common_bf.cpp:107-160 on resident genomes, filters re-zeroed every step
        t_hp = time.perf_counter()
        # zero-fills, insert x G, AND x (G - 2); the last level stays apart and the sketches look candidates up in both
        # filters (nts_bf_build_common_lazy / nts_sketch2) instead of one more pass over 2 x 14.8 GB
        apart = common.build_common(level, [gen_list[i] for i in size_sorted], K, lazy=True)
        phase["bf_wall_ms"] = phase.get("bf_wall_ms", 0.0) + (time.perf_counter() - t_hp) * 1e3
        be = pipeline.CudaBackend(ctx, [gen_list[i] for i in order], [names[i] for i in order], [wl.names] * G,
                                  [[int(x) for x in gen_list[i].lengths] for i in order], K, common=common,
                                  common2=level if apart else None)
        eng = SyntenyEngine(be, K, W, ps["w_rounds"], ps["indel"], ps["merge"], ps["block_size"], write_files=False,
                            quiet=True)
        text = eng.run()
        be.close()
        phase["total_wall_ms"] = phase.get("total_wall_ms", 0.0) + (time.perf_counter() - t_hp) * 1e3
        phase["calls"] = phase.get("calls", 0) + 1
        phase.setdefault("each_ms", []).append(round((time.perf_counter() - t_hp) * 1e3, 1))
        return text, eng

    gen_list = [gens[g] for g in range(G)]
    # ---- warm-up (untimed)
    for _ in range(max(args.warmup, 0)):
        text, eng = hot_path(gen_list)
    # ---- timed: inputs resident in HBM
    ctx.prof_enable(True)
    ctx.prof_reset()
    launches0 = ctx.launches
    clocks = ClockSampler(dist.local_rank)
    dist.barrier(); ctx.sync()
    ctx.timer_start()
    t_wall = time.perf_counter()
    for _ in range(args.steps):
        text, eng = hot_path(gen_list)
    ms = ctx.timer_stop()
    wall = time.perf_counter() - t_wall
    ctx.sync(); dist.barrier()
    ms = dist.max(ms)
    launches = ctx.launches - launches0
    prof = ctx.prof()
    ctx.prof_enable(False)
    value = dist.sum(total_bp) * args.steps / (ms / 1e3)

    # ---- e2e: packed genomes start in pinned host memory; result text ends on the host
    packed_host = []
    for g in range(G):
        pk = gens[g].to_packed()
        pin = device.PinnedU64(len(pk.words))
        pin.array[:] = pk.words
        pk.words = pin.array
        packed_host.append((pk, pin))
    ctx.prof_reset()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))

    def e2e_step():
        fresh = [ctx.upload(pk, async_copy=True) for pk, _ in packed_host]     # H2D of genome i+1 overlaps insert i
        out, _ = hot_path(fresh)
        for f in fresh:
            f.close()
        return out
    e2e_step()                                   # untimed: first-use costs of this leg (copy stream, device blocks)
    ctx.prof_reset()                             # (also zeroes the transfer counters)
    dist.barrier(); ctx.sync()
    ctx.timer_start()
    for _ in range(e2e_steps):
        text_e2e = e2e_step()
    ms_e2e = dist.max(ctx.timer_stop())
    h2d, d2h = ctx.xfer_bytes()
    clk = clocks.stop()
    e2e_value = dist.sum(total_bp) * e2e_steps / (ms_e2e / 1e3)
    assert text_e2e == text, "e2e and resident runs disagree"

    # ---- roofline of the dominant kernel family (CUDA events around every launch, same timed region)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json"), encoding="utf-8") as fh:
            peaks = json.load(fh)
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    # algorithmic bytes per unit: S = 0.25 B per packed base, 32 B sector.  One Bloom insert is the
    # kernel pair bf_bin_kernel + bf_apply_kernel (timed separately on their own streams; a "launch" = one pair);
    # SURVEY's figure for it is S + 64 B per k-mer (sector read + write-back of one random bit set).
    alg = {"bf_insert": 64.25, "sketch": 32.25, "bf_combine": 3.0, "fill": 1.0}
    fams = {f: prof[f] for f in alg if prof.get(f, (0, 0, 0))[2]}
    if prof.get("bf_part1", (0, 0, 0))[2]:
        # one partitioned insert = bf_part1 + bf_part2 + bf_apply(+overflow) (three passes, timed separately)
        fams["bf_insert"] = (prof["bf_part1"][0] + prof["bf_part2"][0] + prof["bf_apply"][0], prof["bf_part1"][1],
                             prof["bf_part1"][2])
    fam = max(fams, key=lambda f: fams[f][0])
    f_ms, f_units, f_n = fams[fam]
    bytes_per_launch = alg[fam] * f_units / f_n
    achieved = bytes_per_launch / ((f_ms / f_n) / 1e3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json"), encoding="utf-8") as fh:
            tj = json.load(fh).get(fam, {})
        if abs(tj.get("genome_mbp", 0) - args.genome_mbp) < 1:
            traffic = tj.get("dram_bytes_per_launch")
    except (OSError, ValueError):
        pass
    three_pass = bool(prof.get("bf_part2", (0, 0, 0))[2])
    insert_name = ("bf_part1_kernel + bf_part2_kernel + bf_apply3_kernel (one Bloom insert, NTS_BF_IMPL=3)" if three_pass
                   else "bf_bin_kernel<512,16> + bf_apply_kernel (one Bloom insert)")
    sm_clock = (clk.get("sm_mhz") or 1965.0) * 1e6
    issue_peak = 148 * 128 * sm_clock                           # thread-instructions per second: 148 SMs x 4 schedulers x 32 lanes
    roofline = {"kernel": {"bf_insert": insert_name, "sketch": "sketch_sparse_kernel<512,16,3072>"}.get(fam, fam),
                "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "frac_of_nominal_8TBs": round(achieved / 8000.0, 4),
                "traffic": traffic, "peak_source": peak_src,
                # what the kernel really moves (ncu dram__bytes of one launch) over the same launch time: the contract figure
                # above assumes one random 32-byte sector read-modify-write per k-mer, which the partitioned insert avoids
                "dram_frac": round(traffic / ((f_ms / f_n) / 1e3) / 1e9 / peak, 4) if traffic else None,
                "algorithmic_bytes_per_launch": bytes_per_launch, "avg_launch_ms": f_ms / f_n,
                "kernel_ms_per_step": {f: round(prof[f][0] / args.steps, 3) for f in prof if prof[f][2]},
                "kernel_share_of_step": round(f_ms / ms, 4),
                "note": ("bf_insert = bf_part1 (binning pass) + bf_apply (the two passes of one Bloom insert, also listed separately); "
                         "the pair is bound by atomic issue (one shared-memory ATOMS and one global RED per k-mer at 2 LSU cycles "
                         "per lane each), not by HBM -- see dram_frac and profiles/")}
    # the sketch kernel is instruction-bound (it looks up 2.4 % of the k-mers): report it against the issue rate
    sk = prof.get("sketch", (0, 0, 0))
    sketch_roof = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json"), encoding="utf-8") as fh:
            ipk = json.load(fh).get("sketch", {}).get("thread_instructions_per_kmer")
    except (OSError, ValueError):
        ipk = None
    if sk[2] and ipk:
        ach = ipk * sk[1] / (sk[0] / 1e3)
        sketch_roof = {"kernel": "sketch_sparse_kernel<512,16,3072> (round-0 sketch of one genome)", "bound": "issue",
                       "thread_instructions_per_kmer": ipk, "achieved": round(ach / 1e12, 3), "peak": round(issue_peak / 1e12, 3),
                       "unit": "T thread-instr/s", "frac": round(ach / issue_peak, 4), "avg_launch_ms": round(sk[0] / sk[2], 3),
                       "note": "instructions per k-mer from the committed ncu capture (profiles/traffic.json); launches include the small masked rounds"}
    roofline["sketch_issue"] = sketch_roof

    # ---- CPU baseline on rank 0 (bounded sample of the same workload)
    cpu = None
    if dist.rank == 0 and N == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        smbp = cpu_sample_mbp(args, G)
        recs = cpu_sample_records(wl, G, smbp)
        dt, tot, _ = cpu_path(recs, file_names, d, threads)
        cpu = {"value": tot / dt, "unit": "bp/s", "cores": threads, "kind": "port",
               "sample": f"first {smbp:g} Mbp of each of the {G} genomes ({tot} bp): oracle/ C+OpenMP "
                         f"Bloom filter and sketch, pure-Python graph stage; {dt:.1f} s"}
    ingest = ingest_probe(wl, 600.0) if (dist.rank == 0 and N == 1 and not args.no_cpu) else None
    line = {
        "metric": METRIC, "value": value, "unit": "bp/s", "n_gpus": N, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"{G} synthetic ~{args.genome_mbp:g} Mbp human-like genomes, d={d:g}, k={K} w={W}, "
                               f"w_rounds {ps['w_rounds']}, 1xB200" if N == 1 else
                               f"{G} synthetic {args.genome_mbp:g} Mbp genomes one-per-GPU, counting-BF NCCL-sum merge",
                   "genomes": G, "genome_bp": total_bp // max(len(my_ids), 1), "k": K, "w": W, "fpr": 0.025,
                   "bloom_bytes": nbytes, "l2": "inputs (bases + filters) are larger than L2; no flush needed",
                   "blocks": text.count("\n") // G, "blocks_sha1": sha1_text(text), "vertices": eng.stats.get("vertices")},
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": "bp/s", "h2d_bytes_per_step": h2d // e2e_steps,
                "d2h_bytes_per_step": d2h // e2e_steps, "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "ingest": ingest,
        "host_wall_ms_per_step": wall * 1e3 / args.steps,
        "graph_stage_phase_ms": {k[2:]: round(v * 1e3, 1) for k, v in eng.stats.items() if k.startswith("t_")},
        "hot_path_wall_ms": {k: (v if k == "each_ms" else round(v / max(phase.get("calls", 1), 1), 1)) for k, v in phase.items() if k != "calls"},
    }
    if dist.rank == 0:
        print(json.dumps(line))


def run_ours_multi(args, dist, ctx):
    nccl_env()
    """N > 1: one genome per GPU (G = N, or a multiple), per-GPU filters merged by NCCL all-reduce(sum)
    of packed counters, owners sketch, tables all-gathered, rank 0 runs the join + graph stage."""
    import numpy as np
    from ntsynt_b200 import device, distributed, pipeline, synth
    from ntsynt_b200.synteny import SyntenyEngine
    N, rank = dist.world, dist.rank
    G = args.genomes or N
    if G % N:
        raise SystemExit("one-genome-per-GPU sharding needs --genomes to be a multiple of the number of GPUs (use --shard contig)")
    d = args.divergence
    ps = presets(d)
    wl = synth.Workload(G, int(args.genome_mbp * 1e6), d, seed=args.seed)
    file_names = [wl.file_name(g) for g in range(G)]
    names = [pipeline.tsv_name(f, K, W) for f in file_names]
    order = pipeline.processing_order(names)
    own = [g for g in range(G) if g % N == rank]
    resident = list(range(G)) if rank == 0 else own          # rank 0 re-sketches the masked rounds of every genome
    gens = {g: wl.materialize(ctx, g) for g in resident}
    sizes = [int(wl.segments(g)[0].sum()) for g in range(G)]
    total_bp = sum(sizes)
    first = sorted(range(G), key=lambda i: file_names[i])[0]
    nbytes = device.BloomFilter.size_for(sizes[first], 0.025)
    mine, level = ctx.bloom(nbytes), (ctx.bloom(nbytes) if len(own) > 1 else None)
    ident = distributed.Comm.new_unique_id() if rank == 0 else b""
    comm = distributed.Comm(ctx, rank, N, dist.bcast_bytes(ident, 128))
    peer = distributed.PeerMerge(mine, rank, N, dist.gather_objects, dist.barrier) if N <= 16 else None
    if peer is not None and not peer.ok:           # no peer access on some rank: the NCCL form is the fallback
        peer = None
    use_p2p = args.merge in ("p2p", "owned") and peer is not None     # (owned not available: the peer-memory merge)

    phase = {}

    def tick(name, t0):
        phase[name] = phase.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
        return time.perf_counter()

    def hot_path(gen_map):
        t0 = time.perf_counter()
        mine.set_genome(gen_map[own[0]], K)
        for g in own[1:]:
            level.clear(); level.insert_genome(gen_map[g], K); mine.iand(level)
        t0 = tick("insert", t0)
        if use_p2p:
            peer.merge("and", comm=comm)                           # NVLink peer loads: reduce-scatter + all-gather
        else:
            comm.allreduce_and(mine)                               # the one bulk exchange (NCCL sum of counters)
        gathered = {}
        for slot, g in enumerate(own):
            t = ctx.sketch(gen_map[g], K, W, common=mine)
            counts = dist.gather_objects(len(t))
            owners = [slot * N + r for r in range(N)]
            tabs = comm.allgather_tables(t, counts, [gen_map.get(x) for x in owners])
            t.close()
            for x, tb in zip(owners, tabs):
                if rank == 0:
                    gathered[x] = tb
                else:
                    tb.close()
        t0 = tick("merge_sketch_gather", t0)                       # (the merge is asynchronous: it is waited for here)
        text, eng = None, None
        if rank == 0:
            be = distributed.GatheredBackend(ctx, [gen_map[i] for i in order], [names[i] for i in order], [wl.names] * G,
                                             [[int(x) for x in gen_map[i].lengths] for i in order], K, mine,
                                             [gathered[i] for i in order])
            eng = SyntenyEngine(be, K, W, ps["w_rounds"], ps["indel"], ps["merge"], ps["block_size"], write_files=False,
                                quiet=True)
            text = eng.run()
            for tb in gathered.values():
                tb.close()
            t0 = tick("graph_rank0", t0)
        dist.barrier()                                             # the step ends when the block table exists
        tick("wait_for_rank0", t0)
        phase["calls"] = phase.get("calls", 0) + 1
        return text, eng

    for _ in range(max(args.warmup, 0)):
        text, eng = hot_path(gens)
    ctx.prof_enable(True); ctx.prof_reset()
    phase.clear()
    launches0 = ctx.launches
    clocks = ClockSampler(dist.local_rank)
    dist.barrier(); ctx.sync()
    ctx.timer_start()
    for _ in range(args.steps):
        text, eng = hot_path(gens)
    ms = ctx.timer_stop()
    phase = dict(phase)                    # freeze: the e2e leg below would add to it
    hot_phase, phase = phase, {}
    ctx.sync(); dist.barrier()
    ms = dist.max(ms)
    launches = dist.sum(ctx.launches - launches0)
    prof = ctx.prof()
    ctx.prof_enable(False)
    value = total_bp * args.steps / (ms / 1e3)
    # e2e: owners (and rank 0 for every genome) start from pinned host memory
    packed_host = {}
    for g in resident:
        pk = gens[g].to_packed()
        pin = device.PinnedU64(len(pk.words)); pin.array[:] = pk.words; pk.words = pin.array
        packed_host[g] = (pk, pin)
    ctx.prof_reset()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    dist.barrier(); ctx.sync()
    ctx.timer_start()
    for _ in range(e2e_steps):
        fresh = {g: ctx.upload(packed_host[g][0], async_copy=True) for g in own + [x for x in resident if x not in own]}
        text_e2e, _ = hot_path(fresh)
        for f in fresh.values():
            f.close()
    ms_e2e = dist.max(ctx.timer_stop())
    h2d, d2h = ctx.xfer_bytes()
    h2d, d2h = dist.sum(h2d), dist.sum(d2h)
    clk = clocks.stop()
    if rank == 0:
        assert text_e2e == text
    # the exchange alone, both implementations (same bits: tests/test_gpu_multi.py), max over ranks
    merge_ms = {}
    for name in ("nccl", "p2p"):
        if name == "p2p" and peer is None:
            continue
        best = None
        for _ in range(2):
            mine.set_genome(gens[own[0]], K)
            ctx.sync(); dist.barrier()
            t0 = time.perf_counter()
            if name == "nccl":
                comm.allreduce_and(mine)
            else:
                peer.merge("and", comm=comm)
            ctx.sync()
            dt = dist.max((time.perf_counter() - t0) * 1e3)
            best = dt if best is None else min(best, dt)
        merge_ms[name] = round(best, 2)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json"), encoding="utf-8") as fh:
            peaks = json.load(fh)
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    alg = {"bf_insert": 64.25, "sketch": 32.25, "bf_combine": 3.0, "fill": 1.0}
    fam = max((f for f in alg if prof[f][2]), key=lambda f: prof[f][0])
    f_ms, f_units, f_n = prof[fam]
    achieved = (alg[fam] * f_units / f_n) / ((f_ms / f_n) / 1e3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json"), encoding="utf-8") as fh:
            tj = json.load(fh).get(fam, {})
        if abs(tj.get("genome_mbp", 0) - args.genome_mbp) < 1:
            traffic = tj.get("dram_bytes_per_launch")
    except (OSError, ValueError):
        pass
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "bp/s", "n_gpus": N, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"{G} synthetic {args.genome_mbp:g} Mbp genomes one-per-GPU, Bloom filters merged over NVLink "
                                   f"({'peer-memory AND kernels' if use_p2p else 'counting-BF NCCL-sum'}), "
                                   f"d={d:g}, k={K} w={W}, w_rounds {ps['w_rounds']}, {N}xB200",
                       "genomes": G, "genome_bp": sizes[0], "k": K, "w": W, "fpr": 0.025, "bloom_bytes": nbytes,
                       "merge": "p2p" if use_p2p else "nccl",
                       "merge_wire_bytes_per_rank": distributed.merge_wire_bytes(nbytes, N),
                       "merge_alone_ms": merge_ms,
                       "l2": "inputs (bases + filters) are larger than L2; no flush needed",
                       "blocks": text.count("\n") // G, "blocks_sha1": sha1_text(text), "vertices": eng.stats.get("vertices")},
            "clocks": clk,
            "e2e": {"value": total_bp * e2e_steps / (ms_e2e / 1e3), "unit": "bp/s", "h2d_bytes_per_step": int(h2d // e2e_steps),
                    "d2h_bytes_per_step": int(d2h // e2e_steps), "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps},
            "gpu_launches": int(launches),
            "roofline": {"kernel": fam, "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "frac_of_nominal_8TBs": round(achieved / 8000.0, 4),
                         "traffic": traffic, "rank": 0,
                         "kernel_ms_per_step": {f: round(prof[f][0] / args.steps, 3) for f in prof if prof[f][2]}},
            "cpu_baseline": None,
            "graph_stage_phase_ms": {k[2:]: round(v * 1e3, 1) for k, v in eng.stats.items() if k.startswith("t_")},
            "phase_ms_rank0": {k: round(v / max(hot_phase.get("calls", 1), 1), 2) for k, v in hot_phase.items() if k != "calls"},
        }))
    if peer is not None:
        peer.close()
    comm.close()



def run_ours_sharded(args, dist, ctx):
    """N > 1, contig-sharded ownership (BASELINE configs 3 and 4): every rank holds its contigs of every
    genome; per genome it builds the bits of those contigs; common = AND over genomes of (OR over ranks) with the
    peer-memory kernel (or, --merge nccl, one NCCL counter all-reduce per genome and a local AND); every rank sketches
    its contigs; the tables are all-gathered and put back in contig order; rank 0 runs the join + graph stage; the masked
    refinement rounds are sketched by every rank on its own contigs as well (distributed.ShardedSketcher)."""
    nccl_env()
    import numpy as np
    from ntsynt_b200 import device, distributed, pipeline, synth
    from ntsynt_b200.synteny import SyntenyEngine
    N, rank = dist.world, dist.rank
    G = args.genomes or 3
    d = args.divergence
    ps = presets(d)
    wl = synth.Workload(G, int(args.genome_mbp * 1e6), d, seed=args.seed)
    file_names = [wl.file_name(g) for g in range(G)]
    names = [pipeline.tsv_name(f, K, W) for f in file_names]
    order = pipeline.processing_order(names)
    n_contigs = len(wl.names)
    own_of = distributed.assign_contigs(wl.anc_lengths, N)
    owner_of_contig = [next(r for r in range(N) if c in own_of[r]) for c in range(n_contigs)]
    mine_c = own_of[rank]
    shards = [wl.materialize(ctx, g, contigs=mine_c) for g in range(G)]
    whole = None                                       # (rank 0 no longer keeps whole genomes: the masked rounds are sharded too)
    contig_lengths = [wl.segments(g)[0] for g in range(G)]
    sizes = [int(x.sum()) for x in contig_lengths]
    total_bp = sum(sizes)
    first = sorted(range(G), key=lambda i: file_names[i])[0]
    nbytes = device.BloomFilter.size_for(sizes[first], 0.025)
    common = ctx.bloom(nbytes)
    ident = distributed.Comm.new_unique_id() if rank == 0 else b""
    comm = distributed.Comm(ctx, rank, N, dist.bcast_bytes(ident, 128))
    owned = None
    if args.merge == "owned" and N <= 16:
        plan_valid = int(dist.max(max(int(x.total_bases) for x in shards)))
        owned = distributed.OwnedBuild(common, ctx.bloom(nbytes), rank, N, K, plan_valid, dist.gather_objects, dist.barrier, comm)
        if not owned.ok:
            owned = None
    parts = [ctx.bloom(nbytes) for _ in range(G)] if owned is None else []
    peer = (distributed.ShardedMerge(parts, common, rank, N, dist.gather_objects, dist.barrier)
            if owned is None and N <= 16 and G <= 8 else None)
    if peer is not None and not peer.ok:
        peer = None
    use_p2p = args.merge in ("p2p", "owned") and peer is not None     # (owned not available: the peer-memory merge)
    phase = {}

    def tick(name, t0):
        phase[name] = phase.get(name, 0.0) + (time.perf_counter() - t0) * 1e3

    def hot_path(shard_list, whole_list):
        t0 = time.perf_counter()
        if owned is not None:
            over = owned.build(shard_list)                         # bin everywhere, apply on the owner over peer memory, all-gather
            if dist.sum(over):
                raise SystemExit("owned build: bucket overflow (heavy-hitter k-mers); use --merge p2p")
            ctx.sync(); tick("owned_build", t0); t0 = time.perf_counter()
        for g in (range(G) if owned is None else ()):
            parts[g].set_genome(shard_list[g], K)                  # bits of my contigs of genome g
        if owned is None:
            ctx.sync(); tick("insert", t0); t0 = time.perf_counter()
        if owned is not None:
            pass
        elif use_p2p:
            peer.merge(comm=comm)                                  # AND_g OR_r over NVLink peer memory, then all-gather
        else:
            for g in range(G):                                     # north-star form: counter all-reduce per genome
                comm.allreduce_or(parts[g])
            common.build_from_and(parts)
        ctx.sync(); tick("merge", t0); t0 = time.perf_counter()
        tables = []
        for g in range(G):
            t = ctx.sketch(shard_list[g], K, W, common=common)
            tables.append(distributed.gather_sharded_table(comm, t, n_contigs, owner_of_contig, dist.gather_objects))
            t.close()
        tick("sketch_gather", t0); t0 = time.perf_counter()
        text, eng = None, None
        # the masked refinement rounds are sketched by every rank on its own contigs (distributed.ShardedSketcher)
        svc = distributed.ShardedSketcher(comm, ctx, shard_list, K, common, n_contigs, owner_of_contig, dist.bcast_object,
                                          dist.gather_objects)
        if rank == 0:
            be = distributed.GatheredBackend(ctx, None, [names[i] for i in order], [wl.names] * G,
                                             [[int(x) for x in contig_lengths[i]] for i in order], K, common,
                                             [tables[i] for i in order],
                                             masked_sketch=lambda a, w_, masks: svc.sketch(order[a], w_, masks))
            eng = SyntenyEngine(be, K, W, ps["w_rounds"], ps["indel"], ps["merge"], ps["block_size"], write_files=False,
                                quiet=True)
            try:
                text = eng.run()
            finally:
                svc.done()                                         # (the other ranks sit in serve(): never leave them there)
            be.close()
        else:
            svc.serve()
        for t in tables:
            t.close()
        dist.barrier()
        tick("graph", t0)
        return text, eng

    for _ in range(max(args.warmup, 0)):
        text, eng = hot_path(shards, whole)
    ctx.prof_enable(True); ctx.prof_reset()
    phase.clear()
    launches0 = ctx.launches
    clocks = ClockSampler(dist.local_rank)
    dist.barrier(); ctx.sync()
    ctx.timer_start()
    for _ in range(args.steps):
        text, eng = hot_path(shards, whole)
    ms = ctx.timer_stop()
    ctx.sync(); dist.barrier()
    ms = dist.max(ms)
    launches = dist.sum(ctx.launches - launches0)
    prof = ctx.prof()
    ctx.prof_enable(False)
    phase_ms = {k: round(v / args.steps, 2) for k, v in phase.items()}
    value = total_bp * args.steps / (ms / 1e3)
    # e2e: the shards start in pinned host memory
    def pin(gen):
        pk = gen.to_packed()
        pb = device.PinnedU64(len(pk.words)); pb.array[:] = pk.words; pk.words = pb.array
        return pk, pb
    pinned_sh = [pin(x) for x in shards]
    ctx.prof_reset()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    dist.barrier(); ctx.sync()
    ctx.timer_start()
    for _ in range(e2e_steps):
        fs = [ctx.upload(pk, async_copy=True) for pk, _ in pinned_sh]
        text_e2e, _ = hot_path(fs, None)
        for f in fs:
            f.close()
    ms_e2e = dist.max(ctx.timer_stop())
    h2d, d2h = ctx.xfer_bytes()
    h2d, d2h = dist.sum(h2d), dist.sum(d2h)
    clk = clocks.stop()
    if rank == 0:
        assert text_e2e == text
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json"), encoding="utf-8") as fh:
            peaks = json.load(fh)
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    pair = prof.get("bf_part1", (0, 0, 0))
    f_ms = pair[0] + prof.get("bf_apply", (0, 0, 0))[0]
    achieved = (64.25 * pair[1] / max(pair[2], 1)) / ((f_ms / max(pair[2], 1)) / 1e3) / 1e9 if pair[2] else 0.0
    wire = nbytes * (N - 1) / N * (G + 1)
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "bp/s", "n_gpus": N, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"{G} synthetic ~{args.genome_mbp:g} Mbp genomes, contigs sharded over {N}xB200, d={d:g}, "
                                   f"k={K} w={W}, w_rounds {ps['w_rounds']}; common = AND_g OR_rank over NVLink "
                                   f"({'owned build: buckets applied by the slice owner over peer memory, no merge' if owned is not None else 'peer-memory kernel' if use_p2p else 'NCCL counter all-reduce per genome'})",
                       "genomes": G, "genome_bp": sizes[0], "k": K, "w": W, "fpr": 0.025, "bloom_bytes": nbytes,
                       "shard": "contig", "contigs_per_rank": [len(x) for x in own_of],
                       "bases_per_rank_of_genome0": [int(sum(int(wl.segments(0)[0][c]) for c in x)) for x in own_of],
                       "merge": "owned" if owned is not None else "p2p" if use_p2p else "nccl",
                       "merge_wire_bytes_per_rank": int(wire) if owned is None else int(nbytes * (N - 1) / N + 4 * total_bp * (N - 1) / N / N),
                       "phase_ms_rank0": phase_ms,
                       "l2": "inputs (bases + filters) are larger than L2; no flush needed",
                       "blocks": text.count("\n") // G, "blocks_sha1": sha1_text(text), "vertices": eng.stats.get("vertices"),
                       "refinement_new_minimizers": eng.stats.get("new_raw"), "unmasked_fraction": eng.stats.get("unmasked")},
            "clocks": clk,
            "e2e": {"value": total_bp * e2e_steps / (ms_e2e / 1e3), "unit": "bp/s", "h2d_bytes_per_step": int(h2d // e2e_steps),
                    "d2h_bytes_per_step": int(d2h // e2e_steps), "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "bf_bin_kernel + bf_apply_kernel (one Bloom insert of this rank's contigs)", "bound": "hbm",
                         "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                         "traffic": None, "rank": 0,
                         "kernel_ms_per_step": {f: round(prof[f][0] / args.steps, 3) for f in prof if prof[f][2]}},
            "cpu_baseline": None,
            "graph_stage_phase_ms": {k[2:]: round(v * 1e3, 1) for k, v in eng.stats.items() if k.startswith("t_")},
        }))
    if peer is not None:
        peer.close()
    if owned is not None:
        owned.close()
    comm.close()

def run_reference(args, dist):
    """CPU restatement of the reference path on a bounded sample of the same workload (rank 0 only).  Nothing of the
    product is imported here: the sample comes from oracle/synth_oracle.c over ntsynt_b200/synth_layout.py (pure numpy),
    the path from oracle/ntsynt_oracle.c and oracle/graph_oracle.py."""
    if dist.rank != 0:
        return
    from ntsynt_b200 import synth_layout
    N = dist.world
    G = args.genomes or (2 if N == 1 else N)
    d = args.divergence
    lay = synth_layout.Layout(G, int(args.genome_mbp * 1e6), d, seed=args.seed)
    smbp = cpu_sample_mbp(args, G)
    recs = cpu_sample_records(lay, G, smbp)
    file_names = [lay.file_name(g) for g in range(G)]
    threads = os.cpu_count() or 1
    for _ in range(args.warmup):
        cpu_path(recs, file_names, d, threads)
    t = 0.0
    for _ in range(args.steps):
        dt, tot, text = cpu_path(recs, file_names, d, threads)
        t += dt
    v = tot * args.steps / t
    sample = (f"first {smbp:g} Mbp of each of the {G} genomes ({tot} bp) per step -- a bp/s extrapolation, not a same-size "
              f"run (the filter is sized for the sample); oracle/ C+OpenMP Bloom filter + sketch with the reference's "
              f"threading structure and pure-Python graph stage (btllib / python-igraph are not installable); the "
              f"per-k-mer substr of indexlr --seq is not charged to this arm")
    ps = presets(d)
    assert "ntsynt_b200._lib" not in sys.modules, "the reference arm must not load the CUDA library"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "bp/s", "n_gpus": N, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"{G} synthetic ~{args.genome_mbp:g} Mbp human-like genomes, d={d:g}, k={K} w={W}, "
                               f"w_rounds {ps['w_rounds']} (bounded sample)", "genomes": G, "k": K, "w": W,
                   "sample_bp": tot, "blocks_sha1_of_sample": sha1_text(text)},
        "cpu_baseline": {"value": v, "unit": "bp/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "bp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--genome-mbp", type=float, default=3000.0, help="size of each synthetic genome")
    ap.add_argument("--genomes", type=int, default=0, help="number of genomes (default 2 at N=1, N otherwise)")
    ap.add_argument("--divergence", type=float, default=1.0)
    ap.add_argument("--seed", type=int, default=20260117)
    ap.add_argument("--cpu-sample-mbp", type=float, default=192.0,
                    help="per-genome sample for the CPU arm at 2 genomes (192 Mbp x 2 is ~12 s of CPU work on 16 cores); "
                         "with G genomes each gets 2/G of it, so the arm costs the same at every N")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--shard", choices=["genome", "contig"], default=None,
                    help="multi-GPU ownership: one genome per GPU (default when --genomes is a multiple of --gpus) or the "
                         "contigs of every genome spread over the GPUs (default otherwise: BASELINE configs 3 and 4)")
    ap.add_argument("--merge", choices=["nccl", "p2p", "owned"], default=None,
                    help="multi-GPU filter merge: peer-memory reduce-scatter/all-gather kernels over NVLink (default; "
                         "bit-identical and ~4x less wire volume), NCCL all-reduce(sum) of packed counters (the "
                         "north-star form; also timed alone in config.merge_alone_ms), or -- contig-sharded runs -- no "
                         "merge at all: hash-range owned build (every GPU bins, the owner of a filter slice applies "
                         "every GPU's buckets over peer memory).  Default: owned for contig-sharded runs, p2p otherwise")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        print("note: fewer than 3 warm-up steps; the number is not reportable", file=sys.stderr)
    dist = Dist()
    try:
        if args.impl == "reference":
            run_reference(args, dist)
        else:
            run_ours(args, dist)
    finally:
        dist.close()


if __name__ == "__main__":
    main()
