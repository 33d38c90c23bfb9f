#!/usr/bin/env python3
"summarise an .ncu-rep (raw page key metrics + top stall lines of the source page) as markdown"
import csv, io, subprocess, sys
rep = sys.argv[1]
sel = ["--kernel-name", "regex:" + sys.argv[2], "--launch-count", "1"] if len(sys.argv) > 2 else []   # one kernel of the report
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", *sel], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]
for r in rows[2:]:
    print(f"### {r[idx['Kernel Name']][:110]}\n")
    print("| metric | value | unit |\n|---|---|---|")
    for w in want:
        if w in idx:
            print(f"| {w} | {r[idx[w]]} | {units[idx[w]]} |")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", *sel], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
if hi:
    h = rows[hi[0]]
    i_src, i_s, i_ex = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
    def num(x):
        try: return int(x)
        except ValueError: return 0
    end = hi[1] if len(hi) > 1 else len(rows)                 # first launch of the selection only
    rows = rows[:end]
    data = [(num(r[i_s]), r[i_src].strip(), num(r[i_ex])) for r in rows[hi[0] + 1:] if len(r) > i_ex and r[0].startswith("0x")]
    tot = sum(d[0] for d in data) or 1
    stall_cols = [(j, c) for j, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    sums = {c: sum(num(r[j]) for r in rows[hi[0] + 1:] if len(r) > j) for j, c in stall_cols}
    print("\nstall reasons (samples): " + ", ".join(f"{c}={v}" for c, v in sorted(sums.items(), key=lambda x: -x[1])[:7]))
    print(f"\ntotal instructions executed (warp-level): {sum(d[2] for d in data)}; SASS lines {len(data)}\n")
    print("| stall % | executed | SASS |\n|---|---|---|")
    for s, text, ex in sorted(data, key=lambda x: -x[0])[:16]:
        print(f"| {100 * s / tot:.1f} | {ex} | `{text[:100]}` |")
