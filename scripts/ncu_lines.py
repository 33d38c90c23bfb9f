#!/usr/bin/env python3
"per-CUDA-source-line instruction and stall shares from an .ncu-rep (needs -lineinfo and --import-source on)"
import csv, io, subprocess, sys
rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, text=True).stdout
cur, hdr, out = None, None, []
for r in csv.reader(io.StringIO(txt)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr and r[0] not in ("", "Function Name") and len(r) >= 8:
        try:
            out.append((cur, int(r[0]), r[1].strip(), int(r[7]), int(r[4])))
        except ValueError:
            pass
ti, ts = sum(o[3] for o in out) or 1, sum(o[4] for o in out) or 1
print(f"total warp instructions {ti}, stall samples {ts}\n")
print("| inst % | stall % | file:line | source |\n|---|---|---|---|")
for f, ln, src, ins, st in sorted(out, key=lambda o: -o[3])[:top]:
    print(f"| {100 * ins / ti:.1f} | {100 * st / ts:.1f} | {f}:{ln} | `{src[:95]}` |")
