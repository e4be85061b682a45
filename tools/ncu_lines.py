"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line.
Usage: python tools/ncu_lines.py report.ncu-rep [top_n]   (prints file:line, warp instructions, share, samples, source)"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None
lines = []
col = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Line No":
        col = {n: i for i, n in enumerate(r)}
    elif r[0] and r[0].isdigit() and cur_file:
        try:
            lines.append((cur_file, int(r[0]), r[1].strip(), int(r[col["Instructions Executed"]]), int(r[col["# Samples"]]),
                          int(r[col["Thread Instructions Executed"]])))
        except (ValueError, KeyError):
            pass
tot = sum(l[3] for l in lines)
smp = sum(l[4] for l in lines)
print(f"total warp instructions {tot}, samples {smp}")
for f, n, s, inst, sa, ti in sorted(lines, key=lambda l: -l[3])[:top]:
    print(f"{f}:{n:<4} inst {inst:>10} {100*inst/tot:5.1f}%  samples {100*sa/max(smp,1):5.1f}%  thr/inst {ti/max(inst,1):4.1f}  | {s[:90]}")
