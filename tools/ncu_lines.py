"""Aggregate an ncu report's source page by source line: samples and executed warp instructions per line.
usage: python tools/ncu_lines.py report.ncu-rep [top]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                     text=True).stdout
cur_file = None
agg = defaultdict(lambda: [0, 0, ""])
for row in csv.reader(io.StringIO(out)):
    if not row:
        continue
    if row[0] == "File Path":
        cur_file = row[1].split("/")[-1]
        continue
    if row[0] == "Function Name":
        continue
    if row[0] == "Line No":
        hdr = row
        i_s, i_n = hdr.index("# Samples"), hdr.index("Instructions Executed")
        continue
    try:
        ln = int(row[0])
    except ValueError:
        continue
    key = (cur_file, ln)
    def num(x):
        try:
            return int(float(x))
        except ValueError:
            return 0
    agg[key][0] += num(row[i_s])
    agg[key][1] += num(row[i_n])
    agg[key][2] = row[1].strip()[:110]
tot_s = sum(v[0] for v in agg.values()) or 1
tot_i = sum(v[1] for v in agg.values()) or 1
print(f"total samples {tot_s}  total warp-instructions {tot_i}")
for (f, ln), (s, n, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*s/tot_s:5.1f}% smp {100*n/tot_i:5.1f}% ins  {f}:{ln}  {src}")
