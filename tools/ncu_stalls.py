"""Top SASS instructions of an ncu report by one stall reason, with their source line.
usage: python tools/ncu_stalls.py report.ncu-rep stall_long_sb [top]"""
import csv
import io
import subprocess
import sys

rep, reason = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur, recs, hdr, mode = None, [], None, None
for row in rows:
    if not row:
        continue
    if row[0] == "File Path":
        cur = row[1].split("/")[-1]
        continue
    if row[0] == "Line No":
        hdr, mode = row, "cuda"
        continue
    if row[0] == "Address":
        hdr, mode = row, "sass"
        continue
    if hdr is None or len(row) != len(hdr):
        continue
    d = dict(zip(hdr, row))
    try:
        v = int(float(d.get(reason, "0") or 0))
    except ValueError:
        continue
    if v:
        recs.append((v, mode, cur if mode == "cuda" else "", d.get("Line No", d.get("Address", "")), d.get("Source", "").strip()[:120]))
for mode in ("cuda", "sass"):
    sel = sorted([r for r in recs if r[1] == mode], reverse=True)[:top]
    tot = sum(r[0] for r in recs if r[1] == mode) or 1
    print(f"== {reason} by {mode} (total {tot})")
    for v, _, f, ln, src in sel:
        print(f"{100*v/tot:5.1f}%  {f}:{ln}  {src}")
