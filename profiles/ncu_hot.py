"""print the hottest SASS instructions (warp-stall samples) of an ncu report: python profiles/ncu_hot.py x.ncu-rep [N]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
si, ai, ci = hdr.index("Source"), hdr.index("Address"), hdr.index("# Samples")
data = []
for i, r in enumerate(rows[2:]):
    try:
        data.append((int(r[ci]), i, r[si].strip()))
    except Exception:
        pass
tot = sum(d[0] for d in data) or 1
print(rows[0][1][:100], "total samples", tot)
for s, i, src in sorted(data, reverse=True)[:top]:
    print(f"{100.0 * s / tot:5.1f}%  line {i:5d}  {src[:110]}")
