"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches / total time / share per kernel.

    python profiles/launches_summary.py gpurun_out/launches.csv > profiles/r2/launches_step_summary.md

Per-launch times under ncu are cold-cache and serialised: only the shares are meaningful (B200_PROFILING.md)."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(r[iu], 1e-6)
    name = re.sub(r"^void ", "", r[ik])
    name = re.sub(r"\(.*$", "", name)[:90]
    tot[name][0] += 1
    tot[name][1] += v
all_ms = sum(v[1] for v in tot.values())
ours = {k: v for k, v in tot.items() if k.startswith("b200::") or "b200" in k.split("<")[0]}
ours_ms = sum(v[1] for v in ours.values())
print(f"{sum(v[0] for v in tot.values())} launches, {all_ms:.1f} ms in total; {sum(v[0] for v in ours.values())} of them are "
      f"this library's kernels ({ours_ms:.1f} ms, {100 * ours_ms / all_ms:.0f} %); the rest is torch set-up work outside "
      "the step (random weights, dtype casts, fills).\n")
print("| kernel | launches | total ms | share of all | share of the library's kernels |\n|---|---|---|---|---|")
for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:40]:
    so = f"{100 * ms / ours_ms:.1f} %" if k in ours else ""
    print(f"| `{k}` | {n} | {ms:.2f} | {100 * ms / all_ms:.1f} % | {so} |")
