"""Per-kernel counts of the Blackwell-native SASS mnemonics in libb200lmd.so (B200_PROFILING.md "What proves a
Blackwell-native kernel"): UTC*MMA = tcgen05.mma, UTMALDG/UTMASTG = TMA load/store, LDTM/STTM = tcgen05.ld/st,
UTCBAR = tcgen05.commit, SYNCS = mbarrier, HMMA = legacy mma.sync (should be 0).
Usage: python profiles/sass_summary.py > profiles/sass_summary.md   (build container; needs cuobjdump + c++filt)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "llm-groundeddiffusion_b200", "libb200lmd.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
pats = collections.OrderedDict([("UTC*MMA", r"\bUTC[A-Z]*MMA"), ("UTMALDG", r"\bUTMALDG"), ("multicast", r"UTMALDG\S*MULTICAST"),
                                ("UTMASTG", r"\bUTMASTG"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
                                ("UTCBAR", r"\bUTCBAR"), ("SYNCS", r"\bSYNCS"), ("HMMA", r"\bHMMA")])
rows, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        rows[cur] = collections.Counter()
        continue
    if cur:
        for k, p in pats.items():
            if re.search(p, line):
                rows[cur][k] += 1
names = list(rows)
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
agg = collections.OrderedDict()
for n, d in zip(names, dem):
    base = re.sub(r"\(.*", "", d).replace("void ", "")
    base = re.sub(r"<.*", "", base)
    a = agg.setdefault(base, [0, collections.Counter()])
    a[0] += 1
    a[1].update(rows[n])
head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
print(f"# SASS summary of libb200lmd.so (built from HEAD {head} + working tree; `cuobjdump -sass`, sm_100a)\n")
print("Counts are summed over the template instantiations of each kernel.\n")
print("| kernel | instantiations | " + " | ".join(pats) + " |")
print("|---|---|" + "---|" * len(pats))
tot = collections.Counter()
for base, (n, c) in sorted(agg.items(), key=lambda kv: -kv[1][1]["UTC*MMA"]):
    print(f"| `{base}` | {n} | " + " | ".join(str(c[k]) for k in pats) + " |")
    tot.update(c)
print(f"| **total** | {sum(a[0] for a in agg.values())} | " + " | ".join(str(tot[k]) for k in pats) + " |")
