#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-native SASS mnemonics in the built library (no GPU needed):

    python scripts/sass_counts.py [path/to/libwaveformer_b200.so] > profiles/r02_sass_tcgen05.txt

UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk (TMA engine, 1-D), UTCBAR = tcgen05.commit,
LDGSTS = cp.async, SYNCS = mbarrier ops; HMMA would be the legacy mma.sync path (none expected).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "waveformer_b200", "lib", "libwaveformer_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda names: subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
pats = collections.OrderedDict([("UTC*MMA", r"\bUTC[A-Z]*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UBLKCP", r"\bUBLKCP"),
                                ("UTMALDG", r"\bUTMALDG"), ("UTCBAR", r"\bUTCBAR"), ("LDGSTS", r"\bLDGSTS"), ("SYNCS", r"\bSYNCS"),
                                ("HMMA", r"\bHMMA")])
counts, order, cur = {}, [], None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        order.append(cur)
        continue
    if cur is None:
        continue
    for k, p in pats.items():
        if re.search(p, line):
            counts[cur][k] += 1
names = dict(zip(order, demangle(order)))
archs = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
print(f"# {os.path.relpath(lib, ROOT)}: {len(order)} kernels, arch {archs}")
print("# kernels that contain tensor-core (tcgen05) or bulk-copy instructions; counts are static SASS instructions")
print(f"{'kernel':110s} " + " ".join(f"{k:>8s}" for k in pats))
tot = collections.Counter()
for f in order:
    c = counts[f]
    tot.update(c)
    if c["UTC*MMA"] or c["LDTM"] or c["UBLKCP"] or c["UTMALDG"]:
        short = re.sub(r"\(.*", "", names[f])[:110]
        print(f"{short:110s} " + " ".join(f"{c[k]:8d}" for k in pats))
print(f"{'TOTAL (all kernels)':110s} " + " ".join(f"{tot[k]:8d}" for k in pats))
