#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-native SASS mnemonics in libidqn_b200.so (cuobjdump -sass):
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA loads/stores, UTCBAR = tcgen05.commit,
SYNCS = mbarrier ops, HMMA = legacy mma.sync (expected 0).   python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "i-dqn_b200", "libidqn_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pat = {"UTC*MMA": r"\bUTC\w*MMA", "LDTM": r"\bLDTM", "STTM": r"\bSTTM", "UTMALDG": r"\bUTMALDG", "UTMASTG": r"\bUTMASTG",
       "UTCBAR": r"\bUTCBAR", "SYNCS": r"\bSYNCS", "HMMA": r"\bHMMA", "LDGSTS": r"\bLDGSTS"}
counts, order, arch, cur = collections.defaultdict(collections.Counter), [], {}, None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("void ", "")
        order.append(cur)
        continue
    m = re.match(r"\s*\.target\s+(\S+)|.*arch = (sm_\w+)", line)
    if m and (m.group(1) or m.group(2)):
        last_arch = m.group(1) or m.group(2)
    if cur:
        for k, p in pat.items():
            if re.search(p, line):
                counts[cur][k] += 1
archs = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
print(f"# {os.path.basename(so)}: {len(order)} kernels, architectures {archs}")
print(f"{'kernel':70s} " + " ".join(f"{k:>8s}" for k in pat))
tot = collections.Counter()
for fn in order:
    c = counts[fn]
    tot.update(c)
    print(f"{fn[:70]:70s} " + " ".join(f"{c[k]:8d}" for k in pat))
print(f"{'TOTAL':70s} " + " ".join(f"{tot[k]:8d}" for k in pat))
