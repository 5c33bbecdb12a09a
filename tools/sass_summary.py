#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-only SASS mnemonics in the built library (no GPU needed):
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor loads/stores, UTCBAR = tcgen05.commit,
SYNCS = mbarrier ops, plus the legacy HMMA (must be 0) and FFMA2.  Usage: python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
so = ROOT / "auto-dynamic-deeplab_b200" / "libadd_b200.so"
out = subprocess.run(["cuobjdump", "-sass", str(so)], stdout=subprocess.PIPE, text=True, check=True).stdout
pats = {"UTC*MMA": r"\bUTC\w*MMA\b", "LDTM": r"\bLDTM\b", "STTM": r"\bSTTM\b", "UTMALDG": r"\bUTMALDG\b", "UTMASTG": r"\bUTMASTG\b",
        "UTCBAR": r"\bUTCBAR\b", "SYNCS": r"\bSYNCS\b", "HMMA": r"\bHMMA\b", "FFMA2": r"\bFFMA2\b", "MATCH": r"\bMATCH\b", "ATOMG/RED": r"\b(ATOMG|RED|ATOM)\b", "ATOMS": r"\bATOMS\b"}
rows, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"\(.*", "", name)
        cur = rows.setdefault(name, collections.Counter())
        continue
    if cur is None:
        continue
    for k, p in pats.items():
        if re.search(p, line):
            cur[k] += 1
arch = re.findall(r"arch = (sm_\w+)", out)
print(f"# cuobjdump -sass {so.relative_to(ROOT)}  (arch: {sorted(set(arch))}; {len(rows)} kernels)")
print(f"# {'kernel':88s} " + " ".join(f"{k:>8s}" for k in pats))
tot = collections.Counter()
for name, c in sorted(rows.items(), key=lambda kv: -(kv[1]['UTC*MMA'] * 1000 + kv[1]['UTMALDG'])):
    tot.update(c)
    if sum(c.values()) == 0:
        continue
    print(f"{name[:90]:90s} " + " ".join(f"{c[k]:8d}" for k in pats))
print(f"{'TOTAL':90s} " + " ".join(f"{tot[k]:8d}" for k in pats))
