"""Per-kernel counts of the Blackwell-specific SASS instructions in the shipped library (evidence that the hot kernels are
tcgen05 / TMEM / TMA kernels, B200_PROFILING.md):

    python tools/sass_counts.py [i-vit_b200/csrc/libivit_b200.so] > profiles/sass_r2.txt

UTCIMMA = tcgen05.mma kind::i8, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA load / store,
SYNCS = mbarrier, IMMA = mma.sync (the general attention path), IDP = dp4a / dp2a."""
import collections
import os
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                        "i-vit_b200", "csrc", "libivit_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pats = ["UTCIMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "SYNCS", "IMMA", "IDP", "UTCATOMSWS"]
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = per.setdefault(re.sub(r"\(.*", "", name), collections.Counter())
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
    if m:
        cur["_all"] += 1
        op = m.group(1)
        if op in pats:
            cur[op] += 1
            if op in ("UTCIMMA", "UTMALDG", "UTMASTG", "UTCBAR"):
                cur[op + m.group(2)] += 1
print("# SASS instruction counts per kernel of %s (cuobjdump -sass)" % os.path.basename(so))
print("# %-78s %6s  %s" % ("kernel", "instr", "Blackwell-specific instructions"))
tot = collections.Counter()
for name, c in per.items():
    keys = [k for k in c if k != "_all"]
    if not any(k in c for k in ("UTCIMMA", "UTMALDG", "UTMASTG", "LDTM", "IMMA")):
        continue
    detail = ", ".join("%s %d" % (k, c[k]) for k in sorted(keys))
    print("%-80s %6d  %s" % (name[:80], c["_all"], detail))
    tot.update({k: c[k] for k in pats if k in c})
print("# totals: " + ", ".join("%s %d" % (k, tot[k]) for k in pats if tot[k]))
