"""Opcode evidence for the shipped library: counts of the tcgen05 / TMA / TMEM SASS mnemonics per kernel.
    python tools/sass_histogram.py > profiles/r2_sass_opcodes.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "popcorn_b200", "libpopcorn_b200.so")
COLS = ["UTCHMMA", "STTM", "LDTM", "UTMALDG", "UTCBAR", "SYNCS", "FFMA", "FFMA2", "LDS", "LD", "STG"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
names, counts, cur = [], {}, None
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        names.append(cur)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m and cur:
        counts[cur][m.group(1)] += 1
        counts[cur]["_n"] += 1
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
print(f"# SASS opcode histogram of popcorn_b200/libpopcorn_b200.so (cuobjdump -sass, sm_100a)\n")
print("UTCHMMA = tcgen05.mma, STTM / LDTM = tcgen05.st / tcgen05.ld, UTMALDG = TMA tensor load, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops;")
print("LD = generic loads (none left: every shared-memory access of the library is an LDS / STS since the base pointers keep their address space).\n")
print("| kernel | " + " | ".join(COLS) + " | instructions |")
print("|---|" + "---|" * (len(COLS) + 1))
tot = collections.Counter()
for n, d in sorted(zip(names, dem), key=lambda t: t[1]):
    c = counts[n]
    tot.update(c)
    short = re.sub(r"\(.*", "", d).replace("void ", "")
    print(f"| `{short}` | " + " | ".join(str(c[k]) for k in COLS) + f" | {c['_n']} |")
print("| **total** | " + " | ".join(str(tot[k]) for k in COLS) + f" | {tot['_n']} |")
