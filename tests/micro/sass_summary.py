"""Developer tool: opcode histogram per kernel of the built library (cuobjdump -sass): which kernels carry tcgen05 (UTC*MMA),
TMA (UTMALDG / UBLKCP), TMEM loads (LDTM), legacy tensor-core (HMMA) instructions.  Usage: python tests/micro/sass_summary.py > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "dyffusion_b200", "libdyffusion_b200.so")
OPS = ["UTCHMMA", "UTCBAR", "UTMALDG", "UBLKCP", "LDTM", "HMMA", "LDGSTS", "LDSM", "SYNCS", "UCGABAR", "MUFU"]
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1] if len(sys.argv) > 1 else LIB], capture_output=True, text=True).stdout
cur, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        hist[cur]["_total"] += 1
        for o in OPS:
            if m.group(1).startswith(o):
                hist[cur][o] += 1
names = subprocess.run(["c++filt"], input="\n".join(hist), capture_output=True, text=True).stdout.splitlines()
print(f"{'kernel':78s} {'SASS':>6s} " + " ".join(f"{o:>8s}" for o in OPS))
tot = collections.Counter()
for (k, c), name in zip(hist.items(), names):
    name = name.replace("(anonymous namespace)::", "").replace("dyf::", "")
    name = re.sub(r"\((?:dyf::)?[A-Za-z_:]*Params.*", "", name)[:78]
    print(f"{name:78s} {c['_total']:6d} " + " ".join(f"{c[o]:8d}" for o in OPS))
    tot.update(c)
print(f"{'TOTAL':78s} {tot['_total']:6d} " + " ".join(f"{tot[o]:8d}" for o in OPS))
