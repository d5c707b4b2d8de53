"""Developer tool (CPU): per CUDA source line, the executed warp instructions and stall samples of the first kernel of an
ncu report captured with --import-source on (-lineinfo build).  usage: ncu_lines.py <report> [top N]"""
import collections
import csv
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
ie, ss = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
agg = collections.OrderedDict()
line, text = None, ""
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    try:  # (source text with embedded quotes can break a CSV row: skip it)
        if r[0]:
            line, text = int(r[0]), r[1]
        vi, vs = float(r[ie] or 0), float(r[ss] or 0)
    except ValueError:
        continue
    a = agg.setdefault(line, [text, 0.0, 0.0])
    a[1] += vi
    a[2] += vs
tot_i = sum(v[1] for v in agg.values()) or 1
tot_s = sum(v[2] for v in agg.values()) or 1
top = sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]
print(f"total warp instructions {tot_i:.3e}, stall samples {tot_s:.0f}")
for ln, (t, i, s) in sorted(top):
    print(f"{ln:5d} {100 * i / tot_i:5.1f}% inst {100 * s / tot_s:5.1f}% stall  {t.strip()[:110]}")
