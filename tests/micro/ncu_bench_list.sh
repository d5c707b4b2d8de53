#!/bin/bash
# Developer tool (GPU): the recipe's launch-list pass (`--metrics gpu__time_duration.sum --clock-control none`) over the
# bench command of one config, and its per-kernel aggregate.  usage: ncu_bench_list.sh <config> <tag> [launch cap]
cfg="$1"; tag="$2"; cap="${3:-4000}"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c "$cap" --csv --log-file gpurun_out/${tag}_launches_ncu.csv \
  python bench.py --config "$cfg" --steps 1 --warmup 1 > gpurun_out/${tag}_bench_under_ncu.log 2>&1
python - "$tag" <<'PY'
import collections, csv, re, sys
tag = sys.argv[1]
rows = list(csv.reader(open(f"gpurun_out/{tag}_launches_ncu.csv")))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict(); n = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    nm = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("dyf::<unnamed>::", "")[:70]
    a = agg.setdefault(nm, [0, 0.0]); a[0] += 1; a[1] += float(r[ix["Metric Value"]].replace(",", "")) / 1000; n += 1
tot = sum(v[1] for v in agg.values())
with open(f"gpurun_out/{tag}_launches_agg.txt", "w") as f:
    f.write(f"{n} launches, {tot / 1000:.2f} ms of kernel time (ncu: cold-cache, serialised; shares, not absolutes)\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k:72s} n={v[0]:5d} {v[1] / 1000:9.3f} ms {100 * v[1] / tot:5.1f}%\n")
print(open(f"gpurun_out/{tag}_launches_agg.txt").read()[:3000])
PY
