import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
recs = [(r[ix["Kernel Name"]], float(r[ix["Metric Value"]].replace(",", ""))) for r in rows[hi + 1:] if len(r) >= len(hdr)]
starts = [k for k, (n, _) in enumerate(recs) if "pack_" in n or "pack_kernel" in n]
fw = recs[starts[-1]:]
agg = collections.OrderedDict()
for n, t in fw:
    nm = re.sub(r"\(.*", "", n).replace("dyf::<unnamed>::", "")[:60]
    a = agg.setdefault(nm, [0, 0.0, []]); a[0] += 1; a[1] += t / 1000; a[2].append(round(t / 1000))
tot = sum(v[1] for v in agg.values())
print("total us", round(tot, 1))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]): print(f"{k:62s} n={v[0]:3d} {v[1]:9.1f} us {100*v[1]/tot:5.1f}%  {v[2][:12]}")
