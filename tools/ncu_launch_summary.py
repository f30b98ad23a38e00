"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel totals and shares."""
import csv, sys, collections
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
tot, cnt = collections.OrderedDict(), collections.Counter()
for r in rows[1:]:
    if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    unit = r[ix["Metric Unit"]]
    v = float(r[ix["Metric Value"]].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[unit]
    name = r[ix["Kernel Name"]][:72]
    tot[name] = tot.get(name, 0.0) + v
    cnt[name] += 1
s = sum(tot.values())
print("kernel,launches,total_ms,share_pct,avg_ms")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f'"{k}",{cnt[k]},{v:.3f},{100*v/s:.2f},{v/cnt[k]:.4f}')
