"""ncu csv (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch) -> per-kernel table."""
import csv, sys, collections
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
acc = collections.OrderedDict()
for r in rows[1:]:
    if len(r) < len(hdr) or not r[ix["ID"]].isdigit():
        continue
    name = r[ix["Kernel Name"]].split("(")[0]
    key = (r[ix["ID"]], name)
    d = acc.setdefault(key, {})
    val = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    mult = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    d[r[ix["Metric Name"]]] = val * mult
print(f"{'kernel':60s} {'ms':>9s} {'rd GB':>8s} {'wr GB':>8s} {'GB/s':>8s}")
for (i, name), d in acc.items():
    t = d.get("gpu__time_duration.sum", 0.0)
    rd, wr = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
    if t <= 0:
        continue
    print(f"{name[:60]:60s} {t*1e3:9.3f} {rd/1e9:8.3f} {wr/1e9:8.3f} {(rd+wr)/t/1e9:8.0f}")
