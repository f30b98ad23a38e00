"""profiles/traffic.json from an `ncu --set full` raw CSV of a bench.py run: DRAM bytes (read + write) per launch of the
rotate_reduce kernel, keyed by the shape the capture was taken at.  bench.py reports it as `roofline.traffic` for that
exact shape only (and names this capture as the source).
    python tools/ncu_traffic.py profiles/r2_ncu_full_c2.csv 5000 500000 6"""
import csv, json, os, sys
src, n, m, S = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
rows = list(csv.reader(open(src)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
def gbytes(r, key):
    v, u = float(r[ix[key]]), units[ix[key]].lower()
    return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0, "tbyte": 1e12}[u]
vals = [gbytes(r, "dram__bytes_read.sum") + gbytes(r, "dram__bytes_write.sum") for r in rows[2:]
        if "ScanEpi" in r[ix["Kernel Name"]]]
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
doc = {"rotate_reduce": []}
if os.path.exists(out):
    doc = json.load(open(out))
doc["rotate_reduce"] = [e for e in doc["rotate_reduce"] if (e["n"], e["m"], e["slices"]) != (n, m, S)]
doc["rotate_reduce"].append({"n": n, "m": m, "slices": S, "dram_bytes": sum(vals) / len(vals), "launches": len(vals),
                             "source": f"ncu --set full, {src}"})
json.dump(doc, open(out, "w"), indent=1)
print(doc)
