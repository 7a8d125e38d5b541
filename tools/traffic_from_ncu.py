"""profiles/traffic.json from an ncu CSV of the trace launches (dram__bytes_read.sum, dram__bytes_write.sum):
average DRAM bytes per k_wf_trace launch over the captured launches — the same averaging bench.py uses for `achieved`."""
import csv
import json
import sys

src, dst = sys.argv[1], sys.argv[2:]
rows = [r for r in csv.reader(open(src)) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
per = {}
for r in rows[1:]:
    v = float(r[ix["Metric Value"]].replace(",", ""))
    u = r[ix["Metric Unit"]].lower()
    v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    per.setdefault(r[ix["ID"]], 0.0)
    per[r[ix["ID"]]] += v
vals = list(per.values())
out = {"dram_bytes_per_launch": sum(vals) / len(vals), "launches_captured": len(vals), "per_launch": vals,
       "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:k_wf_trace (one bench step: rounds 0..%d)" % (len(vals) - 1)}
for d in dst:
    json.dump(out, open(d, "w"))
print(json.dumps(out)[:300])
