"""profiles/traffic.json from an ncu CSV of the trace launches of ONE bench step:

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,\
l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum \
        --clock-control none -k regex:k_wf_trace -c 10 --csv --log-file trace.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline
    python tools/traffic_from_ncu.py trace.csv profiles/traffic.json

Per-launch averages over the captured launches (the same averaging bench.py uses for `achieved`): real DRAM bytes, L2 bytes
(sectors x 32 B), L1 global-load bytes, and the duration-weighted utilisation of the L1TEX data pipe and of the issue slots —
the `limiter` is whichever unit is busiest. bench.py copies these next to the algorithmic-bytes roofline."""
import csv
import json
import sys

src, dst = sys.argv[1], sys.argv[2:]
rows = [r for r in csv.reader(open(src)) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
SCALE = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "sector": 1, "%": 1, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1, "ns": 1e-9, "us": 1e-6, "ms": 1e-3}
per = {}
for r in rows[1:]:
    name = r[ix["Metric Name"]]
    v = float(r[ix["Metric Value"]].replace(",", "")) * SCALE.get(r[ix["Metric Unit"]].lower(), 1)
    per.setdefault(r[ix["ID"]], {})[name] = v
launches = list(per.values())
n = len(launches)


def avg(key):
    vals = [l[key] for l in launches if key in l]
    return sum(vals) / len(vals) if vals else None


def weighted(key):
    w = [(l[key], l.get("gpu__time_duration.sum", 1.0)) for l in launches if key in l]
    return sum(v * t for v, t in w) / sum(t for _, t in w) if w else None


dram = [l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0) for l in launches]
l1pipe = weighted("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed")
issue = weighted("smsp__issue_active.avg.pct_of_peak_sustained_active")
lts = avg("lts__t_sectors.sum")
l1s = avg("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum")
units = {"l1tex_data_pipe": l1pipe or 0.0, "issue_slots": issue or 0.0}
out = {"dram_bytes_per_launch": sum(dram) / n, "launches_captured": n, "per_launch": dram,
       "l2_bytes_per_launch": lts * 32 if lts else None, "l1_global_load_bytes_per_launch": l1s * 32 if l1s else None,
       "l1tex_data_pipe_pct": l1pipe, "issue_active_pct": issue, "limiter": max(units, key=units.get) if any(units.values()) else None,
       "avg_launch_ms_under_ncu": (avg("gpu__time_duration.sum") or 0.0) * 1e3,
       "source": "ncu (see this script's docstring) -k regex:k_wf_trace, one bench step: rounds 0..%d" % (n - 1)}
for d in dst:
    json.dump(out, open(d, "w"))
print(json.dumps(out)[:600])
