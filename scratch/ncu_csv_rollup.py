"""`ncu --csv --log-file x.csv --metrics ...` (long format: one row per launch and metric) ->
per-kernel-family roll-up JSON (launches, total / mean us, share, duration-weighted tensor-pipe
activity, DRAM bytes).   python scratch/ncu_csv_rollup.py in.csv out.json [max_launches]"""
import collections, csv, json, re, sys
src, out = sys.argv[1], sys.argv[2]
limit = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 30
lines = open(src, errors="replace").read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rows = list(csv.DictReader(lines[start:]))
launch = collections.OrderedDict()
for r in rows:
    d = launch.setdefault(r["ID"], {"name": r["Kernel Name"]})
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    u = r["Metric Unit"]
    k = r["Metric Name"]
    if k == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}.get(u, 1.0)
    if k.startswith("dram__bytes"):
        v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
    d[k] = v
fam = collections.OrderedDict()
n = 0
for d in launch.values():
    if n >= limit:
        break
    n += 1
    name = re.sub(r"\(.*", "", d["name"]).replace("void ", "").strip()
    f = fam.setdefault(name, dict(launches=0, us=0.0, tx=0.0, rd=0.0, wr=0.0))
    us = d.get("gpu__time_duration.sum", 0.0)
    f["launches"] += 1
    f["us"] += us
    f["tx"] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0) * us
    f["rd"] += d.get("dram__bytes_read.sum", 0.0)
    f["wr"] += d.get("dram__bytes_write.sum", 0.0)
total = sum(f["us"] for f in fam.values())
res = {"launches": n, "total_us": round(total, 1), "kernels": []}
for k, f in sorted(fam.items(), key=lambda kv: -kv[1]["us"]):
    res["kernels"].append(dict(kernel=k, launches=f["launches"], total_us=round(f["us"], 1), share=round(f["us"] / total, 4),
                               mean_us=round(f["us"] / f["launches"], 2),
                               tensor_pipe_pct_of_elapsed=round(f["tx"] / f["us"], 1) if f["us"] else 0.0,
                               dram_read_mb=round(f["rd"], 1), dram_write_mb=round(f["wr"], 1)))
json.dump(res, open(out, "w"), indent=1)
print(n, "launches,", len(fam), "families, total", round(total / 1e3, 3), "ms ->", out)
