"""ncu-rep -> per-kernel-family roll-up JSON: launches, total / mean duration, duration-weighted tensor
pipe activity, summed DRAM read / write bytes.
    python scratch/ncu_rollup.py gpurun_out/x.ncu-rep profiles/x.json"""
import collections, csv, io, json, re, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {k: i for i, k in enumerate(hdr)}
def num(r, k):
    try:
        return float(r[col[k]].replace(",", ""))
    except Exception:
        return 0.0
def scale(k, target):      # ncu prints per-metric units; normalise durations to us and bytes to MB
    u = units[col[k]] if k in col else ""
    f = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6,
         "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
    return f.get(u, 1.0)
fam = collections.OrderedDict()
T = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"
TA = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
for r in data:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").strip()
    d = fam.setdefault(name, dict(launches=0, us=0.0, tensor_elapsed_x_us=0.0, tensor_active_x_us=0.0, dram_read_mb=0.0, dram_write_mb=0.0))
    us = num(r, "gpu__time_duration.sum") * scale("gpu__time_duration.sum", "us")
    d["launches"] += 1
    d["us"] += us
    d["tensor_elapsed_x_us"] += num(r, T) * us if T in col else 0.0
    d["tensor_active_x_us"] += num(r, TA) * us if TA in col else 0.0
    d["dram_read_mb"] += num(r, "dram__bytes_read.sum") * scale("dram__bytes_read.sum", "MB")
    d["dram_write_mb"] += num(r, "dram__bytes_write.sum") * scale("dram__bytes_write.sum", "MB")
total = sum(d["us"] for d in fam.values())
res = {"total_us": total, "kernels": []}
for k, d in sorted(fam.items(), key=lambda kv: -kv[1]["us"]):
    res["kernels"].append(dict(kernel=k, launches=d["launches"], total_us=round(d["us"], 1), share=round(d["us"] / total, 4),
                               mean_us=round(d["us"] / d["launches"], 2),
                               tensor_pipe_pct_of_elapsed=round(d["tensor_elapsed_x_us"] / d["us"], 1) if d["us"] else 0,
                               tensor_pipe_pct_of_active=round(d["tensor_active_x_us"] / d["us"], 1) if d["us"] else 0,
                               dram_read_mb=round(d["dram_read_mb"], 1), dram_write_mb=round(d["dram_write_mb"], 1)))
json.dump(res, open(out, "w"), indent=1)
print(len(data), "launches,", len(fam), "families ->", out)
