"""ncu launch list (long-format CSV: one row per launch and metric) -> roll-up of ONE eager train
step = every launch from one observe_persistent_fwd_kernel launch up to the next one.
    python scratch/ncu_step_rollup.py gpurun_out/launches.csv profiles/launches_step.json"""
import collections, csv, json, re, sys
src, out = sys.argv[1], sys.argv[2]
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
    u, k = r["Metric Unit"], r["Metric Name"]
    if k == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}.get(u, 1.0)
    if k.startswith("dram__bytes"):
        v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
    d[k] = v
L = list(launch.values())
marks = [i for i, d in enumerate(L) if "observe_persistent_fwd_kernel" in d["name"]]
assert len(marks) >= 2, f"need two observe forward launches in the window, found {len(marks)}"
step = L[marks[0]:marks[1]]
fam = collections.OrderedDict()
for d in step:
    name = re.sub(r"\(.*", "", d["name"]).replace("void ", "").strip()
    f = fam.setdefault(name, {"kernel": name, "launches": 0, "total_us": 0.0, "tp_w": 0.0, "dram_read_mb": 0.0, "dram_write_mb": 0.0})
    t = d.get("gpu__time_duration.sum", 0.0)
    f["launches"] += 1; f["total_us"] += t
    f["tp_w"] += t * d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0)
    f["dram_read_mb"] += d.get("dram__bytes_read.sum", 0.0); f["dram_write_mb"] += d.get("dram__bytes_write.sum", 0.0)
total = sum(f["total_us"] for f in fam.values())
ks = []
for f in sorted(fam.values(), key=lambda f: -f["total_us"]):
    ks.append({"kernel": f["kernel"], "launches": f["launches"], "total_us": round(f["total_us"], 1),
               "share": round(f["total_us"] / total, 4), "mean_us": round(f["total_us"] / f["launches"], 2),
               "tensor_pipe_pct_of_elapsed": round(f["tp_w"] / max(f["total_us"], 1e-9), 1),
               "dram_read_mb": round(f["dram_read_mb"], 1), "dram_write_mb": round(f["dram_write_mb"], 1)})
lib = [k for k in ks if k["kernel"].startswith("dv3::") or "umma" in k["kernel"] or "_kernel" in k["kernel"] and not k["kernel"].startswith("at::")]
gem = [k for k in ks if "umma" in k["kernel"]]
res = {"what": "one eager train step (dmc_proprio 16x64, H=15): every launch between two consecutive "
               "observe_persistent_fwd_kernel launches; ncu --metrics gpu__time_duration.sum,"
               "sm__pipe_tensor_cycles_active...,dram__bytes_* --clock-control none (cold-cache, serialised)",
       "launches": len(step), "library_launches": sum(k["launches"] for k in ks if not k["kernel"].startswith("at::") and "nccl" not in k["kernel"].lower() and "cub::" not in k["kernel"]),
       "total_us": round(total, 1), "kernels": ks,
       "gemm_family": {"launches": sum(k["launches"] for k in gem), "total_us": round(sum(k["total_us"] for k in gem), 1),
                       "share": round(sum(k["total_us"] for k in gem) / total, 4),
                       "dram_read_mb": round(sum(k["dram_read_mb"] for k in gem), 1),
                       "dram_write_mb": round(sum(k["dram_write_mb"] for k in gem), 1)}}
res["other_launches"] = res["launches"] - res["library_launches"]
json.dump(res, open(out, "w"), indent=1)
print(out, res["launches"], res["library_launches"], res["total_us"], res["gemm_family"])
for k in ks[:14]:
    print(f"  {k['kernel'][:80]:80s} n={k['launches']:4d} tot={k['total_us']:8.1f} share={k['share']:.3f} tp={k['tensor_pipe_pct_of_elapsed']}")
