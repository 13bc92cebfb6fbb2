"""ncu-rep -> small JSON (one record per launch) with the metrics profiles/README.md quotes.
    python scratch/ncu_summary.py gpurun_out/x.ncu-rep profiles/x.json [kernel-substring]"""
import csv, io, json, subprocess, sys
KEEP = ["Kernel Name", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "sm__cycles_elapsed.max"]
rep, out = sys.argv[1], sys.argv[2]
sub = sys.argv[3] if len(sys.argv) > 3 else ""
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, data = rows[0], rows[1], rows[2:]
recs = []
for r in data:
    d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
    if sub and sub not in d.get("Kernel Name", ""): continue
    recs.append({k: (d[k] + (" " + u[k] if u.get(k) else "")).strip() for k in KEEP if k in d})
json.dump(recs, open(out, "w"), indent=1)
print(len(recs), "launches ->", out)
