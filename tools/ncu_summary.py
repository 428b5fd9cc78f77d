"""Reads an `ncu --set full` report (.ncu-rep) HERE (no GPU needed) and writes the per-launch summary the bench line and the
profiles/ evidence use:  python tools/ncu_summary.py gpurun_out/r2_pair.ncu-rep profiles/r2_ncu_pre  [kernel-name-substring [workload samples]]
-> <out>.md (table) and <out>_traffic.json (dram bytes per launch of the FIRST matching launch: bench.py's roofline.traffic, for
the bench workload named in the json -- default dtu512, 524 288 samples per captured launch)."""
import csv, io, json, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
want = sys.argv[3] if len(sys.argv) > 3 else "mlp_pair_kernel"
workload = sys.argv[4] if len(sys.argv) > 4 else "dtu512"
samples = int(sys.argv[5]) if len(sys.argv) > 5 else 524288
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
keys = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
lines = ["| kernel | " + " | ".join(k for k in keys if k in col) + " |", "|---|" + "---|" * sum(k in col for k in keys)]
traffic = None
for r in data:
    name = r[col["Kernel Name"]]
    if want not in name:
        continue
    vals = []
    for k in keys:
        if k in col:
            vals.append("%s %s" % (r[col[k]], units[col[k]]))
    lines.append("| %s | %s |" % (name[:60], " | ".join(vals)))
    if traffic is None and "dram__bytes_read.sum" in col:
        def b(k):
            return float(r[col[k]].replace(",", "")) * scale.get(units[col[k]], 1.0)
        traffic = dict(kernel=name, dram_bytes_per_launch=b("dram__bytes_read.sum") + b("dram__bytes_write.sum"),
                       dram_bytes_read=b("dram__bytes_read.sum"), dram_bytes_write=b("dram__bytes_write.sum"),
                       tensor_pipe_active_pct=float(r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]]),
                       duration=r[col["gpu__time_duration.sum"]] + " " + units[col["gpu__time_duration.sum"]], source=rep, workload=workload,
                       samples=samples)
open(out + ".md", "w").write("# ncu --set full summary of %s (read with tools/ncu_summary.py; profiler numbers are for attribution only)\n\n" % rep + "\n".join(lines) + "\n")
if traffic:
    json.dump(traffic, open(out + "_traffic.json", "w"), indent=1)
print("\n".join(lines))
print(traffic)
