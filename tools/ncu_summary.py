"""Per-launch summary of an `ncu --set full` report: tools/ncu_summary.py REPORT.ncu-rep > profiles/NAME.txt
(also prints a traffic JSON -- dram bytes read + written per kernel, first launch of each -- on stderr)."""
import csv, json, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
ki = hdr.index("Kernel Name")
idx = [(w, hdr.index(w)) for w in want if w in hdr]
seen, traffic = {}, {}
for r in rows[2:]:
    name = r[ki].split("(")[0].replace("void ", "").replace("sldm::", "")
    seen[name] = seen.get(name, 0) + 1
    print(f"{name}   [launch #{seen[name]} of this kernel in the capture]")
    vals = {}
    for w, i in idx:
        print(f"    {w:78s} {r[i]} {units[i]}")
        vals[w] = (r[i], units[i])
    if seen[name] == 1 and "dram__bytes_read.sum" in vals:
        def b(v):
            x, u = float(v[0].replace(",", "")), v[1].lower()
            return x * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        traffic[name] = int(b(vals["dram__bytes_read.sum"]) + b(vals["dram__bytes_write.sum"]))
print(json.dumps(traffic, indent=1), file=sys.stderr)
