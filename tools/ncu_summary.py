"""Extract the judged metrics from an .ncu-rep:  python tools/ncu_summary.py rep.ncu-rep > profiles/x.csv"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ("Kernel Name", "gpu__time_duration", "launch__grid_size", "launch__registers", "launch__occupancy",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput", "sm__pipe_tensor_cycles_active.avg.pct",
        "sm__warps_active", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct",
        "sm__throughput.avg.pct", "lts__t_bytes.sum", "l1tex__t_bytes.sum", "smsp__warp_issue_stalled")
keep = [i for i, h in enumerate(hdr) if ("Triage" not in h) and any(k in h for k in KEYS)]
w = csv.writer(sys.stdout)
w.writerow([hdr[i] for i in keep]); w.writerow([units[i] for i in keep])
for r in rows[2:]:
    w.writerow([r[i] for i in keep])
