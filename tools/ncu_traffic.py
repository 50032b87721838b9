"""Summarise an `ncu --set full` report of bench.py's eager step: per kernel name the launches captured, average duration,
DRAM bytes read + written per launch, tensor-pipe / DRAM / L2 utilisation; and write profiles/traffic_by_entry.json, the
per-launch DRAM traffic of the C-ABI entries bench.py names in `roofline` (bench.py reads that file).

  ncu -i gpurun_out/<name>.ncu-rep --page raw --csv > /tmp/raw.csv
  python tools/ncu_traffic.py /tmp/raw.csv profiles/<round>_ncu_full_top_kernels.txt [profiles/traffic_by_entry.json]
"""
import collections
import csv
import json
import re
import sys

# kernel (as ncu prints it, template arguments included) -> "entry shape" key of bench.py's step profile (G32, N = 64)
ENTRY = {
    "tc2_gemm_kernel<0, 4>": "tatt_gemm M128 N1024 K3072 x2",
    "rows_wgrad1_kernel<3>": "tatt_rows_wgrad NB3",
    "rows_wgrad_ws_kernel<3>": "tatt_rows_wgrad NB3",
    "rows_wgrad_ws_kernel<2>": "tatt_rows_wgrad NB2",
    "rows_wgrad_ws_kernel<1>": "tatt_rows_wgrad NB1",
    "conv3x3_roll_kernel<0, 1>": "tatt_conv2d_igemm 3x3 64->64",
    "conv3x3_roll_kernel<0, 1, 0>": "tatt_conv2d_igemm 3x3 64->64",
    "mha_bwd_kernel": "tatt_mha64_bwd ",
    "mha_bwd_mma_kernel": "tatt_mha64_bwd ",
    "rows_wgrad_kernel<1>": "tatt_rows_wgrad NB1",
    "gru32_scan_bwd_mma_kernel<0>": "tatt_gru32_scan_bwd T32",
    "gru32_scan_fwd_mma_kernel<0>": "tatt_gru32_scan_fwd T32",
    "rpe_fwd_persist_kernel": "tatt_rpe_fwd T64 W128 Hd1024",
    "conv3x3_wgrad_tma_kernel": "tatt_conv2d_wgrad 3x3 64->64",
    "tp_declayer_fwd_kernel<1>": "tatt_tp_declayer_fwd ",
    "gru32_scan_bwd_mma_kernel<1>": "tatt_gru32_scan_bwd T128",
    "gru32_scan_fwd_mma_kernel<1>": "tatt_gru32_scan_fwd T128",
    "rows_gemm_kernel<1, 1>": "tatt_rows_gemm K64 N64",
    "rows_gemm_kernel<3, 1>": "tatt_rows_gemm K64 N192",
}
ALIAS = {"conv3x3_roll_kernel<0, 1>": ["tatt_conv3x3_stats 3x3 64->64"],
         "conv3x3_roll_kernel<0, 1, 0>": ["tatt_conv3x3_stats 3x3 64->64"]}      # same kernel behind a second entry point
COLS = {"dur": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
        "tensor": "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "tensor2": "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "dram": "FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l2": "lts__t_sectors.avg.pct_of_peak_sustained_elapsed", "sm": "sm__throughput.avg.pct_of_peak_sustained_elapsed"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6,
        "ns": 1e-3, "us": 1.0, "ms": 1e3, "%": 1.0, "": 1.0}


def num(s):
    try:
        return float(str(s).replace(",", ""))
    except ValueError:
        return None


def main():
    raw, out_txt = sys.argv[1], sys.argv[2]
    out_json = sys.argv[3] if len(sys.argv) > 3 else None
    with open(raw) as f:
        rows = list(csv.reader(l for l in f if not l.startswith("==")))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    kcol = idx.get("Kernel Name")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= kcol:
            continue
        name = re.sub(r"^void ", "", r[kcol]).replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
        name = re.sub(r"\(.*", "", name)
        a = agg.setdefault(name, collections.defaultdict(list))
        for k, col in COLS.items():
            if col in idx:
                v = num(r[idx[col]])
                if v is not None:
                    a[k].append(v * UNIT.get(units[idx[col]], 1.0))
    lines = ["# per kernel: launches captured, avg duration (us), DRAM read / written per launch (MB), tensor pipe %, DRAM GB/s, "
             "L2 %, SM % (ncu --set full, --clock-control none; serialised); GB/s = (read + written) / duration, copy peak 6449",
             "%-44s %4s %9s %9s %9s %7s %6s %6s %6s" % ("kernel", "n", "us", "rd MB", "wr MB", "tensor", "GB/s", "l2", "sm")]
    js = {"_source": out_txt}
    for name, a in sorted(agg.items(), key=lambda kv: -sum(kv[1]["dur"])):
        def av(k):
            return sum(a[k]) / len(a[k]) if a[k] else float("nan")
        t = av("tensor") if a["tensor"] else av("tensor2")
        lines.append("%-44s %4d %9.1f %9.2f %9.2f %7.1f %6.1f %6.1f %6.1f" % (
            name[:44], len(a["dur"]), av("dur"), av("rd") / 1e6, av("wr") / 1e6, t, (av("rd") + av("wr")) / av("dur") / 1e3,
            av("l2"), av("sm")))
        for key in ([ENTRY[name]] if name in ENTRY else []) + ALIAS.get(name, []):
            js[key] = {"kernel": name, "dram_bytes_per_launch": av("rd") + av("wr"), "launches": len(a["dur"]),
                               "avg_us_under_ncu": av("dur"), "tensor_pipe_pct": t,
                               "dram_gbs": (av("rd") + av("wr")) / av("dur") / 1e3}
    open(out_txt, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
    if out_json:
        json.dump(js, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main()
