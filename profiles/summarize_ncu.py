"""Summarise an `ncu --set full` capture of the solve kernel into text + traffic.json.

    python profiles/summarize_ncu.py <report.ncu-rep> <out_dir> --model M --dtype D --batch B --mode converge|fixed

Writes <out_dir>/<report>_summary.txt (the raw metrics the design discussion cites) and appends the
per-launch DRAM traffic to <out_dir>/traffic.json, which bench.py reads for `roofline.traffic`.
"""
import argparse
import csv
import io
import json
import os
import subprocess

KEYS = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit",
        "sm__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled", "sm__cycles_elapsed.avg.per_second", "smsp__cycles_active.avg")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("out_dir")
    ap.add_argument("--model", default="ThreeLinkPlanarManipulator")
    ap.add_argument("--dtype", default="float64")
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--mode", default="converge")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    rec = dict(zip(hdr, zip(units, vals)))
    name = os.path.splitext(os.path.basename(a.report))[0]
    os.makedirs(a.out_dir, exist_ok=True)
    with open(os.path.join(a.out_dir, name + "_summary.txt"), "w") as fh:
        fh.write("ncu --set full --clock-control none capture: %s\nkernel: %s\n\n" % (name, rec.get("Kernel Name", ("", ""))[1]))
        for h in hdr:
            if any(k in h for k in KEYS) and "not_issued" not in h and "pcsamp" not in h:
                fh.write("%-90s %-14s %s\n" % (h, rec[h][0], rec[h][1]))

    def num(key):
        u, v = rec[key]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
        return float(v) * scale

    traffic = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    path = os.path.join(a.out_dir, "traffic.json")
    recs = json.load(open(path)) if os.path.exists(path) else []
    recs = [r for r in recs if (r["model"], r["dtype"], r["batch"], r["mode"]) != (a.model, a.dtype, a.batch, a.mode)]
    recs.append({"model": a.model, "dtype": a.dtype, "batch": a.batch, "mode": a.mode, "dram_bytes": traffic,
                 "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
                 "kernel_ms_under_ncu": float(rec["gpu__time_duration.sum"][1]) * {"ms": 1.0, "us": 1e-3, "s": 1e3,
                                                                                  "ns": 1e-6}[rec["gpu__time_duration.sum"][0]],
                 "source": "%s.ncu-rep (ncu --set full --clock-control none)" % name})
    json.dump(recs, open(path, "w"), indent=1)
    print("traffic %.3f GB -> %s" % (traffic / 1e9, path))


if __name__ == "__main__":
    main()
