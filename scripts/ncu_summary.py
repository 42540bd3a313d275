import os
"""Compact per-launch table (and per-sweep DRAM traffic) from an `ncu --set full` report.

    python scripts/ncu_summary.py report.ncu-rep out.csv [traffic.json]
"""
import csv, io, json, subprocess, sys

WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__waves_per_multiprocessor", "sm__inst_executed_pipe_fp64.sum"]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}


def main(rep, out, traffic=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, rows = rows[0], rows[1], rows[2:]
    idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
    tot = {}
    with open(out, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow([f"{n} [{units[i]}]" if units[i] else n for n, i in idx])
        for r in rows:
            w.writerow([r[i] for _, i in idx])
            name = r[hdr.index("Kernel Name")]
            kind = "forward_sweep_4rhs" if "forward" in name else ("backward_sweep_4rhs" if "backward" in name else "other")

            def val(col):
                i = hdr.index(col)
                return float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)
            t = tot.setdefault(kind, {"launches": 0, "us": 0.0, "dram_bytes": 0.0})
            t["launches"] += 1; t["us"] += val("gpu__time_duration.sum")
            t["dram_bytes"] += val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
    if traffic:
        for t in tot.values():
            t["dram_bytes_per_launch"] = t["dram_bytes"] / max(t["launches"], 1)
        json.dump({"source": rep, "designs": int(os.environ.get("DESIGNS", "12")),
                   "note": "one sweep pair of a forest of config-1 designs, ncu --set full (cold caches, serialised)",
                   "sweeps": tot}, open(traffic, "w"), indent=1)
    print(json.dumps(tot))


if __name__ == "__main__":
    main(*sys.argv[1:4])
