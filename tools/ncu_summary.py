"""Summarise an `ncu --set full` capture of one kernel launch into the small JSON bench.py reads
(roofline.traffic / issue) plus a metrics CSV for profiles/.

    python tools/ncu_summary.py gpurun_out/r2_flat2_c2.ncu-rep profiles/r02_flat2_c2_ncu \
        --envs 4096 --keywords 100 --mean-volume 128

Writes <out>.json and <out>_metrics.csv (the raw page restricted to the metrics that matter)."""
import argparse
import csv
import json
import subprocess

KEEP = ("gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sass__inst_executed_local_loads",
        "sass__inst_executed_local_stores", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("out")
    ap.add_argument("--envs", type=int, required=True)
    ap.add_argument("--keywords", type=int, required=True)
    ap.add_argument("--mean-volume", type=int, default=128, dest="mean_volume")
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}

    def num(name):
        u, v = col[name]
        return float(v.replace(",", "")) * scale.get(u, 1)

    keep = [(h, col[h][0], col[h][1]) for h in hdr
            if h in KEEP or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    with open(a.out + "_metrics.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit", "value"])
        w.writerows(keep)
    out = {
        "kernel": col["Kernel Name"][1], "grid": int(float(col["launch__grid_size"][1])),
        "envs": a.envs, "keywords": a.keywords, "mean_volume": a.mean_volume,
        "duration_us": num("gpu__time_duration.sum"),
        "warp_instructions": int(num("smsp__inst_executed.sum")),
        "dram_bytes": int(num("dram__bytes_read.sum") + num("dram__bytes_write.sum")),
        "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "registers": int(float(col["launch__registers_per_thread"][1])),
        "source": a.rep, "note": a.note or "one launch, ncu --set full --clock-control none (cold-cache replay passes)",
    }
    json.dump(out, open(a.out + ".json", "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
