"""Turns the scratch ncu outputs under gpurun_out/ into the tracked summaries under profiles/.

    python profiles/summarize_ncu.py <round-tag>      e.g.  r1

reads   gpurun_out/launches.csv                 (ncu --metrics gpu__time_duration.sum --clock-control none ... prof_run.py)
        gpurun_out/{tower,kstep,heads}_full.ncu-rep   (ncu --set full --clock-control none --import-source on -k regex:... -c 1)
writes  profiles/<tag>_launches_by_kernel.csv   per-kernel launch count, total / mean time, share of the captured window
        profiles/<tag>_ncu_full_summary.json    the metrics DESIGN.md / bench.py quote (duration, DRAM bytes, tensor-pipe %, ...)
"""
import csv
import io
import json
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_warps", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_uniform.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]

SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}


def launches(tag, src="launches.csv", suffix=""):
    path = os.path.join(OUT, src)
    if not os.path.exists(path):
        return None
    text = [ln for ln in open(path) if ln.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(text))))
    agg = OrderedDict()
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", "")) * SCALE.get(r["Metric Unit"], 1.0)
        k = (r["Kernel Name"].split("(")[0], r["Grid Size"], r["Block Size"])
        a = agg.setdefault(k, [0, 0.0, 1e30, 0.0])
        a[0] += 1
        a[1] += v
        a[2] = min(a[2], v)
        a[3] = max(a[3], v)
    total = sum(a[1] for a in agg.values())
    out = os.path.join(ROOT, "profiles", f"{tag}_launches_by_kernel{suffix}.csv")
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "grid", "block", "launches", "total_us", "mean_us", "min_us", "max_us", "share_of_window"])
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k[0], k[1], k[2], a[0], f"{a[1]:.1f}", f"{a[1] / a[0]:.2f}", f"{a[2]:.2f}", f"{a[3]:.2f}", f"{a[1] / total:.4f}"])
    return out


def full(tag):
    res = OrderedDict()
    for name in ("tower", "kstep", "heads", "tower_cfg3", "tower_cfg4", "tower_cfg5"):
        rep = os.path.join(OUT, f"{name}_full.ncu-rep")
        if not os.path.exists(rep):
            continue
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        if len(rows) < 3:
            continue
        hdr, units, vals = rows[0], rows[1], rows[2]
        m = OrderedDict()
        for h, u, v in zip(hdr, units, vals):
            if h in ("Kernel Name", "Grid Size", "Block Size"):
                m[h] = v
            if h in KEEP and v != "":
                x = float(v.replace(",", ""))
                if u in SCALE and ("bytes" in h or "time" in h):
                    x *= SCALE[u]
                    u = "byte" if "bytes" in h else "us"
                m[f"{h} [{u}]"] = x
        rd, wr = m.get("dram__bytes_read.sum [byte]"), m.get("dram__bytes_write.sum [byte]")
        if rd is not None and wr is not None:
            m["dram_traffic_bytes_per_launch"] = rd + wr
        res[name] = m
    out = os.path.join(ROOT, "profiles", f"{tag}_ncu_full_summary.json")
    with open(out, "w") as f:
        json.dump(res, f, indent=1)
    return out


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    print(launches(tag))
    for cfg in (3, 4, 5):
        print(launches(tag, f"launches_cfg{cfg}.csv", f"_cfg{cfg}"))
    print(full(tag))
