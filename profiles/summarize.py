#!/usr/bin/env python3
"""Turn an .ncu-rep (ncu --set full) or a launch-list CSV (ncu --metrics gpu__time_duration.sum --csv) into the
compact text summaries kept under profiles/.  Usage:
    python profiles/summarize.py rep   gpurun_out/x.ncu-rep  > profiles/r1_x_full.csv
    python profiles/summarize.py list  gpurun_out/launches.csv > profiles/r1_launches.txt
"""
import csv, subprocess, sys, collections

KEEP = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__instruction_throughput.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    idx = [h.index(k) for k in KEEP if k in h]
    w = csv.writer(sys.stdout)
    w.writerow([h[i] for i in idx])
    w.writerow([units[i] for i in idx])
    for r in rows[2:]:
        w.writerow([r[i][:90] for i in idx])


def launch_list(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 14 and r[0].isdigit()]
    agg = collections.OrderedDict()
    tot = 0.0
    for r in rows:
        name = r[4].split("(")[0][:70]
        ns = float(r[14].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += ns; tot += ns
    print(f"# {path}: {len(rows)} launches, {tot/1e6:.3f} ms total kernel time (cold-cache, serialised by ncu)")
    print(f"{'kernel':72s} {'launches':>8s} {'total_ms':>10s} {'mean_us':>9s} {'share':>7s}")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:72s} {n:8d} {ns/1e6:10.3f} {ns/n/1e3:9.1f} {100*ns/tot:6.1f}%")


if __name__ == "__main__":
    {"rep": rep, "list": launch_list}[sys.argv[1]](sys.argv[2])
