#!/usr/bin/env python3
"""Extracts the per-launch counters bench.py reports next to its live timings from an ncu --set full capture:

    python profiles/counters.py gpurun_out/x.ncu-rep KEY=kernel_regex [KEY=kernel_regex ...] >> merges into profiles/r2_kernel_counters.json

e.g. spmv_n119_adpm_1gpu=k_spmv_tma assemble_n119_adpm_1gpu=k_assemble.  Per key: DRAM bytes read/written, warp instructions,
fp64 thread instructions (DADD + DMUL + DFMA, predicated-on), registers, duration under ncu, the capture it came from."""
import csv, json, os, re, subprocess, sys

rep = sys.argv[1]
out_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "r2_kernel_counters.json")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
col = {name: i for i, name in enumerate(h)}


def val(r, name, scale=None):
    i = col.get(name)
    if i is None or r[i] == "":
        return None
    v = float(r[i].replace(",", ""))
    u = units[i]
    mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u, 1.0)
    return v * mult


db = json.load(open(out_path)) if os.path.exists(out_path) else {}
for spec in sys.argv[2:]:
    key, rx = spec.split("=", 1)
    sel = [r for r in rows[2:] if re.search(rx, r[col["Kernel Name"]])]
    if not sel:
        print("no launch matches", rx, file=sys.stderr)
        continue
    n = len(sel)
    mean = lambda name: (sum(val(r, name) for r in sel) / n) if val(sel[0], name) is not None else None
    cyc = mean("sm__cycles_elapsed.max")
    fp = None
    rates = [mean(f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed") for op in ("dadd", "dmul", "dfma")]
    if cyc and all(x is not None for x in rates):
        fp = sum(rates) * cyc
    db[key] = {"kernel": sel[0][col["Kernel Name"]][:90], "launches_averaged": n, "dram_read": mean("dram__bytes_read.sum"),
               "dram_write": mean("dram__bytes_write.sum"), "inst_executed": mean("smsp__inst_executed.sum"),
               "fp64_thread_inst": fp, "registers": mean("launch__registers_per_thread"),
               "duration_under_ncu_s": mean("gpu__time_duration.sum"), "source": os.environ.get("COUNTERS_SOURCE", "profiles/" + os.path.basename(rep).replace(".ncu-rep", "_full.csv"))}
json.dump(db, open(out_path, "w"), indent=1, sort_keys=True)
print(json.dumps(db, indent=1, sort_keys=True))
