#!/usr/bin/env python3
"""Turns an `ncu --set full` report of one kernel launch into the tracked files bench.py reads:

    tools/ncu_export.py REPORT.ncu-rep META.json OUT_PREFIX

  OUT_PREFIX.csv   `ncu -i REPORT --page raw --csv` (every metric of the capture, unedited)
  OUT_PREFIX.json  META.json (written by tools/quick_time.py under QUICK_TIME_META: frames, edge updates of the captured
                   launch) + the kernel name and the handful of metrics quoted in DESIGN.md
"""
import csv, json, subprocess, sys

def main():
    rep, meta_path, prefix = sys.argv[1:4]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    raw = raw[raw.index('"ID"'):]
    open(prefix + ".csv", "w").write(raw)
    rows = list(csv.reader(raw.splitlines()))
    col = {n: i for i, n in enumerate(rows[0])}
    get = lambda k: rows[2][col[k]]
    meta = json.load(open(meta_path))
    meta["kernel"] = get("Kernel Name")
    meta["report"] = rep
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
            "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__cycles_active.avg", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "smsp__warps_eligible.avg.per_cycle_active"]
    meta["metrics"] = {k: "%s %s" % (get(k), rows[1][col[k]]) for k in keys if k in col}
    json.dump(meta, open(prefix + ".json", "w"), indent=1)
    print(json.dumps(meta["metrics"], indent=1))

main()
