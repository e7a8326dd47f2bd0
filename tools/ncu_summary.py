#!/usr/bin/env python
"""Turns ncu outputs brought back in gpurun_out/ into the small text/JSON summaries committed under profiles/.
  python tools/ncu_summary.py launches <launches.csv> <out.md>          per-kernel launch count, device time, share of the step
  python tools/ncu_summary.py rep <file.ncu-rep> <out.md> [out.json]    key counters of every captured launch (ncu --page raw)
"""
import collections
import csv
import json
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__maximum_warps_per_active_cycle_pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "smsp__inst_executed.sum", "sm__inst_issued.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]


def launches(path, out):
    rows = list(csv.DictReader(l for l in open(path) if l.startswith('"')))
    agg = collections.OrderedDict()
    for r in rows:
        k = r["Kernel Name"].split("(")[0].replace("void ", "")
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r["Metric Value"]) / 1e6
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write("| kernel | launches | device ms (ncu, serialised, cold cache) | share |\n|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("| `%s` | %d | %.3f | %.3f |\n" % (k, v[0], v[1], v[1] / tot))
        f.write("| total | %d | %.3f | 1.000 |\n" % (len(rows), tot))
    print(open(out).read())


def rep(path, out, out_json=None):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = collections.OrderedDict(); d["kernel"] = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w); d[w] = [r[i], units[i]]
        res.append(d)
    with open(out, "w") as f:
        for d in res:
            f.write("### %s\n\n| metric | value | unit |\n|---|---|---|\n" % d["kernel"])
            for k, v in d.items():
                if k != "kernel":
                    f.write("| %s | %s | %s |\n" % (k, v[0], v[1]))
            f.write("\n")
    if out_json:
        json.dump(res, open(out_json, "w"), indent=1)
    print(open(out).read())


if __name__ == "__main__":
    {"launches": launches, "rep": rep}[sys.argv[1]](*sys.argv[2:])
