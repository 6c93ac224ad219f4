#!/usr/bin/env python
"""Turn the raw ncu artefacts gpurun brought back (gpurun_out/) into the small text summaries kept
under profiles/.  usage: python profiles/summarize.py <tag> <launches.csv> <prof.ncu-rep>"""
import collections
import csv
import io
import subprocess
import sys

tag, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]
txt = open(launches).read()
txt = txt[txt.index('"ID"'):]
rows = list(csv.DictReader(io.StringIO(txt)))
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r["Kernel Name"].split("(")[0].replace("unnamed>::", "").replace("void ", "")
    agg[name][0] += 1
    agg[name][1] += float(r["Metric Value"]) / 1e3
tot = sum(v[1] for v in agg.values())
with open(f"profiles/{tag}_launches.txt", "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
    f.write(f"# {len(rows)} launches, {tot / 1e3:.3f} ms in total\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k:28s} n={v[0]:4d} total={v[1] / 1e3:8.3f} ms  avg={v[1] / v[0]:8.1f} us  share={100 * v[1] / tot:5.1f}%\n")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hdr, units = rr[0], rr[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
with open(f"profiles/{tag}_ncu_full.txt", "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on : selected raw metrics per captured launch\n")
    for r in rr[2:]:
        f.write("---- " + r[idx["Kernel Name"]][:90] + "\n")
        for w in want:
            if w in idx:
                f.write(f"   {w:82s} {r[idx[w]]:>18s} {units[idx[w]]}\n")
print("wrote", f"profiles/{tag}_launches.txt", f"profiles/{tag}_ncu_full.txt")
