"""Turns an .ncu-rep capture (ncu --set full) of ONE kernel launch into the summaries committed under profiles/:

    python tools/ncu_profile.py <rep> <name> <units> "<what>" [--traverse <workload>]      (ncu must be on PATH)

writes profiles/<name>_ncu_summary.json (selected raw metrics), profiles/<name>_regions.txt (per-SASS-region issue and
lane statistics) and, with --traverse cfgN, profiles/r2_traverse_<cfgN>_ncu.json — the per-workload figures bench.py
reads for its roofline (warp instructions and DRAM bytes per query of round 1, l1tex and issue-slot utilisation).
`units` = queries (or sorted pairs, points, ...) the captured launch processed.
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, name, units, what = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
KEYS = """gpu__time_duration.sum launch__grid_size launch__block_size launch__registers_per_thread
launch__occupancy_limit_registers launch__occupancy_limit_shared_mem launch__occupancy_limit_warps
sm__warps_active.avg.pct_of_peak_sustained_active
smsp__issue_active.avg.pct_of_peak_sustained_active smsp__inst_executed.sum smsp__thread_inst_executed_per_inst_executed.ratio
sm__inst_executed.sum.per_cycle_active dram__bytes_read.sum dram__bytes_write.sum
dram__bytes_read.sum.pct_of_peak_sustained_elapsed dram__bytes_write.sum.pct_of_peak_sustained_elapsed
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
lts__t_sector_hit_rate.pct l1tex__t_sector_hit_rate.pct lts__throughput.avg.pct_of_peak_sustained_elapsed
lts__t_bytes.sum l1tex__throughput.avg.pct_of_peak_sustained_elapsed l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active smsp__warps_eligible.avg.per_cycle_active
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio sm__cycles_elapsed.avg.per_second
sm__cycles_active.avg sm__cycles_active.max""".split()


def to_bytes(v):
    x, u = float(v[0]), v[1].lower()
    return x * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}[u]


raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units_row, vals = rows[0], rows[1], rows[2]
out = {"_what": what, "_units": units, "kernel": vals[hdr.index("Kernel Name")]}
for k in KEYS:
    if k in hdr:
        out[k] = [vals[hdr.index(k)], units_row[hdr.index(k)]]
rd, wr = to_bytes(out["dram__bytes_read.sum"]), to_bytes(out["dram__bytes_write.sum"])
out["_dram_bytes_per_unit"] = (rd + wr) / units
out["_warp_instructions_per_unit"] = float(out["smsp__inst_executed.sum"][0]) / units
json.dump(out, open(os.path.join(ROOT, "profiles", f"{name}_ncu_summary.json"), "w"), indent=1)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
tmp = f"/tmp/{name}_sass.csv"
open(tmp, "w").write(src)
reg = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_regions.py"), tmp, "0.45"], capture_output=True, text=True).stdout
open(os.path.join(ROOT, "profiles", f"{name}_regions.txt"), "w").write(reg)
if "--traverse" in sys.argv:
    wl = sys.argv[sys.argv.index("--traverse") + 1]
    json.dump({"workload": wl, "kernel": out["kernel"], "queries": units, "capture": what,
               "summary": f"profiles/{name}_ncu_summary.json",
               "warp_instructions_per_query": out["_warp_instructions_per_unit"],
               "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_query": out["_dram_bytes_per_unit"],
               "issue_active_pct": float(out["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]),
               "threads_per_inst": float(out["smsp__thread_inst_executed_per_inst_executed.ratio"][0]),
               "l1tex_throughput_pct": float(out["l1tex__throughput.avg.pct_of_peak_sustained_elapsed"][0]),
               "lts_throughput_pct": float(out["lts__throughput.avg.pct_of_peak_sustained_elapsed"][0]),
               "duration_ms_under_ncu": float(out["gpu__time_duration.sum"][0]) * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "msecond": 1, "ms": 1, "nsecond": 1e-6, "second": 1e3}.get(out["gpu__time_duration.sum"][1], 1),
               "registers": int(float(out["launch__registers_per_thread"][0])),
               "warps_active_pct": float(out["sm__warps_active.avg.pct_of_peak_sustained_active"][0])},
              open(os.path.join(ROOT, "profiles", f"r2_traverse_{wl}_ncu.json"), "w"), indent=1)
print(json.dumps(out, indent=1)[:800])
