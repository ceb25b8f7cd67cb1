"""Turns an .ncu-rep capture of the traversal kernel into the summaries committed under profiles/:
   python tools/ncu_summarise.py <rep> <tag> <queries> "<what>" [--issue]   (ncu must be on PATH)
writes profiles/r1_traverse_<tag>_ncu_summary.json, ..._regions.txt, and with --issue refreshes
profiles/traverse_issue_profile.json + traverse_dram_traffic.json (read by bench.py)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, tag, queries, what = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
KEYS = """gpu__time_duration.sum launch__grid_size launch__block_size launch__registers_per_thread
launch__occupancy_limit_registers launch__occupancy_limit_shared_mem sm__warps_active.avg.pct_of_peak_sustained_active
smsp__issue_active.avg.pct_of_peak_sustained_active smsp__inst_executed.sum smsp__thread_inst_executed_per_inst_executed.ratio
sm__inst_executed.sum.per_cycle_active dram__bytes_read.sum dram__bytes_write.sum
dram__bytes_read.sum.pct_of_peak_sustained_elapsed dram__bytes_write.sum.pct_of_peak_sustained_elapsed
lts__t_sector_hit_rate.pct l1tex__t_sector_hit_rate.pct lts__throughput.avg.pct_of_peak_sustained_elapsed
l1tex__throughput.avg.pct_of_peak_sustained_elapsed sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active smsp__warps_eligible.avg.per_cycle_active
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio sm__cycles_elapsed.avg.per_second
sm__cycles_active.avg sm__cycles_active.max""".split()
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
out = {"_what": what, "kernel": vals[hdr.index("Kernel Name")]}
for k in KEYS:
    if k in hdr:
        out[k] = [vals[hdr.index(k)], units[hdr.index(k)]]
json.dump(out, open(os.path.join(ROOT, "profiles", f"r1_traverse_{tag}_ncu_summary.json"), "w"), indent=1)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
tmp = f"/tmp/{tag}_sass.csv"
open(tmp, "w").write(src)
reg = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_regions.py"), tmp, "0.45"], capture_output=True, text=True).stdout
open(os.path.join(ROOT, "profiles", f"r1_traverse_{tag}_regions.txt"), "w").write(reg)
if "--issue" in sys.argv:
    inst = float(out["smsp__inst_executed.sum"][0])
    json.dump({"warp_instructions_per_query": inst / queries, "queries": queries, "smsp__inst_executed.sum": inst,
               "issue_active_pct": float(out["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]),
               "source": f"profiles/r1_traverse_{tag}_ncu_summary.json ({what})"},
              open(os.path.join(ROOT, "profiles", "traverse_issue_profile.json"), "w"), indent=1)
    def to_bytes(v):
        x, u = float(v[0]), v[1].lower()
        return x * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
    rd, wr = to_bytes(out["dram__bytes_read.sum"]), to_bytes(out["dram__bytes_write.sum"])
    json.dump({"kernel": out["kernel"], "capture": f"{what}; profiles/r1_traverse_{tag}_ncu_summary.json", "queries": queries,
               "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_query": (rd + wr) / queries,
               "note": "per query: ~16 B point + ~3 B node compulsory reads, 80 B results written as two 40 B rows "
                       "(sector-granular scatter => ~124 B written)"},
              open(os.path.join(ROOT, "profiles", "traverse_dram_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1)[:600])
