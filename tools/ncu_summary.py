"""Summarise an .ncu-rep (read on the CPU box): key raw metrics, stall mix, hottest source lines.
   python tools/ncu_summary.py gpurun_out/x.ncu-rep [topN]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30


def run(*a):
    return subprocess.run(["ncu", "-i", rep, *a], capture_output=True, text=True).stdout


rows = list(csv.reader(io.StringIO(run("--page", "raw", "--csv"))))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
        "lts__t_bytes.sum", "launch__grid_size", "launch__block_size", "smsp__warps_eligible.avg.per_cycle_active"]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:70])
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w:68s} {r[i]:>16s} {units[i]}")
rows = list(csv.reader(io.StringIO(run("--page", "source", "--csv", "--print-source", "cuda,sass"))))
agg, cur, tot_i, tot_s = collections.OrderedDict(), None, 0.0, 0.0
stall_cols, stalls = {}, collections.Counter()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        stall_cols = {h: i for i, h in enumerate(r) if h.startswith("stall_") and "Not Issued" not in h}
        continue
    if r[0].isdigit():
        try:
            inst, samp = float(r[7]), float(r[6])
        except ValueError:
            continue
        agg[(cur, int(r[0]))] = (inst, samp, r[1].strip())
        tot_i += inst
        tot_s += samp
        for h, i in stall_cols.items():
            try:
                stalls[h] += float(r[i])
            except (ValueError, IndexError):
                pass
print(f"-- warp instructions {tot_i:.3e}, samples {tot_s:.0f}")
print("-- stall mix:", ", ".join(f"{k[6:]} {100 * v / max(tot_s, 1):.0f}%" for k, v in stalls.most_common(9)))
byf = collections.Counter()
for (f, l), (i, s, src) in agg.items():
    byf[f] += i
print("-- by file:", ", ".join(f"{f} {100 * v / tot_i:.0f}%" for f, v in byf.most_common()))
for (f, l), (i, s, src) in sorted(agg.items(), key=lambda x: -x[1][0])[:topn]:
    print(f"{f:16s}{l:5d} {100 * i / tot_i:5.1f}% inst {100 * s / max(tot_s, 1):5.1f}% smp  {src[:96]}")
