"""Summarise an .ncu-rep (read here, no GPU needed): python scripts/ncu_summary.py report.ncu-rep [n_top_lines]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe", "sm__inst_executed_pipe_uniform",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__inst_issued.m",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__average_warps_issue_stalled"]
for vals in rows[2:]:
    name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print("==", name[:100])
    for h, u, v in zip(hdr, units, vals):
        if any(h.startswith(w) for w in WANT) and not any(x in h for x in (".max", ".min", ".sum.p", "per_second", "pct_of_peak_sustained_elapsed")):
            if h.startswith("smsp__average_warps_issue_stalled") and float(v or 0) < 0.05:
                continue
            print(f"  {h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if rows:
    # find header row
    hi = next(i for i, r in enumerate(rows) if "Source" in r)
    h = rows[hi]
    def col(n):
        return h.index(n) if n in h else None
    c_src, c_samp, c_inst = col("Source"), col("# Samples") or col("Warp Stall Sampling (All Samples)"), col("# Instructions Executed") or col("Instructions Executed")
    c_addr = col("Address")
    data = []
    for r in rows[hi + 1:]:
        try:
            data.append((int(r[c_samp] or 0), int(r[c_inst] or 0) if c_inst is not None else 0, r[c_src], r[c_addr] if c_addr is not None else ""))
        except Exception:
            pass
    tot = sum(d[0] for d in data) or 1
    print(f"-- top {ntop} SASS lines by stall samples (total {tot})")
    for s, n, t, a in sorted(data, reverse=True)[:ntop]:
        print(f"  {100*s/tot:5.1f}%  exec {n:9d}  {a[-5:]}  {t[:110]}")
