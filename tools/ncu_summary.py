"""Summarise an .ncu-rep (ncu --set full) into markdown + a traffic json: python tools/ncu_summary.py rep out.md [traffic.json]"""
import csv, io, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__warps_eligible.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]
lines = [f"# ncu --set full summary of `{rep.split('/')[-1]}`", ""]
traffic = None
for r in rows[2:]:
    d = dict(zip(hdr, r))
    lines.append(f"## {d.get('Kernel Name','?')[:110]}  (launch id {d.get('ID','?')})")
    lines.append("")
    lines.append("| metric | value | unit |"); lines.append("|---|---|---|")
    for k in want:
        if k in d:
            lines.append(f"| {k} | {d[k]} | {units[hdr.index(k)]} |")
    lines.append("")
    try:
        def tobytes(k):
            v = float(d[k].replace(",", "")); u = units[hdr.index(k)].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
        traffic = tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum")
    except Exception:
        pass
open(out, "w").write("\n".join(lines) + "\n")
if len(sys.argv) > 3 and traffic is not None:
    json.dump({"dram_bytes_per_launch": traffic, "source": rep.split("/")[-1]}, open(sys.argv[3], "w"))
print("wrote", out, "traffic", traffic)
