"""Print selected raw metrics of the kernels in an ncu report (development tool):
    python tools/ncu_keys.py <report.ncu-rep> <kernel substring> [metric substring ...]"""
import csv, io, subprocess, sys
rep, kern, subs = sys.argv[1], sys.argv[2], sys.argv[3:]
DEFAULT = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct", "issue_stalled", "op_dfma_pred_on.sum.per_cycle_elapsed",
           "op_dmul_pred_on.sum.per_cycle_elapsed", "op_dadd_pred_on.sum.per_cycle_elapsed", "smsp__inst_executed.sum", "bank_conflicts_pipe_lsu_mem_shared.sum",
           "wavefronts_mem_shared.sum ", "dram__bytes_read.sum", "dram__bytes_write.sum", "local_op_ld.sum", "local_op_st.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
           "l1tex__data_pipe_lsu_wavefronts.avg.pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
subs = subs or DEFAULT
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, u = rows[0], rows[1]
ki = h.index("Kernel Name")
for r in rows[2:]:
    if kern not in r[ki]:
        continue
    print("==", r[ki][:90])
    for k, v, un in zip(h, r, u):
        if any(s.strip() in k for s in subs) and "pcsamp" not in k and "not_issued" not in k and v not in ("", "0"):
            print(f"   {k} = {v} {un}")
