"""Summarise ncu reports into profiles/*.md:  python tools/ncu_summary.py <report.ncu-rep> [...]  (run where ncu is installed)."""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe % of peak (active)"),
    ("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "DFMA thread-instructions"),
    ("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "DMUL thread-instructions"),
    ("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "DADD thread-instructions"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long scoreboard / issue"),
    ("derived__smsp__sass_thread_inst_executed_op_dfma_pred_on_x2", "DFMA x2"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers, blocks)"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads per instruction"),
    ("local_load/store: smsp__inst_executed_op_local_ld.sum", ""),
    ("smsp__inst_executed_op_local_ld.sum", "local loads (spills)"),
    ("smsp__inst_executed_op_local_st.sum", "local stores (spills)"),
]


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    for rep in sys.argv[1:]:
        hdr, units, data = rows_of(rep)
        ki = hdr.index("Kernel Name")
        print(f"## {rep}\n")
        seen = {}
        for r in data:
            name = r[ki].split("(")[0]
            seen.setdefault(name, []).append(r)
        for name, rs in seen.items():
            r = rs[-1]          # last captured launch of the kernel
            print(f"### `{name}`  ({len(rs)} launches captured; last one shown)\n")
            print("| metric | value | unit |\n|---|---|---|")
            for key, label in WANT:
                if key in hdr and label:
                    i = hdr.index(key)
                    print(f"| {label} (`{key}`) | {r[i]} | {units[i]} |")
            print()


if __name__ == "__main__":
    main()
