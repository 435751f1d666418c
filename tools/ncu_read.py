"""Reads the `ncu --page raw --csv` exports under gpurun_out/ (or profiles/) and prints the metrics the design notes
quote. usage: python tools/ncu_read.py <prefix> name [name ...]     e.g.  gpurun_out/r2c_ tetration stft"""
import csv, io, sys, json

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__occupancy_limit_registers",
 "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
 "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
 "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
 "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum",
 "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
 "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
 "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
 "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
 "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
 "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
 "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
 "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"]


def raw(path):
    rows = [l for l in open(path) if not l.startswith("==")]
    r = list(csv.reader(io.StringIO("".join(rows))))
    names, units, vals = r[0], r[1], r[2]
    return {n: (v, u) for n, u, v in zip(names, units, vals)}


if __name__ == "__main__":
    prefix = sys.argv[1]
    for name in sys.argv[2:]:
        d = raw(f"{prefix}{name}_raw.csv")
        print("==", name, d.get("Kernel Name", ("", ""))[0][:80])
        for k in KEYS:
            if k in d:
                print(f"   {k:88s} {d[k][0]:>18s} {d[k][1]}")
