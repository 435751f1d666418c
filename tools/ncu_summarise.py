"""Condenses `ncu --page raw --csv` exports (gpurun_out/<prefix><name>_raw.csv) into the small JSON files under
profiles/ that DESIGN.md and bench.py quote. usage: python tools/ncu_summarise.py <prefix> <out.json> name=label ..."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))
from ncu_read import raw

PICK = {
    "gpu__time_duration.sum": "time",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__occupancy_limit_registers": "ctas_per_sm_by_registers",
    "launch__occupancy_limit_shared_mem": "ctas_per_sm_by_shared_memory",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "active_lanes_per_instruction",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "shared_wavefronts",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "shared_wavefronts_pct_of_peak",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "shared_bank_conflicts",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1tex_throughput_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio": "stall_mio_throttle",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe_throttle",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio": "stall_no_instruction",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio": "stall_not_selected",
}
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}


def number(value: str, unit: str):
    try:
        x = float(value.replace(",", ""))
    except ValueError:
        return value
    if unit in UNIT:
        x *= UNIT[unit]
    return x


def summarise(path: str) -> dict:
    d = raw(path)
    out = dict(kernel=d.get("Kernel Name", ("", ""))[0])
    for key, name in PICK.items():
        if key in d:
            v = number(*d[key])
            if name == "time":
                out["time_ms"] = v
            elif name in ("dram_read", "dram_write"):
                out[name + "_bytes"] = v
            else:
                out[name] = v
    if "dram_read_bytes" in out:
        out["dram_bytes_per_launch"] = out["dram_read_bytes"] + out.get("dram_write_bytes", 0.0)
    # FP32 work: thread-level FADD + FMUL + 2*FFMA; the raw page gives them per elapsed SM cycle
    per_cycle = {op: float(d[f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed"][0].replace(",", ""))
                 for op in ("fadd", "fmul", "ffma") if f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed" in d}
    cycles = d.get("smsp__cycles_elapsed.max") or d.get("sm__cycles_elapsed.max")
    if per_cycle and cycles:
        c = float(cycles[0].replace(",", ""))
        flop = (per_cycle.get("fadd", 0) + per_cycle.get("fmul", 0) + 2*per_cycle.get("ffma", 0))*c
        out["fp32_flop_per_launch"] = flop
        if out.get("time_ms"):
            out["fp32_tflops_under_ncu"] = flop/(out["time_ms"]/1e3)/1e12
    return out


if __name__ == "__main__":
    prefix, target = sys.argv[1], Path(sys.argv[2])
    result = json.loads(target.read_text()) if target.exists() else {}
    for item in sys.argv[3:]:
        name, _, label = item.partition("=")
        result[label or name] = dict(summarise(f"{prefix}{name}_raw.csv"), source=f"{Path(prefix).name}{name}_raw.csv (ncu --set full --clock-control none, one launch)")
    target.write_text(json.dumps(result, indent=1))
    print(json.dumps({k: {a: b for a, b in v.items() if a in ("time_ms", "issue_active_pct", "shared_wavefronts_pct_of_peak", "fp32_tflops_under_ncu", "dram_bytes_per_launch")} for k, v in result.items()}, indent=1))
