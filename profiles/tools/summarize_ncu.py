#!/usr/bin/env python
"""Extract the handful of ncu metrics DESIGN.md / bench.py rely on from a .ncu-rep (run HERE, no GPU).

    python profiles/tools/summarize_ncu.py gpurun_out/prof.ncu-rep [units_per_launch] > profiles/<name>.json
"""
import csv
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__registers_per_thread": "regs",
    "launch__shared_mem_per_block_dynamic": "smem_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pipe_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pipe_pct",
    "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active": "adu_pipe_pct",
    "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active": "cbu_pipe_pct",
    "smsp__inst_executed.sum": "warp_inst",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "active_threads_per_inst",
    "sm__cycles_elapsed.avg": "cycles_elapsed",
    "smsp__cycles_active.avg": "cycles_active",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "lts__t_bytes.sum": "l2_bytes",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_throttle",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_sb",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_sb",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio": "stall_branch",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio": "stall_dispatch",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio": "stall_not_selected",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio": "stall_mio",
}
WITH_UNIT = ("duration", "dram_read", "dram_write", "l2_bytes", "smem_dynamic")


def main():
    rep = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(raw.splitlines()) if len(r) > 10]
    hdr, unit = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, unit))
        e = {"kernel": d.get("Kernel Name", "")[:120]}
        for k, name in KEYS.items():
            if k in d and d[k] != "":
                try:
                    e[name] = float(d[k].replace(",", ""))
                except ValueError:
                    e[name] = d[k]
                if u.get(k) and name in WITH_UNIT:
                    e[name + "_unit"] = u[k]
        if units and "warp_inst" in e:
            e["units_per_launch"] = units
            e["thread_inst_per_unit"] = e["warp_inst"] * e.get("active_threads_per_inst", 32.0) / units
            # FP64 pipe: 64 lanes per SM; busy lane-slots (incl. predicated-off lanes of a partially active warp)
            if "fp64_pipe_pct" in e and "cycles_active" in e:
                e["fp64_lane_slots_per_unit"] = e["fp64_pipe_pct"] / 100.0 * e["cycles_active"] * 148 * 64 / units
        out.append(e)
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
