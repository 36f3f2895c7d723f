#!/usr/bin/env python
"""Summarises an ncu report (.ncu-rep, `--set full`) into a small text table for profiles/.

    python tools/ncu_summary.py gpurun_out/<tag>_full.ncu-rep > profiles/<tag>_ncu_summary.txt
"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu_dram_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_static", "smem_static"),
    ("launch__occupancy_limit_registers", "occ_lim_regs"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved_occupancy_pct"),
    ("smsp__issue_active.avg.per_cycle_active", "issue_per_cycle"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu_wavefront_pct"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
    ("smsp__sass_inst_executed_op_global_ld.sum", "global_ld_insts"),
    ("smsp__sass_inst_executed_op_global_st.sum", "global_st_insts"),
    ("smsp__sass_inst_executed_op_shared_ld.sum", "shared_ld_insts"),
    ("smsp__sass_inst_executed_op_shared_st.sum", "shared_st_insts"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global_ld_sectors"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall_mio_throttle"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg_throttle"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall_not_selected"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# {rep}: ncu --set full --clock-control none (per-launch values; cold-cache, serialised)")
    for r in rows[2:]:
        print(f"\n== {r[col['Kernel Name']][:110]}")
        for k, nice in KEYS:
            if k in col and r[col[k]] != "":
                print(f"  {nice:26s} {r[col[k]]:>18s} {units[col[k]]}")
        try:
            rd = float(r[col["dram__bytes_read.sum"]]); wr = float(r[col["dram__bytes_write.sum"]])
            t = float(r[col["gpu__time_duration.sum"]])
            ur, ut = units[col["dram__bytes_read.sum"]], units[col["gpu__time_duration.sum"]]
            scale_b = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[ur]
            scale_t = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}[ut]
            print(f"  {'dram_traffic_bytes':26s} {(rd + wr) * scale_b:18.0f} byte")
            print(f"  {'dram_GB_per_s':26s} {(rd + wr) * scale_b / (t * scale_t) / 1e9:18.1f} GB/s (under ncu)")
        except Exception:
            pass


if __name__ == "__main__":
    main()
