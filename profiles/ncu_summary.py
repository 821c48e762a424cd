"""Key figures of one `ncu --set full` capture (.ncu-rep -> text).  Usage: python profiles/ncu_summary.py file.ncu-rep"""
import csv
import subprocess
import sys

KEYS = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__cycles_active.avg",
        "sm__inst_executed_pipe_xu.sum", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
        "smsp__average_warp_latency_issue_stalled_wait.ratio", "smsp__average_warp_latency_issue_stalled_no_instruction.ratio",
        "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "smsp__average_warp_latency_issue_stalled_membar.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def main(path):
    out = subprocess.check_output(["ncu", "-i", path, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(out.splitlines()))
    head, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[head.index("Kernel Name")])
        for k in KEYS:
            if k in head:
                i = head.index(k)
                print(f"  {k:88s} {r[i]:>18s} {units[i]}")
        rd, wr = (float(r[head.index(k)].replace(",", "")) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        ru, wu = units[head.index("dram__bytes_read.sum")], units[head.index("dram__bytes_write.sum")]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        print(f"  dram traffic (read + write) per launch: {(rd * scale[ru] + wr * scale[wu]) / 1e6:.1f} MB")


if __name__ == "__main__":
    main(sys.argv[1])
