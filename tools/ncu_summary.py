"""Text summary of an `ncu --set full` report (run where ncu is installed; no GPU needed): per profiled launch the duration, DRAM bytes,
pipe utilisation, issue / occupancy figures and global-memory sector efficiency.  Usage: python tools/ncu_summary.py report.ncu-rep [out.txt]"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.pct", "smsp__sass_average_data_bytes_per_sector_mem_global_op_st.pct",
        "l1tex__average_t_sectors_per_request_pipe_lsu_mem_global_op_ld.ratio", "l1tex__average_t_sectors_per_request_pipe_lsu_mem_global_op_st.ratio",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]


def main(path, out=None):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"ncu --set full summary of {path.split('/')[-1]} ({len(rows) - 2} profiled launch(es)); values are per launch, measured under the profiler "
             f"(cold cache, serialised replays): use them for ratios and traffic, not as bench numbers"]
    for r in rows[2:]:
        get = lambda k: r[hdr.index(k)] if k in hdr else "?"
        lines.append("")
        lines.append(f"kernel {get('Kernel Name')}  grid {get('Grid Size')} block {get('Block Size')}")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                lines.append(f"  {w:88s} {r[i]:>16s} {units[i]}")
    text = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(text)
    else:
        print(text)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
