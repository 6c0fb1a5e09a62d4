#!/usr/bin/env python
"""Turn an ``ncu --set full`` report (gpurun_out/*.ncu-rep, scratch) into a small text summary that
is committed under profiles/.  Usage: python profiles/summarize_ncu.py <report.ncu-rep> <out.md> [title]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fma.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else rep
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# {title}\n\nSource: `ncu --set full --clock-control none --import-source on` (one launch per kernel, replayed by ncu;\n"
                "durations under ncu are cold-cache and serialised -- shares, not absolutes).\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write(f"\n## {d.get('Kernel Name', '?')}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in d and d[k] != "":
                    f.write(f"| {k} | {d[k]} | {units[hdr.index(k)]} |\n")
            rd, wr = d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum")
            if rd and wr:
                ur, uw = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
                f.write(f"| dram traffic (read + write) | {rd} {ur} + {wr} {uw} | |\n")
    print(out)


if __name__ == "__main__":
    main()
