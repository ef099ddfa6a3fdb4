#!/usr/bin/env python
"""Summarise an .ncu-rep (one `ncu --set full` capture) into a small JSON + text file under
profiles/.  Usage: tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_summary"""
import csv
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]


def main():
    rep, outbase = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        d = {"kernel": vals[hdr.index("Kernel Name")]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                try:
                    v = float(vals[i].replace(",", ""))
                except ValueError:
                    v = vals[i]
                d[k] = [v, units[i]]
        out.append(d)
    with open(outbase + ".json", "w") as f:
        json.dump(out, f, indent=1)
    with open(outbase + ".txt", "w") as f:
        for d in out:
            f.write("kernel: %s\n" % d["kernel"])
            for k in KEYS:
                if k in d:
                    f.write("  %-85s %s %s\n" % (k, d[k][0], d[k][1]))
    print(open(outbase + ".txt").read())


if __name__ == "__main__":
    main()
