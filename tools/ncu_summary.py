"""Extracts the judged counters from an `ncu --set full` report into a small JSON file.
usage: python tools/ncu_summary.py report.ncu-rep out.json [label]"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_issued.avg.per_cycle_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__icc_request_hit_rate.pct", "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_active.avg",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    label = sys.argv[3] if len(sys.argv) > 3 else rep
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    kernels = []
    for row in rows[2:]:
        d = {"kernel": row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"}
        for h, u, v in zip(hdr, units, row):
            if h in KEYS:
                try:
                    d[h] = {"value": float(v.replace(",", "")), "unit": u}
                except ValueError:
                    d[h] = {"value": v, "unit": u}
        if "dram__bytes_read.sum" in d and "dram__bytes_write.sum" in d:
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            d["traffic_bytes"] = d["dram__bytes_read.sum"]["value"] * mult.get(d["dram__bytes_read.sum"]["unit"], 1) + \
                d["dram__bytes_write.sum"]["value"] * mult.get(d["dram__bytes_write.sum"]["unit"], 1)
        kernels.append(d)
    json.dump({"label": label, "source": "ncu --set full --clock-control none --import-source on", "kernels": kernels},
              open(out, "w"), indent=1)
    print("wrote", out, len(kernels), "kernel(s)")


if __name__ == "__main__":
    main()
