"""Condenses an `ncu --page raw --csv` export into one block per kernel launch with the metrics the
rooflines in DESIGN.md / bench.py cite.  Usage: python tools/ncu_summary.py raw.csv > summary.txt"""
import csv
import re
import sys

KEYS = [
    "gpu__time_duration.sum",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = re.sub(r"\(anonymous namespace\)::|gpfx::|void ", "", r[idx["Kernel Name"]])
        print(f"== {name[:110]}  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}")
        rd = wr = dur = None
        for k in KEYS:
            if k in idx:
                v = r[idx[k]]
                print(f"   {k:86s} {v} {units[idx[k]]}")
                if k == "dram__bytes_read.sum":
                    rd = (float(v), units[idx[k]])
                if k == "dram__bytes_write.sum":
                    wr = (float(v), units[idx[k]])
                if k == "gpu__time_duration.sum":
                    dur = (float(v), units[idx[k]])
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tscale = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}
        if rd and wr and dur and rd[1] in scale and wr[1] in scale and dur[1] in tscale:
            b = rd[0] * scale[rd[1]] + wr[0] * scale[wr[1]]
            print(f"   {'=> DRAM traffic / duration':86s} {b / 1e6:.2f} MB, {b / (dur[0] * tscale[dur[1]]) / 1e9:.1f} GB/s")


if __name__ == "__main__":
    main()
