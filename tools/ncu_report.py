#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, without a GPU) into markdown: per-launch key metrics."""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall membar"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    cols = [(hdr.index(k), label, k) for k, label in WANT if k in hdr]
    name_i = hdr.index("Kernel Name")
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on : %s\n\n" % rep)
        f.write("(profiler replays serialise and cool the caches: durations here are upper bounds; the timed\n"
                "numbers are the CUDA-event ones in README.md)\n\n")
        f.write("| kernel | " + " | ".join(l for _, l, _ in cols) + " |\n")
        f.write("|---|" + "---|" * len(cols) + "\n")
        for r in rows[2:]:
            vals = ["%s %s" % (r[i][:12], units[i]) for i, _, _ in cols]
            f.write("| `%s` | %s |\n" % (r[name_i].split("(")[0][:48], " | ".join(vals)))
    print(open(out).read())


if __name__ == "__main__":
    if len(sys.argv) != 3:
        sys.exit("usage: tools/ncu_report.py <report.ncu-rep> <summary.md>")
    main(sys.argv[1], sys.argv[2])
