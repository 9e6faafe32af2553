"""Per-kernel table from an `ncu --metrics ... --csv` log of the segment kernels (run here, no GPU needed):
python tools/ncu_segments.py gpurun_out/x_metrics.csv SAMPLES"""
import collections
import csv
import sys


def main(path, samples):
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
    d = collections.OrderedDict()
    for r in rows:
        d.setdefault((int(r[0]), r[4]), {})[r[12]] = float(r[14].replace(",", ""))
    g = lambda m, k: m.get(k, float("nan"))  # noqa: E731
    print("%-10s %8s %8s %8s %5s %6s %6s %6s %6s %6s %6s" % ("kernel", "us", "rd MB", "wr MB", "regs", "fp64%", "issue%", "warps%", "dram%", "lsb", "noinst"))
    T = RD = WR = 0.0
    for (i, k), m in d.items():
        t = g(m, "gpu__time_duration.sum") / 1e3
        T += t
        RD += g(m, "dram__bytes_read.sum")
        WR += g(m, "dram__bytes_write.sum")
        print("%-10s %8.1f %8.1f %8.1f %5d %6.1f %6.1f %6.1f %6.1f %6.2f %6.2f" % (
            k, t, g(m, "dram__bytes_read.sum") / 1e6, g(m, "dram__bytes_write.sum") / 1e6, g(m, "launch__registers_per_thread"),
            g(m, "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active"), g(m, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            g(m, "sm__warps_active.avg.pct_of_peak_sustained_active"), g(m, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            g(m, "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
            g(m, "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio")))
    print(f"# {len(d)} kernels, total {T:.1f} us (cold-cache, serialised), dram read {RD / 1e9:.3f} GB + write {WR / 1e9:.3f} GB "
          f"= {(RD + WR) / samples:.0f} B/sample over {samples} samples; {(RD + WR) / T / 1e3:.0f} GB/s while the kernels run")
    return RD + WR


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]))
