"""Prints the metrics that matter from an .ncu-rep (run here, no GPU needed): python tools/ncu_summary.py file.ncu-rep"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.avg", "smsp__average_warp_latency_per_inst_issued.ratio", "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        print("==", d.get("Kernel Name", ("?",))[0][:90])
        for k in KEYS:
            if k in d:
                print(f"  {k:78s} {d[k][0]:>18s} {d[k][1]}")
        st = sorted(((float(v[0].replace(",", "")), h) for h, v in d.items()
                     if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")), reverse=True)
        print("  stalls (warps per issue):", ", ".join(f"{h.split('stalled_')[1].split('_per')[0]}={v:.2f}" for v, h in st[:8]))


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
