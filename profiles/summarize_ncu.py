"""Reads an .ncu-rep here (no GPU needed) and prints the counters DESIGN.md / bench.py's roofline block cite.
usage: python profiles/summarize_ncu.py gpurun_out/r01_syrk_c3.ncu-rep > profiles/r01_syrk_c3.summary.txt"""
import csv
import subprocess
import sys

KEYS = (
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
    "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
)


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print("kernel:", d.get("Kernel Name"))
        for k in hdr:
            if k in KEYS or k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
                print("  %-90s %s %s" % (k, d[k], u.get(k, "")))
        rd, wr = d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum")
        print()


if __name__ == "__main__":
    main()
