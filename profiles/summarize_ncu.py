"""Print the handful of `ncu --set full` metrics the round reports quote, one block per profiled launch.

    python profiles/summarize_ncu.py gpurun_out/prof_x.ncu-rep > profiles/r01_x_ncu_summary.txt
"""
import csv
import re
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg",
]
STALL = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio")


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(r[hdr.index("Kernel Name")])
        for h in WANT:
            if h in hdr:
                print(f"  {h} = {r[hdr.index(h)]} {units[hdr.index(h)]}")
        stalls = sorted(((float(r[i]), STALL.match(h).group(1)) for i, h in enumerate(hdr) if STALL.match(h) and r[i] not in ("", "n/a")), reverse=True)
        print("  top stalls (warps per issue-active cycle): " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:5]))


if __name__ == "__main__":
    main()
