#!/usr/bin/env python
"""Summarise the launches of an `ncu --set full` report into the counters the roofline argument uses.

    python tools/ncu_full_summary.py gpurun_out/prof_gemm.ncu-rep > profiles/rNN_prof_gemm_ncu_full.txt
"""
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
    "launch__cluster_size", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
]


def main(path: str) -> None:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    print(f"{'Kernel Name':88s} {'':18s} {[r[ki][:26] for r in data]}")
    for m in METRICS:
        if m not in hdr:
            continue
        i = hdr.index(m)
        print(f"{m:88s} {units[i]:18s} {[r[i] for r in data]}")


if __name__ == "__main__":
    main(sys.argv[1])
