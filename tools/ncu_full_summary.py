#!/usr/bin/env python
"""Summarise the launches of an `ncu --set full` report into the counters the roofline argument uses.

    python tools/ncu_full_summary.py gpurun_out/prof_gemm.ncu-rep > profiles/rNN_prof_gemm_ncu_full.txt
    python tools/ncu_full_summary.py gpurun_out/prof_gemm.ncu-rep --traffic-json profiles/r02_gemm_traffic.json "<command>"

The second form also writes the mean DRAM bytes (read + write) per captured launch: bench.py reports it as
`roofline.traffic` (ncu cannot run inside the bench; the capture is of the same code on the command given).
"""
import csv
import json
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


def _bytes(value: str, unit: str) -> float:
    v = float(value.replace(",", ""))
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def main(path: str, traffic_json: str | None = None, command: str = "") -> None:
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
    if traffic_json:
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        per = [_bytes(r[ir], units[ir]) + _bytes(r[iw], units[iw]) for r in data]
        with open(traffic_json, "w") as f:
            json.dump({"bytes_per_launch": sum(per) / len(per), "launches": [r[ki][:40] for r in data], "per_launch": per,
                       "source": f"dram__bytes_read.sum + dram__bytes_write.sum, mean over {len(per)} launches of "
                                 f"`ncu --set full --clock-control none` on `{command}` ({path.split('/')[-1]})"}, f, indent=1)
            f.write("\n")


if __name__ == "__main__":
    if len(sys.argv) > 3 and sys.argv[2] == "--traffic-json":
        main(sys.argv[1], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        main(sys.argv[1])
