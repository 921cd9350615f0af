#!/usr/bin/env python
"""Per-source-line totals of an ncu report's source page (needs -lineinfo):

    python tools/ncu_source_lines.py report.ncu-rep [launch-index] [top]

prints executed warp instructions, stall samples and shared-memory wavefronts per CUDA source line, largest first."""
import csv
import subprocess
import sys
from collections import defaultdict


SORT = 1


def main(path, launch=0, top=45):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    blocks, cur = [], []
    for line in out.splitlines():
        if line.startswith('"File Path"') and cur:
            blocks.append(cur)
            cur = []
        cur.append(line)
    blocks.append(cur)
    # blocks alternate per launch: take those of the chosen launch (same kernel captured several times repeats files)
    per_line = defaultdict(lambda: [0, 0, 0, 0, ""])
    seen_files = set()
    n_launch = -1
    for blk in blocks:
        rows = list(csv.reader(blk))
        fpath = rows[0][1]
        if fpath in seen_files and rows[0][0] == "File Path":
            pass
        hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
        hdr = rows[hdr_i]
        if "Instructions Executed" not in hdr:
            continue
        ie, ss = hdr.index("Instructions Executed"), hdr.index("# Samples")
        ws = hdr.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hdr else None
        key0 = (fpath,)
        if fpath.endswith("logmel.cu") or True:
            for r in rows[hdr_i + 1:]:
                if len(r) <= ie or not r[0].isdigit():
                    continue
                k = (fpath.split("/")[-1], int(r[0]))
                e = per_line[k]
                try:
                    e[0] += int(float(r[ie] or 0))
                    e[1] += int(float(r[ss] or 0))
                    if ws is not None:
                        e[2] += int(float(r[ws] or 0))
                except ValueError:
                    continue
                e[4] = r[1][:100]
    tot = sum(v[0] for v in per_line.values()) or 1
    tots = sum(v[1] for v in per_line.values()) or 1
    print(f"total warp instructions {tot}, samples {tots} (all captured launches summed)")
    for k, v in sorted(per_line.items(), key=lambda kv: -kv[1][SORT])[:top]:
        print(f"{k[0]:12s}:{k[1]:4d} inst {100*v[0]/tot:5.1f}%  samples {100*v[1]/tots:5.1f}%  smem-wf {v[2]:9d}  | {v[4]}")


if __name__ == "__main__":
    if "--by-inst" in sys.argv:
        sys.argv.remove("--by-inst")
        SORT = 0
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 45)
