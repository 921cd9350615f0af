#!/usr/bin/env python
"""Per-source-line totals of an ncu report's source page (the kernels are built with -lineinfo):

    python tools/ncu_source_lines.py report.ncu-rep [top] [--by-inst]

prints, for every CUDA source line, its share of the executed warp instructions, its share of the warp-stall samples
and its shared-memory wavefronts, summed over the captured launches; sorted by stall samples, or by instructions with
--by-inst.  This is how the hot spots of the fused log-mel kernel were found (cluster-barrier waits, 64-bit scratch
stores that touched 32 sectors per instruction, bank conflicts of the mel taps)."""
import csv
import subprocess
import sys
from collections import defaultdict


def per_line(path: str) -> dict:
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    blocks, cur = [], []
    for line in out.splitlines():  # one block per source file
        if line.startswith('"File Path"') and cur:
            blocks.append(cur)
            cur = []
        cur.append(line)
    blocks.append(cur)
    acc = defaultdict(lambda: [0, 0, 0, ""])  # (file, line) -> [instructions, samples, smem wavefronts, text]
    for blk in blocks:
        rows = list(csv.reader(blk))
        if not rows or len(rows[0]) < 2:
            continue
        fname = rows[0][1].split("/")[-1]
        hdr_i = next((i for i, r in enumerate(rows) if r and r[0] == "Line No"), None)
        if hdr_i is None or "Instructions Executed" not in rows[hdr_i]:
            continue
        hdr = rows[hdr_i]
        ie, ss = hdr.index("Instructions Executed"), hdr.index("# Samples")
        ws = hdr.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hdr else None
        for r in rows[hdr_i + 1:]:
            if len(r) <= ie or not r[0].isdigit():
                continue
            e = acc[(fname, int(r[0]))]
            try:
                e[0] += int(float(r[ie] or 0))
                e[1] += int(float(r[ss] or 0))
                if ws is not None:
                    e[2] += int(float(r[ws] or 0))
            except ValueError:
                continue
            e[3] = r[1][:100]
    return acc


def main(path: str, top: int = 45, by_inst: bool = False) -> None:
    acc = per_line(path)
    tot_i = sum(v[0] for v in acc.values()) or 1
    tot_s = sum(v[1] for v in acc.values()) or 1
    print(f"total warp instructions {tot_i}, samples {tot_s} (all captured launches summed)")
    key = 0 if by_inst else 1
    for k, v in sorted(acc.items(), key=lambda kv: -kv[1][key])[:top]:
        print(f"{k[0]:12s}:{k[1]:4d} inst {100 * v[0] / tot_i:5.1f}%  samples {100 * v[1] / tot_s:5.1f}%  smem-wf {v[2]:9d}  | {v[3]}")


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if a != "--by-inst"]
    nums = [int(a) for a in args[1:] if a.isdigit()]
    main(args[0], nums[-1] if nums else 45, "--by-inst" in sys.argv)
