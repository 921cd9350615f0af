#!/bin/bash
# final 1-GPU measurements of a round: default bench line, reference arm, launch list and ncu captures
TAG=${1:-r02z}
mkdir -p gpurun_out
python bench.py --impl reference > gpurun_out/${TAG}_reference_arm_line.json 2> gpurun_out/${TAG}_reference.err; echo "reference rc $?"
python bench.py > gpurun_out/${TAG}_bench_line_1gpu.json 2> gpurun_out/${TAG}_bench_1gpu.err; echo "bench rc $?"
tail -2 gpurun_out/${TAG}_bench_1gpu.err
bash tools/gpu_profile.sh > gpurun_out/${TAG}_profile.log 2>&1; echo "profile rc $?"
python - <<PY
import json
for f in ("${TAG}_bench_line_1gpu","${TAG}_reference_arm_line"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["value"], d.get("ms_per_step"), d["e2e"]["value"], d.get("roofline",{}).get("frac"), d.get("clocks")); print({k:(v.get("frac"),v.get("us_per_window"),v.get("sm_mhz")) for k,v in d.get("side_kernels",{}).items()})
    except Exception as e: print(f, "ERR", e)
PY
