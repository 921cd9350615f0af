#!/bin/bash
# default bench line, reference arm and a small --corpus-hours run on one GPU
mkdir -p gpurun_out
TAG=${1:-r02z}
python bench.py > gpurun_out/${TAG}_bench_line_1gpu.json 2> gpurun_out/${TAG}_bench_1gpu.err; echo "bench rc $?"
tail -3 gpurun_out/${TAG}_bench_1gpu.err
python bench.py --workload corpus --corpus-hours 20 --steps 1 --warmup 1 > gpurun_out/${TAG}_corpus20h_line_1gpu.json 2> gpurun_out/${TAG}_corpus20h_1gpu.err; echo "corpus rc $?"
tail -3 gpurun_out/${TAG}_corpus20h_1gpu.err
python - <<PY
import json
for f in ("${TAG}_bench_line_1gpu","${TAG}_corpus20h_line_1gpu"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("roofline",{}).get("frac"), d["clocks"]); print({k:(v.get("frac"),v.get("us_per_window")) for k,v in d.get("side_kernels",{}).items()}); print(d["config"])
    except Exception as e: print(f, "ERR", e)
PY
