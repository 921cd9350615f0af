#!/bin/bash
# BASELINE config 4 at full size: tools/multi_gpu_corpus_full.sh N [hours] -- a corpus of `hours` (1000) hours of
# synthetic audio, files with log-normal durations, sharded by file over N GPUs, one final interval all-gather
N=$1
H=${2:-1000}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 1 --warmup 1 --workload corpus --corpus-hours $H > gpurun_out/r02z_corpus${H}h_${N}gpu.json 2> gpurun_out/r02z_corpus${H}h_${N}gpu.err
echo "rc $?"; tail -3 gpurun_out/r02z_corpus${H}h_${N}gpu.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02z_corpus${H}h_${N}gpu.json")); print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"], d["scaling"], d["clocks"], d["config"]["workload"][:120], d.get("intervals_per_step"))
except Exception as e: print("ERR", e)
PY
