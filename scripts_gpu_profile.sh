#!/bin/bash
# launch list (per-launch device time) and one full capture of the top kernel for a short bench run
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --hours 0.15 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc5 -s 60 -c 4 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
tail -2 gpurun_out/plain.log
ls -la gpurun_out
